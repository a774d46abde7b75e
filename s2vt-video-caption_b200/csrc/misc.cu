// Library plumbing (error buffer, launch counter) and the small memory-bound kernels of the path:
// embedding gather / dense scatter-add, column sums (bias grads), mean cross-entropy with fused
// dlogits, Adam, f32->bf16 casts.
#include "common.cuh"
#include <atomic>
#include <math.h>

namespace s2vt {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

char* err_buf() { return g_err; }
int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------- block reductions
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <bool IS_MAX>
__device__ float block_reduce(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : (IS_MAX ? -INFINITY : 0.f);
  r = IS_MAX ? warp_max(r) : warp_sum(r);
  return r;
}

// ---------------------------------------------------------------- embedding
__global__ void embed_gather_kernel(const float* __restrict__ table, int E, const int64_t* __restrict__ ids, long long ids_ld,
                                    int B, int n_t, float* __restrict__ out, long long out_ld) {
  const int row = blockIdx.x;                 // t*B + b
  const int t = row / B, b = row % B;
  const long long id = ids[(long long)b * ids_ld + t];
  const float* src = table + id * E;
  float* dst = out + (long long)row * out_ld;
  for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
}

__global__ void embed_scatter_add_kernel(float* __restrict__ grad, int E, const int64_t* __restrict__ ids, long long ids_ld,
                                         int B, int n_t, const float* __restrict__ src, long long src_ld) {
  const int row = blockIdx.x;
  const int t = row / B, b = row % B;
  const long long id = ids[(long long)b * ids_ld + t];
  float* dst = grad + id * E;
  const float* s = src + (long long)row * src_ld;
  for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, s[e]);
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// dst[(t*B + b)*N + n] = src[b*src_ld + n]: one [B,N] block repeated over n_t time steps (the constant attention context)
__global__ void bcast_rows_kernel(const float* __restrict__ src, long long src_ld, int B, int N, float* __restrict__ dst) {
  const int row = blockIdx.x;                 // t*B + b
  const float* s = src + (long long)(row % B) * src_ld;
  float* d = dst + (long long)row * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) d[n] = s[n];
}

// ---------------------------------------------------------------- column sums
// grid.x = ceil(N/32); block = 32 x 8.  Row chunks are summed in a fixed order -> deterministic.
__global__ void colsum_kernel(const float* __restrict__ X, long long M, int N, long long ld, float* __restrict__ out, int accumulate) {
  __shared__ float sh[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N)
    for (long long m = threadIdx.y; m < M; m += 8) acc += X[m * ld + n];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    out[n] = accumulate ? out[n] + s : s;
  }
}

// ---------------------------------------------------------------- cross entropy
// one CTA per row: lse = max + log(sum exp(z - max));  row_loss = lse - z[target]
__global__ void ce_row_kernel(const float* __restrict__ logits, int V, const int64_t* __restrict__ targets, RowMap tmap,
                              float* __restrict__ row_loss, float* __restrict__ dlogits, const float* __restrict__ gscale,
                              float inv_rows) {
  __shared__ float sh[32];
  const long long r = blockIdx.x;
  const float* z = logits + r * V;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, z[j]);
  mx = block_reduce<true>(mx, sh);
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf(z[j] - mx);
  s = block_reduce<false>(s, sh);
  const float lse = mx + logf(s);
  const long long tgt = targets[tmap(r)];
  if (threadIdx.x == 0) row_loss[r] = lse - z[tgt];
  __syncthreads();                      // dlogits may alias logits: z[tgt] is read before it is overwritten
  if (dlogits) {
    const float sc = (gscale ? gscale[0] : 1.f) * inv_rows;
    float* d = dlogits + r * V;
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      float pz = expf(z[j] - lse);
      if (j == tgt) pz -= 1.f;
      d[j] = pz * sc;
    }
  }
}

__global__ void mean_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sh[0] / (double)n);
}

// ---------------------------------------------------------------- Adam
// step <- step + 1 and this step's bias-corrected scalars, all on the device: a captured CUDA graph replays it with no host value baked in
__global__ void adam_prepare_kernel(int* __restrict__ step, const float* __restrict__ lr, float beta1, float beta2, float* __restrict__ hyper) {
  const int t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  hyper[0] = (float)((double)*lr / bc1);
  hyper[1] = (float)(1.0 / sqrt(bc2));
}

// A CTA handles chunks of 256 threads x 4 float4 = 4096 consecutive elements: 16-byte accesses, the four loads of every array issued
// before the first use.  Normally one chunk per CTA; under s2vt_set_bulk_cta_cap (update beside a recurrence sweep) a capped grid
// strides over the chunks.
constexpr int ADAM_ELEMS_PER_CTA = 4096;
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float beta1, float beta2, float eps,
                                                   float step_size, float inv_bc2_sqrt, float grad_scale,
                                                   __nv_bfloat16* __restrict__ shadow, const float* __restrict__ hyper) {
  if (hyper) { step_size = hyper[0]; inv_bc2_sqrt = hyper[1]; }
  for (long long base = (long long)blockIdx.x * ADAM_ELEMS_PER_CTA; base < n; base += (long long)gridDim.x * ADAM_ELEMS_PER_CTA) {
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  auto upd = [&](float gi, float& mi, float& vi, float& pi) {
    gi *= grad_scale;
    mi = beta1 * mi + omb1 * gi;
    vi = beta2 * vi + omb2 * gi * gi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!shadow || (reinterpret_cast<uintptr_t>(shadow) & 7) == 0);
  if (vec && base + ADAM_ELEMS_PER_CTA <= n) {
    float4 G[4], M[4], V[4], P[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = base + ((long long)j * 256 + threadIdx.x) * 4;
      G[j] = __ldcs(reinterpret_cast<const float4*>(g + i));
      M[j] = *reinterpret_cast<const float4*>(m + i);
      V[j] = *reinterpret_cast<const float4*>(v + i);
      P[j] = *reinterpret_cast<const float4*>(p + i);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = base + ((long long)j * 256 + threadIdx.x) * 4;
      upd(G[j].x, M[j].x, V[j].x, P[j].x); upd(G[j].y, M[j].y, V[j].y, P[j].y);
      upd(G[j].z, M[j].z, V[j].z, P[j].z); upd(G[j].w, M[j].w, V[j].w, P[j].w);
      *reinterpret_cast<float4*>(m + i) = M[j];
      *reinterpret_cast<float4*>(v + i) = V[j];
      *reinterpret_cast<float4*>(p + i) = P[j];
      if (shadow) {
        __nv_bfloat162 a = __floats2bfloat162_rn(P[j].x, P[j].y), b = __floats2bfloat162_rn(P[j].z, P[j].w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(shadow + i) = pk;
      }
    }
    continue;
  }
  const long long end = base + ADAM_ELEMS_PER_CTA < n ? base + ADAM_ELEMS_PER_CTA : n;
  for (long long i = base + threadIdx.x; i < end; i += 256) {
    float mi = m[i], vi = v[i], pi = p[i];
    upd(g[i], mi, vi, pi);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16(pi);
  }
  }
}

// ---------------------------------------------------------------- casts
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i4);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i4) = pk;
  } else {
    for (long long i = i4; i < n; ++i) dst[i] = __float2bfloat16(src[i]);
  }
}

// dst_t[c, r] = bf16(src[r, c]) through a 32x33 smem tile (and optionally dst[r,c])
__global__ void cast_transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                           __nv_bfloat16* __restrict__ dst_t, long long rows, long long cols) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long r = r0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = src[r * cols + c];
      if (dst) dst[r * cols + c] = __float2bfloat16(v);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst_t[c * rows + r] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}

}  // namespace s2vt

namespace s2vt {
static thread_local int g_bulk_cta_cap = 0;
int bulk_cta_cap() { return g_bulk_cta_cap; }
}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_set_bulk_cta_cap(int n) { g_bulk_cta_cap = n > 0 ? n : 0; return 0; }
extern "C" int s2vt_abi_version(void) { return S2VT_ABI_VERSION; }
extern "C" const char* s2vt_last_error(void) { return err_buf(); }
extern "C" int64_t s2vt_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int64_t s2vt_add_launch_count(int64_t n) { count_launch((int)n); return (int64_t)g_launches.load(); }

extern "C" int s2vt_embed_gather_f32(void* stream, const float* table, int E, const int64_t* ids, int64_t ids_ld,
                                     int B, int n_t, float* out, int64_t out_ld) {
  S2VT_REQUIRE(table && ids && out, "s2vt_embed_gather_f32: null pointer");
  if (B * n_t == 0) return 0;
  embed_gather_kernel<<<B * n_t, 128, 0, (cudaStream_t)stream>>>(table, E, ids, ids_ld, B, n_t, out, out_ld);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_embed_scatter_add_f32(void* stream, float* grad_table, int E, const int64_t* ids, int64_t ids_ld,
                                          int B, int n_t, const float* src, int64_t src_ld) {
  S2VT_REQUIRE(grad_table && ids && src, "s2vt_embed_scatter_add_f32: null pointer");
  if (B * n_t == 0) return 0;
  embed_scatter_add_kernel<<<B * n_t, 128, 0, (cudaStream_t)stream>>>(grad_table, E, ids, ids_ld, B, n_t, src, src_ld);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_bcast_rows_f32(void* stream, const float* src, int64_t src_ld, int B, int N, int n_t, float* dst) {
  S2VT_REQUIRE(src && dst, "s2vt_bcast_rows_f32: null pointer");
  if (B * n_t == 0 || N == 0) return 0;
  bcast_rows_kernel<<<B * n_t, 128, 0, (cudaStream_t)stream>>>(src, src_ld, B, N, dst);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_add_f32(void* stream, const float* a, const float* b, float* out, int64_t n) {
  S2VT_REQUIRE(a && b && out, "s2vt_add_f32: null pointer");
  if (n == 0) return 0;
  add_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_colsum_f32(void* stream, const float* X, int64_t M, int N, int64_t ld, float* out, int accumulate) {
  S2VT_REQUIRE(X && out, "s2vt_colsum_f32: null pointer");
  if (N == 0) return 0;
  colsum_kernel<<<ceil_div(N, 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(X, M, N, ld, out, accumulate);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_ce_f32(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                           float* row_loss, float* loss, float* dlogits, const float* gscale) {
  S2VT_REQUIRE(logits && targets && row_loss && loss, "s2vt_ce_f32: null pointer");
  S2VT_REQUIRE(R > 0 && V > 0, "s2vt_ce_f32: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  ce_row_kernel<<<(unsigned)R, 256, 0, st>>>(logits, V, targets, to_rowmap(tmap), row_loss, dlogits, gscale, 1.0f / (float)R);
  S2VT_CHECK_LAUNCH();
  mean_kernel<<<1, 256, 0, st>>>(row_loss, R, loss);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_adam_f32(void* stream, float* p, const float* g, float* m, float* v, int64_t n,
                             float lr, float beta1, float beta2, float eps, int step_count, float grad_scale,
                             void* bf16_copy) {
  S2VT_REQUIRE(p && g && m && v, "s2vt_adam_f32: null pointer");
  S2VT_REQUIRE(step_count >= 1, "s2vt_adam_f32: step_count is 1-based");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, step_count), bc2 = 1.0 - pow((double)beta2, step_count);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int blocks = (int)ceil_div(n, (int64_t)ADAM_ELEMS_PER_CTA);
  if (bulk_cta_cap() > 0 && blocks > bulk_cta_cap()) blocks = bulk_cta_cap();
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, beta1, beta2, eps, step_size, inv_bc2_sqrt, grad_scale,
                                                        (__nv_bfloat16*)bf16_copy, nullptr);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_adam_prepare(void* stream, int* step_dev, const float* lr_dev, float beta1, float beta2, float* hyper_dev) {
  S2VT_REQUIRE(step_dev && lr_dev && hyper_dev, "s2vt_adam_prepare: null pointer");
  adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, lr_dev, beta1, beta2, hyper_dev);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_adam_f32_dev(void* stream, float* p, const float* g, float* m, float* v, int64_t n,
                                 float beta1, float beta2, float eps, const float* hyper_dev, float grad_scale, void* bf16_copy) {
  S2VT_REQUIRE(p && g && m && v && hyper_dev, "s2vt_adam_f32_dev: null pointer");
  if (n == 0) return 0;
  int blocks = (int)ceil_div(n, (int64_t)ADAM_ELEMS_PER_CTA);
  if (bulk_cta_cap() > 0 && blocks > bulk_cta_cap()) blocks = bulk_cta_cap();
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, beta1, beta2, eps, 0.f, 0.f, grad_scale, (__nv_bfloat16*)bf16_copy,
                                                        hyper_dev);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_cast_bf16(void* stream, const float* src, void* dst, void* dst_t, int64_t rows, int64_t cols) {
  S2VT_REQUIRE(src && (dst || dst_t), "s2vt_cast_bf16: null pointer");
  if (rows * cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_t) {
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
    cast_transpose_bf16_kernel<<<grid, dim3(32, 8), 0, st>>>(src, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_t, rows, cols);
  } else {
    S2VT_REQUIRE(aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "s2vt_cast_bf16: misaligned buffers");
    const long long n = rows * cols;
    cast_bf16_kernel<<<ceil_div(ceil_div(n, 4), 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, n);
  }
  S2VT_CHECK_LAUNCH();
  return 0;
}
