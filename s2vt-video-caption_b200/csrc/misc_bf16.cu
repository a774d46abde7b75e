// Memory-bound helpers of the bf16 training path: bf16 embedding gather, bf16 column sums (bias gradients),
// cross-entropy backward with bf16 dlogits.
#include "common.cuh"
#include <math.h>

namespace s2vt {

// out[(t*B + b), 0:E] = table[ids[b*ids_ld + t], 0:E]   (bf16 rows, 16-byte vectors when E % 8 == 0)
__global__ void embed_gather_bf16_kernel(const __nv_bfloat16* __restrict__ table, int E, const int64_t* __restrict__ ids, long long ids_ld,
                                         int B, __nv_bfloat16* __restrict__ out, long long out_ld) {
  const int row = blockIdx.x;
  const int t = row / B, b = row % B;
  const long long id = ids[(long long)b * ids_ld + t];
  const __nv_bfloat16* src = table + id * E;
  __nv_bfloat16* dst = out + (long long)row * out_ld;
  if ((E & 7) == 0 && (out_ld & 7) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int e = threadIdx.x; e < E / 8; e += blockDim.x) d4[e] = s4[e];
  } else {
    for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = src[e];
  }
}

// out[n] = sum_m X[m*ld + n], X bf16, fp32 accumulation.  grid.x = ceil(N/64); block = 32 x 8; each thread owns 2 columns.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long M, int N, long long ld, float* __restrict__ out) {
  __shared__ float sh[8][65];
  const int n = blockIdx.x * 64 + threadIdx.x * 2;
  float a0 = 0.f, a1 = 0.f;
  if (n + 1 < N && (ld & 1) == 0) {
    for (long long m = threadIdx.y; m < M; m += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(X + m * ld + n);
      a0 += __low2float(v); a1 += __high2float(v);
    }
  } else if (n < N) {
    for (long long m = threadIdx.y; m < M; m += 8) {
      a0 += __bfloat162float(X[m * ld + n]);
      if (n + 1 < N) a1 += __bfloat162float(X[m * ld + n + 1]);
    }
  }
  sh[threadIdx.y][threadIdx.x * 2] = a0;
  sh[threadIdx.y][threadIdx.x * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (n + k < N) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x * 2 + k];
        out[n + k] = s;
      }
    }
  }
}

__device__ __forceinline__ float blk_reduce_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int j = 1; j < (int)(blockDim.x >> 5); ++j) r = fmaxf(r, sh[j]);
  return r;
}
__device__ __forceinline__ float blk_reduce_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int j = 0; j < (int)(blockDim.x >> 5); ++j) r += sh[j];
  return r;
}

// one CTA per row: row_loss = lse - z[target]; optional bf16 dlogits = (softmax - onehot) * gscale / R
__global__ void ce_row_bf16_kernel(const float* __restrict__ logits, int V, const int64_t* __restrict__ targets, RowMap tmap,
                                   float* __restrict__ row_loss, __nv_bfloat16* __restrict__ dlogits, const float* __restrict__ gscale,
                                   float inv_rows) {
  __shared__ float sh[32];
  const long long r = blockIdx.x;
  const float* z = logits + r * V;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, z[j]);
  mx = blk_reduce_max(mx, sh);
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += __expf(z[j] - mx);
  s = blk_reduce_sum(s, sh);
  const float lse = mx + logf(s);
  const long long tgt = targets[tmap(r)];
  if (threadIdx.x == 0 && row_loss) row_loss[r] = lse - z[tgt];
  if (dlogits) {
    const float sc = (gscale ? gscale[0] : 1.f) * inv_rows;
    __nv_bfloat16* d = dlogits + r * V;
    if ((V & 1) == 0) {
      for (int j = threadIdx.x * 2; j < V; j += blockDim.x * 2) {
        float p0 = __expf(z[j] - lse), p1 = __expf(z[j + 1] - lse);
        if (j == tgt) p0 -= 1.f;
        if (j + 1 == tgt) p1 -= 1.f;
        *reinterpret_cast<__nv_bfloat162*>(d + j) = __floats2bfloat162_rn(p0 * sc, p1 * sc);
      }
    } else {
      for (int j = threadIdx.x; j < V; j += blockDim.x) {
        float pz = __expf(z[j] - lse);
        if (j == tgt) pz -= 1.f;
        d[j] = __float2bfloat16(pz * sc);
      }
    }
  }
}

__global__ void mean_f32_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
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

}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_embed_gather_bf16(void* stream, const void* table_bf16, int E, const int64_t* ids, int64_t ids_ld,
                                      int B, int n_t, void* out_bf16, int64_t out_ld) {
  S2VT_REQUIRE(table_bf16 && ids && out_bf16, "s2vt_embed_gather_bf16: null pointer");
  if (B * n_t == 0) return 0;
  embed_gather_bf16_kernel<<<B * n_t, 64, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)table_bf16, E, ids, ids_ld, B,
                                                                   (__nv_bfloat16*)out_bf16, out_ld);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_colsum_bf16(void* stream, const void* X_bf16, int64_t M, int N, int64_t ld, float* out) {
  S2VT_REQUIRE(X_bf16 && out, "s2vt_colsum_bf16: null pointer");
  if (N == 0) return 0;
  colsum_bf16_kernel<<<ceil_div(N, 64), dim3(32, 8), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)X_bf16, M, N, ld, out);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_ce_bf16(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                            float* row_loss, float* loss, void* dlogits_bf16, const float* gscale) {
  S2VT_REQUIRE(logits && targets, "s2vt_ce_bf16: null pointer");
  S2VT_REQUIRE(R > 0 && V > 0, "s2vt_ce_bf16: empty input");
  S2VT_REQUIRE(!loss || row_loss, "s2vt_ce_bf16: loss needs row_loss scratch");
  cudaStream_t st = (cudaStream_t)stream;
  ce_row_bf16_kernel<<<(unsigned)R, 256, 0, st>>>(logits, V, targets, to_rowmap(tmap), row_loss, (__nv_bfloat16*)dlogits_bf16, gscale,
                                                 1.0f / (float)R);
  S2VT_CHECK_LAUNCH();
  if (loss) {
    mean_f32_kernel<<<1, 256, 0, st>>>(row_loss, R, loss);
    S2VT_CHECK_LAUNCH();
  }
  return 0;
}
