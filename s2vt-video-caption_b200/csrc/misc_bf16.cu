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

// out[n] += sum_{m in row chunk} X[m*ld + n], X bf16, fp32 accumulation.  grid = (ceil(N/64), row chunks); block = 32 x 8; each
// thread owns 2 adjacent columns (one 4-byte load per row) and the row chunks combine with one atomicAdd per column.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long M, int N, long long ld, float* __restrict__ out,
                                   float* __restrict__ out2) {
  __shared__ float sh[8][65];
  const int n = blockIdx.x * 64 + threadIdx.x * 2;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long m0 = (long long)blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float a0 = 0.f, a1 = 0.f;
  if (n + 1 < N && (ld & 1) == 0) {
    const __nv_bfloat16* src = X + n;
    long long m = m0 + threadIdx.y;
    for (; m + 24 < m1; m += 32) {                      // 4 independent loads in flight per thread
      const __nv_bfloat162 v0 = *reinterpret_cast<const __nv_bfloat162*>(src + m * ld);
      const __nv_bfloat162 v1 = *reinterpret_cast<const __nv_bfloat162*>(src + (m + 8) * ld);
      const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(src + (m + 16) * ld);
      const __nv_bfloat162 v3 = *reinterpret_cast<const __nv_bfloat162*>(src + (m + 24) * ld);
      a0 += (__low2float(v0) + __low2float(v1)) + (__low2float(v2) + __low2float(v3));
      a1 += (__high2float(v0) + __high2float(v1)) + (__high2float(v2) + __high2float(v3));
    }
    for (; m < m1; m += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + m * ld);
      a0 += __low2float(v); a1 += __high2float(v);
    }
  } else if (n < N) {
    for (long long m = m0 + threadIdx.y; m < m1; m += 8) {
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
        atomicAdd(out + n + k, s);
        if (out2) atomicAdd(out2 + n + k, s);
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

// one CTA per row.  Pass 1 (skipped when the row's log-sum-exp is supplied): online max / sum-exp in ONE read of the row;
// row_loss = lse - z[target].  Pass 2 (optional): bf16 dlogits = (softmax - onehot) * gscale / R.
__global__ void ce_row_bf16_kernel(const float* __restrict__ logits, int V, const int64_t* __restrict__ targets, RowMap tmap,
                                   float* __restrict__ row_loss, float* __restrict__ row_lse, int have_lse,
                                   __nv_bfloat16* __restrict__ dlogits, const float* __restrict__ gscale, float inv_rows, RowMap omap) {
  __shared__ float sh[32];
  const long long r = blockIdx.x;
  const float* z = logits + r * V;
  float lse;
  if (have_lse) {
    lse = row_lse[r];
  } else {
    float mx = -INFINITY, s = 0.f;
    if ((V & 3) == 0) {
      const float4* z4 = reinterpret_cast<const float4*>(z);
      for (int j = threadIdx.x; j < V / 4; j += blockDim.x) {
        const float4 v = z4[j];
        const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (m4 > mx) { s *= __expf(mx - m4); mx = m4; }
        s += (__expf(v.x - mx) + __expf(v.y - mx)) + (__expf(v.z - mx) + __expf(v.w - mx));
      }
    } else {
      for (int j = threadIdx.x; j < V; j += blockDim.x) {
        const float v = z[j];
        if (v > mx) { s *= __expf(mx - v); mx = v; }
        s += __expf(v - mx);
      }
    }
    const float gmx = blk_reduce_max(mx, sh);
    s = (mx == -INFINITY) ? 0.f : s * __expf(mx - gmx);
    s = blk_reduce_sum(s, sh);
    lse = gmx + logf(s);
    if (threadIdx.x == 0 && row_lse) row_lse[r] = lse;
  }
  const long long tgt = targets[tmap(r)];
  if (threadIdx.x == 0 && row_loss) row_loss[r] = lse - z[tgt];
  if (dlogits) {
    const float sc = (gscale ? gscale[0] : 1.f) * inv_rows;
    __nv_bfloat16* d = dlogits + omap(r);
    if ((V & 3) == 0) {
      const float4* z4 = reinterpret_cast<const float4*>(z);
      for (int j = threadIdx.x; j < V / 4; j += blockDim.x) {
        const float4 v = z4[j];
        float p0 = __expf(v.x - lse), p1 = __expf(v.y - lse), p2 = __expf(v.z - lse), p3 = __expf(v.w - lse);
        const long long j0 = 4ll * j;
        if (tgt >= j0 && tgt < j0 + 4) { if (tgt == j0) p0 -= 1.f; else if (tgt == j0 + 1) p1 -= 1.f; else if (tgt == j0 + 2) p2 -= 1.f; else p3 -= 1.f; }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(p0 * sc, p1 * sc), hi = __floats2bfloat162_rn(p2 * sc, p3 * sc);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(d + j0) = pk;
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

namespace s2vt {
// Wide form: a thread owns 8 adjacent columns (one 16-byte load per row, four rows in flight), a warp 256 columns, the 8 warps of a
// block split the rows.  grid = (ceil(N/256), row chunks); partial sums meet in shared memory, one atomicAdd per column and block.
__global__ void __launch_bounds__(256) colsum_bf16_v8_kernel(const __nv_bfloat16* __restrict__ X, long long M, int N, long long ld,
                                                             float* __restrict__ out, float* __restrict__ out2) {
  __shared__ float sh[8][257];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 256 + lane * 8;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long m0 = (long long)blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  auto acc = [&](const uint4& v) {
    a[0] += __uint_as_float(v.x << 16); a[1] += __uint_as_float(v.x & 0xffff0000u);
    a[2] += __uint_as_float(v.y << 16); a[3] += __uint_as_float(v.y & 0xffff0000u);
    a[4] += __uint_as_float(v.z << 16); a[5] += __uint_as_float(v.z & 0xffff0000u);
    a[6] += __uint_as_float(v.w << 16); a[7] += __uint_as_float(v.w & 0xffff0000u);
  };
  if (n < N) {
    const __nv_bfloat16* src = X + n;
    long long m = m0 + warp;
    for (; m + 24 < m1; m += 32) {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(src + m * ld));
      const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(src + (m + 8) * ld));
      const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(src + (m + 16) * ld));
      const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(src + (m + 24) * ld));
      acc(v0); acc(v1); acc(v2); acc(v3);
    }
    for (; m < m1; m += 8) acc(__ldg(reinterpret_cast<const uint4*>(src + m * ld)));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[warp][lane * 8 + j] = a[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    atomicAdd(out + col, t);
    if (out2) atomicAdd(out2 + col, t);
  }
}
}  // namespace s2vt

extern "C" int s2vt_colsum_bf16(void* stream, const void* X_bf16, int64_t M, int N, int64_t ld, float* out, float* out2) {
  S2VT_REQUIRE(X_bf16 && out, "s2vt_colsum_bf16: null pointer");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  S2VT_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (out2) S2VT_CHECK_CUDA(cudaMemsetAsync(out2, 0, sizeof(float) * N, st));
  if (N % 8 == 0 && ld % 8 == 0 && aligned16(X_bf16)) {
    const int gx8 = ceil_div(N, 256);
    int gy8 = ceil_div(148 * 4, gx8);                   // at least ~4 blocks of 256 threads per SM in total ...
    const int gy_short = ceil_div(M, 128);              // ... and at most 128 rows (64 KB) per block: blocks that live a microsecond or two
    if (gy8 < gy_short) gy8 = gy_short;                 // give their SM slots back to the wave front's coupling products at once
    if (bulk_cta_cap() > 0) gy8 = bulk_cta_cap() / gx8;  // beside a recurrence sweep: all blocks resident at once, slots left over
    const int max_gy8 = ceil_div(M, 32);
    if (gy8 > max_gy8) gy8 = max_gy8;
    if (gy8 < 1) gy8 = 1;
    colsum_bf16_v8_kernel<<<dim3(gx8, gy8), 256, 0, st>>>((const __nv_bfloat16*)X_bf16, M, N, ld, out, out2);
    S2VT_CHECK_LAUNCH();
    return 0;
  }
  const int gx = ceil_div(N, 64);
  int gy = ceil_div(148 * 8, gx);                       // ~8 CTAs per SM in total
  const int max_gy = ceil_div(M, 64);
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  colsum_bf16_kernel<<<dim3(gx, gy), dim3(32, 8), 0, st>>>((const __nv_bfloat16*)X_bf16, M, N, ld, out, out2);
  S2VT_CHECK_LAUNCH();
  return 0;
}

extern "C" int s2vt_ce_bf16(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                            float* row_loss, float* loss, float* row_lse, int have_lse, void* dlogits_bf16, const float* gscale) {
  S2VT_REQUIRE(logits && targets, "s2vt_ce_bf16: null pointer");
  S2VT_REQUIRE(R > 0 && V > 0, "s2vt_ce_bf16: empty input");
  S2VT_REQUIRE(!loss || row_loss, "s2vt_ce_bf16: loss needs row_loss scratch");
  S2VT_REQUIRE(!have_lse || row_lse, "s2vt_ce_bf16: have_lse needs row_lse");
  cudaStream_t st = (cudaStream_t)stream;
  ce_row_bf16_kernel<<<(unsigned)R, 256, 0, st>>>(logits, V, targets, to_rowmap(tmap), row_loss, row_lse, have_lse,
                                                 (__nv_bfloat16*)dlogits_bf16, gscale, 1.0f / (float)R, RowMap{1, (long long)V, 0});
  S2VT_CHECK_LAUNCH();
  if (loss) {
    mean_f32_kernel<<<1, 256, 0, st>>>(row_loss, R, loss);
    S2VT_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int s2vt_ce_bf16_mapped(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                                   float* row_loss, float* loss, float* row_lse, int have_lse, void* dlogits_bf16, s2vt_rowmap omap,
                                   const float* gscale) {
  S2VT_REQUIRE(logits && targets, "s2vt_ce_bf16_mapped: null pointer");
  S2VT_REQUIRE(R > 0 && V > 0, "s2vt_ce_bf16_mapped: empty input");
  S2VT_REQUIRE(!loss || row_loss, "s2vt_ce_bf16_mapped: loss needs row_loss scratch");
  S2VT_REQUIRE(!have_lse || row_lse, "s2vt_ce_bf16_mapped: have_lse needs row_lse");
  S2VT_REQUIRE(omap.inner >= 1 && tmap.inner >= 1, "s2vt_ce_bf16_mapped: rowmap.inner must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  ce_row_bf16_kernel<<<(unsigned)R, 256, 0, st>>>(logits, V, targets, to_rowmap(tmap), row_loss, row_lse, have_lse,
                                                 (__nv_bfloat16*)dlogits_bf16, gscale, 1.0f / (float)R, to_rowmap(omap));
  S2VT_CHECK_LAUNCH();
  if (loss) {
    mean_f32_kernel<<<1, 256, 0, st>>>(row_loss, R, loss);
    S2VT_CHECK_LAUNCH();
  }
  return 0;
}
