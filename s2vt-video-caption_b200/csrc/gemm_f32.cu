// Exact fp32 GEMM on CUDA cores (FFMA, fp32 operands, fp32 accumulation, fixed summation order).
// This is the arithmetic behind the fp32 "exact" mode whose greedy/beam token ids must match the
// reference's CPU fp32 module bit for bit; tensor-core work lives in gemm_bf16_sm100.cu.
//
// C[cmap(m), n] = sum_k A(m,k) B(n,k) (+ bias[n]) (+ C)
// Tile: BM x BN x 16 per CTA, 256 threads, register micro-tile TM x TN, smem operands stored
// k-major ([k][m]) so that the inner product reads are 128-bit and bank-conflict free; global
// loads are 128-bit along whichever index is contiguous and double-buffered through registers.
#include "common.cuh"

namespace s2vt {

struct GemmF32Params {
  int M, N, K;
  const float* A; RowMap am; int a_trans;
  const float* B; RowMap bm; int b_trans;
  float* C; RowMap cm;
  const float* bias;
  int accumulate;
  int k_chunk;             // K range per blockIdx.z
  long long split_stride;  // C offset per blockIdx.z
  int a_vec, b_vec, c_vec; // 128-bit access allowed
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(const GemmF32Params p) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile is built from float4 groups");
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_chunk;
  const int kend = min(p.K, kbeg + p.k_chunk);
  float* __restrict__ C = p.C + (long long)blockIdx.z * p.split_stride;

  // ---- global -> register staging: each thread owns LA (LB) float4 slots of the A (B) tile
  constexpr int LA = (BM * BK / 4) / NT, LB = (BN * BK / 4) / NT;
  static_assert((BM * BK / 4) % NT == 0 && (BN * BK / 4) % NT == 0, "tile/threads mismatch");
  float4 ra[LA], rb[LB];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int slot = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!p.a_trans) {                       // K contiguous: slot -> (m, 4 consecutive k)
        const int m = m0 + slot / (BK / 4), k = k0 + (slot % (BK / 4)) * 4;
        if (m < p.M && k < kend) {
          const float* src = p.A + p.am(m) + k;
          if (p.a_vec && k + 3 < kend) v = *reinterpret_cast<const float4*>(src);
          else { v.x = src[0]; if (k + 1 < kend) v.y = src[1]; if (k + 2 < kend) v.z = src[2]; if (k + 3 < kend) v.w = src[3]; }
        }
      } else {                                // M contiguous: slot -> (k, 4 consecutive m)
        const int k = k0 + slot / (BM / 4), m = m0 + (slot % (BM / 4)) * 4;
        if (k < kend && m < p.M) {
          const float* src = p.A + p.am(k) + m;
          if (p.a_vec && m + 3 < p.M) v = *reinterpret_cast<const float4*>(src);
          else { v.x = src[0]; if (m + 1 < p.M) v.y = src[1]; if (m + 2 < p.M) v.z = src[2]; if (m + 3 < p.M) v.w = src[3]; }
        }
      }
      ra[i] = v;
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int slot = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!p.b_trans) {
        const int n = n0 + slot / (BK / 4), k = k0 + (slot % (BK / 4)) * 4;
        if (n < p.N && k < kend) {
          const float* src = p.B + p.bm(n) + k;
          if (p.b_vec && k + 3 < kend) v = *reinterpret_cast<const float4*>(src);
          else { v.x = src[0]; if (k + 1 < kend) v.y = src[1]; if (k + 2 < kend) v.z = src[2]; if (k + 3 < kend) v.w = src[3]; }
        }
      } else {
        const int k = k0 + slot / (BN / 4), n = n0 + (slot % (BN / 4)) * 4;
        if (k < kend && n < p.N) {
          const float* src = p.B + p.bm(k) + n;
          if (p.b_vec && n + 3 < p.N) v = *reinterpret_cast<const float4*>(src);
          else { v.x = src[0]; if (n + 1 < p.N) v.y = src[1]; if (n + 2 < p.N) v.z = src[2]; if (n + 3 < p.N) v.w = src[3]; }
        }
      }
      rb[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int slot = tid + i * NT;
      if (!p.a_trans) {
        const int m = slot / (BK / 4), k = (slot % (BK / 4)) * 4;
        As[k + 0][m] = ra[i].x; As[k + 1][m] = ra[i].y; As[k + 2][m] = ra[i].z; As[k + 3][m] = ra[i].w;
      } else {
        const int k = slot / (BM / 4), m = (slot % (BM / 4)) * 4;
        *reinterpret_cast<float4*>(&As[k][m]) = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int slot = tid + i * NT;
      if (!p.b_trans) {
        const int n = slot / (BK / 4), k = (slot % (BK / 4)) * 4;
        Bs[k + 0][n] = rb[i].x; Bs[k + 1][n] = rb[i].y; Bs[k + 2][n] = rb[i].z; Bs[k + 3][n] = rb[i].w;
      } else {
        const int k = slot / (BN / 4), n = (slot % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[k][n]) = rb[i];
      }
    }
  };

  // ---- micro-tile ownership: TM rows as TM/4 groups of 4 spaced BM/(TM/4) apart (same for columns)
  constexpr int GM = TM / 4, GN = TN / 4;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (kbeg < kend) {
    load_a(kbeg); load_b(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
      __syncthreads();
      store_tiles();
      __syncthreads();
      if (k0 + BK < kend) { load_a(k0 + BK); load_b(k0 + BK); }
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int g = 0; g < GM; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(&As[kk][g * (BM / GM) + ty * 4]);
          a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
        }
#pragma unroll
        for (int g = 0; g < GN; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][g * (BN / GN) + tx * 4]);
          b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue
  const bool plain = (gridDim.z == 1);
#pragma unroll
  for (int gi = 0; gi < GM; ++gi)
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int m = m0 + gi * (BM / GM) + ty * 4 + ii;
      if (m >= p.M) continue;
      float* crow = C + p.cm(m);
#pragma unroll
      for (int gj = 0; gj < GN; ++gj) {
        const int n = n0 + gj * (BN / GN) + tx * 4;
        if (n >= p.N) continue;
        float v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) v[jj] = acc[gi * 4 + ii][gj * 4 + jj];
        if (p.c_vec && n + 3 < p.N) {
          if (plain && p.bias) { const float4 bb = *reinterpret_cast<const float4*>(p.bias + n); v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w; }
          if (plain && p.accumulate) { const float4 cc = *reinterpret_cast<const float4*>(crow + n); v[0] += cc.x; v[1] += cc.y; v[2] += cc.z; v[3] += cc.w; }
          *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            if (n + jj < p.N) {
              float x = v[jj];
              if (plain && p.bias) x += p.bias[n + jj];
              if (plain && p.accumulate) x += crow[n + jj];
              crow[n + jj] = x;
            }
          }
        }
      }
    }
}

int gemm_f32_launch(cudaStream_t st, const GemmF32Params& p0, int split_k) {
  GemmF32Params p = p0;
  if (p.M <= 0 || p.N <= 0) return 0;
  if (split_k < 1) split_k = 1;
  int kc = (p.K + split_k - 1) / split_k;
  kc = ((kc + 15) / 16) * 16;
  p.k_chunk = kc > 0 ? kc : 16;
  // 128-bit access eligibility
  p.a_vec = aligned16(p.A) && (p.am.so % 4 == 0) && (p.am.si % 4 == 0);
  p.b_vec = aligned16(p.B) && (p.bm.so % 4 == 0) && (p.bm.si % 4 == 0);
  p.c_vec = aligned16(p.C) && (p.cm.so % 4 == 0) && (p.cm.si % 4 == 0) && (p.split_stride % 4 == 0) &&
            (!p.bias || aligned16(p.bias));
  // big tile when the grid still fills the machine, small tile otherwise
  const long long tiles_big = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 128) * split_k;
  if (tiles_big >= 120 && p.M > 64) {
    dim3 grid(ceil_div(p.N, 128), ceil_div(p.M, 128), split_k);
    gemm_f32_kernel<128, 128, 8, 8><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), split_k);
    gemm_f32_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(p);
  }
  S2VT_CHECK_LAUNCH();
  return 0;
}

// dense row-major helper used by the recurrent / decode drivers
int gemm_f32_simple(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                    int b_trans, float* C, long long ldc, const float* bias, int accumulate, int split_k,
                    long long split_stride) {
  GemmF32Params p{};
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.am = RowMap{1, lda, 0}; p.a_trans = 0;
  p.B = B; p.bm = RowMap{1, ldb, 0}; p.b_trans = b_trans;
  p.C = C; p.cm = RowMap{1, ldc, 0};
  p.bias = bias; p.accumulate = accumulate; p.split_stride = split_k > 1 ? split_stride : 0;
  return gemm_f32_launch(st, p, split_k);
}

}  // namespace s2vt

extern "C" int s2vt_gemm_f32(void* stream, int M, int N, int K,
                             const float* A, s2vt_rowmap amap, int a_trans,
                             const float* B, s2vt_rowmap bmap, int b_trans,
                             float* C, s2vt_rowmap cmap,
                             const float* bias, int accumulate, int split_k, int64_t split_stride) {
  using namespace s2vt;
  S2VT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "s2vt_gemm_f32: negative dimension");
  S2VT_REQUIRE(A && B && C, "s2vt_gemm_f32: null operand");
  S2VT_REQUIRE(split_k <= 1 || (!bias && !accumulate), "s2vt_gemm_f32: split_k excludes bias/accumulate");
  S2VT_REQUIRE(amap.inner >= 1 && bmap.inner >= 1 && cmap.inner >= 1, "s2vt_gemm_f32: rowmap.inner must be >= 1");
  GemmF32Params p{};
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.am = to_rowmap(amap); p.a_trans = a_trans;
  p.B = B; p.bm = to_rowmap(bmap); p.b_trans = b_trans;
  p.C = C; p.cm = to_rowmap(cmap);
  p.bias = bias; p.accumulate = accumulate; p.split_stride = split_k > 1 ? split_stride : 0;
  return gemm_f32_launch((cudaStream_t)stream, p, split_k);
}
