// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) feeds a
// 3-stage shared-memory ring, one elected thread issues tcgen05.mma (M=128, N=128, K=16) into a TMEM
// accumulator, four epilogue warps drain it with tcgen05.ld and apply bias / accumulate / dtype.
//
//   C[cmap(m), n] = sum_k A(m,k) B(n,k) (+ bias[n]) (+ C)
//
// One 128x128 output tile per CTA; 96 KB of smem and 128 TMEM columns per CTA so that two CTAs share an SM
// and one CTA's epilogue overlaps the other's main loop.  Operands may be K-major ([rows, K], K contiguous)
// or MN-major ([K, rows], used by the weight-gradient products where the reduction runs over time x batch).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "sm100_err.cuh"
#include <string.h>

namespace s2vt {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024;
constexpr int TMEM_COLS = 128;

struct GemmBf16Params {
  int M, N, K, num_kb;
  float* Cf;
  __nv_bfloat16* Cb;
  RowMap cm;
  const float* bias;
  int accumulate;
  int c_vec;
  int tma_store;   // dense C: stage the tile in swizzled smem and write it with cp.async.bulk.tensor (1) or add it into C with
                   // cp.reduce.async.bulk.tensor (2: accumulate and split-K)
  int kb_per_split;
};

namespace ptx {
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
}  // namespace ptx

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 2)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmBf16Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb0 = blockIdx.z * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);   // this CTA's slice of K (split-K)

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp_idx == 1 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&tmem_full_bar), 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 2) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1)) { atomicExch(&g_sm100_error, 1); break; }
        const uint32_t fb = ptx::smem_u32(&full_bar[s]);
        const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + A_BYTES;
        ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        if (!A_MN) {
          ptx::tma_load_2d(sA, &tmA, fb, kb * BK, m0);                 // box {64 k, 128 rows}
        } else {
          ptx::tma_load_2d(sA, &tmA, fb, m0, kb * BK);                 // box {64 m, 64 k}
          ptx::tma_load_2d(sA + A_BYTES / 2, &tmA, fb, m0 + 64, kb * BK);
        }
        if (!B_MN) {
          ptx::tma_load_2d(sB, &tmB, fb, kb * BK, n0);
        } else {
          ptx::tma_load_2d(sB, &tmB, fb, n0, kb * BK);
          ptx::tma_load_2d(sB + B_BYTES / 2, &tmB, fb, n0 + 64, kb * BK);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph)) { atomicExch(&g_sm100_error, 2); break; }
        ptx::tc_fence_after();
        const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major: 16 k-elements = 32 B inside the 128 B swizzled row.  MN-major: 16 k-rows = 2 groups of 1024 B.
          const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sA + k * 2048, A_BYTES / 2, 1024)
                                   : ptx::make_smem_desc_sw128(sA + k * 32, 16, 1024);
          const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sB + k * 2048, B_BYTES / 2, 1024)
                                   : ptx::make_smem_desc_sw128(sB + k * 32, 16, 1024);
          ptx::mma_bf16_ss(tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        ptx::mma_commit(ptx::smem_u32(&empty_bar[s]));        // frees the smem stage when these MMAs retire
      }
      ptx::mma_commit(ptx::smem_u32(&tmem_full_bar));         // accumulator complete
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int e = warp_idx - 4;                               // == warp_idx % 4: the TMEM lane quarter this warp may read
    const int m = m0 + e * 32 + lane;
    bool ok = ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar), 0);
    if (!ok) atomicExch(&g_sm100_error, 3);
    ptx::tc_fence_after();
    const long long crow = (m < p.M) ? p.cm(m) : 0;
    if (p.tma_store) {
      // The pipeline stages are idle once the accumulator is complete: reuse them as the staging tile.  Row r of a box is
      // 128 bytes; its 16-byte chunk j sits at r*128 + ((j ^ (r & 7)) * 16) (the 128B swizzle the C tensor map undoes),
      // so the 32 lanes of a warp -- 32 different rows -- write conflict-free.
      const int rrow = e * 32 + lane;
      const uint32_t row_base = smem_base + (uint32_t)rrow * 128u;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem + ((uint32_t)(e * 32) << 16) + (uint32_t)(c * 32), r);
        ptx::tc_wait_ld();
        const int n = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias && blockIdx.z == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (n + j < p.N) ? __ldg(p.bias + n + j) : 0.f;
        }
        if (p.Cf) {                                            // box c: 128 rows x 32 fp32
          const uint32_t box = row_base + (uint32_t)c * 16384u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(box + (uint32_t)((j ^ (rrow & 7)) * 16)),
                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
        } else {                                               // box c/2: 128 rows x 64 bf16; this pass fills half a row
          const uint32_t box = row_base + (uint32_t)(c >> 1) * 16384u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), t1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), t3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
            const int chunk = (c & 1) * 4 + j;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + (uint32_t)((chunk ^ (rrow & 7)) * 16)),
                         "r"(*reinterpret_cast<uint32_t*>(&t0)), "r"(*reinterpret_cast<uint32_t*>(&t1)),
                         "r"(*reinterpret_cast<uint32_t*>(&t2)), "r"(*reinterpret_cast<uint32_t*>(&t3)) : "memory");
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::epi_bar_sync();
      if (warp_idx == 4 && ptx::elect_one()) {
        if (p.Cf) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (n0 + c * 32 < p.N) {
              if (p.tma_store == 2) ptx::tma_reduce_add_2d(&tmC, smem_base + c * 16384, n0 + c * 32, m0);
              else ptx::tma_store_2d(&tmC, smem_base + c * 16384, n0 + c * 32, m0);
            }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (n0 + c * 64 < p.N) ptx::tma_store_2d(&tmC, smem_base + c * 16384, n0 + c * 64, m0);
        }
        ptx::bulk_commit();
        ptx::bulk_wait_read0();                                // smem must stay valid until the stores have read it
      }
    } else {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem + ((uint32_t)(e * 32) << 16) + (uint32_t)(c * 32), r);
      ptx::tc_wait_ld();
      const int n = n0 + c * 32;
      if (!ok || m >= p.M || n >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const bool full = (n + 31 < p.N);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (full || n + j < p.N) v[j] += __ldg(p.bias + n + j);
      }
      if (p.Cf) {
        float* dst = p.Cf + crow + n;
        if (full && p.c_vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.accumulate) { const float4 old = *reinterpret_cast<const float4*>(dst + j); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.N) dst[j] = p.accumulate ? dst[j] + v[j] : v[j];
        }
      } else {
        __nv_bfloat16* dst = p.Cb + crow + n;
        if (full && p.c_vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
            o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
            o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(dst + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.N) dst[j] = __float2bfloat16(v[j]);
        }
      }
    }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) ptx::tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2D bf16 tensor map over a [outer, inner] row-major view with `ld` elements between rows, 128B swizzle
int make_tmap_any(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                  int esize);
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_any(out, base, inner, outer, ld, box_inner, box_outer, 2);
}
int make_tmap_any(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                  int esize) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * esize) % 16) return fail("TMA operand must be 16-byte aligned with a 16-byte row pitch (ld=%llu)", (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu)", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return 0;
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_has_tcgen05(void) { return 1; }

namespace s2vt {
int lstm_bf16_error_flag();
int lstm_bwd_bf16_error_flag();
int gemm_persist_error_flag();
int xdec_error_flag();
int lstm_bf16_error_clear();
int lstm_bwd_bf16_error_clear();
int gemm_persist_error_clear();
int xdec_error_clear();
int lstm_step_error_flag();
int lstm_step_error_clear();
int launch_gemm_persist(cudaStream_t st, int M, int N, int K, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                        void* C, int64_t ldc, int out_bf16, const float* bias, int accumulate);
void gemm_persist_set_max_ctas(int n);
static thread_local int g_use_persistent = 1;
static thread_local int g_high_priority = 0;      // 1: launch with the device's greatest priority as a launch attribute
}


extern "C" int s2vt_set_launch_priority(int high) { g_high_priority = high ? 1 : 0; return 0; }

extern "C" int s2vt_gemm_bf16_set_mode(int max_ctas, int use_persistent) {
  gemm_persist_set_max_ctas(max_ctas);
  g_use_persistent = use_persistent;
  return 0;
}

extern "C" int s2vt_device_error_flag(void* stream) {
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -2;
  const int a = read_sm100_error_flag(), b = lstm_bf16_error_flag(), c = lstm_bwd_bf16_error_flag(), d = gemm_persist_error_flag();
  const int x = xdec_error_flag(), y = lstm_step_error_flag();
  return a != 0 ? a : (b != 0 ? b : (c != 0 ? c : (d != 0 ? d : (x != 0 ? x : y))));
}

extern "C" int s2vt_device_error_clear(void) {
  return clear_sm100_error_flag() | lstm_bf16_error_clear() | lstm_bwd_bf16_error_clear() | gemm_persist_error_clear() | xdec_error_clear() | lstm_step_error_clear();
}

extern "C" int s2vt_gemm_bf16(void* stream, int M, int N, int K,
                              const void* A, int64_t lda, int a_mn_major,
                              const void* B, int64_t ldb, int b_mn_major,
                              void* C, s2vt_rowmap cmap, int out_bf16,
                              const float* bias, int accumulate) {
  S2VT_REQUIRE(M > 0 && N > 0 && K > 0, "s2vt_gemm_bf16: dimensions must be positive (M=%d N=%d K=%d)", M, N, K);
  S2VT_REQUIRE(A && B && C, "s2vt_gemm_bf16: null operand");
  S2VT_REQUIRE(!(out_bf16 && accumulate), "s2vt_gemm_bf16: accumulate needs an f32 output");
  S2VT_REQUIRE(cmap.inner >= 1, "s2vt_gemm_bf16: rowmap.inner must be >= 1");
  CUtensorMap tmA, tmB;
  int rc;
  rc = a_mn_major ? make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64)
                  : make_tmap_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 128);
  if (rc) return rc;
  rc = b_mn_major ? make_tmap_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64)
                  : make_tmap_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, 128);
  if (rc) return rc;
  GemmBf16Params p{};
  p.M = M; p.N = N; p.K = K; p.num_kb = (K + BK - 1) / BK;
  p.Cf = out_bf16 ? nullptr : (float*)C;
  p.Cb = out_bf16 ? (__nv_bfloat16*)C : nullptr;
  p.cm = to_rowmap(cmap);
  p.bias = bias; p.accumulate = accumulate;
  const int vq = out_bf16 ? 8 : 4;
  p.c_vec = aligned16(C) && (p.cm.so % vq == 0) && (p.cm.si % vq == 0);
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  const bool dense_c = p.cm.inner == 1 && p.cm.si == 0 && p.c_vec && p.cm.so >= N;
  if (dense_c && g_use_persistent) {
    rc = launch_gemm_persist((cudaStream_t)stream, M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, C, (int64_t)p.cm.so, out_bf16, bias, accumulate);
    if (rc >= 0) return rc;
  }
  p.tma_store = dense_c ? ((accumulate && !out_bf16) ? 2 : (accumulate ? 0 : 1)) : 0;
  // split-K when the output has too few tiles to fill the machine (the weight-gradient products: K = time x batch)
  int splits = 1;
  const int tiles = ceil_div(N, BN) * ceil_div(M, BM);
  if (dense_c && !out_bf16 && tiles < 148 && p.num_kb >= 16) {
    splits = 296 / tiles;
    if (splits > 8) splits = 8;
    if (splits > p.num_kb / 8) splits = p.num_kb / 8;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = ceil_div(p.num_kb, splits);
  splits = ceil_div(p.num_kb, p.kb_per_split);                 // no empty slices
  if (splits > 1) {
    if (!accumulate) S2VT_CHECK_CUDA(cudaMemset2DAsync(C, (size_t)p.cm.so * 4, 0, (size_t)N * 4, (size_t)M, (cudaStream_t)stream));
    p.tma_store = 2;
  }
  if (p.tma_store) {
    rc = make_tmap_any(&tmC, C, (uint64_t)N, (uint64_t)M, (uint64_t)p.cm.so, out_bf16 ? 64 : 32, 128, out_bf16 ? 2 : 4);
    if (rc) return rc;
  }
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
  cudaStream_t st = (cudaStream_t)stream;
#define S2VT_LAUNCH_GEMM(AM, BMJ)                                                                                   \
  do {                                                                                                              \
    S2VT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<AM, BMJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM)); \
    cudaLaunchConfig_t cfg{};                                                                                       \
    cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = GEMM_SMEM; cfg.stream = st;                \
    cudaLaunchAttribute attr[1];                                                                                    \
    if (g_high_priority) {                                                                                          \
      int least = 0, greatest = 0;                                                                                  \
      S2VT_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));                                         \
      attr[0].id = cudaLaunchAttributePriority; attr[0].val.priority = greatest;                                    \
      cfg.attrs = attr; cfg.numAttrs = 1;                                                                           \
    }                                                                                                               \
    S2VT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<AM, BMJ>, tmA, tmB, tmC, p));                         \
  } while (0)
  if (!a_mn_major && !b_mn_major) S2VT_LAUNCH_GEMM(false, false);
  else if (a_mn_major && !b_mn_major) S2VT_LAUNCH_GEMM(true, false);
  else if (!a_mn_major && b_mn_major) S2VT_LAUNCH_GEMM(false, true);
  else S2VT_LAUNCH_GEMM(true, true);
#undef S2VT_LAUNCH_GEMM
  S2VT_CHECK_LAUNCH();
  return 0;
}
