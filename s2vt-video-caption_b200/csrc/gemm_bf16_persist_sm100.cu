// Persistent bf16 GEMM for dense outputs: one CTA per SM walks a dynamic list of work units (128 x 256 output tile x K slice).
//
//   warp 0 (one thread)  scheduler + TMA producer: atomically claims the next unit, publishes it through a 4-deep shared-memory ring,
//                        streams its K blocks into a 4-stage ring (A 128x64, B 256x64 bf16, 128B swizzle)
//   warp 1 (one thread)  tcgen05.mma M=128 N=256 K=16 (12 KB of operands per 128 tensor cycles = 96 B/clk of shared-memory reads,
//                        below the 128 B/clk port limit that a 128x128 tile sits on) into one of TWO 256-column TMEM accumulators
//   warps 2..5           epilogue: drain the other accumulator while the next unit's MMAs run -- tcgen05.ld -> bias -> swizzled
//                        staging -> cp.async.bulk.tensor store (or cp.reduce add for accumulate / split-K); each warp owns its
//                        32 rows end to end (own staging, own bulk groups), so the epilogue has no CTA-level barrier at all
//
// Units are claimed with atomicAdd, not assigned statically: when part of the machine is held by the persistent recurrence
// clusters of another stream, the CTAs that do get an SM finish all the work and late CTAs exit at once.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "sm100_err.cuh"
#include <string.h>
#include <atomic>

namespace s2vt {

int make_tmap_any(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer, int esize);

int gemm_persist_error_flag() { return read_sm100_error_flag(); }

constexpr int PBM = 128, PBN = 256, PBK = 64, PSTAGES = 4, PRING = 4;
constexpr int PA_BYTES = PBM * PBK * 2, PB_BYTES = PBN * PBK * 2, PSTAGE_BYTES = PA_BYTES + PB_BYTES;
constexpr int PSTG_WARP = 2 * 4096;                                   // per epilogue warp: two 32-row x 128-byte staging boxes
constexpr int PSMEM = PSTAGES * PSTAGE_BYTES + 4 * PSTG_WARP + 1024;
constexpr int SCHED_SLOTS = 1024;
__device__ unsigned int g_gemm_sched[2 * SCHED_SLOTS];                // per launch slot: {next unit, finished CTAs}; self-resetting
static std::atomic<unsigned> g_launch_seq{0};

struct GemmPersistParams {
  int M, N, K, num_kb, kb_per_split;
  int tiles_n, tiles, total;        // units: u -> split = u / tiles, tile = u % tiles (n fastest)
  int out_bf16, reduce;
  const float* bias;
  unsigned int* sched;
};

namespace ptx {
__device__ __forceinline__ void p_tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void p_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void p_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}  // namespace ptx

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmPersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[PSTAGES], empty_bar[PSTAGES], acc_full[2], acc_empty[2], ring_full[PRING], ring_empty[PRING];
  __shared__ int ring_unit[PRING];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + PSTAGES * PSTAGE_BYTES;
  const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
  }
  if (warp_idx == 1 && ptx::elect_one()) {
    for (int s = 0; s < PSTAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(ptx::smem_u32(&acc_full[a]), 1);
      ptx::mbar_init(ptx::smem_u32(&acc_empty[a]), 4);
    }
    for (int r = 0; r < PRING; ++r) {
      ptx::mbar_init(ptx::smem_u32(&ring_full[r]), 1);
      ptx::mbar_init(ptx::smem_u32(&ring_empty[r]), 5);        // MMA thread + one lane of each epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 2) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp_idx == 0) {
    // ===================== scheduler + TMA producer =====================
    if (ptx::elect_one()) {
      uint32_t it = 0, ri = 0;
      for (;;) {
        const unsigned u = atomicAdd(&p.sched[0], 1u);
        const int unit = u < (unsigned)p.total ? (int)u : -1;
        const int slot = ri % PRING;
        if (!ptx::mbar_wait(ptx::smem_u32(&ring_empty[slot]), ((ri / PRING) & 1) ^ 1)) { atomicExch(&g_sm100_error, 31); break; }
        ring_unit[slot] = unit;
        ptx::mbar_arrive(ptx::smem_u32(&ring_full[slot]));
        ++ri;
        if (unit < 0) break;
        const int split = unit / p.tiles, tile = unit % p.tiles;
        const int n0 = (tile % p.tiles_n) * PBN, m0 = (tile / p.tiles_n) * PBM;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        bool ok = true;
        for (int kb = kb0; kb < kb1 && ok; ++kb, ++it) {
          const int s = it % PSTAGES;
          ok = ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ((it / PSTAGES) & 1) ^ 1);
          if (!ok) { atomicExch(&g_sm100_error, 32); break; }
          const uint32_t fb = ptx::smem_u32(&full_bar[s]);
          const uint32_t sA = smem_base + s * PSTAGE_BYTES, sB = sA + PA_BYTES;
          ptx::mbar_arrive_expect_tx(fb, PSTAGE_BYTES);
          if (!A_MN) {
            ptx::tma_load_2d(sA, &tmA, fb, kb * PBK, m0);                       // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < PBM / 64; ++j) ptx::tma_load_2d(sA + j * 8192, &tmA, fb, m0 + 64 * j, kb * PBK);   // box {64 m, 64 k}
          }
          if (!B_MN) {
            ptx::tma_load_2d(sB, &tmB, fb, kb * PBK, n0);                       // box {64 k, 256 rows}
          } else {
#pragma unroll
            for (int j = 0; j < PBN / 64; ++j) ptx::tma_load_2d(sB + j * 8192, &tmB, fb, n0 + 64 * j, kb * PBK);
          }
        }
        if (!ok) break;
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(PBM, PBN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t it = 0, ri = 0, ai = 0;
      for (;;) {
        const int slot = ri % PRING;
        if (!ptx::mbar_wait(ptx::smem_u32(&ring_full[slot]), (ri / PRING) & 1)) { atomicExch(&g_sm100_error, 33); break; }
        const int unit = ring_unit[slot];
        ptx::mbar_arrive(ptx::smem_u32(&ring_empty[slot]));
        ++ri;
        if (unit < 0) break;
        const int split = unit / p.tiles;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const uint32_t a = ai & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&acc_empty[a]), ((ai >> 1) & 1) ^ 1)) { atomicExch(&g_sm100_error, 34); break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem + a * PBN;
        bool ok = true;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % PSTAGES;
          ok = ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), (it / PSTAGES) & 1);
          if (!ok) { atomicExch(&g_sm100_error, 35); break; }
          ptx::tc_fence_after();
          const uint32_t sA = smem_base + s * PSTAGE_BYTES, sB = sA + PA_BYTES;
#pragma unroll
          for (int k = 0; k < PBK / 16; ++k) {
            const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sA + k * 2048, 8192, 1024) : ptx::make_smem_desc_sw128(sA + k * 32, 16, 1024);
            const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sB + k * 2048, 8192, 1024) : ptx::make_smem_desc_sw128(sB + k * 32, 16, 1024);
            ptx::mma_bf16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::mma_commit(ptx::smem_u32(&empty_bar[s]));
        }
        if (!ok) break;
        ptx::mma_commit(ptx::smem_u32(&acc_full[a]));
        ++ai;
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp_idx & 3;                                   // TMEM lane quarter this warp may read
    const uint32_t stg = stg_base + (uint32_t)(warp_idx - 2) * PSTG_WARP;
    const uint32_t my_row = stg + (uint32_t)lane * 128u;
    uint32_t ri = 0, ai = 0, nstore = 0;
    for (;;) {
      const int slot = ri % PRING;
      if (!ptx::mbar_wait(ptx::smem_u32(&ring_full[slot]), (ri / PRING) & 1)) { atomicExch(&g_sm100_error, 36); break; }
      const int unit = ring_unit[slot];
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ring_empty[slot]));
      ++ri;
      if (unit < 0) break;
      const int split = unit / p.tiles, tile = unit % p.tiles;
      const int n0 = (tile % p.tiles_n) * PBN, m0 = (tile / p.tiles_n) * PBM;
      const uint32_t a = ai & 1;
      if (!ptx::mbar_wait(ptx::smem_u32(&acc_full[a]), (ai >> 1) & 1)) { atomicExch(&g_sm100_error, 37); break; }
      ++ai;
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem + ((uint32_t)(q * 32) << 16) + a * PBN;
      const bool add_bias = p.bias != nullptr && split == 0;
      const int row0 = m0 + q * 32;                                // first row of this warp's slab
      if (!p.out_bf16) {
        const int nchunk = min(PBN / 32, (p.N - n0 + 31) / 32);
        for (int c = 0; c < nchunk; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_acc + (uint32_t)(c * 32), r);
          ptx::tc_wait_ld();
          if (c == nchunk - 1) {                                   // accumulator drained: hand it back to the MMA thread
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[a]));
          }
          const int n = n0 + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (add_bias) {
            if (n + 32 <= p.N) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (n + j < p.N) ? __ldg(p.bias + n + j) : 0.f;
            }
          }
          if (lane == 0) ptx::p_bulk_wait_read1();                 // the store that used this staging box two chunks ago has read it
          __syncwarp();
          const uint32_t box = my_row + (nstore & 1) * 4096u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(box + (uint32_t)((j ^ (lane & 7)) * 16)),
                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            if (p.reduce) ptx::p_tma_reduce_add_2d(&tmC, stg + (nstore & 1) * 4096u, n, row0);
            else ptx::p_tma_store_2d(&tmC, stg + (nstore & 1) * 4096u, n, row0);
            ptx::p_bulk_commit();
          } else if (lane == 0) {
            ptx::p_bulk_commit();                                   // keep the group count in step with nstore
          }
          ++nstore;
        }
      } else {
        const int nchunk = min(PBN / 64, (p.N - n0 + 63) / 64);
        for (int c = 0; c < nchunk; ++c) {
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_32x32(t_acc + (uint32_t)(c * 64), r0);
          ptx::tmem_ld_32x32(t_acc + (uint32_t)(c * 64 + 32), r1);
          ptx::tc_wait_ld();
          if (c == nchunk - 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[a]));
          }
          const int n = n0 + c * 64;
          float v[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          if (add_bias) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] += (n + j < p.N) ? __ldg(p.bias + n + j) : 0.f;
          }
          if (lane == 0) ptx::p_bulk_wait_read1();
          __syncwarp();
          const uint32_t box = my_row + (nstore & 1) * 4096u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), t1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), t3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + (uint32_t)((j ^ (lane & 7)) * 16)),
                         "r"(*reinterpret_cast<uint32_t*>(&t0)), "r"(*reinterpret_cast<uint32_t*>(&t1)),
                         "r"(*reinterpret_cast<uint32_t*>(&t2)), "r"(*reinterpret_cast<uint32_t*>(&t3)) : "memory");
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (row0 < p.M) ptx::p_tma_store_2d(&tmC, stg + (nstore & 1) * 4096u, n, row0);
            ptx::p_bulk_commit();
          }
          ++nstore;
        }
      }
    }
    if (lane == 0) ptx::p_bulk_wait0();                            // all stores of this warp have landed before the CTA retires
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) ptx::tmem_dealloc(tmem, 512);
  if (threadIdx.x == 0) {                                          // last CTA out re-arms the scheduler slot for a later launch
    __threadfence();
    if (atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
  }
}

static thread_local int g_max_ctas = 0;      // 0 = all SMs
static int g_num_sms = 0;

// Returns 0 on success, < 0 when this kernel does not cover the case (caller falls back), > 0 on error.
int launch_gemm_persist(cudaStream_t st, int M, int N, int K, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                        void* C, int64_t ldc, int out_bf16, const float* bias, int accumulate) {
  if (out_bf16 && accumulate) return -1;
  if (!g_num_sms) {
    int dev = 0;
    S2VT_CHECK_CUDA(cudaGetDevice(&dev));
    S2VT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int G = g_max_ctas > 0 ? (g_max_ctas < g_num_sms ? g_max_ctas : g_num_sms) : g_num_sms;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  rc = a_mn ? make_tmap_any(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64, 2)
            : make_tmap_any(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, PBM, 2);
  if (rc) return rc;
  rc = b_mn ? make_tmap_any(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64, 2)
            : make_tmap_any(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, PBN, 2);
  if (rc) return rc;
  rc = make_tmap_any(&tmC, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, out_bf16 ? 64 : 32, 32, out_bf16 ? 2 : 4);
  if (rc) return rc;
  GemmPersistParams p{};
  p.M = M; p.N = N; p.K = K; p.num_kb = (K + PBK - 1) / PBK;
  p.tiles_n = ceil_div(N, PBN);
  p.tiles = p.tiles_n * ceil_div(M, PBM);
  p.out_bf16 = out_bf16; p.bias = bias;
  // split K so that the unit count fills whole rounds of the G resident CTAs (fp32 outputs only: partials are summed in L2 by TMA)
  int best = 1;
  if (!out_bf16) {
    double best_score = -1.0;
    for (int s = 1; s <= 16; ++s) {
      if (s > 1 && p.num_kb / s < 8) break;
      const int kps = ceil_div(p.num_kb, s);
      if (ceil_div(p.num_kb, kps) != s) continue;                 // no empty slices
      const long long units = (long long)p.tiles * s;
      const long long rounds = (units + G - 1) / G;
      const double eff = (double)units / (double)(rounds * G);
      // every extra slice re-reads nothing but adds M*N*4 bytes of reduction traffic and a drain: charge it against the K work it saves
      const double score = eff - 0.02 * (s - 1) - (s > 1 ? 0.03 : 0.0);
      if (score > best_score + 1e-9) { best_score = score; best = s; }
    }
  }
  p.kb_per_split = ceil_div(p.num_kb, best);
  const int splits = ceil_div(p.num_kb, p.kb_per_split);
  p.total = p.tiles * splits;
  p.reduce = (accumulate || splits > 1) ? 1 : 0;
  if (splits > 1 && !accumulate)
    S2VT_CHECK_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  void* sym = nullptr;
  S2VT_CHECK_CUDA(cudaGetSymbolAddress(&sym, g_gemm_sched));
  p.sched = reinterpret_cast<unsigned int*>(sym) + 2 * (g_launch_seq.fetch_add(1) % SCHED_SLOTS);
  const int grid = p.total < G ? p.total : G;
#define S2VT_LAUNCH_PGEMM(AM, BMJ)                                                                                              \
  do {                                                                                                                          \
    S2VT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_persist_kernel<AM, BMJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM)); \
    gemm_bf16_persist_kernel<AM, BMJ><<<grid, 192, PSMEM, st>>>(tmA, tmB, tmC, p);                                               \
  } while (0)
  if (!a_mn && !b_mn) S2VT_LAUNCH_PGEMM(false, false);
  else if (a_mn && !b_mn) S2VT_LAUNCH_PGEMM(true, false);
  else if (!a_mn && b_mn) S2VT_LAUNCH_PGEMM(false, true);
  else S2VT_LAUNCH_PGEMM(true, true);
#undef S2VT_LAUNCH_PGEMM
  S2VT_CHECK_LAUNCH();
  return 0;
}

void gemm_persist_set_max_ctas(int n) { g_max_ctas = n; }

}  // namespace s2vt
