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
int gemm_persist_error_clear() { return clear_sm100_error_flag(); }

constexpr int PBM = 128, PBN = 256, PBK = 64, PSTAGES = 4, PRING = 4, PNSTG = 2;
constexpr int PA_BYTES = PBM * PBK * 2, PB_BYTES = PBN * PBK * 2, PSTAGE_BYTES = PA_BYTES + PB_BYTES;
constexpr int PSTG_WARP = PNSTG * 4096;                               // per epilogue warp: PNSTG 32-row x 128-byte staging boxes
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
  // fused cross-entropy statistics (vocab projection, bf16 output only): per (row, n-tile) online-softmax partials of the fp32
  // logits BEFORE they are rounded to bf16, and the fp32 logit of each row's target
  float2* ce_part;                  // [M][tiles_n] (max, sum exp(z - max)) or null
  float* ce_ztgt;                   // [M]
  const int64_t* ce_targets;
  RowMap ce_tmap;
  // wave-front gating (s2vt_gemm_bf16_gated): the rows of A are produced chunk by chunk by a recurrence sweep running beside this
  // kernel, and the rows of C are consumed chunk by chunk by a second sweep.  Chunk k = rows [sync_row[k], sync_row[k+1]).
  int n_sync, reverse_m, tiles_m;
  int sync_row[S2VT_MAX_SYNC + 1];
  int sync_expect[S2VT_MAX_SYNC];   // epilogue-warp completions after which chunk k of C is whole (4 x the units overlapping it)
  const unsigned int* sync_wait;    // [n_sync] a unit reads A only once sync_wait[k] >= sync_wait_val for every chunk it overlaps
  unsigned int sync_wait_val;
  unsigned int* sync_done;          // [n_sync] completion counters (zeroed by the caller)
  unsigned int* sync_ready;         // [n_sync] += 1 when chunk k of C is whole and visible device-wide
};

// first row of the unit's tile: tiles are walked from the last row block to the first when the producing sweep runs backwards in time
__device__ __forceinline__ int unit_m0(const GemmPersistParams& p, int tile) {
  const int mt = tile / p.tiles_n;
  return (p.reverse_m ? p.tiles_m - 1 - mt : mt) * PBM;
}

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
__device__ __forceinline__ void p_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PNSTG - 1) : "memory"); }
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
        const int n0 = (tile % p.tiles_n) * PBN, m0 = unit_m0(p, tile);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        bool ok = true;
        if (p.n_sync) {                                              // the sweep beside us has published these rows of A?
          const int m1 = min(p.M, m0 + PBM);
          for (int k = 0; k < p.n_sync && ok; ++k)
            if (p.sync_row[k] < m1 && p.sync_row[k + 1] > m0) ok = ptx::wait_counter_geq(p.sync_wait + k, p.sync_wait_val);
          if (!ok) { atomicExch(&g_sm100_error, 38); break; }
          asm volatile("fence.proxy.async;" ::: "memory");          // acquire (generic proxy) before the TMA reads (async proxy)
        }
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
      const int n0 = (tile % p.tiles_n) * PBN, m0 = unit_m0(p, tile);
      const uint32_t a = ai & 1;
      // bias of the tile's 256 columns, spread over the lanes (lane l holds columns 32k + l); fetched before the accumulator wait,
      // handed to every row with shuffles: a load inside the chunk loop would put an L2 round trip on each chunk's critical path
      const bool add_bias = p.bias != nullptr && split == 0;
      float breg[PBN / 32];
#pragma unroll
      for (int k = 0; k < PBN / 32; ++k) {
        const int col = n0 + 32 * k + lane;
        breg[k] = (add_bias && col < p.N) ? __ldg(p.bias + col) : 0.f;
      }
      if (!ptx::mbar_wait(ptx::smem_u32(&acc_full[a]), (ai >> 1) & 1)) { atomicExch(&g_sm100_error, 37); break; }
      ++ai;
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem + ((uint32_t)(q * 32) << 16) + a * PBN;
      const int row0 = m0 + q * 32;                                // first row of this warp's slab
      if (!p.out_bf16) {
        const int nchunk = min(PBN / 32, (p.N - n0 + 31) / 32);
#pragma unroll
        for (int c = 0; c < PBN / 32; ++c) {
          if (c >= nchunk) break;
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
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, breg[c], j);
          }
          if (lane == 0) ptx::p_bulk_wait_read1();                 // the store that used this staging box two chunks ago has read it
          __syncwarp();
          const uint32_t box = my_row + (nstore % PNSTG) * 4096u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(box + (uint32_t)((j ^ (lane & 7)) * 16)),
                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            if (p.reduce) ptx::p_tma_reduce_add_2d(&tmC, stg + (nstore % PNSTG) * 4096u, n, row0);
            else ptx::p_tma_store_2d(&tmC, stg + (nstore % PNSTG) * 4096u, n, row0);
            ptx::p_bulk_commit();
          } else if (lane == 0) {
            ptx::p_bulk_commit();                                   // keep the group count in step with nstore
          }
          ++nstore;
        }
      } else {
        const int nchunk = min(PBN / 64, (p.N - n0 + 63) / 64);
        const int my_m = row0 + lane;
        float ce_m = -INFINITY, ce_s = 0.f;
        long long ce_tgt = -1;
        if (p.ce_part && my_m < p.M) ce_tgt = p.ce_targets[p.ce_tmap(my_m)];
#pragma unroll
        for (int c = 0; c < PBN / 64; ++c) {
          if (c >= nchunk) break;
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
            for (int j = 0; j < 32; ++j) {
              v[j] += __shfl_sync(0xffffffffu, breg[2 * c], j);
              v[32 + j] += __shfl_sync(0xffffffffu, breg[2 * c + 1], j);
            }
          }
          if (p.ce_part) {
            constexpr float LOG2E = 1.4426950408889634f;
            if (n + 64 > p.N) {
#pragma unroll
              for (int j = 0; j < 64; ++j) if (n + j >= p.N) v[j] = -INFINITY;
            }
            float cm = v[0];
#pragma unroll
            for (int j = 1; j < 64; ++j) cm = fmaxf(cm, v[j]);
            const float mn = fmaxf(ce_m, cm);
            const float off = mn * LOG2E;
            float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              cs0 += exp2f(fmaf(v[j], LOG2E, -off));
              cs1 += exp2f(fmaf(v[j + 1], LOG2E, -off));
            }
            ce_s = ce_s * exp2f((ce_m - mn) * LOG2E) + (cs0 + cs1);
            ce_m = mn;
            if ((unsigned long long)(ce_tgt - n) < 64ull) {          // rare: exactly one chunk per row holds the target column
              float zt = 0.f;
#pragma unroll
              for (int j = 0; j < 64; ++j) if (n + j == ce_tgt) zt = v[j];
              p.ce_ztgt[my_m] = zt;
            }
          }
          if (lane == 0) ptx::p_bulk_wait_read1();
          __syncwarp();
          const uint32_t box = my_row + (nstore % PNSTG) * 4096u;
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
            if (row0 < p.M) ptx::p_tma_store_2d(&tmC, stg + (nstore % PNSTG) * 4096u, n, row0);
            ptx::p_bulk_commit();
          }
          ++nstore;
        }
        if (p.ce_part && my_m < p.M) p.ce_part[(long long)my_m * p.tiles_n + (tile % p.tiles_n)] = make_float2(ce_m, ce_s);
      }
      if (p.n_sync) {                                              // publish: this warp's 32 x 256 block of C has landed
        if (lane == 0) {
          ptx::p_bulk_wait0();
          asm volatile("fence.proxy.async;" ::: "memory");        // async-proxy writes before the generic-proxy release below
          __threadfence();
          const int m1 = min(p.M, m0 + PBM);
          for (int k = 0; k < p.n_sync; ++k) {
            if (p.sync_row[k] < m1 && p.sync_row[k + 1] > m0) {
              if (atomicAdd(p.sync_done + k, 1u) == (unsigned)p.sync_expect[k] - 1u) {
                __threadfence();
                ptx::red_release_gpu_add(p.sync_ready + k, 1u);
              }
            }
          }
        }
        __syncwarp();
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
struct GemmGate {                    // s2vt_gemm_bf16_gated
  int n_sync, reverse_m, max_ctas;
  const int* sync_row;
  const unsigned int* wait;
  unsigned int wait_val;
  unsigned int *done, *ready;
};

int launch_gemm_persist_ce(cudaStream_t st, int M, int N, int K, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                           void* C, int64_t ldc, int out_bf16, const float* bias, int accumulate,
                           float2* ce_part, float* ce_ztgt, const int64_t* ce_targets, RowMap ce_tmap, const GemmGate* gate = nullptr) {
  if (out_bf16 && accumulate) return -1;
  if (!g_num_sms) {
    int dev = 0;
    S2VT_CHECK_CUDA(cudaGetDevice(&dev));
    S2VT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int want_ctas = gate ? gate->max_ctas : g_max_ctas;
  const int G = want_ctas > 0 ? (want_ctas < g_num_sms ? want_ctas : g_num_sms) : g_num_sms;
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
  p.tiles_m = ceil_div(M, PBM);
  p.tiles = p.tiles_n * p.tiles_m;
  p.out_bf16 = out_bf16; p.bias = bias;
  if (gate && gate->n_sync > 0) {
    p.n_sync = gate->n_sync; p.reverse_m = gate->reverse_m;
    p.sync_wait = gate->wait; p.sync_wait_val = gate->wait_val; p.sync_done = gate->done; p.sync_ready = gate->ready;
    for (int k = 0; k <= gate->n_sync; ++k) p.sync_row[k] = gate->sync_row[k];
    for (int k = 0; k < gate->n_sync; ++k) {
      int tiles_over = 0;
      for (int mt = 0; mt < p.tiles_m; ++mt) {
        const int m0 = mt * PBM, m1 = M < m0 + PBM ? M : m0 + PBM;
        if (p.sync_row[k] < m1 && p.sync_row[k + 1] > m0) ++tiles_over;
      }
      p.sync_expect[k] = 4 * tiles_over * p.tiles_n;
    }
  }
  p.ce_part = ce_part; p.ce_ztgt = ce_ztgt; p.ce_targets = ce_targets; p.ce_tmap = ce_tmap;
  // split K so that the unit count fills whole rounds of the G resident CTAs (fp32 outputs only: partials are summed in L2 by TMA)
  int best = 1;
  if (!out_bf16 && !(gate && gate->n_sync > 0)) {                 // (a gated product publishes whole tiles: no K slices)
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

int launch_gemm_persist(cudaStream_t st, int M, int N, int K, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                        void* C, int64_t ldc, int out_bf16, const float* bias, int accumulate) {
  return launch_gemm_persist_ce(st, M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, out_bf16, bias, accumulate, nullptr, nullptr, nullptr, RowMap{1, 1, 0});
}

void gemm_persist_set_max_ctas(int n) { g_max_ctas = n; }

int launch_gemm_persist_gated(cudaStream_t st, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, int b_mn,
                              void* C, int64_t ldc, const float* bias, int accumulate, const GemmGate& gate) {
  return launch_gemm_persist_ce(st, M, N, K, A, lda, 0, B, ldb, b_mn, C, ldc, 0, bias, accumulate, nullptr, nullptr, nullptr, RowMap{1, 1, 0}, &gate);
}

// one warp per row: merge the per-tile (max, sum-exp) partials -> log-sum-exp; row loss = lse - z[target]
__global__ void ce_combine_kernel(const float2* __restrict__ part, int tiles_n, const float* __restrict__ ztgt, long long R,
                                  float* __restrict__ row_lse, float* __restrict__ row_loss) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  float m = -INFINITY, s = 0.f;
  for (int j = lane; j < tiles_n; j += 32) {
    const float2 v = part[r * tiles_n + j];
    const float mn = fmaxf(m, v.x);
    s = s * __expf(m - mn) + v.y * __expf(v.x - mn);
    m = mn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) {
    const float lse = m + logf(s);
    row_lse[r] = lse;
    if (row_loss) row_loss[r] = lse - ztgt[r];
  }
}

__global__ void ce_mean_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  float a = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float b = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (threadIdx.x == 0) out[0] = b / (float)n;
  }
}

// in place: z (bf16 logits) -> (softmax - onehot) * gscale / R as bf16; one block per row, 16-byte accesses
__global__ void __launch_bounds__(256) ce_dlogits_inplace_kernel(__nv_bfloat16* __restrict__ z, int V, long long ld, const float* __restrict__ row_lse,
                                                                 const int64_t* __restrict__ targets, RowMap tmap, const float* __restrict__ gscale,
                                                                 float inv_rows) {
  const long long r = blockIdx.x;
  __nv_bfloat16* row = z + r * ld;
  const float lse = row_lse[r];
  const long long tgt = targets[tmap(r)];
  const float sc = (gscale ? gscale[0] : 1.f) * inv_rows;
  constexpr float LOG2E = 1.4426950408889634f;
  const float off = lse * LOG2E;
  const int nv = V / 8;
  uint4* row4 = reinterpret_cast<uint4*>(row);
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    uint4 v = row4[j];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const long long j0 = 8ll * j;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = exp2f(fmaf(__uint_as_float(w[k] << 16), LOG2E, -off));
      float b = exp2f(fmaf(__uint_as_float(w[k] & 0xffff0000u), LOG2E, -off));
      if (j0 + 2 * k == tgt) a -= 1.f;
      if (j0 + 2 * k + 1 == tgt) b -= 1.f;
      const __nv_bfloat162 t2 = __floats2bfloat162_rn(a * sc, b * sc);
      w[k] = *reinterpret_cast<const uint32_t*>(&t2);
    }
    row4[j] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int j = nv * 8 + threadIdx.x; j < V; j += blockDim.x) {
    float a = exp2f(fmaf(__bfloat162float(row[j]), LOG2E, -off));
    if (j == tgt) a -= 1.f;
    row[j] = __float2bfloat16(a * sc);
  }
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_vocab_ce_fwd_bf16(void* stream, int R, int V, int K, const void* A_bf16, int64_t lda, const void* W_bf16, int64_t ldw,
                                      const float* bias, void* logits_bf16, int64_t ldl, const int64_t* targets, s2vt_rowmap tmap,
                                      void* part_ws, float* ztgt_ws, float* row_lse, float* row_loss, float* loss) {
  S2VT_REQUIRE(R > 0 && V > 0 && K > 0, "s2vt_vocab_ce_fwd_bf16: bad dims");
  S2VT_REQUIRE(A_bf16 && W_bf16 && logits_bf16 && targets && part_ws && ztgt_ws && row_lse, "s2vt_vocab_ce_fwd_bf16: null pointer");
  S2VT_REQUIRE(!loss || row_loss, "s2vt_vocab_ce_fwd_bf16: loss needs row_loss scratch");
  S2VT_REQUIRE(aligned16(logits_bf16) && ldl % 8 == 0 && ldl >= V, "s2vt_vocab_ce_fwd_bf16: logits must be 16-byte aligned with ld %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_gemm_persist_ce(st, R, V, K, A_bf16, lda, 0, W_bf16, ldw, 0, logits_bf16, ldl, 1, bias, 0, (float2*)part_ws, ztgt_ws, targets,
                                  to_rowmap(tmap));
  if (rc) return rc < 0 ? fail("s2vt_vocab_ce_fwd_bf16: unsupported configuration") : rc;
  const int tiles_n = ceil_div(V, PBN);
  ce_combine_kernel<<<ceil_div(R, 8), 256, 0, st>>>((const float2*)part_ws, tiles_n, ztgt_ws, R, row_lse, row_loss);
  S2VT_CHECK_LAUNCH();
  if (loss) {
    ce_mean_kernel<<<1, 1024, 0, st>>>(row_loss, R, loss);
    S2VT_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int s2vt_gemm_bf16_gated(void* stream, int M, int N, int K, const void* A_bf16, int64_t lda, const void* B_bf16, int64_t ldb,
                                    int b_mn_major, float* C, int64_t ldc, const float* bias, int accumulate, int max_ctas, int reverse_m,
                                    int n_sync, const int* sync_row, const unsigned int* wait, unsigned int wait_val,
                                    unsigned int* done, unsigned int* ready) {
  S2VT_REQUIRE(M > 0 && N > 0 && K > 0, "s2vt_gemm_bf16_gated: dimensions must be positive (M=%d N=%d K=%d)", M, N, K);
  S2VT_REQUIRE(A_bf16 && B_bf16 && C, "s2vt_gemm_bf16_gated: null operand");
  S2VT_REQUIRE(n_sync >= 1 && n_sync <= S2VT_MAX_SYNC && sync_row && wait && done && ready,
               "s2vt_gemm_bf16_gated: needs 1..%d chunks with their row bounds and counters", S2VT_MAX_SYNC);
  S2VT_REQUIRE(sync_row[0] == 0 && sync_row[n_sync] == M, "s2vt_gemm_bf16_gated: sync_row must run from 0 to M");
  for (int k = 0; k < n_sync; ++k) S2VT_REQUIRE(sync_row[k + 1] > sync_row[k], "s2vt_gemm_bf16_gated: sync_row must increase");
  S2VT_REQUIRE(aligned16(C) && ldc % 4 == 0 && ldc >= N, "s2vt_gemm_bf16_gated: C must be dense, 16-byte aligned rows");
  S2VT_REQUIRE(max_ctas >= 1, "s2vt_gemm_bf16_gated: max_ctas must be >= 1");
  GemmGate gate{n_sync, reverse_m ? 1 : 0, max_ctas, sync_row, wait, wait_val, done, ready};
  const int rc = launch_gemm_persist_gated((cudaStream_t)stream, M, N, K, A_bf16, lda, B_bf16, ldb, b_mn_major, C, ldc, bias, accumulate, gate);
  if (rc < 0) return fail("s2vt_gemm_bf16_gated: unsupported configuration");
  return rc;
}

extern "C" int64_t s2vt_vocab_ce_ws_bytes(int R, int V) { return (int64_t)R * ceil_div(V, PBN) * (int64_t)sizeof(float2); }

extern "C" int s2vt_ce_dlogits_inplace_bf16(void* stream, void* logits_bf16, int64_t R, int V, int64_t ld, const float* row_lse,
                                            const int64_t* targets, s2vt_rowmap tmap, const float* gscale) {
  S2VT_REQUIRE(logits_bf16 && row_lse && targets, "s2vt_ce_dlogits_inplace_bf16: null pointer");
  S2VT_REQUIRE(aligned16(logits_bf16) && ld % 8 == 0, "s2vt_ce_dlogits_inplace_bf16: logits must be 16-byte aligned with ld %% 8 == 0");
  if (R == 0) return 0;
  ce_dlogits_inplace_kernel<<<(unsigned)R, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)logits_bf16, V, ld, row_lse, targets, to_rowmap(tmap), gscale,
                                                                          1.0f / (float)R);
  S2VT_CHECK_LAUNCH();
  return 0;
}
