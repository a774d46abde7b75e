// Persistent LSTM recurrence on Blackwell tensor cores: one thread-block CLUSTER steps a batch tile through all T
// timesteps without leaving the chip.
//
//   cluster = H/32 CTAs (16 for H = 512); CTA c owns hidden units [32c, 32c+32) = 128 gate rows (i,f,g,o x 32)
//   W_hh slice  [128 rows x H] bf16 : loaded ONCE into TENSOR MEMORY (tcgen05.st, 128 lanes x H/2 columns) and used as the
//                                     A operand of every step's MMAs -- the weights never touch shared memory again
//   h_{t-1}^T   [16 batch x H] bf16 : double-buffered in shared memory in the canonical no-swizzle K-major layout (B operand).
//                                     NB = 16 real columns per cluster, or NB = 8 ("wide": twice the clusters, half the exchange
//                                     volume and epilogue work per step; the MMA still runs N = 16 with 8 idle columns)
//   per step:   gates[128 x NB] = W_slice . h_{t-1}^T      tcgen05.mma (A in TMEM) M=128 N=NB K=16, fp32 accumulator in TMEM
//               epilogue warps: tcgen05.ld -> + input-side pre-activation (prefetched one step ahead) -> one MUFU.TANH per gate
//               -> gate exchange through smem -> c,h update (c stays in registers for the whole sequence)
//               -> each warp pushes its 16-byte chunks of h_t straight from registers into the h buffer of EVERY cluster
//                  peer with st.async (DSMEM), completing on the peer's mbarrier: the only inter-CTA sync per step.
//
// There is no grid-wide barrier and no global-memory round trip on the recurrent critical path; the stash for BPTT
// (post-activation gates bf16, cell state fp32) streams to HBM in a kernel-private, fully coalesced layout.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "sm100_err.cuh"
#include <string.h>

namespace s2vt {

int lstm_bf16_error_flag() { return read_sm100_error_flag(); }
int lstm_bf16_error_clear() { return clear_sm100_error_flag(); }
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer);

namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem), "r"(rank));
  return r;
}
// 16 bytes from registers into a PEER's shared memory; completes (complete_tx 16) on the peer's mbarrier
__device__ __forceinline__ void st_async_16(uint32_t dst_cluster, uint4 v, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst_cluster), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// generic smem descriptor: layout_type 0 = no swizzle (core matrices of 8 rows x 16 B), 2 = 128B swizzle
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one MUFU op per activation: tanh.approx.f32 (max rel. error ~2^-11, below bf16 resolution);
// sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace ptx

constexpr int LSTM_NB = 16;                 // batch columns per cluster of the stash layout and of the narrow kernel
constexpr int LSTM_MN = 16;                 // MMA N (tcgen05 M=128 needs N % 16 == 0)

struct LstmFwdParams {
  int T, B, H, n_pre;
  const float* pre;                // [n_pre, B, 4H]
  const float* bias;               // [4H]
  const __nv_bfloat16* w;          // [4H, H] bf16 (TMEM-resident path reads it directly)
  const float* h0;                 // [B,H] or null
  const float* c0;                 // [B,H] or null
  __nv_bfloat16* out;              // [T,B,H]
  __nv_bfloat16* gates;            // kernel-private stash [T][nbt][CS][NB][32][4] or null
  float* cells;                    // kernel-private stash [T][nbt][CS][NB][32] or null
  float* hT;                       // [B,H] or null
  float* cT;                       // [B,H] or null
  long long* trace;                // debug: [TRACE_STEPS][8] clock64 stamps of CTA 0, or null
  int dbg_flags;
  int reverse;                     // 1: processing step s reads / writes time index T-1-s (the reverse direction of a bidirectional LSTM)
  // wave-front coupling of two sweeps that run side by side, each launched once (never with reverse):
  int n_sync;                      // time chunks: chunk k = steps [sync_t[k], sync_t[k+1])
  int sync_t[S2VT_MAX_SYNC + 1];
  unsigned int* signal;            // [n_sync] counters, += 1 per (CTA, batch tile) once its out / stash rows of chunk k are visible, or null
  const unsigned int* wait;        // [n_sync] counters advanced by the producer of `pre`: chunk k is read once wait[k] >= wait_val, or null
  unsigned int wait_val;
};
constexpr int TRACE_STEPS = 32, TRACE_T0 = 16;
__device__ long long g_lstm_trace[TRACE_STEPS * 8];
static bool g_trace_enabled = false;
static int g_last_max_clusters[2] = {-1, -1};   // [wide, narrow] result of cudaOccupancyMaxActiveClusters (debug)
static thread_local int g_tiles_per_cluster = 1;   // 2: two batch tiles per cluster (half the SMs per sweep)
static int g_dbg_flags = 0;                 // 8 = keep W in shared memory (the v1 data path) instead of TMEM
#define S2VT_TRACE(slot)                                                                              \
  do {                                                                                                \
    if (p.trace && blockIdx.x == 0 && t >= TRACE_T0 && t < TRACE_T0 + TRACE_STEPS)                    \
      p.trace[(t - TRACE_T0) * 8 + (slot)] = clock64();                                               \
  } while (0)

// NTL = batch tiles per cluster.  NTL = 2 runs two independent 16-column recurrences on the same resident weight slice, each with its
// own four epilogue warps, barriers, h buffers and accumulator: while one tile's h_t is in flight through the cluster, the other
// tile's MMAs / activations run, so the same sweep needs half the SMs (B = 64 -> 2 clusters instead of 4).
template <int NB, bool W_TMEM, int NTL>
__global__ void __launch_bounds__(32 * (4 * NTL + 1), 1)
lstm_fwd_cluster_kernel(const __grid_constant__ CUtensorMap tmW, const LstmFwdParams p) {
  static_assert(NB == 16 || NB == 8, "thread mapping below assumes 16 or 8 batch columns per tile (4 or 2 per epilogue warp)");
  static_assert(NTL == 1 || (NTL == 2 && NB == 16 && W_TMEM), "two tiles per cluster: 16-column tiles, weights in tensor memory");
  constexpr int MN = LSTM_MN;
  constexpr int CPT = NB / 4;                       // phase-2 columns per thread
  constexpr int NCH = NB;                           // 16-byte h chunks (8 units x 1 column) produced per warp and step
  constexpr int PPL = NCH / 2;                      // peers each lane serves: 32 / NCH lanes share a chunk and split the 16 peers
  constexpr int CTRL = 4 * NTL, EPI_THREADS = 128 * NTL;
  constexpr uint32_t LBO_H = (MN / 8) * 128;        // K-direction stride between 8x16B core matrices of h^T
  constexpr uint32_t SLICE_BYTES = 4 * LBO_H;       // one CTA's 32 hidden units x MN batch columns, bf16
  constexpr uint32_t XCHG_BYTES = 4 * (NB / 8) * 128;   // bytes of it that carry real columns and are exchanged
  constexpr uint32_t TILE_SMEM = 2 * 4 * NB * 32 * 4 + 2 * NB * 32 * 4 * 2 + 4 * 256;   // sG + sSt + sPk of one tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full, h_full[NTL][2], mma_done[NTL];
  __shared__ uint32_t tmem_slot;

  const int H = p.H, KC = H / 64, CS = H / 32;
  const uint32_t W_BYTES = W_TMEM ? 0u : 128u * (uint32_t)H * 2u, HBUF_BYTES = (uint32_t)MN * (uint32_t)H * 2u;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));            // generic pointer to the aligned base

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tl = (NTL > 1 && warp < CTRL) ? (warp >> 2) : 0;                           // batch tile served by this epilogue warp
  const int wq = warp & 3;                                                // its TMEM lane quarter / gate / column group
  const uint32_t c = ptx::cluster_ctarank();                              // hidden-unit slice of this CTA
  const int bt = blockIdx.x / CS;                                         // cluster index
  const int T = p.T;
  const uint32_t tmem_cols = W_TMEM ? (H >= 512 ? 512u : (H >= 256 ? 256u : (H >= 128 ? 128u : 64u))) : 32u;
  // per-tile shared memory: [tile][2 h buffers] first (the MMA's B operands), then each tile's staging areas
  const uint32_t sH_all = sW + W_BYTES;
  uint8_t* gH_all = gen + W_BYTES;
  const uint32_t sH0 = sH_all + (uint32_t)tl * 2 * HBUF_BYTES;
  uint8_t* gH0 = gH_all + (size_t)tl * 2 * HBUF_BYTES;
  uint8_t* gTile = gH_all + (size_t)NTL * 2 * HBUF_BYTES + (size_t)tl * TILE_SMEM;
  float* sG = reinterpret_cast<float*>(gTile);                            // [2][4 gates][NB][32 units] fp32
  __nv_bfloat16* sSt = reinterpret_cast<__nv_bfloat16*>(sG + 2 * 4 * NB * 32);   // [2][NB][32][4] bf16 stash staging
  uint8_t* sPk = reinterpret_cast<uint8_t*>(sSt + 2 * NB * 32 * 4);      // [4 warps][16 chunks][16 B] h packing
  const int b0 = (bt * NTL + tl) * NB;                                    // first batch column of this warp's tile
  const bool tile_on = b0 < p.B;

  if (warp == CTRL && ptx::elect_one()) {
    if (!W_TMEM) ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(ptx::smem_u32(&w_full), 1);
    for (int i = 0; i < NTL; ++i) {
      ptx::mbar_init(ptx::smem_u32(&h_full[i][0]), 1);
      ptx::mbar_init(ptx::smem_u32(&h_full[i][1]), 1);
      ptx::mbar_init(ptx::smem_u32(&mma_done[i]), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), tmem_cols);
    ptx::tmem_relinquish();
  }
  // initial h^T buffer (step 0 input): zeros or h0 in the canonical layout
  if (warp < CTRL) {
    for (int idx = (int)threadIdx.x - 128 * tl; idx < MN * H / 8; idx += 128) {    // one 16-byte chunk (8 k-elements) per iteration
      const int kblk = idx / MN, b = idx % MN;
      uint4 v = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(gH0 + HBUF_BYTES + (size_t)(kblk * (MN / 8) + b / 8) * 128 + (b % 8) * 16) = v;   // idle columns stay 0
      if (p.h0 && b < NB && b0 + b < p.B) {
        const float* src = p.h0 + (long long)(b0 + b) * H + kblk * 8;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(src[0], src[1]), t1 = __floats2bfloat162_rn(src[2], src[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(src[4], src[5]), t3 = __floats2bfloat162_rn(src[6], src[7]);
        v.x = *reinterpret_cast<uint32_t*>(&t0); v.y = *reinterpret_cast<uint32_t*>(&t1);
        v.z = *reinterpret_cast<uint32_t*>(&t2); v.w = *reinterpret_cast<uint32_t*>(&t3);
      }
      *reinterpret_cast<uint4*>(gH0 + (size_t)(kblk * (MN / 8) + b / 8) * 128 + (b % 8) * 16) = v;
    }
    ptx::fence_proxy_async();                                             // generic writes -> visible to the MMA (async proxy)
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_acc0 = W_TMEM ? tmem + (uint32_t)(H / 2) : tmem;    // accumulator columns sit after the weight columns
  const uint32_t tmem_acc = tmem_acc0 + (uint32_t)(tl * MN);
  if (W_TMEM && warp < CTRL) {
    // thread (gate g = wq, unit u = lane) owns TMEM lane 32g+u = gate row g*H + 32c + u of W_hh: two bf16 per 32-bit column
    const uint4* src = reinterpret_cast<const uint4*>(p.w + ((long long)wq * H + 32 * (int)c + lane) * H);
    for (int cc = tl; cc < H / 64; cc += NTL) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(src + cc * 8 + i);
        r[4 * i + 0] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      ptx::tmem_st_32x32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cc * 32), r);
    }
    ptx::tc_wait_st();
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();
  ptx::cluster_arrive();                                                  // every peer's barriers are initialised before
  ptx::cluster_wait();                                                    // anyone signals them remotely

  if (warp == CTRL) {
    // ===================== control thread: (weight load,) per-step MMA issue for every tile =====================
    if (ptx::elect_one()) {
      bool ok = true;
      if (!W_TMEM) {
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&w_full), W_BYTES);
        for (int kc = 0; kc < KC; ++kc)
          for (int g = 0; g < 4; ++g)                                     // rows [g*H + 32c, +32) -> tile rows [32g, 32g+32)
            ptx::tma_load_2d(sW + kc * 16384 + g * 4096, &tmW, ptx::smem_u32(&w_full), kc * 64, g * H + 32 * (int)c);
        ok = ptx::mbar_wait(ptx::smem_u32(&w_full), 0);
        if (!ok) atomicExch(&g_sm100_error, 11);
      }
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, MN, 0, 0);
      uint32_t ph[NTL][2];
      bool on[NTL];
      for (int i = 0; i < NTL; ++i) { ph[i][0] = ph[i][1] = 0; on[i] = (bt * NTL + i) * NB < p.B; }
      const bool have_h0 = p.h0 != nullptr;
      for (int t = 0; t < T && ok; ++t) {
        const int pb = t & 1;
#pragma unroll
        for (int i = 0; i < NTL; ++i) {
          if (!on[i]) continue;
          const uint32_t hb = sH_all + (uint32_t)i * 2 * HBUF_BYTES;
          if (t + 1 < T) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&h_full[i][pb ^ 1]), (uint32_t)CS * XCHG_BYTES);   // h_t lands here
          if (t > 0) {
            ok = ptx::mbar_wait(ptx::smem_u32(&h_full[i][pb]), ph[i][pb]);
            ph[i][pb] ^= 1;
            if (!ok) { atomicExch(&g_sm100_error, 12); break; }
            ptx::fence_proxy_async();                                     // peers' st.async data -> visible to the tensor core
          }
          if (i == 0) S2VT_TRACE(0);
          if (t > 0 || have_h0) {
            ptx::tc_fence_after();
            // descriptors advance by constants: one 16-wide k-step = 2 core-matrix columns of h^T (2*LBO_H bytes)
            uint64_t db = ptx::make_smem_desc(hb + (uint32_t)pb * HBUF_BYTES, LBO_H, 128, 0);
            uint32_t ta = tmem;
            const uint32_t acc = tmem_acc0 + (uint32_t)(i * MN);
#pragma unroll 4
            for (int ks = 0; ks < 4 * KC; ++ks) {
              if (W_TMEM) {
                ptx::mma_bf16_ts(acc, ta, db, idesc, ks != 0 ? 1u : 0u);
              } else {
                const uint64_t da = ptx::make_smem_desc(sW + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2);
                ptx::mma_bf16_ss(acc, da, db, idesc, ks != 0 ? 1u : 0u);
              }
              db += (2 * LBO_H) >> 4;
              ta += 8;
            }
            ptx::mma_commit(ptx::smem_u32(&mma_done[i]));
            if (i == 0) S2VT_TRACE(1);
          } else {
            ptx::mbar_arrive(ptx::smem_u32(&mma_done[i]));                // h_{-1} = 0: nothing to multiply
          }
        }
      }
    }
  } else if (tile_on) {
    // ===================== epilogue warps: tile tl, quarter wq =====================
    const bool tr = (tl == 0) && threadIdx.x == 0;                        // trace stamps come from one thread
    const int g = wq;                         // phase 1: gate row block of this warp (i,f,g,o) == TMEM lane quarter
    const int u = lane;                       // hidden unit within the CTA slice
    const int unit = 32 * (int)c + u;
    const int q = wq;                         // phase 2: column group (columns CPT*q .. CPT*q + CPT - 1)
    float creg[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int b = b0 + q * CPT + j;
      creg[j] = (p.c0 && b < p.B) ? p.c0[(long long)b * H + unit] : 0.f;
    }
    const float bias_g = p.bias[g * H + unit];
    const float sc = (g == 2) ? 1.0f : 0.5f, sh = (g == 2) ? 0.0f : 0.5f;     // tanh for the g gate, sigmoid otherwise
    float pre_cur[NB];
    // columns past the batch are clamped to the last row (finite garbage that is never stored) so that the loads stay
    // unconditional: a select on the loaded value would stall this warp for a full memory latency
    const float* pre_row0 = p.pre + (long long)min(b0, p.B - 1) * 4 * H + g * H + unit;
    const int row_stride = (b0 + NB <= p.B) ? 4 * H : 0;               // ragged last tile: every column reads one valid row...
    const long long step_stride = (long long)p.B * 4 * H;
    auto tm = [&](int s) { return p.reverse ? T - 1 - s : s; };               // processing step -> time index in global memory
    auto load_pre = [&](int s, float (&dst)[NB]) {
      const int t = tm(s);
      if (t < p.n_pre) {
        const float* src = pre_row0 + t * step_stride;
        if (row_stride != 0) {
#pragma unroll
          for (int j = 0; j < NB; ++j) dst[j] = __ldcg(src + j * row_stride);   // (L2: a chunk of pre may be produced while this kernel runs)
        } else {                                                        // ...unless it exists (slow path, partial tile only)
#pragma unroll
          for (int j = 0; j < NB; ++j) dst[j] = __ldcg(src + (long long)(min(b0 + j, p.B - 1) - min(b0, p.B - 1)) * 4 * H);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) dst[j] = bias_g;
      }
    };
    // wave-front coupling: wk / sk = next chunk to wait for / to signal
    int wk = 0, sk = 0;
    bool ok = true;
    auto wait_chunk = [&](int s) {                                        // before the first read of pre at step s
      if (p.wait && wk < p.n_sync && s == p.sync_t[wk]) {
        if (!ptx::wait_counter_geq(p.wait + wk, p.wait_val)) { atomicExch(&g_sm100_error, 14); ok = false; }
        ++wk;
      }
    };
    wait_chunk(0);
    load_pre(0, pre_cur);
    // st.async targets: this lane serves h chunk (m = octet of units, colL = column within the warp's CPT) to PPL peers
    const int chunk = lane & (NCH - 1), colL = chunk >> 2, m = chunk & 3;
    const int bcol = q * CPT + colL;
    const uint32_t chunk_off = c * SLICE_BYTES + (uint32_t)((m * (MN / 8) + bcol / 8) * 128 + (bcol % 8) * 16);
    const int peer0 = (lane / NCH) * PPL;
    uint32_t peer_h[PPL], peer_bar[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      const uint32_t peer = (uint32_t)min(peer0 + i, CS - 1);
      peer_h[i] = ptx::mapa(sH0, peer) + chunk_off;
      peer_bar[i] = ptx::mapa(ptx::smem_u32(&h_full[tl][0]), peer);
    }
    const uint32_t bar_stride = ptx::smem_u32(&h_full[0][1]) - ptx::smem_u32(&h_full[0][0]);
    uint8_t* myPk = sPk + wq * 256;
    // the stash keeps the 16-column block geometry of the BPTT kernel; a wide cluster fills half a block
    const long long stash_blk = (long long)((p.B + LSTM_NB - 1) / LSTM_NB) * CS;      // blocks per timestep
    const int bt16 = b0 / LSTM_NB, colbase = b0 % LSTM_NB;
    const bool have_h0 = p.h0 != nullptr;
    const uint32_t mma_bar = ptx::smem_u32(&mma_done[tl]);
    for (int t = 0; t < T; ++t) {
      const int sb = t & 1;
      float* sGb = sG + sb * (4 * NB * 32);
      __nv_bfloat16* sStb = sSt + sb * (NB * 32 * 4);
      // ---- phase 1: accumulator + pre-activation -> activation -> smem
      ok = ok && ptx::mbar_wait(mma_bar, (uint32_t)(t & 1));
      if (!ok) { atomicExch(&g_sm100_error, 13); break; }
      if (tr) S2VT_TRACE(2);
      float x[NB];
      if (t > 0 || have_h0) {
        ptx::tc_fence_after();
        uint32_t r[NB];
        if constexpr (NB == 16) ptx::tmem_ld_32x16(tmem_acc + ((uint32_t)(wq * 32) << 16), r);
        else ptx::tmem_ld_32x8(tmem_acc + ((uint32_t)(wq * 32) << 16), r);
        ptx::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < NB; ++j) x[j] = __uint_as_float(r[j]) + pre_cur[j];
        ptx::tc_fence_before();
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) x[j] = pre_cur[j];
      }
      if (t + 1 < T) {
        wait_chunk(t + 1);
        load_pre(t + 1, pre_cur);                                         // prefetch: latency hides behind the rest of the step
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float a = fmaf(sc, ptx::fast_tanh(sc * x[j]), sh);
        sGb[(g * NB + j) * 32 + u] = a;
        sStb[(j * 32 + u) * 4 + g] = __float2bfloat16(a);
      }
      if (tr) S2VT_TRACE(3);
      ptx::named_bar_sync(1 + tl, 128);
      if (tr) S2VT_TRACE(4);
      // ---- phase 2: cell / hidden update for (unit u, columns CPT*q ..)
      float hval[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int col = q * CPT + j;
        const float gi = sGb[(0 * NB + col) * 32 + u], gf = sGb[(1 * NB + col) * 32 + u];
        const float gg = sGb[(2 * NB + col) * 32 + u], go = sGb[(3 * NB + col) * 32 + u];
        const float cn = gf * creg[j] + gi * gg;
        creg[j] = cn;
        hval[j] = go * ptx::fast_tanh(cn);
        *reinterpret_cast<__nv_bfloat16*>(myPk + ((j * 4 + (u >> 3)) * 8 + (u & 7)) * 2) = __float2bfloat16(hval[j]);
      }
      __syncwarp();
      const uint4 hchunk = *reinterpret_cast<const uint4*>(myPk + chunk * 16);    // 8 units x 1 column, bf16
      if (tr) S2VT_TRACE(5);
      // ---- h_t to every peer's next-step buffer (critical path first)
      if (t + 1 < T) {
        const uint32_t boff = (uint32_t)((t + 1) & 1);
#pragma unroll
        for (int i = 0; i < PPL; ++i)
          if (peer0 + i < CS) ptx::st_async_16(peer_h[i] + boff * HBUF_BYTES, hchunk, peer_bar[i] + boff * bar_stride);
      }
      if (tr) S2VT_TRACE(6);
      // ---- off the critical path: h_t, c_t and the gate activations to HBM
      if (lane < NCH && b0 + bcol < p.B)
        *reinterpret_cast<uint4*>(p.out + ((long long)tm(t) * p.B + b0 + bcol) * H + 32 * (int)c + 8 * m) = hchunk;
      const long long blk = (long long)tm(t) * stash_blk + (long long)bt16 * CS + c;
      if (p.cells) {
        float* cdst = p.cells + blk * (LSTM_NB * 32) + colbase * 32 + u;
#pragma unroll
        for (int j = 0; j < CPT; ++j) cdst[(q * CPT + j) * 32] = creg[j];
      }
      if (p.gates) {
        const uint4* ssrc = reinterpret_cast<const uint4*>(sStb);
        uint4* gdst = reinterpret_cast<uint4*>(p.gates + blk * (LSTM_NB * 32 * 4) + colbase * 32 * 4);
        const int tid = (int)threadIdx.x - 128 * tl;
#pragma unroll
        for (int r = 0; r < NB / 8; ++r) gdst[tid + 128 * r] = ssrc[tid + 128 * r];
      }
      if (t == T - 1) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int b = b0 + q * CPT + j;
          if (b < p.B) {
            if (p.hT) p.hT[(long long)b * H + unit] = hval[j];
            if (p.cT) p.cT[(long long)b * H + unit] = creg[j];
          }
        }
      }
      if (p.signal && sk < p.n_sync && t + 1 == p.sync_t[sk + 1]) {       // chunk complete: publish it to the consumer of out
        __threadfence();
        ptx::named_bar_sync(1 + tl, 128);
        if ((int)threadIdx.x == 128 * tl) ptx::red_release_gpu_add(p.signal + sk, 1u);
        ++sk;
      }
      if (tr) S2VT_TRACE(7);
    }
    if (p.signal && (int)threadIdx.x == 128 * tl)                         // (error exit: never leave a consumer waiting)
      for (; sk < p.n_sync; ++sk) ptx::red_release_gpu_add(p.signal + sk, 1u);
  }
  // no CTA may exit while peers can still write into its shared memory
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 0) ptx::tmem_dealloc(tmem, tmem_cols);
}

template <int NB, bool W_TMEM, int NTL>
static int launch_lstm_fwd(cudaStream_t st, const CUtensorMap& tmW, const LstmFwdParams& p) {
  const int H = p.H, CS = H / 32;
  const size_t tile_smem = 2 * (size_t)4 * NB * 32 * 4 + 2 * (size_t)NB * 32 * 4 * 2 + 4 * 256;
  const size_t smem_need = 1024 + (W_TMEM ? 0 : (size_t)128 * H * 2) + (size_t)NTL * (2 * (size_t)LSTM_MN * H * 2 + tile_smem);
  // This CTA owns all of the SM's tensor memory (512 columns for H = 512): a co-resident GEMM CTA from another stream would block
  // in tcgen05.alloc until the sweep ends and would contend for the SM meanwhile.  Asking for > (227 - 97) KB keeps them out.
  const size_t smem = smem_need < (size_t)136 * 1024 ? (size_t)136 * 1024 : smem_need;
  auto kern = lstm_fwd_cluster_kernel<NB, W_TMEM, NTL>;
  S2VT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) S2VT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  const int n_clusters = ceil_div(p.B, NB * NTL);
  cfg.gridDim = dim3(CS * n_clusters);
  cfg.blockDim = dim3(32 * (4 * NTL + 1));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  S2VT_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  g_last_max_clusters[NB == 8 ? 0 : 1] = max_clusters;
  S2VT_REQUIRE(max_clusters >= 1, "s2vt_lstm_fwd_bf16: a cluster of %d CTAs with %zu B of shared memory cannot be scheduled on this device", CS, smem);
  if (NB < LSTM_NB && max_clusters < n_clusters) return -1;             // the wide form must be fully co-resident; caller falls back
  S2VT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmW, p));
  count_launch();
  return 0;
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int64_t s2vt_lstm_bf16_batch_pad(int B) { return (int64_t)ceil_div(B, LSTM_NB) * LSTM_NB; }

extern "C" int s2vt_lstm_fwd_bf16(void* stream, int T, int B, int H, int n_pre,
                                  const float* pre, const float* bias_sum, const void* w_hh_bf16,
                                  const float* h0, const float* c0,
                                  void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT) {
  return s2vt_lstm_fwd_bf16_dir(stream, T, B, H, n_pre, pre, bias_sum, w_hh_bf16, h0, c0, out_bf16, gates_bf16, cells, hT, cT, 0);
}

extern "C" int s2vt_lstm_fwd_bf16_dir(void* stream, int T, int B, int H, int n_pre,
                                      const float* pre, const float* bias_sum, const void* w_hh_bf16,
                                      const float* h0, const float* c0,
                                      void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT, int reverse) {
  return s2vt_lstm_fwd_bf16_sync(stream, T, B, H, n_pre, pre, bias_sum, w_hh_bf16, h0, c0, out_bf16, gates_bf16, cells, hT, cT, reverse,
                                 g_tiles_per_cluster, 0, nullptr, nullptr, nullptr, 0);
}

extern "C" int s2vt_lstm_fwd_bf16_sync(void* stream, int T, int B, int H, int n_pre,
                                       const float* pre, const float* bias_sum, const void* w_hh_bf16,
                                       const float* h0, const float* c0,
                                       void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT, int reverse,
                                       int tiles_per_cluster, int n_sync, const int* sync_t, unsigned int* signal,
                                       const unsigned int* wait, unsigned int wait_val) {
  S2VT_REQUIRE(T >= 1 && B >= 1, "s2vt_lstm_fwd_bf16: bad dims");
  S2VT_REQUIRE(n_sync >= 0 && n_sync <= S2VT_MAX_SYNC, "s2vt_lstm_fwd_bf16_sync: at most %d chunks", S2VT_MAX_SYNC);
  S2VT_REQUIRE(n_sync == 0 || (sync_t && (signal || wait) && !reverse), "s2vt_lstm_fwd_bf16_sync: chunks need sync_t and a counter array, and the forward direction");
  S2VT_REQUIRE(tiles_per_cluster == 1 || tiles_per_cluster == 2, "s2vt_lstm_fwd_bf16_sync: tiles_per_cluster must be 1 or 2");
  S2VT_REQUIRE(H % 64 == 0 && H >= 64 && H <= 512, "s2vt_lstm_fwd_bf16: the cluster-resident kernel needs H %% 64 == 0 and 64 <= H <= 512 (got %d)", H);
  S2VT_REQUIRE(bias_sum && w_hh_bf16 && out_bf16, "s2vt_lstm_fwd_bf16: null pointer");
  S2VT_REQUIRE(n_pre <= 0 || pre, "s2vt_lstm_fwd_bf16: pre is null but n_pre > 0");
  S2VT_REQUIRE((h0 == nullptr) == (c0 == nullptr), "s2vt_lstm_fwd_bf16: h0 and c0 must be given together");
  S2VT_REQUIRE(aligned16(w_hh_bf16) && aligned16(out_bf16) && (!gates_bf16 || aligned16(gates_bf16)), "s2vt_lstm_fwd_bf16: buffers must be 16-byte aligned");
  LstmFwdParams p{};
  p.T = T; p.B = B; p.H = H; p.n_pre = n_pre < 0 ? 0 : n_pre;
  p.pre = pre; p.bias = bias_sum; p.w = (const __nv_bfloat16*)w_hh_bf16; p.h0 = h0; p.c0 = c0;
  p.out = (__nv_bfloat16*)out_bf16; p.gates = (__nv_bfloat16*)gates_bf16; p.cells = cells; p.hT = hT; p.cT = cT;
  p.trace = nullptr;
  p.dbg_flags = g_dbg_flags;
  p.reverse = reverse ? 1 : 0;
  p.n_sync = n_sync; p.signal = n_sync ? signal : nullptr; p.wait = n_sync ? wait : nullptr; p.wait_val = wait_val;
  for (int i = 0; i <= n_sync && n_sync > 0; ++i) {
    S2VT_REQUIRE(sync_t[i] >= 0 && sync_t[i] <= T && (i == 0 || sync_t[i] > sync_t[i - 1]), "s2vt_lstm_fwd_bf16_sync: sync_t must increase from 0 to T");
    p.sync_t[i] = sync_t[i];
  }
  S2VT_REQUIRE(n_sync == 0 || (p.sync_t[0] == 0 && p.sync_t[n_sync] == T), "s2vt_lstm_fwd_bf16_sync: sync_t must start at 0 and end at T");
  if (g_trace_enabled) {
    void* sym = nullptr;
    S2VT_CHECK_CUDA(cudaGetSymbolAddress(&sym, g_lstm_trace));
    p.trace = (long long*)sym;
  }
  CUtensorMap tmW;
  if (g_dbg_flags & 8) {
    int rc = make_tmap_bf16(&tmW, w_hh_bf16, (uint64_t)H, (uint64_t)4 * H, (uint64_t)H, 64, 32);
    if (rc) return rc;
    return launch_lstm_fwd<LSTM_NB, false, 1>((cudaStream_t)stream, tmW, p);
  }
  memset(&tmW, 0, sizeof(tmW));
  // wide form (8 columns per cluster): half the DSMEM exchange and epilogue work per step, when all ceil(B/8) clusters fit at once
  if (!(g_dbg_flags & 32) && B > 8 && ceil_div(B, 8) * (H / 32) <= 128 && n_sync == 0 && tiles_per_cluster == 1) {
    const int rc = launch_lstm_fwd<8, true, 1>((cudaStream_t)stream, tmW, p);
    if (rc >= 0) return rc;
  }
  if ((g_dbg_flags & 64) || tiles_per_cluster == 2) return launch_lstm_fwd<LSTM_NB, true, 2>((cudaStream_t)stream, tmW, p);
  return launch_lstm_fwd<LSTM_NB, true, 1>((cudaStream_t)stream, tmW, p);
}

// debug aids (not part of the product path): per-step clock64 stamps of CTA 0 for steps [16, 48)
extern "C" int s2vt_lstm_bf16_set_tiles_per_cluster(int n) { g_tiles_per_cluster = (n == 2) ? 2 : 1; return 0; }
extern "C" int s2vt_debug_max_clusters(int which) { return g_last_max_clusters[which & 1]; }
extern "C" int s2vt_debug_trace_enable(int on) { g_trace_enabled = on != 0; return 0; }
extern "C" int s2vt_debug_set_flags(int flags) { g_dbg_flags = flags; return 0; }
extern "C" int s2vt_debug_trace_read(long long* host_out, int n) {
  if (n > TRACE_STEPS * 8) n = TRACE_STEPS * 8;
  S2VT_CHECK_CUDA(cudaDeviceSynchronize());
  S2VT_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_lstm_trace, sizeof(long long) * n));
  return 0;
}
