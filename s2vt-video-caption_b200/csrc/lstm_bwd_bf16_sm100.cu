// Persistent BPTT of one LSTM layer on Blackwell tensor cores: the backward twin of lstm_bf16_sm100.cu.
//
//   cluster = H/32 CTAs; CTA c owns hidden units [32c, 32c+32) = 128 gate rows r (i,f,g,o x 32), batch tile of 16 columns
//   per step t = T-1 .. 0
//     dh_t   = dL/dh_t (from the layer above, fp32) + dh_rec          dh_rec: recurrent gradient, reduced from 16 partials
//     dgates = pointwise(dh_t, dc, stash_t)                           fused gate gradients, dc stays in registers
//     -> bf16 dgates_t go (a) to HBM in GEMM layout [T*B, 4H] for the time-batched weight-gradient products and
//                         (b) into a 4 KB shared-memory B operand (canonical no-swizzle K-major layout)
//     partial[j, b] = sum_{r in rows_c} W_hh[r, j] * dgates[b, r]     tcgen05.mma, A = W_hh[rows_c, :]^T resident in TMEM
//                                                                     (H/128 tiles of M=128 x K=128), fp32 accumulators in TMEM
//     reduce-scatter: the 32 x 16 block of partial that belongs to CTA d's units is rounded to bf16 (the cluster's DSMEM fabric,
//                     not the tensor pipe, bounds the step: 16 KB per CTA per step instead of 32 KB) and pushed into d's receive
//                     buffer with st.async, completing on d's mbarrier; d sums the 16 blocks in fp32 at the start of step t-1.
//
// As in the forward kernel there is no grid barrier and no global-memory round trip on the critical path.
#include "common.cuh"
#include "sm100_cluster.cuh"
#include "sm100_err.cuh"

namespace s2vt {

int lstm_bwd_bf16_error_flag() { return read_sm100_error_flag(); }
int lstm_bwd_bf16_error_clear() { return clear_sm100_error_flag(); }

constexpr int BWD_NB = 16;

struct LstmBwdParams {
  int T, B, H, dout_t0;
  const float* dout;               // [T,B,H] fp32 or null; rows t < dout_t0 are zero and never read
  const __nv_bfloat16* gates;      // forward stash [T][nbt][CS][16][32][4]
  const float* cells;              // forward stash [T][nbt][CS][16][32]
  const __nv_bfloat16* w_t;        // W_hh^T bf16 [H, 4H]
  __nv_bfloat16* dgates;           // [T,B,4H] bf16 out
  int reverse;                     // 1: processing step s <-> time index T-1-s in every global buffer (reverse direction of a BiLSTM)
  // chaining a long sweep through several launches (time chunks, latest chunk first):
  const float* dh_in;              // [B,H] recurrent gradient flowing into the chunk's last step (from the chunk after it) or null
  const float* dc_in;              // [B,H] cell gradient flowing into the chunk's last step or null
  float* dh_out;                   // [B,H] recurrent gradient leaving the chunk's first step (null: the sweep ends here, not computed)
  float* dc_out;                   // [B,H]
  int has_prev;                    // 1: the stash holds a step before this chunk's first one (its cells are c_{t-1} of step 0)
  // wave-front coupling of two sweeps that run side by side, each launched once (never with reverse); chunks are walked latest first:
  int n_sync;                      // time chunks: chunk k = steps [sync_t[k], sync_t[k+1])
  int sync_t[S2VT_MAX_SYNC + 1];
  unsigned int* signal;            // [n_sync] counters, += 1 per (CTA, batch tile) once its dgates rows of chunk k are visible, or null
  const unsigned int* wait;        // [n_sync] counters advanced by the producer of `dout`: chunk k is read once wait[k] >= wait_val, or null
  unsigned int wait_val;
};

// NTL = batch tiles per cluster.  NTL = 2 runs two independent 16-column sweeps on the same resident weight slice (own epilogue warps,
// barriers, receive buffers, B operand and accumulators per tile): one tile's reduce-scatter is in flight while the other computes, so a
// sweep needs half the SMs and two layers' sweeps fit on the machine side by side.
template <int NTL>
__global__ void __launch_bounds__(32 * (4 * NTL + 1), 1)
lstm_bwd_cluster_kernel(const LstmBwdParams p) {
  constexpr int NB = BWD_NB, CPT = NB / 4, CTRL = 4 * NTL;
  constexpr uint32_t LBO_B = (NB / 8) * 128;         // K-direction stride between core matrices of dgates^T
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t recv_full[NTL][2], b_ready[NTL], mma_done[NTL];
  __shared__ uint32_t tmem_slot;

  const int H = p.H, CS = H / 32, NT = H / 128;      // NT = 128-row output tiles of dh
  const uint32_t RECV_BYTES = (uint32_t)CS * 32u * NB * 2u;        // [src CTA][column half][unit][8 bf16]
  const uint32_t base = (cl::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tl = (NTL > 1 && warp < CTRL) ? (warp >> 2) : 0;      // batch tile served by this epilogue warp
  const int wq = warp & 3;                           // its TMEM lane quarter / column group
  // shared memory: [tile][2 receive buffers], then [tile] B operand
  const uint32_t sRecv0 = base + (uint32_t)tl * 2 * RECV_BYTES, sB_all = base + (uint32_t)NTL * 2 * RECV_BYTES;
  uint8_t* gen = smem_raw + (base - cl::smem_u32(smem_raw));
  const uint2* gRecv0 = reinterpret_cast<const uint2*>(gen + (size_t)tl * 2 * RECV_BYTES);
  uint8_t* gB = gen + (size_t)NTL * 2 * RECV_BYTES + (size_t)tl * (NB * 128 * 2);
  const uint32_t c = cl::cluster_ctarank();
  const int bt = blockIdx.x / CS, nbt = (p.B + NB - 1) / NB;
  const int tile = bt * NTL + tl;
  const int b0 = tile * NB;
  const bool tile_on = b0 < p.B;
  const int T = p.T;
  const bool chain_out = p.dh_out != nullptr;        // the step at local t = 0 still feeds a predecessor (in an earlier-time chunk)
  const uint32_t need_cols = (uint32_t)(H / 2 + NTL * NT * NB);
  const uint32_t tmem_cols = need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u));

  if (warp == CTRL && ptx::elect_one()) {
    for (int i = 0; i < NTL; ++i) {
      ptx::mbar_init(cl::smem_u32(&recv_full[i][0]), 1);
      ptx::mbar_init(cl::smem_u32(&recv_full[i][1]), 1);
      ptx::mbar_init(cl::smem_u32(&b_ready[i]), 1);
      ptx::mbar_init(cl::smem_u32(&mma_done[i]), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(cl::smem_u32(&tmem_slot), tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_acc0 = tmem + (uint32_t)(H / 2);
  const uint32_t tmem_acc = tmem_acc0 + (uint32_t)(tl * NT * NB);
  if (warp < CTRL) {
    // A tile i, TMEM lane m = 32*wq + lane  <->  output unit j = 128 i + m;  K index k = g*32 + u  <->  gate row g*H + 32c + u.
    // W_hh^T[j, g*H + 32c .. +32] is 64 contiguous bytes, so each (tile, gate) is four 16-byte loads = 16 TMEM columns.
    for (int i = tl; i < NT; i += NTL) {
      const __nv_bfloat16* row = p.w_t + (long long)(128 * i + 32 * wq + lane) * 4 * H + 32 * (int)c;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const uint4* src = reinterpret_cast<const uint4*>(row + (long long)(half * 2 + gg) * H);
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const uint4 v = __ldg(src + v4);
            r[gg * 16 + v4 * 4 + 0] = v.x; r[gg * 16 + v4 * 4 + 1] = v.y; r[gg * 16 + v4 * 4 + 2] = v.z; r[gg * 16 + v4 * 4 + 3] = v.w;
          }
        }
        cl::tmem_st_32x32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(i * 64 + half * 32), r);
      }
    }
    cl::tc_wait_st();
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();
  cl::cluster_arrive();
  cl::cluster_wait();

  if (warp == CTRL) {
    // ===================== control thread: one batch of MMAs per step and tile =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NB, 0, 0);
      bool on[NTL];
      for (int i = 0; i < NTL; ++i) on[i] = (bt * NTL + i) * NB < p.B;
      bool ok = true;
      const int t_last = chain_out ? 0 : 1;                    // without a predecessor chunk, step 0 has nobody to feed
      for (int t = T - 1; t >= t_last && ok; --t) {
#pragma unroll
        for (int tlc = 0; tlc < NTL; ++tlc) {
          if (!on[tlc]) continue;
          ptx::mbar_arrive_expect_tx(cl::smem_u32(&recv_full[tlc][t & 1]), RECV_BYTES);   // partials of step t land in recv[t&1]
          ok = ptx::mbar_wait(cl::smem_u32(&b_ready[tlc]), (uint32_t)((T - 1 - t) & 1));
          if (!ok) { atomicExch(&g_sm100_error, 21); break; }
          ptx::tc_fence_after();
          const uint64_t db_base = cl::make_smem_desc(sB_all + (uint32_t)tlc * (NB * 128 * 2), LBO_B, 128, 0);
          for (int i = 0; i < NT; ++i) {
            uint64_t db = db_base;
            uint32_t ta = tmem + (uint32_t)(i * 64);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              cl::mma_bf16_ts(tmem_acc0 + (uint32_t)((tlc * NT + i) * NB), ta, db, idesc, ks != 0 ? 1u : 0u);
              db += (2 * LBO_B) >> 4;
              ta += 8;
            }
          }
          ptx::mma_commit(cl::smem_u32(&mma_done[tlc]));
        }
      }
    }
  } else if (tile_on) {
    // ===================== epilogue warps of tile tl: thread = (unit u = lane, column group q = wq) =====================
    const int u = lane, q = wq;
    const int tid = (int)threadIdx.x - 128 * tl;             // 0..127 within the tile's warps
    const int unit = 32 * (int)c + u;
    float dc[CPT];
    const long long stash_blk = (long long)nbt * CS;
    const long long my_blk0 = (long long)tile * CS + c;
    // prefetched per-step operands
    uint2 gate_raw[CPT];
    float c_t[CPT], c_prev[CPT], dout_v[CPT];
    int rowoff[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      rowoff[j] = min(b0 + q * CPT + j, p.B - 1);
      dc[j] = p.dc_in ? __ldg(p.dc_in + (long long)rowoff[j] * H + unit) : 0.f;
    }
    auto tm = [&](int s) { return p.reverse ? T - 1 - s : s; };
    auto load_step = [&](int s, bool first) {
      const int t = s;                                           // (processing step; `tt` below is its time index)
      const int tt = tm(s);
      const long long blk = (long long)tt * stash_blk + my_blk0;
      const uint2* gsrc = reinterpret_cast<const uint2*>(p.gates + blk * (NB * 32 * 4));
      const float* csrc = p.cells + blk * (NB * 32);
      // (s == 0 with has_prev: the step before this chunk, one stash block row below the pointer; never with reverse)
      const long long tprev = s > 0 ? (long long)tm(s - 1) : (p.has_prev ? -1ll : 0ll);
      const float* cprev_src = p.cells + (tprev * stash_blk + my_blk0) * (NB * 32);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int col = q * CPT + j;
        gate_raw[j] = __ldg(gsrc + col * 32 + u);
        if (first) c_t[j] = __ldg(csrc + col * 32 + u);
        else c_t[j] = c_prev[j];                                 // c_t of step t == c_{t-1} loaded for step t+1
        c_prev[j] = (t > 0 || p.has_prev) ? __ldg(cprev_src + col * 32 + u) : 0.f;
        // (L2: a chunk of dout may be produced while this kernel runs)
        dout_v[j] = (p.dout && tt >= p.dout_t0) ? __ldcg(p.dout + ((long long)tt * p.B + rowoff[j]) * H + unit) : 0.f;
      }
    };
    // wave-front coupling: wk / sk = next chunk to wait for / to signal (walking down from the last one)
    int wk = p.n_sync - 1, sk = p.n_sync - 1;
    bool ok = true;
    auto wait_chunk = [&](int s) {                                 // before the first read of dout at step s
      if (p.wait && wk >= 0 && s == p.sync_t[wk + 1] - 1) {
        if (!ptx::wait_counter_geq(p.wait + wk, p.wait_val)) { atomicExch(&g_sm100_error, 25); ok = false; }
        --wk;
      }
    };
    wait_chunk(T - 1);
    load_step(T - 1, true);
    // reduce-scatter targets: this thread's TMEM lane in tile i holds dh partials of unit (128 i + 32 warp + lane), owned by CTA 4i + warp
    uint32_t dst_recv[4], dst_bar[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t d = (uint32_t)min(4 * i + wq, CS - 1);
      dst_recv[i] = cl::mapa(sRecv0, d) + (uint32_t)(((c * 2 + 0) * 32 + lane) * 16);
      dst_bar[i] = cl::mapa(cl::smem_u32(&recv_full[tl][0]), d);
    }
    const uint32_t bar_stride = cl::smem_u32(&recv_full[0][1]) - cl::smem_u32(&recv_full[0][0]);
    const uint32_t my_b_ready = cl::smem_u32(&b_ready[tl]), my_mma_done = cl::smem_u32(&mma_done[tl]);
    // sum of the CS partial blocks that arrived in receive buffer rb for (unit u, columns 4q .. 4q+3): half (q >> 1) of each source
    // block, 8-byte pair (q & 1) of the unit's 16-byte chunk -- the sender swapped the two pairs for odd (u >> 3), so a half-warp's
    // 64-bit reads hit 32 different banks
    auto sum_partials = [&](int rb, float (&dh)[CPT]) {
      const uint2* rsrc = gRecv0 + (size_t)rb * (RECV_BYTES / 8) + ((q >> 1) * 32 + u) * 2 + ((q ^ (u >> 3)) & 1);
      float acc2[CPT] = {0.f, 0.f, 0.f, 0.f};                  // two independent chains: the 16 loads pipeline instead of serialising
#pragma unroll 4
      for (int s = 0; s < CS; s += 2) {
        const uint2 v = rsrc[s * 128], w = rsrc[(s + 1) * 128];
        dh[0] += __uint_as_float(v.x << 16); dh[1] += __uint_as_float(v.x & 0xffff0000u);
        dh[2] += __uint_as_float(v.y << 16); dh[3] += __uint_as_float(v.y & 0xffff0000u);
        acc2[0] += __uint_as_float(w.x << 16); acc2[1] += __uint_as_float(w.x & 0xffff0000u);
        acc2[2] += __uint_as_float(w.y << 16); acc2[3] += __uint_as_float(w.y & 0xffff0000u);
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) dh[j] += acc2[j];
    };
    uint32_t rph[2] = {0, 0};
    for (int t = T - 1; t >= 0; --t) {
      // ---- A. recurrent gradient: sum the CS partial blocks that arrived for this CTA's units
      float dh[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) dh[j] = dout_v[j];
      if (t < T - 1) {
        const int rb = (t + 1) & 1;
        ok = ok && ptx::mbar_wait(cl::smem_u32(&recv_full[tl][rb]), rph[rb]);
        rph[rb] ^= 1;
        if (!ok) { atomicExch(&g_sm100_error, 22); break; }
        sum_partials(rb, dh);
      } else if (p.dh_in) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) dh[j] += __ldg(p.dh_in + (long long)rowoff[j] * H + unit);
      }
      // ---- B. fused gate gradients
      float dgv[CPT][4];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const __nv_bfloat162 g01 = *reinterpret_cast<const __nv_bfloat162*>(&gate_raw[j].x);
        const __nv_bfloat162 g23 = *reinterpret_cast<const __nv_bfloat162*>(&gate_raw[j].y);
        const float gi = __low2float(g01), gf = __high2float(g01), gg = __low2float(g23), go = __high2float(g23);
        const float tc = cl::fast_tanh(c_t[j]);
        const float d_o = dh[j] * tc;
        const float dcv = dc[j] + dh[j] * go * (1.f - tc * tc);
        dgv[j][0] = dcv * gg * gi * (1.f - gi);
        dgv[j][1] = dcv * c_prev[j] * gf * (1.f - gf);
        dgv[j][2] = dcv * gi * (1.f - gg * gg);
        dgv[j][3] = d_o * go * (1.f - go);
        dc[j] = dcv * gf;
      }
      // ---- C. bf16 dgates into the B operand: element (b = col, k = g*32 + u)
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int col = q * CPT + j;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<__nv_bfloat16*>(gB + ((g * 4 + (u >> 3)) * (NB / 8) + col / 8) * 128 + (col % 8) * 16 + (u & 7) * 2) =
              __float2bfloat16(dgv[j][g]);
      }
      ptx::fence_proxy_async();
      if (t > 0) {
        wait_chunk(t - 1);
        load_step(t - 1, false);                                // prefetch the next step's stash / dout
      }
      cl::named_bar_sync(1 + tl, 128);                          // every thread's operand writes are fenced: ONE arrival releases the MMAs
      if ((t > 0 || chain_out) && tid == 0) ptx::mbar_arrive(my_b_ready);
      // ---- D. dgates_t for HBM (GEMM layout): read the 256 chunks of 8 units x 1 column now, store them after the scatter
      uint4 dgv4[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int x = tid + 128 * r;
        dgv4[r] = *reinterpret_cast<const uint4*>(gB + ((x >> 4) * (NB / 8) + (x & 15) / 8) * 128 + ((x & 15) % 8) * 16);
      }
      // ---- E. scatter this CTA's partial dh to the owners of each unit
      if (t > 0 || chain_out) {
        ok = ok && ptx::mbar_wait(my_mma_done, (uint32_t)((T - 1 - t) & 1));
        if (!ok) { atomicExch(&g_sm100_error, 23); break; }
        ptx::tc_fence_after();
        const uint32_t boff = (uint32_t)(t & 1);
        const bool swap_pairs = ((lane >> 3) & 1) != 0;
        auto scatter = [&](int i, const uint32_t (&r)[16]) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            pk[j] = *reinterpret_cast<const uint32_t*>(&t2);
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t a0 = pk[4 * hh], a1 = pk[4 * hh + 1], a2 = pk[4 * hh + 2], a3 = pk[4 * hh + 3];
            cl::st_async_16(dst_recv[i] + boff * RECV_BYTES + (uint32_t)(hh * 32 * 16), swap_pairs ? a2 : a0, swap_pairs ? a3 : a1,
                            swap_pairs ? a0 : a2, swap_pairs ? a1 : a3, dst_bar[i] + boff * bar_stride);
          }
        };
        const uint32_t t_lane = tmem_acc + ((uint32_t)(wq * 32) << 16);
        if (NT == 4) {                                           // all four tiles in flight behind one wait
          uint32_t r0[16], r1[16], r2[16], r3[16];
          ptx::tmem_ld_32x16(t_lane, r0);
          ptx::tmem_ld_32x16(t_lane + NB, r1);
          ptx::tmem_ld_32x16(t_lane + 2 * NB, r2);
          ptx::tmem_ld_32x16(t_lane + 3 * NB, r3);
          ptx::tc_wait_ld();
          scatter(0, r0); scatter(1, r1); scatter(2, r2); scatter(3, r3);
        } else {
          for (int i = 0; i < NT; ++i) {
            uint32_t r[16];
            ptx::tmem_ld_32x16(t_lane + (uint32_t)(i * NB), r);
            ptx::tc_wait_ld();
            scatter(i, r);
          }
        }
        ptx::tc_fence_before();
        cl::named_bar_sync(1 + tl, 128);                         // B operand / accumulators are free for the next step
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int x = tid + 128 * r;
        const int kblk = x >> 4, b = x & 15;
        if (b0 + b < p.B)
          *reinterpret_cast<uint4*>(p.dgates + ((long long)tm(t) * p.B + b0 + b) * 4 * H + (kblk >> 2) * H + 32 * (int)c + 8 * (kblk & 3)) = dgv4[r];
      }
      if (p.signal && sk >= 0 && t == p.sync_t[sk]) {            // chunk complete: publish it to the consumer of dgates
        __threadfence();
        cl::named_bar_sync(1 + tl, 128);
        if (tid == 0) ptx::red_release_gpu_add(p.signal + sk, 1u);
        --sk;
      }
    }
    if (p.signal && tid == 0)                                    // (error exit: never leave a consumer waiting)
      for (; sk >= 0; --sk) ptx::red_release_gpu_add(p.signal + sk, 1u);
    if (chain_out && ok) {
      // state for the chunk before this one: the recurrent gradient scattered by local step 0 (in receive buffer 0) and dc
      ok = ptx::mbar_wait(cl::smem_u32(&recv_full[tl][0]), rph[0]);
      if (!ok) atomicExch(&g_sm100_error, 24);
      float dh[CPT] = {0.f, 0.f, 0.f, 0.f};
      if (ok) sum_partials(0, dh);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        if (b0 + q * CPT + j < p.B) {
          p.dh_out[(long long)(b0 + q * CPT + j) * H + unit] = dh[j];
          p.dc_out[(long long)(b0 + q * CPT + j) * H + unit] = dc[j];
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  cl::cluster_arrive();
  cl::cluster_wait();
  if (warp == 0) ptx::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace s2vt

using namespace s2vt;

template <int NTL>
static int launch_lstm_bwd(cudaStream_t st, const LstmBwdParams& p) {
  const int H = p.H, CS = H / 32;
  const size_t smem_need = 1024 + (size_t)NTL * (2 * (size_t)CS * 32 * BWD_NB * 2 + (size_t)BWD_NB * 128 * 2);
  // keep GEMM CTAs of other streams (97 KB each) off the SMs of the cluster: this CTA owns the SM's tensor memory
  const size_t smem = smem_need < (size_t)136 * 1024 ? (size_t)136 * 1024 : smem_need;
  auto kern = lstm_bwd_cluster_kernel<NTL>;
  S2VT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) S2VT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CS * ceil_div(p.B, BWD_NB * NTL));
  cfg.blockDim = dim3(32 * (4 * NTL + 1));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  S2VT_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  S2VT_REQUIRE(max_clusters >= 1, "s2vt_lstm_bwd_bf16: a cluster of %d CTAs with %zu B of shared memory cannot be scheduled on this device", CS, smem);
  S2VT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  return 0;
}

extern "C" int s2vt_lstm_bwd_bf16(void* stream, int T, int B, int H, int dout_t0,
                                  const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                                  void* dgates_bf16) {
  return s2vt_lstm_bwd_bf16_dir(stream, T, B, H, dout_t0, dout, gates_bf16, cells, w_hh_t_bf16, dgates_bf16, 0);
}

extern "C" int s2vt_lstm_bwd_bf16_dir(void* stream, int T, int B, int H, int dout_t0,
                                      const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                                      void* dgates_bf16, int reverse) {
  return s2vt_lstm_bwd_bf16_chunk(stream, T, B, H, dout_t0, dout, gates_bf16, cells, w_hh_t_bf16, dgates_bf16, reverse,
                                  nullptr, nullptr, nullptr, nullptr, 0, 1);
}

extern "C" int s2vt_lstm_bwd_bf16_chunk(void* stream, int T, int B, int H, int dout_t0,
                                        const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                                        void* dgates_bf16, int reverse, const float* dh_in, const float* dc_in, float* dh_out,
                                        float* dc_out, int has_prev, int tiles_per_cluster) {
  return s2vt_lstm_bwd_bf16_sync(stream, T, B, H, dout_t0, dout, gates_bf16, cells, w_hh_t_bf16, dgates_bf16, reverse, dh_in, dc_in, dh_out,
                                 dc_out, has_prev, tiles_per_cluster, 0, nullptr, nullptr, nullptr, 0);
}

extern "C" int s2vt_lstm_bwd_bf16_sync(void* stream, int T, int B, int H, int dout_t0,
                                       const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                                       void* dgates_bf16, int reverse, const float* dh_in, const float* dc_in, float* dh_out,
                                       float* dc_out, int has_prev, int tiles_per_cluster, int n_sync, const int* sync_t,
                                       unsigned int* signal, const unsigned int* wait, unsigned int wait_val) {
  S2VT_REQUIRE(T >= 1 && B >= 1, "s2vt_lstm_bwd_bf16: bad dims");
  S2VT_REQUIRE(n_sync >= 0 && n_sync <= S2VT_MAX_SYNC, "s2vt_lstm_bwd_bf16_sync: at most %d chunks", S2VT_MAX_SYNC);
  S2VT_REQUIRE(n_sync == 0 || (sync_t && (signal || wait) && !reverse), "s2vt_lstm_bwd_bf16_sync: chunks need sync_t and a counter array, and the forward direction");
  S2VT_REQUIRE(H % 128 == 0 && H >= 128 && H <= 512, "s2vt_lstm_bwd_bf16: the cluster-resident kernel needs H %% 128 == 0 and 128 <= H <= 512 (got %d)", H);
  S2VT_REQUIRE(gates_bf16 && cells && w_hh_t_bf16 && dgates_bf16, "s2vt_lstm_bwd_bf16: null pointer");
  S2VT_REQUIRE(aligned16(gates_bf16) && aligned16(w_hh_t_bf16) && aligned16(dgates_bf16), "s2vt_lstm_bwd_bf16: buffers must be 16-byte aligned");
  S2VT_REQUIRE((dh_out == nullptr) == (dc_out == nullptr), "s2vt_lstm_bwd_bf16_chunk: dh_out and dc_out come together");
  S2VT_REQUIRE(!(dh_out && !has_prev), "s2vt_lstm_bwd_bf16_chunk: dh_out asks for the gradient into a step before the chunk, has_prev says there is none");
  S2VT_REQUIRE(!(reverse && (dh_in || dc_in || dh_out || has_prev)), "s2vt_lstm_bwd_bf16_chunk: chunk chaining is not available for the reverse direction");
  S2VT_REQUIRE(tiles_per_cluster == 1 || tiles_per_cluster == 2, "s2vt_lstm_bwd_bf16_chunk: tiles_per_cluster must be 1 or 2");
  LstmBwdParams p{};
  p.T = T; p.B = B; p.H = H; p.dout_t0 = dout_t0 < 0 ? 0 : dout_t0;
  p.dout = dout; p.gates = (const __nv_bfloat16*)gates_bf16; p.cells = cells; p.w_t = (const __nv_bfloat16*)w_hh_t_bf16;
  p.dgates = (__nv_bfloat16*)dgates_bf16;
  p.reverse = reverse ? 1 : 0;
  p.dh_in = dh_in; p.dc_in = dc_in; p.dh_out = dh_out; p.dc_out = dc_out; p.has_prev = has_prev ? 1 : 0;
  p.n_sync = n_sync; p.signal = n_sync ? signal : nullptr; p.wait = n_sync ? wait : nullptr; p.wait_val = wait_val;
  for (int i = 0; i <= n_sync && n_sync > 0; ++i) {
    S2VT_REQUIRE(sync_t[i] >= 0 && sync_t[i] <= T && (i == 0 || sync_t[i] > sync_t[i - 1]), "s2vt_lstm_bwd_bf16_sync: sync_t must increase from 0 to T");
    p.sync_t[i] = sync_t[i];
  }
  S2VT_REQUIRE(n_sync == 0 || (p.sync_t[0] == 0 && p.sync_t[n_sync] == T), "s2vt_lstm_bwd_bf16_sync: sync_t must start at 0 and end at T");
  if (tiles_per_cluster == 2) return launch_lstm_bwd<2>((cudaStream_t)stream, p);
  return launch_lstm_bwd<1>((cudaStream_t)stream, p);
}
