// Exact-grade decode on the tensor cores ("x" path): greedy (S2VTModel.py:82-110) and beam search
// (S2VTModel.py:149-240) with every product on tcgen05, fp32-grade accuracy, no [B,V] logits in memory.
//
// Arithmetic.  An fp32 operand x is scaled by a power of two s (per tensor, max|x*s| in [2^14, 2^15)) and
// split into two fp16 planes hi = fp16(x*s), lo = fp16(x*s - hi): hi + lo == x*s up to 2^-23 |x*s| (an
// element below 2^-3 has a subnormal residual: absolute error <= 2^-25, i.e. 2^-39 of the tensor's
// maximum).  A product is three tcgen05.mma.kind::f16 passes per K block,
//     D_corr += A_hi*B_lo + A_lo*B_hi        D_main[kb % 3] += A_hi*B_hi        (A_lo*B_lo, 2^-22 relative, is dropped)
// and the epilogue adds the four fp32 TMEM accumulators in registers and multiplies by the exact power of
// two 1/(sa*sb).  Every fp16 x fp16 product is exact in fp32; what limits the accuracy is the tensor core's
// accumulate step, which truncates (measured: one accumulator for everything loses ~0.5 ulp of the running
// sum per MMA, 2.5x the error of a sequential fp32 FMA chain at K = 512).  Hence the split: the corrections,
// 2^-11 of the result, truncate at 2^-11 of that, and each main accumulator takes a third of the K blocks.
// tests/test_gpu_xdec.py measures the result against an fp64 product next to the FFMA kernel it replaces.
//
// Kernel.  One 128 x BN output tile per CTA (BN = 128 or 64), TMA (3-D maps: k, row, plane; 128B swizzle)
// into a 3/4-stage ring, one elected thread issuing the MMAs, four epilogue warps draining TMEM with
// tcgen05.ld.  Epilogues:
//   STORE   C = acc/(sa sb) + bias (+ C)                      time-batched products (feat_linear, input gates)
//   LSTM    gates -> (c, h); h leaves as fp16 planes          one recurrence step of vid_rnn / word_rnn; the
//           (the A operand of the next step); the input-side  gate rows are interleaved (4u+g) so a thread
//           pre-activation may add a gathered row of          owns all four gates of its units; the previous
//           EW = embedding . W_ih[:, :E]^T, indexed by the    step's argmax is resolved from per-tile partials
//           previous token                                     while the MMAs run
//   ARGMAX  per (row, tile) best logit + index                greedy: out_linear + argmax, logits never stored
//   BEAM    per (row, tile) max, sum exp, top-8               beam: out_linear + log_softmax + top-k partials
// The step kernels are enqueued back to back by the C entry points below (two streams: vid_rnn leads,
// word_rnn + decode trail by a chunk of steps); nothing returns to the host between steps.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "sm100_err.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define X_TRY(expr) do { rc = (expr); if (rc) return rc; } while (0)

namespace s2vt {

int make_tmap_planes(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t ld, uint64_t plane_stride,
                     uint32_t box_rows);

namespace xd {

constexpr int BM = 128, BK = 64;
constexpr int PLANE_A = BM * BK * 2;            // one fp16 plane of the A tile: 16 KB
constexpr int KC = 8;                           // candidates kept per (row, tile) by the beam epilogue
enum { EPI_STORE = 0, EPI_LSTM = 1, EPI_ARGMAX = 2, EPI_BEAM = 3 };

struct XParams {
  int M, N, K, num_kb;
  const float* a_inv;        // device scalars: 1/scale of each operand (powers of two)
  const float* b_inv;
  const float* a_inv_row;    // optional [M]: one scale per row of A instead (the feature matrix: scaled and split in ONE pass)
  const float* bias;         // [N] (LSTM: used when pre == nullptr)
  // STORE
  float* C; RowMap cm; int accumulate; int c_vec;
  // LSTM (N = 4*HP, column 4u+g)
  int HP;
  const float* pre; long long pre_ld;
  const float* gtab; long long gtab_ld; const int* gidx;       // optional: + gtab[gidx[m]] (the embedding half of word_rnn's input)
  const float* c_in; float* c_out;
  float* h_f32; long long h_ld;
  __half* hp; long long hp_ld; long long hp_plane;             // h planes (scale 2^15)
  // BEAM partial outputs (two per tile: one per epilogue warp group)
  float* o_val; int* o_idx; float* o_ms;
  unsigned int* row_thr; int kc;                               // BEAM: [M][KC] per-row slot maxima (order_f32 bits, 0 = none yet)
  int m_fastest;                                               // grid is (row tiles, column tiles) instead of (column tiles, row tiles)
  unsigned long long* o_key;                                   // ARGMAX: per-row packed (value, index) maximum, zeroed by the consumer
  // LSTM: the previous step's argmax keys -> token (gidx == nullptr); n-tile 0 records it and clears the keys of the next step
  const unsigned long long* key_in; unsigned long long* key_clear; int64_t* tok_out; long long tok_ld;
  unsigned long long* trace;   // debug: 8 %globaltimer stamps of one CTA (tools/trace_xdec.py), or nullptr
  int trace_bx, trace_by;      // which CTA (default (0,0); S2VT_XDEC_TRACE_CTA=x,y clamps to the grid)
};

// Programmatic dependent launch: a step kernel is launched while its predecessor in the stream still runs; everything that
// reads the predecessor's output sits behind pdl_wait(), everything before it (barrier / TMEM set-up, weight tiles) overlaps.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// kind::f16 instruction descriptor, fp16 x fp16 -> fp32: D format F32 (bit 4), A/B format F16 (0), K-major operands
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// order-preserving map float -> uint32 (0 is below every float) and back
__device__ __forceinline__ uint32_t order_f32(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long pack_key(float v, int idx) {
  const uint32_t b = __float_as_uint(v);
  const uint32_t u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}
__device__ __forceinline__ int key_index(unsigned long long k) { return (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)); }

// r[j] = bits of (main0 + main1 + main2) + corr for 32 columns of this warp's 32 rows (fp32 adds, round to nearest)
template <int BN>
__device__ __forceinline__ void load_acc(uint32_t taddr, int nmain, uint32_t (&r)[32]) {
  uint32_t q[32];
  ptx::tmem_ld_32x32(taddr + BN, r);
  ptx::tmem_ld_32x32(taddr, q);
  ptx::tc_wait_ld();
  if (nmain > 1) {
    uint32_t m1[32];
    ptx::tmem_ld_32x32(taddr + 2 * BN, m1);
    ptx::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(m1[j]));
    if (nmain > 2) {
      ptx::tmem_ld_32x32(taddr + 3 * BN, m1);
      ptx::tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(m1[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
}

// MINB = 2: two CTAs share an SM (64-column tiles, a 2-stage ring of 96 KB and 256 TMEM columns each), so that one CTA's epilogue
// runs under the other's tile loads and MMAs.  Measured on the beam vocab product: slower than one CTA with 128-column tiles
// (225 vs 191 us per depth: the epilogue's cost is per row, not per column), so no caller uses it at present.
template <int BN, int EPI, int MINB>
__global__ void __launch_bounds__(256, MINB)
xgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const XParams p) {
  constexpr int STAGES = (MINB == 2) ? 2 : ((BN == 128) ? 3 : 4);
  constexpr int PLANE_B = BN * BK * 2;
  constexpr int STAGE_BYTES = 2 * PLANE_A + 2 * PLANE_B;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // row-tile-fastest grids (vocab products of the beam step): CTAs that run together cover all row tiles of a few column tiles, so
  // a row's later column tiles see the candidate threshold its earlier ones published
  const int tile_m = p.m_fastest ? blockIdx.x : blockIdx.y, tile_n = p.m_fastest ? blockIdx.y : blockIdx.x;
  const int n_tiles = p.m_fastest ? gridDim.y : gridDim.x;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  unsigned long long* const trace = (p.trace && blockIdx.x == min(p.trace_bx, (int)gridDim.x - 1) && blockIdx.y == min(p.trace_by, (int)gridDim.y - 1)) ? p.trace : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = ptx::globaltimer_ns();

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
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), 4 * BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  pdl_launch_dependents();

  if (warp_idx == 0) {
    // ===================== TMA producer: both planes of a tile arrive with one 3-D box each =====================
    if (ptx::elect_one()) {
      // the B operand is always a weight: its first tiles are requested before the previous kernel's results are awaited
      const int npre = p.num_kb < STAGES ? p.num_kb : STAGES;
      for (int kb = 0; kb < npre; ++kb) {
        const uint32_t fb = ptx::smem_u32(&full_bar[kb]);
        ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        tma_load_3d(smem_base + kb * STAGE_BYTES + 2 * PLANE_A, &tmB, fb, kb * BK, n0, 0);
      }
      if (trace) trace[1] = ptx::globaltimer_ns();
      pdl_wait();
      if (trace) trace[2] = ptx::globaltimer_ns();
      for (int kb = 0; kb < npre; ++kb)
        tma_load_3d(smem_base + kb * STAGE_BYTES, &tmA, ptx::smem_u32(&full_bar[kb]), kb * BK, m0, 0);
      for (int kb = npre; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1)) { atomicExch(&g_sm100_error, 11); break; }
        const uint32_t fb = ptx::smem_u32(&full_bar[s]);
        const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + 2 * PLANE_A;
        ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        tma_load_3d(sA, &tmA, fb, kb * BK, m0, 0);
        tma_load_3d(sB, &tmB, fb, kb * BK, n0, 0);
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph)) { atomicExch(&g_sm100_error, 12); break; }
        ptx::tc_fence_after();
        if (trace && kb == 0) trace[3] = ptx::globaltimer_ns();
        if (trace && kb == p.num_kb - 1) trace[4] = ptx::globaltimer_ns();
        const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + 2 * PLANE_A;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t a_hi = ptx::make_smem_desc_sw128(sA + k * 32, 16, 1024);
          const uint64_t a_lo = ptx::make_smem_desc_sw128(sA + PLANE_A + k * 32, 16, 1024);
          const uint64_t b_hi = ptx::make_smem_desc_sw128(sB + k * 32, 16, 1024);
          const uint64_t b_lo = ptx::make_smem_desc_sw128(sB + PLANE_B + k * 32, 16, 1024);
          // kind::f16; the idesc selects fp16 operands.  TMEM columns: [0,BN) corrections, [BN,4BN) three main accumulators
          ptx::mma_bf16_ss(tmem, a_hi, b_lo, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          ptx::mma_bf16_ss(tmem, a_lo, b_hi, idesc, 1u);
          ptx::mma_bf16_ss(tmem + (uint32_t)((1 + kb % 3) * BN), a_hi, b_hi, idesc, (kb >= 3 || k > 0) ? 1u : 0u);
        }
        ptx::mma_commit(ptx::smem_u32(&empty_bar[s]));
      }
      ptx::mma_commit(ptx::smem_u32(&tmem_full_bar));
    }
    __syncwarp();
  }

  // ===================== epilogue: all eight warps =====================
  // A warp may read the TMEM lanes of its quarter (warp_idx % 4) only: warps q and q+4 share 32 rows and split the tile's
  // 32-column chunks between them (`half` 0 takes the even chunks, 1 the odd ones).
  {
    const int q = warp_idx & 3, half = warp_idx >> 2;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const float sc = (p.a_inv_row ? (row_ok ? __ldg(p.a_inv_row + m) : 1.f) : (p.a_inv ? __ldg(p.a_inv) : 1.f)) * (p.b_inv ? __ldg(p.b_inv) : 1.f);
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int nmain = p.num_kb < 3 ? p.num_kb : 3;
    pdl_wait();

    // ---- LSTM: everything the gates need besides the accumulators is fetched while the MMAs run
    float pin[32], cp[8];
    int tok = 0;
    auto lstm_inputs = [&](int n) {
      const float4* src = reinterpret_cast<const float4*>(p.pre ? p.pre + (long long)m * p.pre_ld + n : p.bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float4 t = __ldg(src + j); pin[4 * j] = t.x; pin[4 * j + 1] = t.y; pin[4 * j + 2] = t.z; pin[4 * j + 3] = t.w; }
      if (p.gtab) {
        const float4* g = reinterpret_cast<const float4*>(p.gtab + (long long)tok * p.gtab_ld + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float4 t = __ldg(g + j); pin[4 * j] += t.x; pin[4 * j + 1] += t.y; pin[4 * j + 2] += t.z; pin[4 * j + 3] += t.w; }
      }
      if (p.c_in) {
        const float4* c4 = reinterpret_cast<const float4*>(p.c_in + (long long)m * p.HP + (n >> 2));
        const float4 t0 = c4[0], t1 = c4[1];
        cp[0] = t0.x; cp[1] = t0.y; cp[2] = t0.z; cp[3] = t0.w; cp[4] = t1.x; cp[5] = t1.y; cp[6] = t1.z; cp[7] = t1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) cp[j] = 0.f;
      }
    };
    if (EPI == EPI_LSTM) {
      if (p.gtab && row_ok) {
        if (p.gidx) tok = p.gidx[m];
        else {
          tok = key_index(__ldcg(p.key_in + m));
          if (tile_n == 0 && half == 0) {
            if (p.tok_out) p.tok_out[(long long)m * p.tok_ld] = tok;
            if (p.key_clear) p.key_clear[m] = 0ull;
          }
        }
      }
      if (row_ok && n0 + half * 32 < p.N) lstm_inputs(n0 + half * 32);
    }

    bool ok = ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar), 0);
    if (!ok) atomicExch(&g_sm100_error, 13);
    ptx::tc_fence_after();
    if (trace && threadIdx.x == 128) trace[5] = ptx::globaltimer_ns();

    if (EPI == EPI_STORE) {
      const long long crow = row_ok ? p.cm(m) : 0;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;
        uint32_t r[32];
        load_acc<BN>(trow + (uint32_t)(c * 32), nmain, r);
        if (!row_ok) continue;
        float v[32];
        const bool full = (n + 31 < p.N);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * sc;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (full || n + j < p.N) v[j] += __ldg(p.bias + n + j);
        }
        float* dst = p.C + crow + n;
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
      }
    } else if (EPI == EPI_LSTM) {
      // columns n..n+31 = units u0..u0+7, gates (i,f,g,o) adjacent; N = 4*HP is a multiple of 32
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;
        uint32_t r[32];
        load_acc<BN>(trow + (uint32_t)(c * 32), nmain, r);
        if (!row_ok) continue;
        if (c != half) lstm_inputs(n);
        const int u0 = n >> 2;
        float cn[8], h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xi = __uint_as_float(r[4 * j]) * sc + pin[4 * j], xf = __uint_as_float(r[4 * j + 1]) * sc + pin[4 * j + 1];
          const float xg = __uint_as_float(r[4 * j + 2]) * sc + pin[4 * j + 2], xo = __uint_as_float(r[4 * j + 3]) * sc + pin[4 * j + 3];
          const float ig = sigmoidf_exact(xi), fg = sigmoidf_exact(xf), gg = tanhf(xg), og = sigmoidf_exact(xo);
          cn[j] = fg * cp[j] + ig * gg;
          h[j] = og * tanhf(cn[j]);
        }
        {
          float4* dst = reinterpret_cast<float4*>(p.c_out + (long long)m * p.HP + u0);
          dst[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
          dst[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
        }
        if (p.h_f32) {
          float4* dst = reinterpret_cast<float4*>(p.h_f32 + (long long)m * p.h_ld + u0);
          dst[0] = make_float4(h[0], h[1], h[2], h[3]);
          dst[1] = make_float4(h[4], h[5], h[6], h[7]);
        }
        if (p.hp) {
          __align__(16) __half hi[8];
          __align__(16) __half lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xs = h[j] * 32768.f;
            hi[j] = __float2half_rn(xs);
            lo[j] = __float2half_rn(xs - __half2float(hi[j]));
          }
          __half* dst = p.hp + (long long)m * p.hp_ld + u0;
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
          *reinterpret_cast<uint4*>(dst + p.hp_plane) = *reinterpret_cast<const uint4*>(lo);
        }
      }
    } else if (EPI == EPI_ARGMAX) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;
        uint32_t r[32];
        load_acc<BN>(trow + (uint32_t)(c * 32), nmain, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (n + j < p.N) {
            const float v = __uint_as_float(r[j]) * sc + __ldg(p.bias + n + j);
            if (v > bv || bi == 0x7fffffff) { bv = v; bi = n + j; }
          }
        }
      }
      // one 64-bit atomicMax per row and warp: key = (order-preserving bits of the value) : (~index), so the maximum is the
      // highest value and, among equal values, the lowest index (torch.argmax)
      if (row_ok && bi != 0x7fffffff) atomicMax(p.o_key + m, pack_key(bv, bi));
    } else {   // EPI_BEAM
      // Per (row, tile half): running max / sum exp of the logits and the best KC candidates.  Every tile half publishes its maximum
      // into one of kc per-row slots (slot = part % kc, atomicMax): the slots hold kc DIFFERENT elements of the row, so their minimum is
      // a lower bound of the row's kc-th best logit, and a candidate below it is skipped (>=: an equal value may still win on its
      // index).  With the row-tile-fastest grid every wave after the first sees a bound from thousands of columns: few insertions.
      float mx = -INFINITY, sum = 0.f;
      float tv[KC];
      int ti[KC];
#pragma unroll
      for (int qq = 0; qq < KC; ++qq) { tv[qq] = -INFINITY; ti[qq] = 0x7fffffff; }
      float thr0 = -INFINITY;
      if (row_ok && p.row_thr) {
        uint32_t lo_bits = 0xffffffffu;
        for (int g = 0; g < p.kc; ++g) lo_bits = min(lo_bits, __ldcg(p.row_thr + (long long)m * KC + g));
        if (lo_bits != 0u) thr0 = unorder_f32(lo_bits);      // 0 = a slot nobody has written yet: no bound
      }
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;
        uint32_t r[32];
        load_acc<BN>(trow + (uint32_t)(c * 32), nmain, r);
        float v[32];
        float cmx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (n + j < p.N) ? __uint_as_float(r[j]) * sc + __ldg(p.bias + n + j) : -INFINITY;
          cmx = fmaxf(cmx, v[j]);
        }
        const float nmx = fmaxf(mx, cmx);
        float s = sum * expf(mx - nmx);                 // mx = -inf on the first chunk: sum = 0 stays 0
#pragma unroll
        for (int j = 0; j < 32; ++j) s += expf(v[j] - nmx);   // masked columns: exp(-inf) = 0
        sum = s; mx = nmx;
        if (cmx >= thr0 && cmx > tv[KC - 1]) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (v[j] > tv[KC - 1] && v[j] >= thr0) {      // strict vs the list: an equal value keeps the earlier (lower) index ahead
              tv[KC - 1] = v[j]; ti[KC - 1] = n + j;
#pragma unroll
              for (int qq = KC - 1; qq > 0; --qq) {
                if (tv[qq] > tv[qq - 1]) {
                  const float fv = tv[qq]; tv[qq] = tv[qq - 1]; tv[qq - 1] = fv;
                  const int iv = ti[qq]; ti[qq] = ti[qq - 1]; ti[qq - 1] = iv;
                }
              }
            }
          }
        }
      }
      if (row_ok) {                                     // partials are [M][part], part = 2*tile + half (a half may be empty: -inf / 0)
        const long long o = (long long)m * (2 * n_tiles) + 2 * tile_n + half;
        p.o_ms[2 * o] = mx; p.o_ms[2 * o + 1] = sum;
#pragma unroll
        for (int qq = 0; qq < KC; ++qq) { p.o_val[o * KC + qq] = tv[qq]; p.o_idx[o * KC + qq] = ti[qq]; }
        if (p.row_thr && mx > thr0) atomicMax(p.row_thr + (long long)m * KC + (2 * tile_n + half) % p.kc, order_f32(mx));
      }
    }
  }
  if (trace && threadIdx.x == 128) trace[6] = ptx::globaltimer_ns();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) ptx::tmem_dealloc(tmem, 4 * BN);
}

// ------------------------------------------------------------------ fp32 -> fp16 planes
// max |x| as float bits (non-negative floats order like unsigned integers)
__global__ void absmax_kernel(const float* __restrict__ x, long long rows, int cols, long long ld, unsigned int* __restrict__ bits) {
  float mx = 0.f;
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long rr = i / cols;
    const int cc = (int)(i - rr * cols);
    mx = fmaxf(mx, fabsf(x[rr * ld + cc]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(bits, __float_as_uint(mx));
}

// power-of-two scale that puts max|x| into [2^14, 2^15)
__device__ __forceinline__ float scale_from_bits(unsigned int bits) {
  int e = (int)((bits >> 23) & 0xff) - 127;                  // floor(log2(max)) for normal numbers
  if ((bits & 0x7f800000u) == 0u || (bits & 0x7f800000u) == 0x7f800000u) return 1.f;   // zero / denormal / inf / nan
  int s = 14 - e;
  s = s > 100 ? 100 : (s < -100 ? -100 : s);
  return __uint_as_float((unsigned int)(s + 127) << 23);
}

// dst row r' <- src row r:  il_H == 0: r' = r;  il_H > 0 (gate interleave): src row g*il_H + u -> dst row 4u + g.
// Padding (rows / columns beyond the source) is zeroed by the caller beforehand.
__global__ void split_kernel(const float* __restrict__ x, long long rows, int cols, long long ld, int il_H,
                             const unsigned int* __restrict__ bits, float fixed_scale,
                             __half* __restrict__ planes, long long out_ld, long long plane_stride, float* __restrict__ inv_out) {
  const float s = bits ? scale_from_bits(*bits) : fixed_scale;
  if (inv_out && blockIdx.x == 0 && threadIdx.x == 0) *inv_out = 1.f / s;      // exact: s is a power of two
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long rr = i / cols;
    const int cc = (int)(i - rr * cols);
    long long ro = rr;
    if (il_H > 0) { const long long g = rr / il_H, u = rr - g * il_H; ro = 4 * u + g; }
    const float xs = x[rr * ld + cc] * s;
    const __half hi = __float2half_rn(xs);
    const __half lo = __float2half_rn(xs - __half2float(hi));
    planes[ro * out_ld + cc] = hi;
    planes[plane_stride + ro * out_ld + cc] = lo;
  }
}

// One block per row: the row's own power-of-two scale (max |x| into [2^14, 2^15)), both planes and 1/scale -- a single pass over
// the matrix (the per-tensor form needs the maximum first: two passes over 670 MB of features per 512 videos).
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, int cols, long long ld, __half* __restrict__ planes,
                                                         long long out_ld, long long plane_stride, float* __restrict__ inv_row) {
  __shared__ float sh[8];
  const long long r = blockIdx.x;
  const float* src = x + r * ld;
  float mx = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, fabsf(src[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = sh[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sh[w]);
  const float s = scale_from_bits(__float_as_uint(mx));
  if (threadIdx.x == 0) inv_row[r] = 1.f / s;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {          // (second read of the row: L1 / L2 hit)
    const float xs = src[c] * s;
    const __half hi = __float2half_rn(xs);
    planes[r * out_ld + c] = hi;
    planes[plane_stride + r * out_ld + c] = __float2half_rn(xs - __half2float(hi));
  }
}

// dst[4u+g] = a[g*H+u] + b[g*H+u] (b may be null); dst is zero-padded by the caller
__global__ void bias_interleave_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int il, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = il ? 4 * H : H;
  if (i >= n) return;
  const float v = a[i] + (b ? b[i] : 0.f);
  if (il) { const int g = i / H, u = i - g * H; dst[4 * u + g] = v; } else dst[i] = v;
}

__global__ void fill_i32_kernel(int* __restrict__ p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// the last greedy step's token (every earlier one is resolved by the following step's LSTM epilogue)
__global__ void keys_to_tokens_kernel(int M, const unsigned long long* __restrict__ key, int64_t* __restrict__ tok_out, long long tok_ld) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) tok_out[(long long)m * tok_ld] = key_index(key[m]);
}

// ------------------------------------------------------------------ beam bookkeeping
struct BeamMeta { float* key; int* tok; int* len; int* fin; int* hist; };

__global__ void beam_init_kernel(int B, int bw, int D1, int sos, BeamMeta m, int* nbeam, int* done, int64_t* out_tokens, int* out_len) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= B) return;
  for (int j = 0; j < bw; ++j) {
    const int s = v * bw + j;
    m.key[s] = j == 0 ? -0.0f : INFINITY;
    m.tok[s] = j == 0 ? sos : 0;
    m.len[s] = 1;
    m.fin[s] = 0;
    for (int d = 0; d < D1; ++d) m.hist[(long long)s * D1 + d] = (d == 0 && j == 0) ? sos : -1;
  }
  nbeam[v] = 1; done[v] = 0;
  for (int d = 0; d < D1; ++d) out_tokens[(long long)v * D1 + d] = d == 0 ? sos : -1;
  out_len[v] = 1;
}

// ---- one depth's bookkeeping in ONE kernel: a CTA per video, a warp per beam slot.
//   1. every warp merges its slot's per-tile (max, sum exp, top-KC) partials into the slot's best kc (log-prob, token) pairs: each
//      lane keeps the best KC of its share of the candidates (most lists are empty: the epilogue filters against the row's running
//      bound), then kc rounds pick the best lane head (highest value, lowest index on ties) and pop it
//   2. thread 0 runs the PriorityQueue step of S2VTModel.py:186-238 for the video on the candidates in shared memory.  `topk` is
//      the reference's expansion width (only used to count queue entries for the stop rule); kc <= topk candidates per slot are
//      enough because at most beam_width entries leave the queue per depth
//   3. every warp builds its NEW slot: token history (its parent's row + the token just appended; slot 0 is also the video's answer
//      so far) and the LSTM state copied from the parent slot; the candidate bounds of the next depth are reset
// (Four separate launches in the first version: merge 28 us, queue 12 us, histories 5 us, re-order 9 us one after the other.)
__global__ void beam_finish_kernel(int B, int bw, int topk, int kc, int D1, int eos, int HP, int n_part, const float* __restrict__ len_pen,
                                   BeamMeta old_, BeamMeta new_, const float* __restrict__ ms, const float* __restrict__ tv,
                                   const int* __restrict__ ti, int* __restrict__ nbeam, int* __restrict__ done, int* __restrict__ n_done,
                                   int64_t* __restrict__ out_tokens, int* __restrict__ out_len,
                                   __half* __restrict__ a1, long long a1_plane, __half* __restrict__ x, long long x_plane,
                                   const __half* __restrict__ h2n, long long h2n_plane,
                                   float* __restrict__ c1, const float* __restrict__ c1n, float* __restrict__ c2, const float* __restrict__ c2n,
                                   unsigned int* __restrict__ row_thr) {
  __shared__ float s_lp[KC * KC];
  __shared__ int s_tok[KC * KC];
  __shared__ int s_parent[KC], s_was_done;
  const int v = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = v * bw, s = base + w;
  const bool was_done = done[v] != 0;              // (read by every thread before thread 0 may set it below: barrier in between)
  const bool stamp = (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0);        // debug: phase times of one CTA (S2VT_XDEC_EVENTS=1)
  if (stamp) n_done[2] = (int)(ptx::globaltimer_ns() & 0x7fffffff);
  __shared__ float m_key[KC], m_pen[KC];
  __shared__ int m_tok[KC], m_len[KC], m_fin[KC], m_nb;
  if (threadIdx.x < bw) {                          // the video's queue state, fetched in parallel for the serial step below
    const int q = base + threadIdx.x;
    m_key[threadIdx.x] = old_.key[q]; m_tok[threadIdx.x] = old_.tok[q]; m_len[threadIdx.x] = old_.len[q]; m_fin[threadIdx.x] = old_.fin[q];
    m_pen[threadIdx.x] = len_pen[old_.len[q] + 1];
    if (threadIdx.x == 0) m_nb = nbeam[v];
  }
  // ---- 1. merge (skipped for a finished video: its candidates are never looked at)
  if (!was_done) {
    float mx = -INFINITY;
    for (int j = lane; j < n_part; j += 32) mx = fmaxf(mx, ms[2 * ((long long)s * n_part + j)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < n_part; j += 32) {
      const long long o = 2 * ((long long)s * n_part + j);
      sum += ms[o + 1] * expf(ms[o] - mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float logsum = logf(sum);
    // Candidate lists are sorted and mostly EMPTY (the vocab epilogue filters against the row's running bound): a lane first looks at
    // the head of each of its lists (one round trip for all of them), then reads only the non-empty ones.
    float lv[KC];
    int li[KC];
#pragma unroll
    for (int q = 0; q < KC; ++q) { lv[q] = -INFINITY; li[q] = 0x7fffffff; }
    constexpr int MAXL = 16;                                 // lists per lane handled per pass (n_part <= 512 in one pass)
    for (int l0 = lane; l0 < n_part; l0 += 32 * MAXL) {
      int head[MAXL];
#pragma unroll
      for (int u = 0; u < MAXL; ++u) {
        const int l = l0 + 32 * u;
        head[u] = l < n_part ? __ldcg(ti + ((long long)s * n_part + l) * KC) : 0x7fffffff;
      }
#pragma unroll
      for (int u = 0; u < MAXL; ++u) {
        if (head[u] == 0x7fffffff) continue;
        const long long o = ((long long)s * n_part + (l0 + 32 * u)) * KC;
        int ids[KC];
        float vs[KC];
#pragma unroll
        for (int q = 0; q < KC; ++q) { ids[q] = __ldcg(ti + o + q); vs[q] = __ldcg(tv + o + q); }
#pragma unroll
        for (int q = 0; q < KC; ++q) {
          const int id = ids[q];
          const float val = vs[q];
          if (id != 0x7fffffff && (val > lv[KC - 1] || (val == lv[KC - 1] && id < li[KC - 1]))) {
            lv[KC - 1] = val; li[KC - 1] = id;
#pragma unroll
            for (int qq = KC - 1; qq > 0; --qq) {
              if (lv[qq] > lv[qq - 1] || (lv[qq] == lv[qq - 1] && li[qq] < li[qq - 1])) {
                const float fv = lv[qq]; lv[qq] = lv[qq - 1]; lv[qq - 1] = fv;
                const int iv = li[qq]; li[qq] = li[qq - 1]; li[qq - 1] = iv;
              }
            }
          }
        }
      }
    }
    for (int r = 0; r < kc; ++r) {
      float bv = lv[0];
      int bi = li[0], bl = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bl = ol; }
      }
      if (lane == 0) {
        s_lp[w * KC + r] = (bv - mx) - logsum;                    // log_softmax as torch computes it: (z - max) - log(sum exp)
        s_tok[w * KC + r] = bi != 0x7fffffff ? bi : 0;
      }
      if (lane == bl && bi != 0x7fffffff) {
#pragma unroll
        for (int q = 0; q < KC - 1; ++q) { lv[q] = lv[q + 1]; li[q] = li[q + 1]; }
        lv[KC - 1] = -INFINITY; li[KC - 1] = 0x7fffffff;
      }
    }
  }
  __syncthreads();
  if (stamp) n_done[3] = (int)(ptx::globaltimer_ns() & 0x7fffffff);
  // ---- 2. the PriorityQueue step of S2VTModel.py:186-238
  if (threadIdx.x == 0) {
    s_was_done = was_done ? 1 : 0;
    if (was_done) {
      for (int j = 0; j < bw; ++j) {
        const int q = base + j;
        new_.key[q] = m_key[j]; new_.tok[q] = m_tok[j]; new_.len[q] = m_len[j]; new_.fin[q] = m_fin[j];
        s_parent[j] = q;
      }
    } else {
      const int nb = m_nb;
      int ptr[KC];
      long long count = 0;
      for (int j = 0; j < nb; ++j) { ptr[j] = 0; count += m_fin[j] ? 1 : topk; }
      const bool last = (count <= bw);
      const int take = count < bw ? (int)count : bw;
      for (int r = 0; r < take; ++r) {
        float bestk = INFINITY; int bj = -1;
        for (int j = 0; j < nb; ++j) {
          float k;
          if (m_fin[j]) { if (ptr[j] > 0) continue; k = m_key[j]; }
          else {
            if (ptr[j] >= kc) continue;
            k = -(s_lp[j * KC + ptr[j]] / m_pen[j]);
          }
          if (bj < 0 || k < bestk) { bestk = k; bj = j; }
        }
        const int d = base + r;
        if (bj < 0) {
          new_.key[d] = INFINITY; new_.tok[d] = 0; new_.len[d] = 1; new_.fin[d] = 0; s_parent[r] = d;
          continue;
        }
        const int q = base + bj;
        const int ln = m_len[bj];
        if (m_fin[bj]) {
          new_.key[d] = m_key[bj]; new_.tok[d] = m_tok[bj]; new_.len[d] = ln; new_.fin[d] = 1;
        } else {
          const int tk = s_tok[bj * KC + ptr[bj]];
          new_.key[d] = bestk; new_.tok[d] = tk; new_.len[d] = ln + 1; new_.fin[d] = (tk == eos) ? 1 : 0;
        }
        ptr[bj] += 1;
        s_parent[r] = q;
        if (r == 0) out_len[v] = new_.len[d];
      }
      for (int r = take; r < bw; ++r) {
        const int d = base + r;
        new_.key[d] = INFINITY; new_.tok[d] = 0; new_.len[d] = 1; new_.fin[d] = 0;
        s_parent[r] = d;
      }
      nbeam[v] = take;
      if (last) { done[v] = 1; atomicAdd(n_done, 1); }
    }
    __threadfence_block();
  }
  __syncthreads();
  if (stamp) n_done[4] = (int)(ptx::globaltimer_ns() & 0x7fffffff);
  // ---- 3. the new slot d = base + w: history, answer so far, state from its parent
  const int d = s;
  if (lane < KC) row_thr[(long long)d * KC + lane] = 0u;            // next depth's candidate bounds start from "none"
  if (s_was_done) return;                                           // frozen: outputs final; its state buffers keep finite values
  const int ps = s_parent[w];
  {
    const bool unused = (new_.key[d] == INFINITY);
    const bool fresh = !unused && !m_fin[ps - base];
    const int Ln = new_.len[d];
    for (int q = lane; q < D1; q += 32) {
      int t = unused ? -1 : old_.hist[(long long)ps * D1 + q];
      if (fresh && q == Ln - 1) t = new_.tok[d];
      new_.hist[(long long)d * D1 + q] = t;
      if (w == 0) out_tokens[(long long)v * D1 + q] = (q < Ln) ? t : -1;
    }
  }
  // 16-byte copies (HP is a multiple of 8: rows of fp16 planes and fp32 cells are 16-byte aligned); all loads of a row first
  {
    const int nh = HP / 8, nf = HP / 4;                     // uint4 per plane row / per cell row
    const uint4* xs0 = reinterpret_cast<const uint4*>(x + (long long)ps * 2 * HP);
    const uint4* xs1 = reinterpret_cast<const uint4*>(x + x_plane + (long long)ps * 2 * HP);
    const uint4* hs0 = reinterpret_cast<const uint4*>(h2n + (long long)ps * HP);
    const uint4* hs1 = reinterpret_cast<const uint4*>(h2n + h2n_plane + (long long)ps * HP);
    uint4* ad0 = reinterpret_cast<uint4*>(a1 + (long long)d * HP);
    uint4* ad1 = reinterpret_cast<uint4*>(a1 + a1_plane + (long long)d * HP);
    // (x[:, HP:] is written for slot d while another warp may still read x[:, :HP] of slot d as ITS parent: different columns)
    uint4* xd0 = reinterpret_cast<uint4*>(x + (long long)d * 2 * HP + HP);
    uint4* xd1 = reinterpret_cast<uint4*>(x + x_plane + (long long)d * 2 * HP + HP);
    for (int u = lane; u < nh; u += 32) {
      const uint4 v0 = xs0[u], v1 = xs1[u], v2 = hs0[u], v3 = hs1[u];
      ad0[u] = v0; ad1[u] = v1; xd0[u] = v2; xd1[u] = v3;
    }
    const uint4* cs1 = reinterpret_cast<const uint4*>(c1n + (long long)ps * HP);
    const uint4* cs2 = reinterpret_cast<const uint4*>(c2n + (long long)ps * HP);
    uint4* cd1 = reinterpret_cast<uint4*>(c1 + (long long)d * HP);
    uint4* cd2 = reinterpret_cast<uint4*>(c2 + (long long)d * HP);
    for (int u = lane; u < nf; u += 32) {
      const uint4 v0 = cs1[u], v1 = cs2[u];
      cd1[u] = v0; cd2[u] = v1;
    }
  }
  if (stamp) n_done[5] = (int)(ptx::globaltimer_ns() & 0x7fffffff);
}

// slot 0 of each video <- the video's encode state; the other slots start from zero
__global__ void beam_state_init_kernel(int B, int bw, int HP, const __half* __restrict__ h1, long long h1_plane, const float* __restrict__ c1e,
                                       const __half* __restrict__ h2, long long h2_plane, const float* __restrict__ c2e,
                                       __half* __restrict__ a1, long long a1_plane, __half* __restrict__ x, long long x_plane,
                                       float* __restrict__ c1, float* __restrict__ c2) {
  const int s = blockIdx.x;
  const int v = s / bw, j = s % bw;
  const __half z = __float2half_rn(0.f);
  for (int u = threadIdx.x; u < HP; u += blockDim.x) {
    const bool on = (j == 0);
    a1[(long long)s * HP + u] = on ? h1[(long long)v * HP + u] : z;
    a1[a1_plane + (long long)s * HP + u] = on ? h1[h1_plane + (long long)v * HP + u] : z;
    x[(long long)s * 2 * HP + HP + u] = on ? h2[(long long)v * HP + u] : z;
    x[x_plane + (long long)s * 2 * HP + HP + u] = on ? h2[h2_plane + (long long)v * HP + u] : z;
    c1[(long long)s * HP + u] = on ? c1e[(long long)v * HP + u] : 0.f;
    c2[(long long)s * HP + u] = on ? c2e[(long long)v * HP + u] : 0.f;
  }
}

// ------------------------------------------------------------------ host helpers
static inline long long rup(long long x, long long a) { return (x + a - 1) / a * a; }

struct Planes {            // an fp16 (hi, lo) operand: rows x k, leading dimension ld, planes `plane` elements apart
  const __half* p; long long ld; long long plane; const float* inv;
  const float* inv_row = nullptr;     // per-row 1/scale instead of `inv` (A operands only)
};

static unsigned long long* g_trace_buf = nullptr;     // s2vt_xdec_set_trace: [max][8] stamps, one record per xgemm launch
static int g_trace_max = 0, g_trace_n = 0;

template <int BN, int EPI, int MINB = 1>
static int launch_x(cudaStream_t st, int M, int N, int K, const Planes& A, const Planes& B, XParams p) {
  constexpr int STAGES = (MINB == 2) ? 2 : ((BN == 128) ? 3 : 4);
  constexpr int SMEM = STAGES * (2 * PLANE_A + 2 * BN * BK * 2) + 1024;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_planes(&tmA, A.p, (uint64_t)K, (uint64_t)M, (uint64_t)A.ld, (uint64_t)A.plane, BM);
  if (rc) return rc;
  rc = make_tmap_planes(&tmB, B.p, (uint64_t)K, (uint64_t)N, (uint64_t)B.ld, (uint64_t)B.plane, BN);
  if (rc) return rc;
  p.M = M; p.N = N; p.K = K; p.num_kb = (K + BK - 1) / BK;
  p.a_inv = A.inv; p.b_inv = B.inv; p.a_inv_row = A.inv_row;
  if (g_trace_buf && g_trace_n < g_trace_max) {
    p.trace = g_trace_buf + 8 * (size_t)g_trace_n++;
    static int tbx = -1, tby = 0;
    if (tbx < 0) { const char* e = getenv("S2VT_XDEC_TRACE_CTA"); tbx = 0; if (e) sscanf(e, "%d,%d", &tbx, &tby); }
    p.trace_bx = tbx; p.trace_by = tby;
  }
  static bool attr_set = false;
  if (!attr_set) {
    S2VT_CHECK_CUDA(cudaFuncSetAttribute(xgemm_kernel<BN, EPI, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("S2VT_XDEC_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = p.m_fastest ? dim3(ceil_div(M, BM), ceil_div(N, BN)) : dim3(ceil_div(N, BN), ceil_div(M, BM));
  cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  S2VT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, xgemm_kernel<BN, EPI, MINB>, tmA, tmB, p));
  S2VT_CHECK_LAUNCH();
  return 0;
}

static int x_store(cudaStream_t st, int M, int N, int K, const Planes& A, const Planes& B, float* C, RowMap cm, const float* bias,
                   int accumulate) {
  XParams p{};
  p.C = C; p.cm = cm; p.bias = bias; p.accumulate = accumulate;
  p.c_vec = aligned16(C) && (cm.so % 4 == 0) && (cm.si % 4 == 0);
  return launch_x<128, EPI_STORE>(st, M, N, K, A, B, p);
}

static int x_split(cudaStream_t st, const float* x, long long rows, int cols, long long ld, int il_H, unsigned int* bits, float fixed_scale,
                   __half* planes, long long out_ld, long long plane_stride, float* inv_out, bool do_absmax = true) {
  const long long n = rows * cols;
  if (n <= 0) return 0;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (bits && do_absmax) {
    S2VT_CHECK_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned int), st));
    absmax_kernel<<<blocks, 256, 0, st>>>(x, rows, cols, ld, bits);
    S2VT_CHECK_LAUNCH();
  }
  split_kernel<<<blocks, 256, 0, st>>>(x, rows, cols, ld, il_H, bits, fixed_scale, planes, out_ld, plane_stride, inv_out);
  S2VT_CHECK_LAUNCH();
  return 0;
}

// ---- prepared weights: one caller-owned buffer, carved identically by every entry point
struct Cfg { int V, F, L, H, E, sos, eos; int HP, FP, EP; };
static Cfg make_cfg(const s2vt_xdec_cfg& c) {
  Cfg g{c.vocab_size, c.feat_dim, c.length, c.dim_hid, c.dim_embed, c.sos_ix, c.eos_ix, 0, 0, 0};
  g.HP = (int)rup(g.H, 8); g.FP = (int)rup(g.F, 8); g.EP = (int)rup(g.E, 8);
  return g;
}
enum { INV_FEAT = 0, INV_IH1, INV_HH1, INV_IH2V, INV_IH2E, INV_HH2, INV_CAT2, INV_OUT, INV_EMB, INV_H, N_INV };

struct Weights {
  float* inv;                 // [N_INV] inverse scales (+ N_INV absmax words behind them)
  unsigned int* bits;
  __half *feat, *ih1, *hh1, *ih2v, *ih2e, *hh2, *cat2, *out, *emb;
  float *bf, *b1, *b2, *bout, *ew;
  size_t bytes;
};
static Weights carve_weights(char* base, const Cfg& g) {
  Weights w{};
  size_t off = 0;
  auto take = [&](size_t n) { char* p = base ? base + off : nullptr; off += (size_t)rup((long long)n, 256); return p; };
  const size_t G = 4 * (size_t)g.HP;
  w.inv = (float*)take(sizeof(float) * 32);
  w.bits = (unsigned int*)take(sizeof(unsigned int) * 32);
  w.feat = (__half*)take(2 * 2 * (size_t)g.HP * g.FP);
  w.ih1 = (__half*)take(2 * 2 * G * g.HP);
  w.hh1 = (__half*)take(2 * 2 * G * g.HP);
  w.ih2v = (__half*)take(2 * 2 * G * g.HP);
  w.ih2e = (__half*)take(2 * 2 * G * g.EP);
  w.hh2 = (__half*)take(2 * 2 * G * g.HP);
  w.cat2 = (__half*)take(2 * 2 * G * 2 * g.HP);
  w.out = (__half*)take(2 * 2 * (size_t)g.V * g.HP);
  w.emb = (__half*)take(2 * 2 * (size_t)g.V * g.EP);
  w.bf = (float*)take(sizeof(float) * g.HP);
  w.b1 = (float*)take(sizeof(float) * G);
  w.b2 = (float*)take(sizeof(float) * G);
  w.bout = (float*)take(sizeof(float) * g.V);
  w.ew = (float*)take(sizeof(float) * (size_t)g.V * G);
  w.bytes = off;
  return w;
}

static thread_local cudaStream_t g_side = nullptr;
static thread_local cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static thread_local cudaEvent_t g_bev[4] = {nullptr, nullptr, nullptr, nullptr};
static int side_stream(cudaStream_t* out) {
  if (!g_side) {
    S2VT_CHECK_CUDA(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) S2VT_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) S2VT_CHECK_CUDA(cudaEventCreateWithFlags(&g_bev[i], cudaEventDisableTiming));
  }
  *out = g_side;
  return 0;
}

// one recurrence step (vid_rnn or word_rnn): gates = A . W^T (+ pre | bias) (+ gtab[token]) -> c, h planes
struct StepArgs {
  const float* pre; long long pre_ld; const float* bias;
  const float* gtab; long long gtab_ld; const int* gidx;
  const unsigned long long* key_in; unsigned long long* key_clear; int64_t* tok_out; long long tok_ld;
  const float* c_in; float* c_out; __half* hp; long long hp_ld; long long hp_plane; float* h_f32; long long h_ld;
};
static int x_lstm_step(cudaStream_t st, int M, int HP, int K, const Planes& A, const Planes& W, const StepArgs& a) {
  XParams p{};
  p.HP = HP; p.pre = a.pre; p.pre_ld = a.pre_ld; p.bias = a.bias;
  p.gtab = a.gtab; p.gtab_ld = a.gtab_ld; p.gidx = a.gidx;
  p.key_in = a.key_in; p.key_clear = a.key_clear; p.tok_out = a.tok_out; p.tok_ld = a.tok_ld;
  p.c_in = a.c_in; p.c_out = a.c_out; p.hp = a.hp; p.hp_ld = a.hp_ld; p.hp_plane = a.hp_plane; p.h_f32 = a.h_f32; p.h_ld = a.h_ld;
  // up to four row tiles: narrow tiles put the step on twice as many SMs (one wave, one 32-column chunk per epilogue warp)
  if (M <= 512) return launch_x<64, EPI_LSTM>(st, M, 4 * HP, K, A, W, p);
  return launch_x<128, EPI_LSTM>(st, M, 4 * HP, K, A, W, p);
}

}  // namespace xd

// 3-D tensor map (k, row, plane) over an fp16 (hi, lo) operand; box = {64 k, box_rows, 2 planes}, 128B swizzle, zero fill
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// Encoded maps are a pure function of the arguments; a decode call encodes ~1100 of them and repeats the same ones on every
// call with the same workspace, so they are memoised per host thread (cuTensorMapEncodeTiled costs ~1.5 us).
struct TmapKey { const void* base; uint64_t k, rows, ld, plane; uint32_t box; };
struct TmapSlot { TmapKey key; CUtensorMap map; bool used; };
static thread_local TmapSlot* g_tmap_cache = nullptr;
constexpr uint32_t TMAP_SLOTS = 8192;

int make_tmap_planes(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t ld, uint64_t plane_stride,
                     uint32_t box_rows) {
  if (!g_tmap_cache) g_tmap_cache = new TmapSlot[TMAP_SLOTS]();
  uint64_t h = reinterpret_cast<uintptr_t>(base) * 0x9E3779B97F4A7C15ull ^ (k * 0xC2B2AE3D27D4EB4Full) ^ (rows * 0x165667B19E3779F9ull) ^
               (ld << 17) ^ (plane_stride << 3) ^ box_rows;
  h ^= h >> 29;
  TmapSlot& slot = g_tmap_cache[(uint32_t)(h % TMAP_SLOTS)];
  if (slot.used && slot.key.base == base && slot.key.k == k && slot.key.rows == rows && slot.key.ld == ld && slot.key.plane == plane_stride &&
      slot.key.box == box_rows) {
    *out = slot.map;
    return 0;
  }
  static EncodeTiledFn3 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFn3)fp;
  }
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) || (plane_stride % 8))
    return fail("fp16 plane operand must be 16-byte aligned (ld=%llu plane=%llu)", (unsigned long long)ld, (unsigned long long)plane_stride);
  cuuint64_t dims[3] = {k, rows, 2};
  cuuint64_t strides[2] = {ld * 2, plane_stride * 2};
  cuuint32_t box[3] = {64, box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (planes) failed with CUresult %d (k=%llu rows=%llu ld=%llu plane=%llu)", (int)r,
                                     (unsigned long long)k, (unsigned long long)rows, (unsigned long long)ld, (unsigned long long)plane_stride);
  slot.key = TmapKey{base, k, rows, ld, plane_stride, box_rows};
  slot.map = *out;
  slot.used = true;
  return 0;
}

int xdec_error_flag() { return read_sm100_error_flag(); }
int xdec_error_clear() { return clear_sm100_error_flag(); }

}  // namespace s2vt

using namespace s2vt;
using namespace s2vt::xd;

// ====================================================================== C ABI
/* Debug: stamp %globaltimer at the phases of CTA (0,0) of every following xgemm launch into buf[max_records][8]
 * (0 start, 1 weights requested, 2 predecessor done, 3 first tile landed, 4 last tile landed, 5 accumulators complete,
 * 6 epilogue done); buf = NULL stops.  Returns the number of records handed out so far. */
extern "C" int s2vt_xdec_set_trace(void* buf, int max_records) {
  const int n = g_trace_n;
  g_trace_buf = (unsigned long long*)buf; g_trace_max = buf ? max_records : 0; g_trace_n = 0;
  return n;
}

extern "C" int64_t s2vt_xgemm_ws_bytes(int M, int N, int K) {
  const long long KP = rup(K, 8);
  return (int64_t)(2 * 2 * ((long long)M + N) * KP + 1024 + 256);
}

// C[cmap(m), n] = sum_k A[m,k] B[n,k] (+ bias[n]) (+ C): fp32 operands, split on the fly (test / utility entry)
extern "C" int s2vt_xgemm_f32(void* stream, int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb,
                              float* C, s2vt_rowmap cmap, const float* bias, int accumulate, void* ws) {
  cudaStream_t st = (cudaStream_t)stream;
  S2VT_REQUIRE(M > 0 && N > 0 && K > 0, "s2vt_xgemm_f32: dimensions must be positive");
  S2VT_REQUIRE(A && B && C && ws, "s2vt_xgemm_f32: null pointer");
  S2VT_REQUIRE(cmap.inner >= 1, "s2vt_xgemm_f32: rowmap.inner must be >= 1");
  const long long KP = rup(K, 8);
  char* w = (char*)ws;
  float* inv = (float*)w;                               // [2]
  unsigned int* bits = (unsigned int*)(w + 64);         // [2]
  __half* pa = (__half*)(w + 256);
  __half* pb = pa + 2 * (long long)M * KP;
  S2VT_CHECK_CUDA(cudaMemsetAsync(pa, 0, 2 * 2 * ((size_t)M + N) * KP, st));
  int rc = x_split(st, A, M, K, lda, 0, bits, 1.f, pa, KP, (long long)M * KP, inv);
  if (rc) return rc;
  rc = x_split(st, B, N, K, ldb, 0, bits + 1, 1.f, pb, KP, (long long)N * KP, inv + 1);
  if (rc) return rc;
  Planes PA{pa, KP, (long long)M * KP, inv}, PB{pb, KP, (long long)N * KP, inv + 1};
  return x_store(st, M, N, (int)KP, PA, PB, C, to_rowmap(cmap), bias, accumulate);
}

extern "C" int64_t s2vt_xdec_weights_bytes(s2vt_xdec_cfg cfg) {
  return (int64_t)carve_weights(nullptr, make_cfg(cfg)).bytes + 256;
}

// params: the 13 tensors of the state_dict in registration order (S2VTModel.py:19-28)
extern "C" int s2vt_xdec_prepare(void* stream, s2vt_xdec_cfg cfg, const float* const* params, void* wbuf) {
  cudaStream_t st = (cudaStream_t)stream;
  const Cfg g = make_cfg(cfg);
  S2VT_REQUIRE(g.V > 0 && g.F > 0 && g.L > 1 && g.H > 0 && g.E > 0, "s2vt_xdec_prepare: bad config");
  S2VT_REQUIRE(params && wbuf, "s2vt_xdec_prepare: null pointer");
  for (int i = 0; i < 13; ++i) S2VT_REQUIRE(params[i], "s2vt_xdec_prepare: null parameter %d", i);
  const float *w_ih1 = params[0], *w_hh1 = params[1], *b_ih1 = params[2], *b_hh1 = params[3];
  const float *w_ih2 = params[4], *w_hh2 = params[5], *b_ih2 = params[6], *b_hh2 = params[7];
  const float *w_f = params[8], *b_f = params[9], *w_o = params[10], *b_o = params[11], *emb = params[12];
  Weights w = carve_weights((char*)wbuf, g);
  const int H = g.H, HP = g.HP, G = 4 * g.HP;
  S2VT_CHECK_CUDA(cudaMemsetAsync(wbuf, 0, w.bytes - (size_t)rup((long long)sizeof(float) * g.V * G, 256), st));   // everything but EW
  int rc;
  X_TRY(x_split(st, w_f, H, g.F, g.F, 0, w.bits + INV_FEAT, 1.f, w.feat, g.FP, (long long)HP * g.FP, w.inv + INV_FEAT));
  X_TRY(x_split(st, w_ih1, 4 * H, H, H, H, w.bits + INV_IH1, 1.f, w.ih1, HP, (long long)G * HP, w.inv + INV_IH1));
  X_TRY(x_split(st, w_hh1, 4 * H, H, H, H, w.bits + INV_HH1, 1.f, w.hh1, HP, (long long)G * HP, w.inv + INV_HH1));
  X_TRY(x_split(st, w_ih2 + g.E, 4 * H, H, g.E + H, H, w.bits + INV_IH2V, 1.f, w.ih2v, HP, (long long)G * HP, w.inv + INV_IH2V));
  X_TRY(x_split(st, w_ih2, 4 * H, g.E, g.E + H, H, w.bits + INV_IH2E, 1.f, w.ih2e, g.EP, (long long)G * g.EP, w.inv + INV_IH2E));
  X_TRY(x_split(st, w_hh2, 4 * H, H, H, H, w.bits + INV_HH2, 1.f, w.hh2, HP, (long long)G * HP, w.inv + INV_HH2));
  // cat2 = [ W_ih2[:, E:] | W_hh2 ] shares one scale: max over both halves
  {
    S2VT_CHECK_CUDA(cudaMemsetAsync(w.bits + INV_CAT2, 0, sizeof(unsigned int), st));
    const int blocks = 148 * 4;
    absmax_kernel<<<blocks, 256, 0, st>>>(w_ih2 + g.E, 4 * H, H, g.E + H, w.bits + INV_CAT2);
    S2VT_CHECK_LAUNCH();
    absmax_kernel<<<blocks, 256, 0, st>>>(w_hh2, 4 * H, H, H, w.bits + INV_CAT2);
    S2VT_CHECK_LAUNCH();
    X_TRY(x_split(st, w_ih2 + g.E, 4 * H, H, g.E + H, H, w.bits + INV_CAT2, 1.f, w.cat2, 2 * HP, (long long)G * 2 * HP, w.inv + INV_CAT2, false));
    X_TRY(x_split(st, w_hh2, 4 * H, H, H, H, w.bits + INV_CAT2, 1.f, w.cat2 + HP, 2 * HP, (long long)G * 2 * HP, nullptr, false));
  }
  X_TRY(x_split(st, w_o, g.V, H, H, 0, w.bits + INV_OUT, 1.f, w.out, HP, (long long)g.V * HP, w.inv + INV_OUT));
  X_TRY(x_split(st, emb, g.V, g.E, g.E, 0, w.bits + INV_EMB, 1.f, w.emb, g.EP, (long long)g.V * g.EP, w.inv + INV_EMB));
  {
    const float hinv = 1.f / 32768.f;
    S2VT_CHECK_CUDA(cudaMemcpyAsync(w.inv + INV_H, &hinv, sizeof(float), cudaMemcpyHostToDevice, st));
  }
  bias_interleave_kernel<<<ceil_div(H, 128), 128, 0, st>>>(b_f, nullptr, H, 0, w.bf);
  S2VT_CHECK_LAUNCH();
  bias_interleave_kernel<<<ceil_div(4 * H, 128), 128, 0, st>>>(b_ih1, b_hh1, H, 1, w.b1);
  S2VT_CHECK_LAUNCH();
  bias_interleave_kernel<<<ceil_div(4 * H, 128), 128, 0, st>>>(b_ih2, b_hh2, H, 1, w.b2);
  S2VT_CHECK_LAUNCH();
  S2VT_CHECK_CUDA(cudaMemcpyAsync(w.bout, b_o, sizeof(float) * g.V, cudaMemcpyDeviceToDevice, st));
  // EW[v, 4u+g] = sum_e embedding[v, e] W_ih2[g*H+u, e]: the embedding half of word_rnn's input product for every token
  Planes PE{w.emb, g.EP, (long long)g.V * g.EP, w.inv + INV_EMB}, PW{w.ih2e, g.EP, (long long)G * g.EP, w.inv + INV_IH2E};
  X_TRY(x_store(st, g.V, G, g.EP, PE, PW, w.ew, RowMap{1, (long long)G, 0}, nullptr, 0));
  return 0;
}

namespace {
struct GreedyWs {
  float* inv; unsigned int* bits;
  float* finv;                   // [L*B] per-row 1/scale of the feature planes
  __half *fa, *xp, *o1p, *h2p;
  float *xproj, *pre1, *pre2, *c1, *c2;
  int* sos;
  unsigned long long* key;       // [3][B] argmax keys, rotating over the decode steps (written k, read k+1, cleared k+2)
  size_t bytes;
};
GreedyWs carve_greedy(char* base, const Cfg& g, int B, int T1 /* vid_rnn steps */) {
  GreedyWs w{};
  size_t off = 0;
  auto take = [&](size_t n) { char* p = base ? base + off : nullptr; off += (size_t)rup((long long)n, 256); return p; };
  const size_t G = 4 * (size_t)g.HP, LB = (size_t)g.L * B;
  const int n_part = ceil_div(g.V, 128);
  w.inv = (float*)take(64); w.bits = (unsigned int*)take(64);
  w.finv = (float*)take(4 * LB);
  w.fa = (__half*)take(2 * 2 * LB * g.FP);
  w.xproj = (float*)take(4 * LB * g.HP);
  w.xp = (__half*)take(2 * 2 * LB * g.HP);
  w.pre1 = (float*)take(4 * LB * G);
  w.o1p = (__half*)take(2 * 2 * (size_t)(T1 + 1) * B * g.HP);
  w.pre2 = (float*)take(4 * (size_t)T1 * B * G);
  w.h2p = (__half*)take(2 * 2 * 2 * (size_t)B * g.HP);
  w.c1 = (float*)take(4 * 2 * (size_t)B * g.HP);
  w.c2 = (float*)take(4 * 2 * (size_t)B * g.HP);
  w.sos = (int*)take(4 * (size_t)B);
  w.key = (unsigned long long*)take(8 * 3 * (size_t)B);
  w.bytes = off;
  return w;
}

// Shared encode: feat_linear, vid_rnn over T1 steps (the first L with real frames), word_rnn over the first L steps.
// vid_rnn runs on `st`; the word_rnn input products and steps trail on the side stream, one chunk of CH steps behind.
// On return (host side) the side stream holds the tail of the work; the caller continues on it and joins at the end.
int encode(cudaStream_t st, cudaStream_t sd, const Cfg& g, const Weights& W, const GreedyWs& w, int B, const float* feats, int T1) {
  const int L = g.L, HP = g.HP, G = 4 * g.HP;
  const long long BH = (long long)B * HP;
  int rc;
  // features -> fp16 planes; xproj = feat_linear(feats), rows re-ordered batch-major -> time-major
  if (g.FP != g.F) S2VT_CHECK_CUDA(cudaMemsetAsync(w.fa, 0, 2 * 2 * (size_t)L * B * g.FP, st));
  split_rows_kernel<<<(unsigned)(B * L), 256, 0, st>>>(feats, g.F, g.F, w.fa, g.FP, (long long)B * L * g.FP, w.finv);
  S2VT_CHECK_LAUNCH();
  Planes PF{w.fa, g.FP, (long long)B * L * g.FP, nullptr, w.finv}, WF{W.feat, g.FP, (long long)HP * g.FP, W.inv + INV_FEAT};
  X_TRY(x_store(st, B * L, HP, g.FP, PF, WF, w.xproj, RowMap{L, (long long)HP, BH}, W.bf, 0));
  X_TRY(x_split(st, w.xproj, (long long)L * B, HP, HP, 0, w.bits + 1, 1.f, w.xp, HP, (long long)L * B * HP, w.inv + 1));
  Planes PX{w.xp, HP, (long long)L * B * HP, w.inv + 1}, WI1{W.ih1, HP, (long long)G * HP, W.inv + INV_IH1};
  X_TRY(x_store(st, L * B, G, HP, PX, WI1, w.pre1, RowMap{1, (long long)G, 0}, W.b1, 0));
  // recurrence state: slot 0 of the h-plane history is h_{-1} = 0
  const long long o1_plane = (long long)(T1 + 1) * BH;
  S2VT_CHECK_CUDA(cudaMemsetAsync(w.o1p, 0, 2 * (size_t)BH, st));
  S2VT_CHECK_CUDA(cudaMemsetAsync(w.o1p + o1_plane, 0, 2 * (size_t)BH, st));
  S2VT_CHECK_CUDA(cudaMemsetAsync(w.h2p, 0, 2 * 2 * 2 * (size_t)BH, st));
  S2VT_CHECK_CUDA(cudaEventRecord(g_ev[0], st));
  S2VT_CHECK_CUDA(cudaStreamWaitEvent(sd, g_ev[0], 0));
  const Planes WH1{W.hh1, HP, (long long)G * HP, W.inv + INV_HH1}, WH2{W.hh2, HP, (long long)G * HP, W.inv + INV_HH2};
  const Planes WI2{W.ih2v, HP, (long long)G * HP, W.inv + INV_IH2V};
  const int CH = 16;
  for (int t0 = 0; t0 < T1; t0 += CH) {
    const int t1 = t0 + CH < T1 ? t0 + CH : T1;
    for (int t = t0; t < t1; ++t) {
      StepArgs a{};
      a.pre = t < L ? w.pre1 + (long long)t * B * G : nullptr; a.pre_ld = G; a.bias = W.b1;
      a.c_in = t == 0 ? nullptr : w.c1 + (long long)((t - 1) & 1) * BH; a.c_out = w.c1 + (long long)(t & 1) * BH;
      a.hp = w.o1p + (long long)(t + 1) * BH; a.hp_ld = HP; a.hp_plane = o1_plane;
      Planes A{w.o1p + (long long)t * BH, HP, o1_plane, W.inv + INV_H};
      X_TRY(x_lstm_step(st, B, HP, HP, A, WH1, a));
    }
    S2VT_CHECK_CUDA(cudaEventRecord(g_ev[1], st));
    S2VT_CHECK_CUDA(cudaStreamWaitEvent(sd, g_ev[1], 0));
    // vid half of word_rnn's input for the chunk: pre2[t] = out1[t] . W_ih2[:, E:]^T + b2
    Planes A{w.o1p + (long long)(t0 + 1) * BH, HP, o1_plane, W.inv + INV_H};
    X_TRY(x_store(sd, (t1 - t0) * B, G, HP, A, WI2, w.pre2 + (long long)t0 * B * G, RowMap{1, (long long)G, 0}, W.b2, 0));
    for (int t = t0; t < t1 && t < L; ++t) {
      StepArgs a{};
      a.pre = w.pre2 + (long long)t * B * G; a.pre_ld = G; a.bias = W.b2;
      a.c_in = t == 0 ? nullptr : w.c2 + (long long)((t - 1) & 1) * BH; a.c_out = w.c2 + (long long)(t & 1) * BH;
      a.hp = w.h2p + (long long)((t + 1) & 1) * BH; a.hp_ld = HP; a.hp_plane = 2 * BH;
      Planes A2{w.h2p + (long long)(t & 1) * BH, HP, 2 * BH, W.inv + INV_H};
      X_TRY(x_lstm_step(sd, B, HP, HP, A2, WH2, a));
    }
  }
  return 0;
}
}  // namespace

extern "C" int64_t s2vt_xdec_greedy_ws_bytes(s2vt_xdec_cfg cfg, int B) {
  const Cfg g = make_cfg(cfg);
  return (int64_t)carve_greedy(nullptr, g, B, 2 * g.L - 1).bytes + 256;
}

// S2VT.forward(mode='test'), S2VTModel.py:82-110.  feats [B, L, F] f32 -> tokens [B, L-1] i64.
extern "C" int s2vt_xdec_greedy(void* stream, s2vt_xdec_cfg cfg, const void* wbuf, int B, const float* feats, int64_t* tokens, void* ws) {
  cudaStream_t st = (cudaStream_t)stream, sd;
  const Cfg g = make_cfg(cfg);
  S2VT_REQUIRE(B > 0 && wbuf && feats && tokens && ws, "s2vt_xdec_greedy: bad arguments");
  S2VT_REQUIRE(g.sos >= 0 && g.sos < g.V, "s2vt_xdec_greedy: sos_ix out of range");
  int rc = side_stream(&sd);
  if (rc) return rc;
  const Weights W = carve_weights((char*)const_cast<void*>(wbuf), g);
  const int L = g.L, T = 2 * L - 1, HP = g.HP, G = 4 * g.HP;
  const GreedyWs w = carve_greedy((char*)ws, g, B, T);
  const long long BH = (long long)B * HP;
  const int n_part = ceil_div(g.V, 128);
  fill_i32_kernel<<<ceil_div(B, 256), 256, 0, st>>>(w.sos, B, g.sos);
  S2VT_CHECK_LAUNCH();
  S2VT_CHECK_CUDA(cudaMemsetAsync(w.key, 0, 8 * 3 * (size_t)B, st));
  X_TRY(encode(st, sd, g, W, w, B, feats, T));
  // decode steps on the side stream (it already holds word_rnn's encode steps): word_rnn step -> out_linear + argmax
  const Planes WH2{W.hh2, HP, (long long)G * HP, W.inv + INV_HH2}, WO{W.out, HP, (long long)g.V * HP, W.inv + INV_OUT};
  for (int k = 0; k < L - 1; ++k) {
    const int t = L + k;
    StepArgs a{};
    a.pre = w.pre2 + (long long)t * B * G; a.pre_ld = G; a.bias = W.b2;
    a.gtab = W.ew; a.gtab_ld = G;
    if (k == 0) a.gidx = w.sos;
    else { a.key_in = w.key + (size_t)((k - 1) % 3) * B; a.tok_out = tokens + (k - 1); a.tok_ld = L - 1; }
    a.key_clear = w.key + (size_t)((k + 1) % 3) * B;          // the keys step k+1 will accumulate into (last read by step k-1)
    a.c_in = w.c2 + (long long)((t - 1) & 1) * BH; a.c_out = w.c2 + (long long)(t & 1) * BH;
    a.hp = w.h2p + (long long)((t + 1) & 1) * BH; a.hp_ld = HP; a.hp_plane = 2 * BH;
    Planes A2{w.h2p + (long long)(t & 1) * BH, HP, 2 * BH, W.inv + INV_H};
    X_TRY(x_lstm_step(sd, B, HP, HP, A2, WH2, a));
    XParams p{};
    p.bias = W.bout; p.o_key = w.key + (size_t)(k % 3) * B;
    Planes AH{w.h2p + (long long)((t + 1) & 1) * BH, HP, 2 * BH, W.inv + INV_H};
    X_TRY((launch_x<128, EPI_ARGMAX>(sd, B, g.V, HP, AH, WO, p)));
  }
  keys_to_tokens_kernel<<<ceil_div(B, 128), 128, 0, sd>>>(B, w.key + (size_t)((L - 2) % 3) * B, tokens + (L - 2), L - 1);
  S2VT_CHECK_LAUNCH();
  S2VT_CHECK_CUDA(cudaEventRecord(g_ev[2], sd));
  S2VT_CHECK_CUDA(cudaStreamWaitEvent(st, g_ev[2], 0));
  return 0;
}

namespace {
struct BeamWs {
  __half *a1, *x, *h2n;
  float *c1, *c1n, *c2, *c2n, *ms, *tv;
  int *ti, *nbeam, *done, *n_done;
  unsigned int* row_thr;
  xd::BeamMeta meta[2];
  size_t bytes;
};
BeamWs carve_beam(char* base, const Cfg& g, int B, int bw, int D1) {
  BeamWs w{};
  size_t off = 0;
  auto take = [&](size_t n) { char* p = base ? base + off : nullptr; off += (size_t)rup((long long)n, 256); return p; };
  const size_t S = (size_t)B * bw, HP = g.HP;
  const size_t n_part = ceil_div(g.V, 128);
  w.a1 = (__half*)take(2 * 2 * S * HP);
  w.x = (__half*)take(2 * 2 * S * 2 * HP);
  w.h2n = (__half*)take(2 * 2 * S * HP);
  w.c1 = (float*)take(4 * S * HP); w.c1n = (float*)take(4 * S * HP);
  w.c2 = (float*)take(4 * S * HP); w.c2n = (float*)take(4 * S * HP);
  const size_t n_part2 = 2 * (size_t)ceil_div(g.V, 128);
  w.ms = (float*)take(4 * 2 * S * n_part2);
  w.tv = (float*)take(4 * S * n_part2 * KC);
  w.ti = (int*)take(4 * S * n_part2 * KC);
  w.nbeam = (int*)take(4 * (size_t)B); w.done = (int*)take(4 * (size_t)B);
  w.n_done = (int*)take(64);
  w.row_thr = (unsigned int*)take(4 * S * KC);
  for (int i = 0; i < 2; ++i) {
    w.meta[i].key = (float*)take(4 * S); w.meta[i].tok = (int*)take(4 * S); w.meta[i].len = (int*)take(4 * S);
    w.meta[i].fin = (int*)take(4 * S); w.meta[i].hist = (int*)take(4 * S * D1);
  }
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" int64_t s2vt_xdec_beam_ws_bytes(s2vt_xdec_cfg cfg, int B, int beam_width, int max_depth) {
  const Cfg g = make_cfg(cfg);
  return (int64_t)(rup((long long)carve_greedy(nullptr, g, B, g.L).bytes, 256) + carve_beam(nullptr, g, B, beam_width, max_depth + 1).bytes + 512);
}

// S2VT.forward(mode='beam_search'), S2VTModel.py:56-61,149-240, lock-step over videos x beams.
//   len_pen [max_depth+2] f32: float(pow(float(n), 0.7)) (BeamSearchNode.eval), computed by the host in double precision
//   out_tokens [B, max_depth+1] i64 (-1 padded, <sos> first), out_len [B] i32
//   check_every > 0: every that many depths the host reads the number of finished videos and stops early when all are
//   (one stream synchronisation per check); 0 = always run max_depth depths.  host_flag: pinned int32 for that read.
extern "C" int s2vt_xdec_beam(void* stream, s2vt_xdec_cfg cfg, const void* wbuf, int B, const float* feats, int beam_width, int max_depth,
                              int topk, const float* len_pen, int64_t* out_tokens, int32_t* out_len, void* ws, int check_every,
                              int32_t* host_flag) {
  cudaStream_t st = (cudaStream_t)stream, sd;
  const Cfg g = make_cfg(cfg);
  S2VT_REQUIRE(B > 0 && wbuf && feats && len_pen && out_tokens && out_len && ws, "s2vt_xdec_beam: bad arguments");
  S2VT_REQUIRE(beam_width >= 1 && beam_width <= KC, "s2vt_xdec_beam: beam_width must be in [1,%d] on the tensor-core path", KC);
  S2VT_REQUIRE(topk >= 1 && topk <= g.V, "s2vt_xdec_beam: topk must be in [1,V] (the reference's topk(20) needs V >= 20)");
  S2VT_REQUIRE(max_depth >= 1, "s2vt_xdec_beam: max_depth must be >= 1");
  S2VT_REQUIRE(check_every <= 0 || host_flag, "s2vt_xdec_beam: check_every needs a pinned host flag");
  int rc = side_stream(&sd);
  if (rc) return rc;
  const Weights W = carve_weights((char*)const_cast<void*>(wbuf), g);
  const int L = g.L, HP = g.HP, G = 4 * g.HP, D1 = max_depth + 1;
  const GreedyWs e = carve_greedy((char*)ws, g, B, L);
  const BeamWs w = carve_beam((char*)ws + rup((long long)e.bytes, 256), g, B, beam_width, D1);
  const int S = B * beam_width;
  const long long BH = (long long)B * HP, SH = (long long)S * HP;
  const int n_part = ceil_div(g.V, 128);
  const int kc = topk < beam_width ? topk : beam_width;
  X_TRY(encode(st, sd, g, W, e, B, feats, L));
  S2VT_CHECK_CUDA(cudaEventRecord(g_ev[2], sd));
  S2VT_CHECK_CUDA(cudaStreamWaitEvent(st, g_ev[2], 0));
  // encode state: h1 = last slot of the vid_rnn history, h2 / c in the ping-pong buffers written by step L-1
  beam_init_kernel<<<ceil_div(B, 128), 128, 0, st>>>(B, beam_width, D1, g.sos, w.meta[0], w.nbeam, w.done, out_tokens, out_len);
  S2VT_CHECK_LAUNCH();
  S2VT_CHECK_CUDA(cudaMemsetAsync(w.n_done, 0, sizeof(int), st));
  S2VT_CHECK_CUDA(cudaMemsetAsync(w.row_thr, 0, 4 * (size_t)S * KC, st));
  beam_state_init_kernel<<<S, 128, 0, st>>>(B, beam_width, HP, e.o1p + (long long)L * BH, (long long)(L + 1) * BH, e.c1 + (long long)((L - 1) & 1) * BH,
                                            e.h2p + (long long)(L & 1) * BH, 2 * BH, e.c2 + (long long)((L - 1) & 1) * BH,
                                            w.a1, SH, w.x, 2 * SH, w.c1, w.c2);
  S2VT_CHECK_LAUNCH();
  const Planes WH1{W.hh1, HP, (long long)G * HP, W.inv + INV_HH1}, WC2{W.cat2, 2 * HP, (long long)G * 2 * HP, W.inv + INV_CAT2};
  const Planes WO{W.out, HP, (long long)g.V * HP, W.inv + INV_OUT};
  const int n_part2 = 2 * ceil_div(g.V, 128);                 // candidate lists per row: one per tile and epilogue warp group
  int chunk = 0;
  // S2VT_XDEC_EVENTS=1 (debug): CUDA events around the four kernels of depth 10, printed after the call
  static int want_events = -1;
  if (want_events < 0) { const char* e = getenv("S2VT_XDEC_EVENTS"); want_events = (e && e[0] == '1') ? 1 : 0; }
  cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (want_events) for (int i = 0; i < 5; ++i) cudaEventCreate(&tev[i]);
#define X_STAMP(i) do { if (want_events && depth == 10) cudaEventRecord(tev[i], st); } while (0)
  for (int depth = 0; depth < max_depth; ++depth) {
    const xd::BeamMeta& mo = w.meta[depth & 1];
    const xd::BeamMeta& mn = w.meta[(depth + 1) & 1];
    X_STAMP(0);
    {   // vid_rnn step on the zero pad (S2VTModel.py:208-210): h1' -> x[:, :HP]
      StepArgs a{};
      a.bias = W.b1; a.c_in = w.c1; a.c_out = w.c1n; a.hp = w.x; a.hp_ld = 2 * HP; a.hp_plane = 2 * SH;
      Planes A{w.a1, HP, SH, W.inv + INV_H};
      X_TRY(x_lstm_step(st, S, HP, HP, A, WH1, a));
    }
    X_STAMP(1);
    {   // word_rnn step on [embed(word) | vid_out] (S2VTModel.py:207,211-212): K runs over [h1' | h2]
      StepArgs a{};
      a.bias = W.b2; a.gtab = W.ew; a.gtab_ld = G; a.gidx = mo.tok;
      a.c_in = w.c2; a.c_out = w.c2n; a.hp = w.h2n; a.hp_ld = HP; a.hp_plane = SH;
      Planes A{w.x, 2 * HP, 2 * SH, W.inv + INV_H};
      X_TRY(x_lstm_step(st, S, HP, 2 * HP, A, WC2, a));
    }
    X_STAMP(2);
    {   // out_linear + log_softmax + top-k partials (S2VTModel.py:213-216)
      XParams p{};
      p.bias = W.bout; p.o_ms = w.ms; p.o_val = w.tv; p.o_idx = w.ti;
      p.row_thr = w.row_thr; p.kc = kc; p.m_fastest = 1;
      Planes A{w.h2n, HP, SH, W.inv + INV_H};
      X_TRY((launch_x<128, EPI_BEAM>(st, S, g.V, HP, A, WO, p)));
    }
    X_STAMP(3);
    beam_finish_kernel<<<B, 32 * beam_width, 0, st>>>(B, beam_width, topk, kc, D1, g.eos, HP, n_part2, len_pen, mo, mn, w.ms, w.tv, w.ti,
                                                                  w.nbeam, w.done, w.n_done, out_tokens, out_len, w.a1, SH, w.x, 2 * SH, w.h2n, SH,
                                                                  w.c1, w.c1n, w.c2, w.c2n, w.row_thr);
    S2VT_CHECK_LAUNCH();
    X_STAMP(4);
    if (check_every > 0 && (depth + 1) % check_every == 0 && depth + 1 < max_depth) {
      // Early exit without draining the stream: the count of finished videos after this chunk of depths goes to a pinned slot, and the
      // host waits for the PREVIOUS chunk's count before it enqueues the next one -- one chunk is always queued behind the running
      // one.  Depths run past the point where every video has finished change nothing (finished videos are frozen).
      const int slot = chunk & 3;
      S2VT_CHECK_CUDA(cudaMemcpyAsync(host_flag + slot, w.n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
      S2VT_CHECK_CUDA(cudaEventRecord(g_bev[slot], st));
      if (chunk >= 1) {
        const int prev = (chunk - 1) & 3;
        S2VT_CHECK_CUDA(cudaEventSynchronize(g_bev[prev]));
        if (host_flag[prev] >= B) break;
      }
      ++chunk;
    }
  }
  if (want_events && max_depth > 10) {
    cudaStreamSynchronize(st);
    float ms[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], tev[i], tev[i + 1]);
    fprintf(stderr, "[xdec beam depth 10] vid_rnn step %.1f us, word_rnn step %.1f us, vocab + top-k %.1f us, queue/merge/re-order %.1f us\n",
            1e3f * ms[0], 1e3f * ms[1], 1e3f * ms[2], 1e3f * ms[3]);
    int st4[8];
    if (cudaMemcpy(st4, w.n_done, sizeof(st4), cudaMemcpyDeviceToHost) == cudaSuccess)
      fprintf(stderr, "[xdec beam, last depth, one CTA of the bookkeeping kernel] merge %.2f us, queue step %.2f us, histories + state copies %.2f us\n",
              1e-3f * (st4[3] - st4[2]), 1e-3f * (st4[4] - st4[3]), 1e-3f * (st4[5] - st4[4]));
    for (int i = 0; i < 5; ++i) cudaEventDestroy(tev[i]);
  }
#undef X_STAMP
  return 0;
}
#undef X_TRY
