// LSTM recurrence and BPTT for ANY hidden size (H % 8 == 0) on the tensor cores, one kernel launch per time step.
//
// The persistent cluster kernels (lstm_bf16_sm100.cu, lstm_bwd_bf16_sm100.cu) keep W_hh on chip, which caps them at H <= 512
// (a 128 KB slice per CTA, 16 CTAs per cluster).  The paper sizing of the reference (S2VTModel.py:11 defaults dim_hid = 500;
// BASELINE configs[3]: H = 1000, batch 256) does not fit that scheme, so its recurrences run here: per step ONE tcgen05 GEMM
// (bf16 operands, fp32 accumulation in TMEM, TMA-fed 6-stage ring, 128 x 64 tiles so that a batch of 256 puts the step on
// 126 SMs) whose epilogue is the LSTM cell (forward) or the gate-gradient arithmetic (backward).  At batch 256 a step's launch
// and pipeline fill are shared by 256 videos; programmatic dependent launch overlaps the next step's set-up and weight-tile
// requests with the running step.
//
//   forward   gates[B, 4H] = h_{t-1}[B, H] . W_hh_il[4H, H]^T + pre_t      (W_hh rows interleaved 4u+g: a thread owns whole units)
//             -> i,f,g,o -> c_t, h_t (bf16, the next step's operand); gates and c_t are stashed for BPTT
//   backward  dh[B, H] = dgates_{t+1}[B, 4H] . W_hh[4H, H] + dout_t        (W_hh^T [H, 4H] as the K-major B operand)
//             -> dgates_t (bf16, natural gate order g*H+u: the layout of the time-batched weight-gradient products), dc_t
// replaces: nn.LSTM (S2VTModel.py:19-22,67,77) and its autograd (train.py:124) for shapes outside the cluster kernels' range.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "sm100_err.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace s2vt {

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer);

namespace st {

constexpr int BM = 128, BN = 64, BK = 64, STAGES = 6;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM = STAGES * STAGE_BYTES + 1024;
enum { EPI_FWD = 0, EPI_BWD = 1 };

struct StepParams {
  int M, N, H, num_kb;
  // forward (N = 4H, column 4u+g)
  const float* pre; long long pre_ld; const float* bias;
  const float* c_in; float* c_out;
  __nv_bfloat16* h_out; __nv_bfloat16* gates_out;
  // backward (N = H)
  const float* dout; const __nv_bfloat16* gates; const float* c_t; const float* c_prev;
  float* dc; __nv_bfloat16* dgates;
  // backward split-K: gridDim.z CTAs share a tile, each over kb_per_split K blocks; partial sums meet in `part` [z][M][N] and the
  // last CTA of a tile to arrive (tile_count, self-resetting) adds them up and runs the epilogue
  int kb_per_split; float* part; unsigned int* tile_count;
};

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int EPI>
__global__ void __launch_bounds__(256, 1)
lstm_step_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const StepParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb0 = (EPI == EPI_BWD && gridDim.z > 1) ? (int)blockIdx.z * p.kb_per_split : 0;
  const int kb1 = (EPI == EPI_BWD && gridDim.z > 1) ? min(p.num_kb, kb0 + p.kb_per_split) : p.num_kb;
  const int nkb = kb1 - kb0;                             // K blocks of this CTA
  __shared__ int s_last;

  if (warp_idx == 0 && ptx::elect_one() && p.num_kb > 0) {
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
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  pdl_launch_dependents();

  if (warp_idx == 0) {
    if (ptx::elect_one()) {
      // weight tiles first (they do not depend on the previous step), the state operand after the dependency resolves
      const int npre = nkb < STAGES ? nkb : STAGES;
      for (int i = 0; i < npre; ++i) {
        const uint32_t fb = ptx::smem_u32(&full_bar[i]);
        ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        ptx::tma_load_2d(smem_base + i * STAGE_BYTES + A_BYTES, &tmB, fb, (kb0 + i) * BK, n0);
      }
      pdl_wait();
      for (int i = 0; i < npre; ++i)
        ptx::tma_load_2d(smem_base + i * STAGE_BYTES, &tmA, ptx::smem_u32(&full_bar[i]), (kb0 + i) * BK, m0);
      for (int i = npre; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1)) { atomicExch(&g_sm100_error, 21); break; }
        const uint32_t fb = ptx::smem_u32(&full_bar[s]);
        const uint32_t sA = smem_base + s * STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
        ptx::tma_load_2d(sA, &tmA, fb, (kb0 + i) * BK, m0);
        ptx::tma_load_2d(sA + A_BYTES, &tmB, fb, (kb0 + i) * BK, n0);
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        if (!ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph)) { atomicExch(&g_sm100_error, 22); break; }
        ptx::tc_fence_after();
        const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          ptx::mma_bf16_ss(tmem, ptx::make_smem_desc_sw128(sA + k * 32, 16, 1024), ptx::make_smem_desc_sw128(sB + k * 32, 16, 1024), idesc,
                           (i > 0 || k > 0) ? 1u : 0u);
        ptx::mma_commit(ptx::smem_u32(&empty_bar[s]));
      }
      ptx::mma_commit(ptx::smem_u32(&tmem_full_bar));
    }
    __syncwarp();
  }

  // ===================== epilogue: eight warps, one 32-column chunk each (warps q and q+4 share the rows of TMEM quarter q)
  {
    const int q = warp_idx & 3, half = warp_idx >> 2;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const int n = n0 + half * 32;
    const bool col_ok = n < p.N;                         // N is a multiple of 32 (H % 8 == 0 forward, units in 32s masked below)
    pdl_wait();
    if (EPI == EPI_FWD) {
      const int H = p.H, u0 = n >> 2;
      float pin[32], cp[8];
      if (row_ok && col_ok) {
        const float4* src = reinterpret_cast<const float4*>(p.pre ? p.pre + (long long)m * p.pre_ld + n : p.bias + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float4 t = __ldg(src + j); pin[4 * j] = t.x; pin[4 * j + 1] = t.y; pin[4 * j + 2] = t.z; pin[4 * j + 3] = t.w; }
        if (p.c_in) {
          const float4* c4 = reinterpret_cast<const float4*>(p.c_in + (long long)m * H + u0);
          const float4 t0 = c4[0], t1 = c4[1];
          cp[0] = t0.x; cp[1] = t0.y; cp[2] = t0.z; cp[3] = t0.w; cp[4] = t1.x; cp[5] = t1.y; cp[6] = t1.z; cp[7] = t1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) cp[j] = 0.f;
        }
      }
      bool ok = ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar), 0);
      if (!ok) atomicExch(&g_sm100_error, 23);
      ptx::tc_fence_after();
      if (col_ok) {
        uint32_t r[32];
        if (p.num_kb > 0) {
          ptx::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32), r);
          ptx::tc_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (row_ok) {
          float cn[8], h[8];
          __align__(16) __nv_bfloat16 gs[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ig = sigmoidf_exact(__uint_as_float(r[4 * j]) + pin[4 * j]), fg = sigmoidf_exact(__uint_as_float(r[4 * j + 1]) + pin[4 * j + 1]);
            const float gg = tanhf(__uint_as_float(r[4 * j + 2]) + pin[4 * j + 2]), og = sigmoidf_exact(__uint_as_float(r[4 * j + 3]) + pin[4 * j + 3]);
            cn[j] = fg * cp[j] + ig * gg;
            h[j] = og * tanhf(cn[j]);
            gs[4 * j] = __float2bfloat16(ig); gs[4 * j + 1] = __float2bfloat16(fg); gs[4 * j + 2] = __float2bfloat16(gg); gs[4 * j + 3] = __float2bfloat16(og);
          }
          float4* cd = reinterpret_cast<float4*>(p.c_out + (long long)m * H + u0);
          cd[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
          cd[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
          __align__(16) __nv_bfloat16 hb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) hb[j] = __float2bfloat16(h[j]);
          *reinterpret_cast<uint4*>(p.h_out + (long long)m * H + u0) = *reinterpret_cast<const uint4*>(hb);
          if (p.gates_out) {
            uint4* gd = reinterpret_cast<uint4*>(p.gates_out + (long long)m * 4 * H + n);
            const uint4* gsrc = reinterpret_cast<const uint4*>(gs);
            gd[0] = gsrc[0]; gd[1] = gsrc[1]; gd[2] = gsrc[2]; gd[3] = gsrc[3];
          }
        }
      }
    } else {   // EPI_BWD: columns are hidden units u = n .. n+31 (masked against H in groups of 8: H % 8 == 0)
      const int H = p.H;
      bool ok = ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar), 0);
      if (!ok) atomicExch(&g_sm100_error, 24);
      ptx::tc_fence_after();
      uint32_t r[32];
      if (nkb > 0) {
        ptx::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32), r);
        ptx::tc_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      bool mine = true;                                    // does this CTA run the epilogue?
      if (gridDim.z > 1) {
        // split-K: park the partial sums, count arrivals; only the tile's last CTA goes on (the others' sums are in L2 by then)
        const long long slab = (long long)p.M * p.N;
        if (row_ok && col_ok) {
          float* dst = p.part + (long long)blockIdx.z * slab + (long long)m * p.N + n;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            if (n + 8 * g8 < H) {
              reinterpret_cast<float4*>(dst + 8 * g8)[0] = make_float4(__uint_as_float(r[8 * g8]), __uint_as_float(r[8 * g8 + 1]), __uint_as_float(r[8 * g8 + 2]), __uint_as_float(r[8 * g8 + 3]));
              reinterpret_cast<float4*>(dst + 8 * g8)[1] = make_float4(__uint_as_float(r[8 * g8 + 4]), __uint_as_float(r[8 * g8 + 5]), __uint_as_float(r[8 * g8 + 6]), __uint_as_float(r[8 * g8 + 7]));
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
          unsigned int* cnt = p.tile_count + blockIdx.y * gridDim.x + blockIdx.x;
          const unsigned int old = atomicAdd(cnt, 1u);
          s_last = (old == gridDim.z - 1);
          if (s_last) *cnt = 0u;                             // ready for the next step's launch
        }
        __syncthreads();
        mine = s_last != 0;
        if (mine && row_ok && col_ok) {
          __threadfence();
          for (int z = 0; z < (int)gridDim.z; ++z) {
            if (z == (int)blockIdx.z) continue;
            const float* src = p.part + (long long)z * slab + (long long)m * p.N + n;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8)
              if (n + 8 * g8 < H) {
                const float4 a = __ldcg(reinterpret_cast<const float4*>(src + 8 * g8)), b = __ldcg(reinterpret_cast<const float4*>(src + 8 * g8) + 1);
                r[8 * g8] = __float_as_uint(__uint_as_float(r[8 * g8]) + a.x); r[8 * g8 + 1] = __float_as_uint(__uint_as_float(r[8 * g8 + 1]) + a.y);
                r[8 * g8 + 2] = __float_as_uint(__uint_as_float(r[8 * g8 + 2]) + a.z); r[8 * g8 + 3] = __float_as_uint(__uint_as_float(r[8 * g8 + 3]) + a.w);
                r[8 * g8 + 4] = __float_as_uint(__uint_as_float(r[8 * g8 + 4]) + b.x); r[8 * g8 + 5] = __float_as_uint(__uint_as_float(r[8 * g8 + 5]) + b.y);
                r[8 * g8 + 6] = __float_as_uint(__uint_as_float(r[8 * g8 + 6]) + b.z); r[8 * g8 + 7] = __float_as_uint(__uint_as_float(r[8 * g8 + 7]) + b.w);
              }
          }
        }
      }
      if (col_ok && mine) {
        if (row_ok) {
#pragma unroll 4
          for (int g8 = 0; g8 < 4; ++g8) {                 // 8 units at a time (unrolled: the four groups' loads go out together)
            const int u = n + 8 * g8;
            if (u >= H) break;
            const long long o = (long long)m * H + u;
            float dh[8], ct[8], cpv[8], dcv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) dh[j] = __uint_as_float(r[8 * g8 + j]);
            if (p.dout) {
              const float4 a = *reinterpret_cast<const float4*>(p.dout + o), b = *reinterpret_cast<const float4*>(p.dout + o + 4);
              dh[0] += a.x; dh[1] += a.y; dh[2] += a.z; dh[3] += a.w; dh[4] += b.x; dh[5] += b.y; dh[6] += b.z; dh[7] += b.w;
            }
            {
              const float4 a = *reinterpret_cast<const float4*>(p.c_t + o), b = *reinterpret_cast<const float4*>(p.c_t + o + 4);
              ct[0] = a.x; ct[1] = a.y; ct[2] = a.z; ct[3] = a.w; ct[4] = b.x; ct[5] = b.y; ct[6] = b.z; ct[7] = b.w;
            }
            if (p.c_prev) {
              const float4 a = *reinterpret_cast<const float4*>(p.c_prev + o), b = *reinterpret_cast<const float4*>(p.c_prev + o + 4);
              cpv[0] = a.x; cpv[1] = a.y; cpv[2] = a.z; cpv[3] = a.w; cpv[4] = b.x; cpv[5] = b.y; cpv[6] = b.z; cpv[7] = b.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) cpv[j] = 0.f;
            }
            {
              const float4 a = *reinterpret_cast<const float4*>(p.dc + o), b = *reinterpret_cast<const float4*>(p.dc + o + 4);
              dcv[0] = a.x; dcv[1] = a.y; dcv[2] = a.z; dcv[3] = a.w; dcv[4] = b.x; dcv[5] = b.y; dcv[6] = b.z; dcv[7] = b.w;
            }
            // stashed gates of these 8 units: 32 bf16, interleaved (i,f,g,o per unit)
            __align__(16) __nv_bfloat16 gs[32];
            {
              const uint4* gsrc = reinterpret_cast<const uint4*>(p.gates + (long long)m * 4 * H + 4 * u);
              uint4* gd = reinterpret_cast<uint4*>(gs);
              gd[0] = gsrc[0]; gd[1] = gsrc[1]; gd[2] = gsrc[2]; gd[3] = gsrc[3];
            }
            __align__(16) __nv_bfloat16 di[8], df[8], dg[8], dO[8];
            float dcn[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float ig = __bfloat162float(gs[4 * j]), fg = __bfloat162float(gs[4 * j + 1]), gg = __bfloat162float(gs[4 * j + 2]),
                          og = __bfloat162float(gs[4 * j + 3]);
              const float tc = tanhf(ct[j]);
              const float d_o = dh[j] * tc;
              const float dcc = dcv[j] + dh[j] * og * (1.f - tc * tc);
              di[j] = __float2bfloat16(dcc * gg * ig * (1.f - ig));
              df[j] = __float2bfloat16(dcc * cpv[j] * fg * (1.f - fg));
              dg[j] = __float2bfloat16(dcc * ig * (1.f - gg * gg));
              dO[j] = __float2bfloat16(d_o * og * (1.f - og));
              dcn[j] = dcc * fg;
            }
            __nv_bfloat16* drow = p.dgates + (long long)m * 4 * H + u;        // natural gate order: column g*H + u
            *reinterpret_cast<uint4*>(drow) = *reinterpret_cast<const uint4*>(di);
            *reinterpret_cast<uint4*>(drow + H) = *reinterpret_cast<const uint4*>(df);
            *reinterpret_cast<uint4*>(drow + 2 * H) = *reinterpret_cast<const uint4*>(dg);
            *reinterpret_cast<uint4*>(drow + 3 * H) = *reinterpret_cast<const uint4*>(dO);
            float4* dd = reinterpret_cast<float4*>(p.dc + o);
            dd[0] = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
            dd[1] = make_float4(dcn[4], dcn[5], dcn[6], dcn[7]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) ptx::tmem_dealloc(tmem, BN);
}

template <int EPI>
static int launch_step(cudaStream_t stream, int M, int N, int K, const void* A, long long lda, const void* B, long long ldb, StepParams p,
                       int splits = 1) {
  CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmB, 0, sizeof(tmB));
  p.num_kb = (A && K > 0) ? (K + BK - 1) / BK : 0;
  if (p.num_kb > 0) {
    int rc = make_tmap_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
    if (rc) return rc;
  }
  p.M = M; p.N = N;
  if (splits > 1 && p.num_kb >= 2) {
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;      // no empty slice
  } else {
    splits = 1;
    p.kb_per_split = p.num_kb;
  }
  static bool attr_set = false;
  if (!attr_set) {
    S2VT_CHECK_CUDA(cudaFuncSetAttribute(lstm_step_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ceil_div(N, BN), ceil_div(M, BM), splits); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = SMEM; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  S2VT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_step_kernel<EPI>, tmA, tmB, p));
  S2VT_CHECK_LAUNCH();
  return 0;
}

}  // namespace st

int lstm_step_error_flag() { return read_sm100_error_flag(); }
int lstm_step_error_clear() { return clear_sm100_error_flag(); }

}  // namespace s2vt

using namespace s2vt;
using namespace s2vt::st;

// T forward steps.  Layouts (time-major, row = t*B + b):
//   pre      [n_pre, B, 4H] f32   input-side pre-activations incl. both biases, columns interleaved 4u+g; steps >= n_pre use bias_il
//   w_hh_il  [4H, H] bf16         W_hh with rows interleaved 4u+g
//   out      [T, B, H] bf16       h_t
//   gates    [T, B, 4H] bf16      post-activation i,f,g,o, interleaved (nullable: inference)
//   cells    [T, B, H] f32        c_t
// Zero initial state (the reference never passes one on this path).
extern "C" int s2vt_lstm_steps_fwd_bf16(void* stream, int T, int B, int H, int n_pre, const float* pre, const float* bias_il,
                                        const void* w_hh_il, void* out, void* gates, float* cells) {
  cudaStream_t s = (cudaStream_t)stream;
  S2VT_REQUIRE(T >= 0 && B > 0 && H > 0 && H % 8 == 0, "s2vt_lstm_steps_fwd_bf16: needs H %% 8 == 0 (H=%d)", H);
  S2VT_REQUIRE(w_hh_il && out && cells && (n_pre >= T || bias_il) && (n_pre <= 0 || pre), "s2vt_lstm_steps_fwd_bf16: null pointer");
  const long long BH = (long long)B * H;
  for (int t = 0; t < T; ++t) {
    StepParams p{};
    p.H = H;
    p.pre = t < n_pre ? pre + (long long)t * 4 * BH : nullptr; p.pre_ld = 4 * H; p.bias = bias_il;
    p.c_in = t > 0 ? cells + (long long)(t - 1) * BH : nullptr;
    p.c_out = cells + (long long)t * BH;
    p.h_out = (__nv_bfloat16*)out + (long long)t * BH;
    p.gates_out = gates ? (__nv_bfloat16*)gates + (long long)t * 4 * BH : nullptr;
    const void* A = t > 0 ? (const void*)((const __nv_bfloat16*)out + (long long)(t - 1) * BH) : nullptr;
    int rc = launch_step<EPI_FWD>(s, B, 4 * H, H, A, H, w_hh_il, H, p);
    if (rc) return rc;
  }
  return 0;
}

// T backward steps, t = T-1 .. 0.
//   dout    [T, B, H] f32         dL/dh_t from above (rows t < dout_t0 are never read; nullptr = none)
//   gates, cells                  the forward stash
//   w_hh_t  [H, 4H] bf16          W_hh^T (natural gate order along K)
//   dgates  [T, B, 4H] bf16       out: pre-activation gradients, natural gate order (column g*H + u)
//   ws      >= s2vt_lstm_steps_bwd_ws_bytes(B, H): running dL/dc, split-K partial sums and tile counters
// K = 4H runs over few output tiles (batch 256 x H 1000: 32), so the product is split over K until the step fills the machine.
static int bwd_splits(int B, int H) {
  const int tiles = ceil_div(H, BN) * ceil_div(B, BM);
  int sp = 128 / (tiles > 0 ? tiles : 1);
  const int num_kb = ceil_div(4 * H, BK);
  if (sp > 8) sp = 8;
  if (sp > num_kb / 2) sp = num_kb / 2;
  return sp < 1 ? 1 : sp;
}
extern "C" int64_t s2vt_lstm_steps_bwd_ws_bytes(int B, int H) {
  const int sp = bwd_splits(B, H);
  return (int64_t)sizeof(float) * B * H * (1 + sp) + 4 * (int64_t)ceil_div(H, BN) * ceil_div(B, BM) + 512;
}
extern "C" int s2vt_lstm_steps_bwd_bf16(void* stream, int T, int B, int H, int dout_t0, const float* dout, const void* gates,
                                        const float* cells, const void* w_hh_t, void* dgates, void* ws) {
  cudaStream_t s = (cudaStream_t)stream;
  S2VT_REQUIRE(T >= 0 && B > 0 && H > 0 && H % 8 == 0, "s2vt_lstm_steps_bwd_bf16: needs H %% 8 == 0 (H=%d)", H);
  S2VT_REQUIRE(gates && cells && w_hh_t && dgates && ws, "s2vt_lstm_steps_bwd_bf16: null pointer");
  const long long BH = (long long)B * H;
  const int sp = bwd_splits(B, H);
  float* dc_ws = (float*)ws;
  float* part = dc_ws + BH;
  unsigned int* tile_count = (unsigned int*)(((uintptr_t)(part + (long long)sp * BH) + 255) & ~(uintptr_t)255);
  S2VT_CHECK_CUDA(cudaMemsetAsync(dc_ws, 0, sizeof(float) * BH, s));
  S2VT_CHECK_CUDA(cudaMemsetAsync(tile_count, 0, 4 * (size_t)ceil_div(H, BN) * ceil_div(B, BM), s));
  for (int t = T - 1; t >= 0; --t) {
    StepParams p{};
    p.H = H;
    p.dout = (dout && t >= dout_t0) ? dout + (long long)t * BH : nullptr;
    p.gates = (const __nv_bfloat16*)gates + (long long)t * 4 * BH;
    p.c_t = cells + (long long)t * BH;
    p.c_prev = t > 0 ? cells + (long long)(t - 1) * BH : nullptr;
    p.dc = dc_ws;
    p.dgates = (__nv_bfloat16*)dgates + (long long)t * 4 * BH;
    p.part = part; p.tile_count = tile_count;
    const void* A = t < T - 1 ? (const void*)((const __nv_bfloat16*)dgates + (long long)(t + 1) * 4 * BH) : nullptr;
    int rc = launch_step<EPI_BWD>(s, B, H, 4 * H, A, 4 * H, w_hh_t, 4 * H, p, sp);
    if (rc) return rc;
  }
  return 0;
}
