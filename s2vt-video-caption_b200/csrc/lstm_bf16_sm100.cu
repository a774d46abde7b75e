// Persistent LSTM recurrence on Blackwell tensor cores: one thread-block CLUSTER steps a batch tile through all T
// timesteps without leaving the chip.
//
//   cluster = H/32 CTAs (16 for H = 512); CTA c owns hidden units [32c, 32c+32) = 128 gate rows (i,f,g,o x 32)
//   W_hh slice  [128 rows x H] bf16 : loaded ONCE by TMA (128B swizzle), resident in shared memory for all T steps
//   h_{t-1}^T   [NB batch x H] bf16 : double-buffered in shared memory in the canonical no-swizzle K-major layout
//   per step:   gates[128 x NB] = W_slice . h_{t-1}^T      tcgen05.mma M=128 N=NB K=16, fp32 accumulator in TMEM
//               epilogue warps: tcgen05.ld -> + input-side pre-activation (prefetched from HBM/L2) -> sigmoid/tanh
//               -> gate exchange through 8 KB of smem -> c,h update (c stays in registers for the whole sequence)
//               -> the CTA's 32 x NB slice of h_t is pushed to ALL cluster peers with cp.async.bulk (DSMEM),
//                  completing on each peer's mbarrier: the only inter-CTA synchronisation per step.
//
// There is no grid-wide barrier and no global-memory round trip on the recurrent critical path; the stash for BPTT
// (post-activation gates bf16, cell state fp32, h bf16) streams to HBM off the critical path.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "sm100_err.cuh"

namespace s2vt {

int lstm_bf16_error_flag() { return read_sm100_error_flag(); }
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
// local smem -> peer smem bulk copy; completes (complete_tx) on the PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_local, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_local), "r"(bytes), "r"(bar_cluster) : "memory");
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
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // tanh(x) = 2*sigmoid(2x) - 1, with the exponent clamped so that __expf never overflows
  const float e = __expf(-2.0f * fminf(fmaxf(x, -15.0f), 15.0f));
  return __fdividef(1.0f - e, 1.0f + e);
}
}  // namespace ptx

struct LstmFwdParams {
  int T, B, H, n_pre;
  const float* pre;        // [n_pre, B, 4H]
  const float* bias;       // [4H]
  const float* h0;         // [B,H] or null
  const float* c0;         // [B,H] or null
  __nv_bfloat16* out;      // [T,B,H]
  __nv_bfloat16* gates;    // [T,B,4H] or null
  float* cells;            // [T,B,H] or null
  float* hT;               // [B,H] or null
  float* cT;               // [B,H] or null
};

template <int NB>
__global__ void __launch_bounds__(160, 1)
lstm_fwd_cluster_kernel(const __grid_constant__ CUtensorMap tmW, const LstmFwdParams p) {
  static_assert(NB % 16 == 0 && NB <= 64, "tcgen05 M=128 needs N % 16 == 0");
  constexpr int COLS_PER_THREAD = NB / 4;          // phase-2 columns per thread
  constexpr uint32_t LBO_H = (NB / 8) * 128;       // K-direction stride between 8x16B core matrices of h^T
  constexpr uint32_t SLICE_BYTES = 4 * LBO_H;      // one CTA's 32 hidden units x NB batch, bf16
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full, h_full[2], mma_done;
  __shared__ uint32_t tmem_slot;

  const int H = p.H, KC = H / 64, CS = H / 32;
  const uint32_t W_BYTES = 128u * (uint32_t)H * 2u, HBUF_BYTES = (uint32_t)NB * (uint32_t)H * 2u;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base, sH0 = sW + W_BYTES, sStage0 = sH0 + 2 * HBUF_BYTES, sGu = sStage0 + 2 * SLICE_BYTES;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));            // generic pointer to the aligned base
  uint8_t* gH0 = gen + W_BYTES;
  uint8_t* gStage0 = gH0 + 2 * HBUF_BYTES;
  float* sG = reinterpret_cast<float*>(gStage0 + 2 * SLICE_BYTES);      // [4 gates][NB][32 units]
  (void)sGu;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t c = ptx::cluster_ctarank();                              // hidden-unit slice of this CTA
  const int b0 = (blockIdx.x / CS) * NB;                                  // batch tile of this cluster
  const int T = p.T;

  if (warp == 4 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(ptx::smem_u32(&w_full), 1);
    ptx::mbar_init(ptx::smem_u32(&h_full[0]), 1);
    ptx::mbar_init(ptx::smem_u32(&h_full[1]), 1);
    ptx::mbar_init(ptx::smem_u32(&mma_done), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 32);
    ptx::tmem_relinquish();
  }
  // initial h^T buffer (step 0 input): zeros or h0 in the canonical layout; c state into registers (below)
  if (warp < 4) {
    for (int idx = threadIdx.x; idx < NB * H / 8; idx += 128) {           // one 16-byte chunk (8 k-elements) per iteration
      const int kblk = idx / NB, b = idx % NB;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (p.h0 && b0 + b < p.B) {
        const float* src = p.h0 + (long long)(b0 + b) * H + kblk * 8;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(src[0], src[1]), t1 = __floats2bfloat162_rn(src[2], src[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(src[4], src[5]), t3 = __floats2bfloat162_rn(src[6], src[7]);
        v.x = *reinterpret_cast<uint32_t*>(&t0); v.y = *reinterpret_cast<uint32_t*>(&t1);
        v.z = *reinterpret_cast<uint32_t*>(&t2); v.w = *reinterpret_cast<uint32_t*>(&t3);
      }
      *reinterpret_cast<uint4*>(gH0 + (size_t)(kblk * (NB / 8) + b / 8) * 128 + (b % 8) * 16) = v;
    }
    ptx::fence_proxy_async();                                             // generic writes -> visible to the MMA (async proxy)
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::cluster_arrive();                                                  // every peer's barriers are initialised before
  ptx::cluster_wait();                                                    // anyone signals them remotely
  const uint32_t tmem = tmem_slot;

  if (warp == 4) {
    // ===================== control thread: weight load, per-step MMA issue =====================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&w_full), W_BYTES);
      for (int kc = 0; kc < KC; ++kc)
        for (int g = 0; g < 4; ++g)                                       // rows [g*H + 32c, +32) -> tile rows [32g, 32g+32)
          ptx::tma_load_2d(sW + kc * 16384 + g * 4096, &tmW, ptx::smem_u32(&w_full), kc * 64, g * H + 32 * (int)c);
      bool ok = ptx::mbar_wait(ptx::smem_u32(&w_full), 0);
      if (!ok) atomicExch(&g_sm100_error, 11);
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NB, 0, 0);
      uint32_t ph[2] = {0, 0};
      const bool have_h0 = p.h0 != nullptr;
      for (int t = 0; t < T && ok; ++t) {
        const int pb = t & 1;
        if (t + 1 < T) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&h_full[pb ^ 1]), (uint32_t)CS * SLICE_BYTES);   // h_t lands here
        if (t > 0) {
          ok = ptx::mbar_wait(ptx::smem_u32(&h_full[pb]), ph[pb]);
          ph[pb] ^= 1;
          if (!ok) { atomicExch(&g_sm100_error, 12); break; }
        }
        if (t > 0 || have_h0) {
          ptx::tc_fence_after();
          const uint32_t sHp = sH0 + pb * HBUF_BYTES;
          for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t da = ptx::make_smem_desc(sW + kc * 16384 + k4 * 32, 16, 1024, 2);
              const uint64_t db = ptx::make_smem_desc(sHp + (kc * 8 + k4 * 2) * LBO_H, LBO_H, 128, 0);
              ptx::mma_bf16_ss(tmem, da, db, idesc, (kc | k4) != 0 ? 1u : 0u);
            }
          }
          ptx::mma_commit(ptx::smem_u32(&mma_done));
        } else {
          ptx::mbar_arrive(ptx::smem_u32(&mma_done));                     // h_{-1} = 0: nothing to multiply
        }
      }
    }
  } else {
    // ===================== epilogue warps 0..3 =====================
    const int g = warp;                       // phase 1: gate row block of this warp (i,f,g,o) == TMEM lane quarter
    const int u = lane;                       // hidden unit within the CTA slice
    const int unit = 32 * (int)c + u;
    const int q = warp;                       // phase 2: column group
    float creg[COLS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < COLS_PER_THREAD; ++j) {
      const int b = b0 + q * COLS_PER_THREAD + j;
      creg[j] = (p.c0 && b < p.B) ? p.c0[(long long)b * H + unit] : 0.f;
    }
    const float bias_g = p.bias[g * H + unit];
    float pre_cur[NB];
    auto load_pre = [&](int t, float (&dst)[NB]) {
      if (t < p.n_pre) {
        const float* src = p.pre + ((long long)t * p.B + b0) * 4 * H + g * H + unit;
#pragma unroll
        for (int j = 0; j < NB; ++j) dst[j] = (b0 + j < p.B) ? __ldg(src + (long long)j * 4 * H) : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) dst[j] = bias_g;
      }
    };
    load_pre(0, pre_cur);
    const bool have_h0 = p.h0 != nullptr;
    bool ok = true;
    for (int t = 0; t < T; ++t) {
      // ---- phase 1: accumulator + pre-activation -> activation -> smem
      ok = ok && ptx::mbar_wait(ptx::smem_u32(&mma_done), (uint32_t)(t & 1));
      if (!ok) { atomicExch(&g_sm100_error, 13); break; }
      float x[NB];
      if (t > 0 || have_h0) {
        ptx::tc_fence_after();
        uint32_t r[16];
#pragma unroll
        for (int cc = 0; cc < NB / 16; ++cc) {
          ptx::tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + cc * 16, r);
          ptx::tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) x[cc * 16 + j] = __uint_as_float(r[j]) + pre_cur[cc * 16 + j];
        }
        ptx::tc_fence_before();
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) x[j] = pre_cur[j];
      }
      if (t + 1 < T) load_pre(t + 1, pre_cur);                            // prefetch: latency hides behind the rest of the step
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float a = (g == 2) ? ptx::fast_tanh(x[j]) : ptx::fast_sigmoid(x[j]);
        sG[(g * NB + j) * 32 + u] = a;
        x[j] = a;
      }
      if (p.gates) {
        __nv_bfloat16* gdst = p.gates + ((long long)t * p.B + b0) * 4 * H + g * H + unit;
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (b0 + j < p.B) gdst[(long long)j * 4 * H] = __float2bfloat16(x[j]);
      }
      ptx::named_bar_sync(1, 128);
      // ---- phase 2: cell / hidden update for (unit u, columns q*CPT .. +CPT)
      uint8_t* stage = gStage0 + (t & 1) * SLICE_BYTES;
#pragma unroll
      for (int j = 0; j < COLS_PER_THREAD; ++j) {
        const int col = q * COLS_PER_THREAD + j;
        const float gi = sG[(0 * NB + col) * 32 + u], gf = sG[(1 * NB + col) * 32 + u];
        const float gg = sG[(2 * NB + col) * 32 + u], go = sG[(3 * NB + col) * 32 + u];
        const float cn = gf * creg[j] + gi * gg;
        creg[j] = cn;
        const float h = go * ptx::fast_tanh(cn);
        const __nv_bfloat16 hb = __float2bfloat16(h);
        *reinterpret_cast<__nv_bfloat16*>(stage + ((u / 8) * (NB / 8) + col / 8) * 128 + (col % 8) * 16 + (u % 8) * 2) = hb;
        const int b = b0 + col;
        if (b < p.B) {
          const long long o = ((long long)t * p.B + b) * H + unit;
          p.out[o] = hb;
          if (p.cells) p.cells[o] = cn;
          if (t == T - 1) {
            if (p.hT) p.hT[(long long)b * H + unit] = h;
            if (p.cT) p.cT[(long long)b * H + unit] = cn;
          }
        }
      }
      if (t + 1 < T) {
        ptx::fence_proxy_async();                                         // staging writes -> visible to the bulk-copy engine
        ptx::named_bar_sync(1, 128);
        if (warp == 0 && lane < CS) {
          const uint32_t dst = ptx::mapa(sH0 + ((t + 1) & 1) * HBUF_BYTES + c * SLICE_BYTES, (uint32_t)lane);
          const uint32_t bar = ptx::mapa(ptx::smem_u32(&h_full[(t + 1) & 1]), (uint32_t)lane);
          ptx::bulk_copy_to_peer(dst, sStage0 + (t & 1) * SLICE_BYTES, SLICE_BYTES, bar);
        }
      }
    }
  }
  // no CTA may exit while peers can still write into its shared memory
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 0) ptx::tmem_dealloc(tmem, 32);
}

template <int NB>
static int launch_lstm_fwd(cudaStream_t st, const CUtensorMap& tmW, const LstmFwdParams& p) {
  const int H = p.H, CS = H / 32;
  const size_t smem = 1024 + (size_t)128 * H * 2 + 2 * (size_t)NB * H * 2 + 2 * (size_t)(4 * (NB / 8) * 128) + (size_t)4 * NB * 32 * 4;
  auto kern = lstm_fwd_cluster_kernel<NB>;
  S2VT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) S2VT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CS * ceil_div(p.B, NB));
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  S2VT_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  S2VT_REQUIRE(max_clusters >= 1, "s2vt_lstm_fwd_bf16: a cluster of %d CTAs with %zu B of shared memory cannot be scheduled on this device", CS, smem);
  S2VT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmW, p));
  count_launch();
  return 0;
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int s2vt_lstm_fwd_bf16(void* stream, int T, int B, int H, int n_pre,
                                  const float* pre, const float* bias_sum, const void* w_hh_bf16,
                                  const float* h0, const float* c0,
                                  void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT) {
  S2VT_REQUIRE(T >= 1 && B >= 1, "s2vt_lstm_fwd_bf16: bad dims");
  S2VT_REQUIRE(H % 64 == 0 && H >= 64 && H <= 512, "s2vt_lstm_fwd_bf16: the cluster-resident kernel needs H %% 64 == 0 and 64 <= H <= 512 (got %d)", H);
  S2VT_REQUIRE(bias_sum && w_hh_bf16 && out_bf16, "s2vt_lstm_fwd_bf16: null pointer");
  S2VT_REQUIRE(n_pre <= 0 || pre, "s2vt_lstm_fwd_bf16: pre is null but n_pre > 0");
  S2VT_REQUIRE((h0 == nullptr) == (c0 == nullptr), "s2vt_lstm_fwd_bf16: h0 and c0 must be given together");
  CUtensorMap tmW;
  int rc = make_tmap_bf16(&tmW, w_hh_bf16, (uint64_t)H, (uint64_t)4 * H, (uint64_t)H, 64, 32);
  if (rc) return rc;
  LstmFwdParams p{};
  p.T = T; p.B = B; p.H = H; p.n_pre = n_pre < 0 ? 0 : n_pre;
  p.pre = pre; p.bias = bias_sum; p.h0 = h0; p.c0 = c0;
  p.out = (__nv_bfloat16*)out_bf16; p.gates = (__nv_bfloat16*)gates_bf16; p.cells = cells; p.hT = hT; p.cT = cT;
  return launch_lstm_fwd<16>((cudaStream_t)stream, tmW, p);
}
