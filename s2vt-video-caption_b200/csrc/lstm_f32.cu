// Exact fp32 LSTM recurrence (forward + BPTT) for one nn.LSTM layer.
// Per step: a split-K fp32 GEMM of the recurrent product into partial sums, then one fused pointwise
// kernel that sums the partials, adds the input-side pre-activation, applies the gates and updates
// (h, c) -- and, in backward, turns dL/dh_t into the pre-activation gradients.  Launch-per-step keeps
// this path free of inter-CTA spin waits; the persistent tensor-core recurrence is lstm_bf16_sm100.cu.
#include "common.cuh"

namespace s2vt {

struct GemmF32Params;
int gemm_f32_simple(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                    int b_trans, float* C, long long ldc, const float* bias, int accumulate, int split_k,
                    long long split_stride);

// gates(b,u) for g in i,f,g,o:  x = (pre ? pre[b,gH+u] : bias[gH+u]) + sum_z part[z][b][gH+u]
__global__ void lstm_pointwise_fwd_kernel(int B, int H, const float* __restrict__ pre, const float* __restrict__ bias,
                                          const float* __restrict__ part, int nsplit, long long split_stride,
                                          const float* __restrict__ c_prev, float* __restrict__ h_out, long long h_ld,
                                          float* __restrict__ c_out, float* __restrict__ gates_out,
                                          float* __restrict__ h_out2, long long h2_ld) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  float x[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const long long off = (long long)b * 4 * H + g * H + u;
    float v = pre ? pre[off] : bias[g * H + u];
    for (int z = 0; z < nsplit; ++z) v += part[z * split_stride + off];
    x[g] = v;
  }
  const float i = sigmoidf_exact(x[0]), f = sigmoidf_exact(x[1]), g = tanhf(x[2]), o = sigmoidf_exact(x[3]);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float c = f * cp + i * g;
  const float h = o * tanhf(c);
  h_out[(long long)b * h_ld + u] = h;
  if (h_out2) h_out2[(long long)b * h2_ld + u] = h;
  c_out[idx] = c;
  if (gates_out) {
    float* go = gates_out + (long long)b * 4 * H + u;
    go[0] = i; go[H] = f; go[2 * H] = g; go[3 * H] = o;
  }
}

// dh = dout + sum_z part[z];   produces dgates[b, 4H] and updates dc in place
__global__ void lstm_pointwise_bwd_kernel(int B, int H, const float* __restrict__ dout, const float* __restrict__ part,
                                          int nsplit, long long split_stride, const float* __restrict__ gates,
                                          const float* __restrict__ c_t, const float* __restrict__ c_prev,
                                          float* __restrict__ dc, float* __restrict__ dgates) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  float dh = dout ? dout[idx] : 0.f;
  for (int z = 0; z < nsplit; ++z) dh += part[z * split_stride + idx];
  const float* gi = gates + (long long)b * 4 * H + u;
  const float i = gi[0], f = gi[H], g = gi[2 * H], o = gi[3 * H];
  const float tc = tanhf(c_t[idx]);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float d_o = dh * tc;
  const float dcv = dc[idx] + dh * o * (1.f - tc * tc);
  float* dg = dgates + (long long)b * 4 * H + u;
  dg[0] = dcv * g * i * (1.f - i);
  dg[H] = dcv * cp * f * (1.f - f);
  dg[2 * H] = dcv * i * (1.f - g * g);
  dg[3 * H] = d_o * o * (1.f - o);
  dc[idx] = dcv * f;
}

int lstm_pointwise_fwd(cudaStream_t st, int B, int H, const float* pre, const float* bias, const float* part, int nsplit,
                       long long split_stride, const float* c_prev, float* h_out, long long h_ld, float* c_out,
                       float* gates_out, float* h_out2, long long h2_ld) {
  lstm_pointwise_fwd_kernel<<<ceil_div((long long)B * H, 256), 256, 0, st>>>(B, H, pre, bias, part, nsplit, split_stride, c_prev,
                                                                            h_out, h_ld, c_out, gates_out, h_out2, h2_ld);
  S2VT_CHECK_LAUNCH();
  return 0;
}

static int pick_split(int M, int N, int K) {
  const int tiles = ceil_div(M, 64) * ceil_div(N, 64);
  int s = 148 / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > 8) s = 8;
  while (s > 1 && K / s < 64) --s;
  return s;
}

}  // namespace s2vt

using namespace s2vt;

extern "C" int64_t s2vt_lstm_ws_bytes(int B, int H) {
  // 8 split-K partial planes of [B,4H] + 4 state planes of [B,H]
  return (int64_t)sizeof(float) * ((int64_t)8 * B * 4 * H + (int64_t)4 * B * H) + 256;
}

extern "C" int s2vt_lstm_fwd_f32(void* stream, int T, int B, int H, int n_pre,
                                 const float* pre, const float* bias_sum, const float* w_hh,
                                 const float* h0, const float* c0,
                                 float* out, float* gates, float* cells, float* hT, float* cT, void* ws) {
  cudaStream_t st = (cudaStream_t)stream;
  S2VT_REQUIRE(T >= 0 && B > 0 && H > 0, "s2vt_lstm_fwd_f32: bad dims");
  S2VT_REQUIRE(out && w_hh && ws, "s2vt_lstm_fwd_f32: null pointer");
  S2VT_REQUIRE(n_pre <= 0 || pre, "s2vt_lstm_fwd_f32: pre is null but n_pre > 0");
  S2VT_REQUIRE(n_pre >= T || bias_sum, "s2vt_lstm_fwd_f32: bias_sum needed for steps >= n_pre");
  float* part = (float*)ws;
  const long long ps = (long long)B * 4 * H;
  float* cbuf = part + 8 * ps;                     // [2][B][H] ping-pong when no cell stash is kept
  const int S = pick_split(B, 4 * H, H);
  const int threads = 256, blocks = ceil_div((long long)B * H, threads);
  for (int t = 0; t < T; ++t) {
    const float* hp = t == 0 ? h0 : out + (long long)(t - 1) * B * H;
    const float* cp = t == 0 ? c0 : (cells ? cells + (long long)(t - 1) * B * H : cbuf + ((t - 1) & 1) * (long long)B * H);
    float* cn = cells ? cells + (long long)t * B * H : cbuf + (t & 1) * (long long)B * H;
    int ns = 0;
    if (hp) {
      int rc = gemm_f32_simple(st, B, 4 * H, H, hp, H, w_hh, H, 0, part, 4 * H, nullptr, 0, S, ps);
      if (rc) return rc;
      ns = S;
    }
    lstm_pointwise_fwd_kernel<<<blocks, threads, 0, st>>>(
        B, H, t < n_pre ? pre + (long long)t * ps : nullptr, bias_sum, part, ns, ps, cp,
        out + (long long)t * B * H, H, cn, gates ? gates + (long long)t * ps : nullptr, nullptr, 0);
    S2VT_CHECK_LAUNCH();
  }
  if (T > 0) {
    if (hT) S2VT_CHECK_CUDA(cudaMemcpyAsync(hT, out + (long long)(T - 1) * B * H, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
    if (cT) {
      const float* cl = cells ? cells + (long long)(T - 1) * B * H : cbuf + ((T - 1) & 1) * (long long)B * H;
      S2VT_CHECK_CUDA(cudaMemcpyAsync(cT, cl, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
    }
  }
  return 0;
}

extern "C" int s2vt_lstm_bwd_f32(void* stream, int T, int B, int H, int dout_t0,
                                 const float* dout, const float* gates, const float* cells, const float* w_hh,
                                 float* dgates, void* ws) {
  cudaStream_t st = (cudaStream_t)stream;
  S2VT_REQUIRE(T >= 0 && B > 0 && H > 0, "s2vt_lstm_bwd_f32: bad dims");
  S2VT_REQUIRE(gates && cells && w_hh && dgates && ws, "s2vt_lstm_bwd_f32: null pointer");
  float* part = (float*)ws;
  const long long ps = (long long)B * 4 * H;
  const long long hs = (long long)B * H;
  float* dc = part + 8 * ps;
  S2VT_CHECK_CUDA(cudaMemsetAsync(dc, 0, sizeof(float) * hs, st));
  const int S = pick_split(B, H, 4 * H);
  const int threads = 256, blocks = ceil_div(hs, threads);
  for (int t = T - 1; t >= 0; --t) {
    const int ns = (t == T - 1) ? 0 : S;
    lstm_pointwise_bwd_kernel<<<blocks, threads, 0, st>>>(
        B, H, (dout && t >= dout_t0) ? dout + (long long)t * hs : nullptr, part, ns, hs,
        gates + (long long)t * ps, cells + (long long)t * hs, t > 0 ? cells + (long long)(t - 1) * hs : nullptr,
        dc, dgates + (long long)t * ps);
    S2VT_CHECK_LAUNCH();
    if (t > 0) {
      // dh_rec[B,H] = dgates_t[B,4H] . W_hh[4H,H]
      int rc = gemm_f32_simple(st, B, H, 4 * H, dgates + (long long)t * ps, 4 * H, w_hh, H, 1, part, H, nullptr, 0, S, hs);
      if (rc) return rc;
    }
  }
  return 0;
}
