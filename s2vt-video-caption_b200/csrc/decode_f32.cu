// Exact fp32 decode: the greedy loop of S2VT.forward(mode='test') (S2VTModel.py:88-110) and the
// beam search of S2VT.beam_search (S2VTModel.py:149-240), the latter restated in lock-step over all
// videos x beams so that one set of launches per depth serves the whole batch.
#include "common.cuh"
#include <math.h>

namespace s2vt {

int gemm_f32_simple(cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                    int b_trans, float* C, long long ldc, const float* bias, int accumulate, int split_k,
                    long long split_stride);

int lstm_pointwise_fwd(cudaStream_t st, int B, int H, const float* pre, const float* bias, const float* part, int nsplit,
                       long long split_stride, const float* c_prev, float* h_out, long long h_ld, float* c_out,
                       float* gates_out, float* h_out2, long long h2_ld);

static int pick_split_dec(int M, int N, int K) {
  const int tiles = ceil_div(M, 64) * ceil_div(N, 64);
  int s = 148 / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > 8) s = 8;
  while (s > 1 && K / s < 64) --s;
  return s;
}

struct ArgBest { float v; int i; };
__device__ __forceinline__ ArgBest better(ArgBest a, ArgBest b) {   // max value, ties -> lowest index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ ArgBest block_argmax(ArgBest x, ArgBest* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgBest y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = x;
  __syncthreads();
  ArgBest r = sh[0];
  for (int j = 1; j < nw; ++j) r = better(r, sh[j]);
  return r;
}

// ------------------------------------------------------------------ greedy
__global__ void greedy_init_kernel(int B, int H, int E, const float* __restrict__ emb, int sos, const float* __restrict__ h2,
                                   float* __restrict__ acat) {
  const int b = blockIdx.x;
  float* row = acat + (long long)b * (E + H);
  for (int e = threadIdx.x; e < E; e += blockDim.x) row[e] = emb[(long long)sos * E + e];
  for (int u = threadIdx.x; u < H; u += blockDim.x) row[E + u] = h2[(long long)b * H + u];
}

// argmax over one row of logits; writes the token (batch-major) and the next step's embedding
__global__ void greedy_argmax_kernel(const float* __restrict__ logits, int V, int E, const float* __restrict__ emb,
                                     int64_t* __restrict__ tokens, int n_steps, int step, float* __restrict__ acat, int acat_ld) {
  __shared__ ArgBest sh[32];
  const int b = blockIdx.x;
  const float* z = logits + (long long)b * V;
  ArgBest best{-INFINITY, 0x7fffffff};
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float v = z[j];
    if (v > best.v || best.i == 0x7fffffff) { best.v = v; best.i = j; }
  }
  best = block_argmax(best, sh);
  if (threadIdx.x == 0) tokens[(long long)b * n_steps + step] = best.i;
  float* row = acat + (long long)b * acat_ld;
  for (int e = threadIdx.x; e < E; e += blockDim.x) row[e] = emb[(long long)best.i * E + e];
}

// ------------------------------------------------------------------ beam search
struct BeamMeta {            // one buffer of per-slot bookkeeping
  float* key; int* tok; int* len; int* fin; int* hist;
};

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

// replicate the per-video encode state into slot 0 of each video (other slots zero)
__global__ void beam_state_init_kernel(int B, int bw, int H, const float* __restrict__ state, float* __restrict__ cur) {
  const long long n = (long long)B * bw * H;
  for (int p = 0; p < 4; ++p)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const long long s = i / H; const int u = (int)(i % H);
      const int v = (int)(s / bw), j = (int)(s % bw);
      cur[p * n + i] = (j == 0) ? state[(long long)p * B * H + (long long)v * H + u] : 0.f;
    }
}

// A_cat2[s] = [ emb[tok[s]] | (h1' written by the LSTM-1 pointwise) | h2[s] ]
__global__ void beam_assemble_kernel(int H, int E, const float* __restrict__ emb, const int* __restrict__ tok,
                                     const float* __restrict__ h2, float* __restrict__ acat) {
  const int s = blockIdx.x;
  float* row = acat + (long long)s * (E + 2 * H);
  const float* er = emb + (long long)tok[s] * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) row[e] = er[e];
  for (int u = threadIdx.x; u < H; u += blockDim.x) row[E + H + u] = h2[(long long)s * H + u];
}

// log_softmax + top-k of one slot's logits.  256 threads; the row lives in dynamic smem.
__global__ void beam_lsm_topk_kernel(const float* __restrict__ logits, int V, int topk, float* __restrict__ cand_lp,
                                     int* __restrict__ cand_tok) {
  extern __shared__ float zrow[];
  __shared__ ArgBest sh[32];
  __shared__ float shf[32];
  const int s = blockIdx.x;
  const float* z = logits + (long long)s * V;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += blockDim.x) { const float v = z[j]; zrow[j] = v; mx = fmaxf(mx, v); }
  // block max
  {
    float v = mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = v;
    __syncthreads();
    v = shf[0];
    for (int j = 1; j < (int)(blockDim.x >> 5); ++j) v = fmaxf(v, shf[j]);
    mx = v;
    __syncthreads();
  }
  float sum = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) sum += expf(zrow[j] - mx);
  {
    float v = sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = v;
    __syncthreads();
    v = 0.f;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) v += shf[j];
    sum = v;
  }
  const float logsum = logf(sum);
  auto rescan = [&]() {
    ArgBest b{-INFINITY, 0x7fffffff};
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      const float v = zrow[j];
      if (v > b.v || (b.i == 0x7fffffff && v == b.v)) { b.v = v; b.i = j; }
    }
    return b;
  };
  ArgBest mine = rescan();
  for (int r = 0; r < topk; ++r) {
    const ArgBest w = block_argmax(mine, sh);
    if (threadIdx.x == 0) {
      cand_lp[(long long)s * topk + r] = (w.v - mx) - logsum;      // log_softmax as torch computes it
      cand_tok[(long long)s * topk + r] = w.i;
    }
    if (w.i != 0x7fffffff && (w.i % (int)blockDim.x) == (int)threadIdx.x) {
      zrow[w.i] = -INFINITY;
      mine = rescan();
      if (mine.v == -INFINITY) mine.i = 0x7fffffff;               // exhausted
    }
  }
}

// One thread per video: the PriorityQueue logic of S2VTModel.py:186-238 on this depth's candidates.
__global__ void beam_select_kernel(int B, int bw, int topk, int D1, int eos, const float* __restrict__ len_pen,
                                   BeamMeta old_, BeamMeta new_, const float* __restrict__ cand_lp, const int* __restrict__ cand_tok,
                                   int* __restrict__ nbeam, int* __restrict__ done, int* __restrict__ parent,
                                   int64_t* __restrict__ out_tokens, int* __restrict__ out_len) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= B) return;
  const int base = v * bw;
  if (done[v]) {                   // frozen: carry the bookkeeping so that later depths read valid token ids
    for (int j = 0; j < bw; ++j) {
      const int s = base + j;
      new_.key[s] = old_.key[s]; new_.tok[s] = old_.tok[s]; new_.len[s] = old_.len[s]; new_.fin[s] = old_.fin[s];
      parent[s] = s;
    }
    return;
  }
  const int nb = nbeam[v];
  int ptr[32];                     // next unused candidate per beam slot (bw <= 32)
  int count = 0;
  for (int j = 0; j < nb; ++j) { ptr[j] = 0; count += old_.fin[base + j] ? 1 : topk; }
  const bool last = (count <= bw);
  const int take = count < bw ? count : bw;
  for (int r = 0; r < take; ++r) {
    // head of the merge: smallest key among the remaining entries (ties: lowest slot, lowest rank)
    float bestk = INFINITY; int bj = -1;
    for (int j = 0; j < nb; ++j) {
      const int s = base + j;
      float k;
      if (old_.fin[s]) { if (ptr[j] > 0) continue; k = old_.key[s]; }
      else {
        if (ptr[j] >= topk) continue;
        k = -(cand_lp[(long long)s * topk + ptr[j]] / len_pen[old_.len[s] + 1]);
      }
      if (bj < 0 || k < bestk) { bestk = k; bj = j; }
    }
    const int s = base + bj, d = base + r;
    const int ln = old_.len[s];
    int* hn = new_.hist + (long long)d * D1;
    const int* ho = old_.hist + (long long)s * D1;
    for (int q = 0; q < D1; ++q) hn[q] = ho[q];
    if (old_.fin[s]) {
      new_.key[d] = old_.key[s]; new_.tok[d] = old_.tok[s]; new_.len[d] = ln; new_.fin[d] = 1;
    } else {
      const int tk = cand_tok[(long long)s * topk + ptr[bj]];
      new_.key[d] = bestk; new_.tok[d] = tk; new_.len[d] = ln + 1; new_.fin[d] = (tk == eos) ? 1 : 0;
      if (ln < D1) hn[ln] = tk;
    }
    ptr[bj] += 1;
    parent[d] = s;
    if (r == 0) {                  // best entry of this queue = the answer if the search stops here
      const int L = new_.len[d];
      for (int q = 0; q < D1; ++q) out_tokens[(long long)v * D1 + q] = q < L ? hn[q] : -1;
      out_len[v] = L;
    }
  }
  for (int r = take; r < bw; ++r) {
    const int d = base + r;
    new_.key[d] = INFINITY; new_.tok[d] = 0; new_.len[d] = 1; new_.fin[d] = 0;
    parent[d] = d;
  }
  nbeam[v] = take;
  if (last) done[v] = 1;
}

// cur[p][s] = nxt[p][parent[s]] for the four state planes
__global__ void beam_gather_kernel(int S, int H, const int* __restrict__ parent, const float* __restrict__ nxt, float* __restrict__ cur) {
  const int s = blockIdx.x, p = blockIdx.y;
  const float* src = nxt + ((long long)p * S + parent[s]) * H;
  float* dst = cur + ((long long)p * S + s) * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) dst[u] = src[u];
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace s2vt

using namespace s2vt;

extern "C" int64_t s2vt_greedy_ws_bytes(int B, int H, int E, int V) {
  size_t n = 0;
  n += align_up(sizeof(float) * (size_t)B * (E + H));
  n += align_up(sizeof(float) * (size_t)8 * B * 4 * H);
  n += align_up(sizeof(float) * (size_t)B * V);
  return (int64_t)n + 256;
}

extern "C" int s2vt_greedy_decode_f32(void* stream, int B, int H, int E, int V, int n_steps, int sos_ix,
                                      const float* pre2_vid, const float* w_cat, const float* emb,
                                      const float* w_out, const float* b_out,
                                      float* h2, float* c2, int64_t* tokens, void* ws) {
  cudaStream_t st = (cudaStream_t)stream;
  S2VT_REQUIRE(B > 0 && H > 0 && E > 0 && V > 0 && n_steps >= 0, "s2vt_greedy_decode_f32: bad dims");
  S2VT_REQUIRE(pre2_vid && w_cat && emb && w_out && b_out && h2 && c2 && tokens && ws, "s2vt_greedy_decode_f32: null pointer");
  S2VT_REQUIRE(sos_ix >= 0 && sos_ix < V, "s2vt_greedy_decode_f32: sos_ix out of range");
  char* w = (char*)ws;
  float* acat = (float*)w; w += align_up(sizeof(float) * (size_t)B * (E + H));
  float* part = (float*)w; w += align_up(sizeof(float) * (size_t)8 * B * 4 * H);
  float* logits = (float*)w;
  const long long ps = (long long)B * 4 * H;
  const int K = E + H;
  const int S = pick_split_dec(B, 4 * H, K);
  greedy_init_kernel<<<B, 128, 0, st>>>(B, H, E, emb, sos_ix, h2, acat);
  S2VT_CHECK_LAUNCH();
  for (int k = 0; k < n_steps; ++k) {
    int rc = gemm_f32_simple(st, B, 4 * H, K, acat, K, w_cat, K, 0, part, 4 * H, nullptr, 0, S, ps);
    if (rc) return rc;
    rc = lstm_pointwise_fwd(st, B, H, pre2_vid + (long long)k * ps, nullptr, part, S, ps, c2, acat + E, K, c2, nullptr, h2, H);
    if (rc) return rc;
    rc = gemm_f32_simple(st, B, V, H, acat + E, K, w_out, H, 0, logits, V, b_out, 0, 1, 0);
    if (rc) return rc;
    greedy_argmax_kernel<<<B, 256, 0, st>>>(logits, V, E, emb, tokens, n_steps, k, acat, K);
    S2VT_CHECK_LAUNCH();
  }
  return 0;
}

namespace {
struct BeamWs {
  float *cur, *nxt, *acat, *part, *logits, *cand_lp;
  int *cand_tok, *nbeam, *done, *parent;
  s2vt::BeamMeta meta[2];
  size_t bytes;
};
BeamWs carve_beam(char* base, int B, int H, int E, int V, int bw, int D1, int topk) {
  BeamWs w{};
  const size_t S = (size_t)B * bw;
  size_t off = 0;
  auto take = [&](size_t n) { char* p = base ? base + off : nullptr; off += s2vt::align_up(n); return p; };
  w.cur = (float*)take(sizeof(float) * 4 * S * H);
  w.nxt = (float*)take(sizeof(float) * 4 * S * H);
  w.acat = (float*)take(sizeof(float) * S * (E + 2 * (size_t)H));
  w.part = (float*)take(sizeof(float) * 8 * S * 4 * H);
  w.logits = (float*)take(sizeof(float) * S * V);
  w.cand_lp = (float*)take(sizeof(float) * S * topk);
  w.cand_tok = (int*)take(sizeof(int) * S * topk);
  w.nbeam = (int*)take(sizeof(int) * B);
  w.done = (int*)take(sizeof(int) * B);
  w.parent = (int*)take(sizeof(int) * S);
  for (int i = 0; i < 2; ++i) {
    w.meta[i].key = (float*)take(sizeof(float) * S);
    w.meta[i].tok = (int*)take(sizeof(int) * S);
    w.meta[i].len = (int*)take(sizeof(int) * S);
    w.meta[i].fin = (int*)take(sizeof(int) * S);
    w.meta[i].hist = (int*)take(sizeof(int) * S * D1);
  }
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" int64_t s2vt_beam_ws_bytes(int B, int H, int E, int V, int beam_width, int max_depth, int topk) {
  return (int64_t)carve_beam(nullptr, B, H, E, V, beam_width, max_depth + 1, topk).bytes + 256;
}

extern "C" int s2vt_beam_search_f32(void* stream, int B, int H, int E, int V, int beam_width, int max_depth, int topk,
                                    int sos_ix, int eos_ix,
                                    const float* state, const float* bias1, const float* w_hh1,
                                    const float* w_cat2, const float* bias2, const float* emb,
                                    const float* w_out, const float* b_out, const float* len_pen,
                                    int64_t* out_tokens, int32_t* out_len, void* ws) {
  cudaStream_t st = (cudaStream_t)stream;
  S2VT_REQUIRE(B > 0 && H > 0 && E > 0 && V > 0, "s2vt_beam_search_f32: bad dims");
  S2VT_REQUIRE(beam_width >= 1 && beam_width <= 32, "s2vt_beam_search_f32: beam_width must be in [1,32]");
  S2VT_REQUIRE(topk >= 1 && topk <= V, "s2vt_beam_search_f32: topk must be in [1,V] (the reference's topk(20) needs V >= 20)");
  S2VT_REQUIRE(max_depth >= 1, "s2vt_beam_search_f32: max_depth must be >= 1");
  S2VT_REQUIRE((size_t)V * sizeof(float) <= 200 * 1024, "s2vt_beam_search_f32: vocabulary too large for the smem top-k (V <= 51200)");
  S2VT_REQUIRE(state && bias1 && w_hh1 && w_cat2 && bias2 && emb && w_out && b_out && len_pen && out_tokens && out_len && ws,
               "s2vt_beam_search_f32: null pointer");
  const int D1 = max_depth + 1;
  const int S = B * beam_width;
  BeamWs w = carve_beam((char*)ws, B, H, E, V, beam_width, D1, topk);
  const long long SH = (long long)S * H;
  float *h1 = w.cur, *c1 = w.cur + SH, *h2 = w.cur + 2 * SH, *c2 = w.cur + 3 * SH;
  float *h1n = w.nxt, *c1n = w.nxt + SH, *h2n = w.nxt + 2 * SH, *c2n = w.nxt + 3 * SH;
  const long long ps = (long long)S * 4 * H;
  const int K2 = E + 2 * H;
  S2VT_CHECK_CUDA(cudaFuncSetAttribute(beam_lsm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  beam_init_kernel<<<ceil_div(B, 128), 128, 0, st>>>(B, beam_width, D1, sos_ix, w.meta[0], w.nbeam, w.done, out_tokens, out_len);
  S2VT_CHECK_LAUNCH();
  beam_state_init_kernel<<<148 * 2, 256, 0, st>>>(B, beam_width, H, state, w.cur);
  S2VT_CHECK_LAUNCH();
  const int S1 = pick_split_dec(S, 4 * H, H), S2 = pick_split_dec(S, 4 * H, K2);
  for (int depth = 0; depth < max_depth; ++depth) {
    const BeamMeta& mo = w.meta[depth & 1];
    const BeamMeta& mn = w.meta[(depth + 1) & 1];
    // vid_rnn step on the zero pad (S2VTModel.py:208-210)
    int rc = gemm_f32_simple(st, S, 4 * H, H, h1, H, w_hh1, H, 0, w.part, 4 * H, nullptr, 0, S1, ps);
    if (rc) return rc;
    rc = lstm_pointwise_fwd(st, S, H, nullptr, bias1, w.part, S1, ps, c1, h1n, H, c1n, nullptr, w.acat + E, K2);
    if (rc) return rc;
    // word_rnn step on [embed(word) | vid_out] (S2VTModel.py:207,211-212)
    beam_assemble_kernel<<<S, 128, 0, st>>>(H, E, emb, mo.tok, h2, w.acat);
    S2VT_CHECK_LAUNCH();
    rc = gemm_f32_simple(st, S, 4 * H, K2, w.acat, K2, w_cat2, K2, 0, w.part, 4 * H, nullptr, 0, S2, ps);
    if (rc) return rc;
    rc = lstm_pointwise_fwd(st, S, H, nullptr, bias2, w.part, S2, ps, c2, h2n, H, c2n, nullptr, nullptr, 0);
    if (rc) return rc;
    // out_linear + log_softmax + top-k (S2VTModel.py:213-216)
    rc = gemm_f32_simple(st, S, V, H, h2n, H, w_out, H, 0, w.logits, V, b_out, 0, 1, 0);
    if (rc) return rc;
    beam_lsm_topk_kernel<<<S, 256, sizeof(float) * (size_t)V, st>>>(w.logits, V, topk, w.cand_lp, w.cand_tok);
    S2VT_CHECK_LAUNCH();
    beam_select_kernel<<<ceil_div(B, 64), 64, 0, st>>>(B, beam_width, topk, D1, eos_ix, len_pen, mo, mn, w.cand_lp, w.cand_tok,
                                                      w.nbeam, w.done, w.parent, out_tokens, out_len);
    S2VT_CHECK_LAUNCH();
    beam_gather_kernel<<<dim3(S, 4), 128, 0, st>>>(S, H, w.parent, w.nxt, w.cur);
    S2VT_CHECK_LAUNCH();
  }
  return 0;
}
