/*
 * s2vt_b200.h -- C ABI of libs2vt_b200.so, the sm_100a kernel library underneath the S2VT drop-in.
 *
 * The reference (Kamino666/S2VT-video-caption) has no FFI: its hot path is a Python nn.Module whose
 * arithmetic is delegated to PyTorch library calls (nn.LSTM -> cuDNN/oneDNN, nn.Linear -> cuBLAS/MKL,
 * nn.CrossEntropyLoss, optim.Adam).  Each entry point below replaces one of those library call sites;
 * the citation after "replaces:" is the reference file:line (relative to the reference root).
 *
 * Conventions
 *   - plain C, raw DEVICE pointers + sizes, no torch types.  `stream` is a cudaStream_t passed as void*.
 *   - every call returns 0 on success, non-zero on error; s2vt_last_error() gives the message
 *     (thread-local).  Nothing throws, nothing calls exit(), nothing synchronises the device.
 *   - the library allocates nothing: all workspaces are caller-provided.
 *   - "time-major" means rows ordered (t, b): row = t*B + b.
 *   - f32 entry points are the exact path (CUDA-core FMA, fp32 operands and accumulation); bf16 entry
 *     points use tcgen05 tensor cores with fp32 accumulation in TMEM.
 */
#ifndef S2VT_B200_H_
#define S2VT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2VT_ABI_VERSION 1

/* Row addressing of a matrix operand: row m starts at element
 *     (m / inner) * stride_outer + (m % inner) * stride_inner.
 * Identity for a dense row-major matrix with leading dimension ld: {1, ld, 0}.
 * Used to read batch-major [B,L,F] features in time-major order and to write time-major rows into the
 * batch-major [B,L-1,V] logits the reference API returns (S2VTModel.py:78-81). */
typedef struct s2vt_rowmap {
  int32_t inner;
  int64_t stride_outer;
  int64_t stride_inner;
} s2vt_rowmap;

int         s2vt_abi_version(void);
const char* s2vt_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t     s2vt_launch_count(void);
/* 1 when the binary carries sm_100a code for the tcgen05 kernels. */
int         s2vt_has_tcgen05(void);
/* Debug aid: synchronises `stream` and returns the device-side error flag of the tensor-core kernels
 * (0 = none; non-zero = an mbarrier wait timed out, which is never expected). */
int         s2vt_device_error_flag(void* stream);
/* Resets that flag (after the caller has reported it). */
int         s2vt_device_error_clear(void);

/* ------------------------------------------------------------------ exact fp32 GEMM (CUDA cores)
 * C[cmap(m), n] = sum_k A(m,k) * B(n,k) (+ bias[n]) (+ C if accumulate)
 *   A(m,k) = A[amap(m) + k]            if !a_trans  (K contiguous)
 *          = A[amap(k) + m]            if  a_trans  (M contiguous; amap addresses the K index)
 *   B(n,k) = B[bmap(n) + k]            if !b_trans  (weight layout [N,K], as nn.Linear stores it)
 *          = B[bmap(k) + n]            if  b_trans
 *   split_k > 1: slice z of K writes its partial product to C + z * split_stride (no bias/accumulate).
 * replaces: nn.Linear / addmm at S2VTModel.py:54,80 and the input-side and recurrent products inside
 * nn.LSTM at S2VTModel.py:57,60,67,77,86,93,103, plus their autograd transposes (train.py:124). */
int s2vt_gemm_f32(void* stream, int M, int N, int K,
                  const float* A, s2vt_rowmap amap, int a_trans,
                  const float* B, s2vt_rowmap bmap, int b_trans,
                  float* C, s2vt_rowmap cmap,
                  const float* bias, int accumulate, int split_k, int64_t split_stride);

/* ------------------------------------------------------------------ bf16 tcgen05 GEMM (tensor cores)
 * Same contract as s2vt_gemm_f32 with bf16 operands, fp32 accumulation in TMEM, fed by TMA.
 *   a_mn_major / b_mn_major: operand stored with its M (resp. N) index contiguous ([K,M] / [K,N]).
 *   out_bf16: C is bf16 instead of f32.   lda / ldb are leading dimensions in elements.
 * tensor maps are encoded per call on the host (cuTensorMapEncodeTiled) -- pointers must be 16 B
 * aligned and leading dimensions multiples of 8 elements. */
int s2vt_gemm_bf16(void* stream, int M, int N, int K,
                   const void* A, int64_t lda, int a_mn_major,
                   const void* B, int64_t ldb, int b_mn_major,
                   void* C, s2vt_rowmap cmap, int out_bf16,
                   const float* bias, int accumulate);

/* Scheduling knobs of s2vt_gemm_bf16 for the calling thread: max_ctas > 0 caps the persistent kernel's grid (used when the
 * recurrence clusters of another stream hold part of the machine), use_persistent = 0 forces the one-tile-per-CTA kernel. */
int s2vt_gemm_bf16_set_mode(int max_ctas, int use_persistent);

/* f32 -> bf16 cast (optionally also writes the transpose: dst_t[c, r] = src[r, c]). */
int s2vt_cast_bf16(void* stream, const float* src, void* dst, void* dst_t, int64_t rows, int64_t cols);

/* ------------------------------------------------------------------ LSTM recurrence, exact fp32
 * One nn.LSTM layer (num_layers=1, unidirectional, batch_first handled by the caller) over T steps.
 *   pre      [n_pre, B, 4H]  input-side pre-activations W_ih x_t + b_ih + b_hh for t < n_pre
 *   bias_sum [4H]            b_ih + b_hh, used alone for t >= n_pre (the reference's zero padding,
 *                            S2VTModel.py:64-65 / 208-210)
 *   w_hh     [4H, H]         gate row blocks i,f,g,o
 *   h0,c0    [B,H] or NULL (zero state)
 *   out      [T, B, H]       h_t, time-major
 *   gates    [T, B, 4H] or NULL   post-activation i,f,g,o stash for BPTT
 *   cells    [T, B, H]  or NULL   c_t stash for BPTT
 *   hT,cT    [B,H] or NULL   final state
 *   ws       >= s2vt_lstm_ws_bytes(B,H) bytes of scratch
 * replaces: self.vid_rnn(...) / self.word_rnn(...) at S2VTModel.py:57,60,67,77,86. */
int64_t s2vt_lstm_ws_bytes(int B, int H);
int s2vt_lstm_fwd_f32(void* stream, int T, int B, int H, int n_pre,
                      const float* pre, const float* bias_sum, const float* w_hh,
                      const float* h0, const float* c0,
                      float* out, float* gates, float* cells, float* hT, float* cT, void* ws);

/* Same recurrence on the tensor cores: a persistent thread-block cluster (H/32 CTAs, 16 batch columns per cluster)
 * keeps the bf16 W_hh slices resident in shared memory for all T steps, multiplies with tcgen05.mma into TMEM and
 * exchanges h_t between CTAs through distributed shared memory.  Needs H % 64 == 0, 64 <= H <= 512.
 *   w_hh_bf16 [4H,H] bf16;  pre / bias_sum / h0 / c0 / hT / cT as above (fp32)
 *   out_bf16 [T,B,H] bf16 (time-major, feeds the next GEMM)
 *   gates_bf16 / cells: BPTT stash in a kernel-private layout shared only with s2vt_lstm_bwd_bf16, or NULL.
 *     With Bp = s2vt_lstm_bf16_batch_pad(B), nbt = Bp/16, CS = H/32:
 *       gates_bf16 [T][nbt][CS][16][32][4] bf16  (T*Bp*4H elements; innermost = i,f,g,o of one unit)
 *       cells      [T][nbt][CS][16][32]    f32   (T*Bp*H elements) */
int64_t s2vt_lstm_bf16_batch_pad(int B);
int s2vt_lstm_fwd_bf16(void* stream, int T, int B, int H, int n_pre,
                       const float* pre, const float* bias_sum, const void* w_hh_bf16,
                       const float* h0, const float* c0,
                       void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT);

/* As above with a direction flag: reverse = 1 makes processing step s read pre / write out and the stash at time index T-1-s,
 * i.e. the reverse direction of a bidirectional nn.LSTM (attention_baseline.py:23) on time-ordered buffers, with no reversed copies. */
int s2vt_lstm_fwd_bf16_dir(void* stream, int T, int B, int H, int n_pre,
                           const float* pre, const float* bias_sum, const void* w_hh_bf16,
                           const float* h0, const float* c0,
                           void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT, int reverse);

/* Tiles per cluster for the calling thread's subsequent recurrence launches: 1 (default) = one 16-column batch tile per cluster,
 * 2 = two tiles share a cluster's resident weight slice (half the SMs per sweep: lets two layers' sweeps run side by side). */
int s2vt_lstm_bf16_set_tiles_per_cluster(int n);

/* Wave-front coupling of two sweeps that run side by side on different streams, each launched ONCE (S2VT's two layers: the second
 * sweep consumes what a GEMM makes of the first one's output, S2VTModel.py:67-77).  The sequence is cut into n_sync <= S2VT_MAX_SYNC time
 * chunks [sync_t[k], sync_t[k+1]) (host array of n_sync + 1 ints, sync_t[0] = 0, sync_t[n_sync] = T).
 *   signal [n_sync] u32 device counters or NULL: every (CTA, 16-column batch tile) adds 1 to signal[k] once its out / stash rows of chunk
 *          k are visible device-wide, i.e. signal[k] reaches (H/32) * ceil(B/16) when the chunk is complete
 *   wait   [n_sync] u32 device counters or NULL: the kernel reads `pre` rows of chunk k only once wait[k] >= wait_val (advanced by
 *          s2vt_stream_write_value32 behind the producing GEMM); a wait of more than 2 s raises the device error flag.
 * The caller zeroes the counters (stream-ordered) before the launches.  tiles_per_cluster as s2vt_lstm_bf16_set_tiles_per_cluster. */
#define S2VT_MAX_SYNC 64
int s2vt_lstm_fwd_bf16_sync(void* stream, int T, int B, int H, int n_pre,
                            const float* pre, const float* bias_sum, const void* w_hh_bf16,
                            const float* h0, const float* c0,
                            void* out_bf16, void* gates_bf16, float* cells, float* hT, float* cT, int reverse,
                            int tiles_per_cluster, int n_sync, const int* sync_t, unsigned int* signal,
                            const unsigned int* wait, unsigned int wait_val);
/* The product that couples two such sweeps, as ONE launch resident beside them:  C[M,N] (f32, dense) (+)= A[M,K] * B (+ bias), with A's
 * rows produced chunk by chunk by the first sweep and C's rows consumed chunk by chunk by the second.  Persistent tcgen05 kernel on at
 * most max_ctas SMs; its scheduler hands out 128-row tiles in row order (reverse_m = 1: from the last row block to the first, for a
 * sweep that walks backwards in time) and, before a tile's TMA loads, waits until wait[k] >= wait_val for every chunk k the tile
 * overlaps (chunk k = rows [sync_row[k], sync_row[k+1]), sync_row[0] = 0, sync_row[n_sync] = M; wait[] = the first sweep's `signal`
 * counters).  When the last tile overlapping chunk k has landed in C (and is visible device-wide) ready[k] is incremented: the second
 * sweep waits for ready[k] >= 1.  done [n_sync] = scratch counters; the caller zeroes done and ready (stream-ordered) beforehand.
 * A is K-major ([M, lda]); B is K-major ([N, ldb], b_mn_major = 0) or MN-major ([K, ldb], b_mn_major = 1).  accumulate = 1: C += (TMA
 * reduce-add, no bias).  Because the kernel sits on its SMs from the start, the coupling never waits for CTA slots behind the bulk
 * products that run beside the sweeps (CTAs are dispatched in launch order whatever the stream priorities: tools/probe_priority.py).
 * Replaces the per-chunk `input2 @ W_ih` slices of S2VTModel.py:75-77 and their autograd transposes. */
int s2vt_gemm_bf16_gated(void* stream, int M, int N, int K, const void* A_bf16, int64_t lda, const void* B_bf16, int64_t ldb,
                         int b_mn_major, float* C, int64_t ldc, const float* bias, int accumulate, int max_ctas, int reverse_m,
                         int n_sync, const int* sync_row, const unsigned int* wait, unsigned int wait_val,
                         unsigned int* done, unsigned int* ready);
/* Stream-ordered operations on such counters (cuStreamWaitValue32 with GEQ / cuStreamWriteValue32): work enqueued on `stream` after
 * the wait starts only once *addr >= value; the write stores `value` once everything enqueued before it has completed. */
int s2vt_stream_wait_value32(void* stream, const unsigned int* addr, unsigned int value);
int s2vt_stream_write_value32(void* stream, unsigned int* addr, unsigned int value);
/* high = 1: the calling thread's following s2vt_gemm_bf16 launches of the one-tile-per-CTA kernel carry the device's greatest priority as a
 * launch attribute (kept by a captured graph's kernel node), so that their CTAs take SM slots ahead of queued CTAs of bulk products
 * running beside a sweep.  high = 0 restores the default. */
int s2vt_set_launch_priority(int high);
/* n > 0: the calling thread's following launches of the memory-bound bulk kernels (s2vt_adam_f32*, s2vt_colsum_bf16) use at most n CTAs
 * (grid-stride).  For work that runs beside a recurrence sweep: CTAs are dispatched in launch order, so a kernel with more CTAs than
 * free SM slots keeps every later kernel -- the wave front's coupling products -- waiting until its last CTA has been placed; a capped
 * kernel is resident at once and leaves slots over.  0 restores full-machine grids.  (The GEMM counterpart is s2vt_gemm_bf16_set_mode.) */
int s2vt_set_bulk_cta_cap(int n);
/* Executable graphs that honour per-node priorities.  `graph` is a cudaGraph_t (e.g. torch.cuda.CUDAGraph(keep_graph=True).raw_cuda_graph()).
 * Stream capture records each kernel node's priority (that of its stream, or its launch attribute), but cudaGraphInstantiate ignores
 * them unless cudaGraphInstantiateFlagUseNodePriority is given -- which s2vt_graph_instantiate(use_node_priority = 1) does.  The caller
 * keeps `graph` (and the memory its nodes reference) alive for as long as the executable graph is launched.
 * s2vt_graph_kernel_priorities: writes the priority of up to max_nodes kernel nodes (node order) and the number of kernel nodes to *n_out. */
int s2vt_graph_instantiate(void* graph, int use_node_priority, void** exec_out);
int s2vt_graph_launch(void* exec, void* stream);
int s2vt_graph_exec_destroy(void* exec);
int s2vt_graph_kernel_priorities(void* graph, int* prio_out, int max_nodes, int* n_out);
/* Measurement aid: a one-thread kernel that stores %globaltimer (ns) into *slot in stream order.  Unlike CUDA events it can be captured
 * into a graph, so tools/timeline_step.py gets the device-side timeline of a REPLAYED step. */
int s2vt_timestamp(void* stream, unsigned long long* slot);

/* Persistent tensor-core BPTT, the backward twin of s2vt_lstm_fwd_bf16 (same cluster shape; needs H % 128 == 0, H <= 512).
 *   dout [T,B,H] f32 (rows t < dout_t0 are zero and never read) or NULL;  gates_bf16 / cells: the forward stash (private layout)
 *   w_hh_t_bf16 [H,4H] bf16 = W_hh transposed;  dgates_bf16 [T,B,4H] bf16 out (time-major GEMM layout)
 * replaces: autograd of nn.LSTM under loss.backward(), train.py:124. */
int s2vt_lstm_bwd_bf16(void* stream, int T, int B, int H, int dout_t0,
                       const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                       void* dgates_bf16);

int s2vt_lstm_bwd_bf16_dir(void* stream, int T, int B, int H, int dout_t0,
                           const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                           void* dgates_bf16, int reverse);

/* One time chunk [t0, t1) of a longer BPTT sweep (chunks are launched latest first and chained through the state gradients), so that
 * the products consuming dgates of a chunk -- and the sweep of the layer below -- can run beside the following chunk's sweep.
 * All buffer pointers are those of the chunk's first step (dout / dgates rows t0.., stash blocks t0..), T = t1 - t0, dout_t0 relative.
 *   dh_in / dc_in  [B,H] f32: gradient w.r.t. h / c flowing into the chunk's last step from the chunk after it, or NULL (zero)
 *   dh_out / dc_out [B,H] f32: the same quantities for the chunk before this one, or NULL when t0 = 0
 *   has_prev: 1 when t0 > 0 (the stash block before `cells` holds c_{t0-1});  tiles_per_cluster: 1, or 2 = two 16-column batch tiles
 *   share a cluster's resident weight slice (half the SMs per sweep).  Not available with reverse = 1. */
int s2vt_lstm_bwd_bf16_chunk(void* stream, int T, int B, int H, int dout_t0,
                             const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                             void* dgates_bf16, int reverse, const float* dh_in, const float* dc_in, float* dh_out, float* dc_out,
                             int has_prev, int tiles_per_cluster);
/* The whole sweep in one launch with the wave-front counters of s2vt_lstm_fwd_bf16_sync (chunks are walked latest first: dgates rows of
 * chunk k are signalled, dout rows of chunk k are awaited). */
int s2vt_lstm_bwd_bf16_sync(void* stream, int T, int B, int H, int dout_t0,
                            const float* dout, const void* gates_bf16, const float* cells, const void* w_hh_t_bf16,
                            void* dgates_bf16, int reverse, const float* dh_in, const float* dc_in, float* dh_out, float* dc_out,
                            int has_prev, int tiles_per_cluster, int n_sync, const int* sync_t, unsigned int* signal,
                            const unsigned int* wait, unsigned int wait_val);

/* BPTT through one layer from a zero final-state gradient.
 *   dout   [T, B, H]  dL/dh_t from above; rows t < dout_t0 are treated as zero (and not read)
 *   gates, cells      the forward stash;  c0 = 0 is assumed (the reference never passes a state in training)
 *   dgates [T, B, 4H] out: gradient w.r.t. the pre-activations
 * replaces: autograd of nn.LSTM under loss.backward(), train.py:124. */
int s2vt_lstm_bwd_f32(void* stream, int T, int B, int H, int dout_t0,
                      const float* dout, const float* gates, const float* cells, const float* w_hh,
                      float* dgates, void* ws);

/* ------------------------------------------------------------------ embedding
 * out[t*B + b, :] = table[ids[b*ids_ld + t], :]   for t < n_t    (time-major gather)
 * replaces: self.embedding(targets), S2VTModel.py:71. */
int s2vt_embed_gather_f32(void* stream, const float* table, int E, const int64_t* ids, int64_t ids_ld,
                          int B, int n_t, float* out, int64_t out_ld);
/* grad_table[ids[b*ids_ld+t], :] += src[t*B + b, 0:E]  (dense grad, like nn.Embedding sparse=False) */
int s2vt_embed_scatter_add_f32(void* stream, float* grad_table, int E, const int64_t* ids, int64_t ids_ld,
                               int B, int n_t, const float* src, int64_t src_ld);

/* dst[(t*B + b)*N + n] = src[b*src_ld + n] for t < n_t: repeats one [B,N] block over time.
 * replaces: the per-step re-use of the (constant) attention context and its broadcast gradient,
 * attention_baseline.py:56,71-77 (torch.bmm of an all-ones weight row + torch.cat per decode step). */
int s2vt_bcast_rows_f32(void* stream, const float* src, int64_t src_ld, int B, int N, int n_t, float* dst);

/* out[i] = a[i] + b[i]  (b_ih + b_hh) */
int s2vt_add_f32(void* stream, const float* a, const float* b, float* out, int64_t n);

/* column sums: out[n] (+)= sum_m X[m*ld + n]; bias gradients. */
int s2vt_colsum_f32(void* stream, const float* X, int64_t M, int N, int64_t ld, float* out, int accumulate);

/* ------------------------------------------------------------------ loss
 * Mean cross entropy over R rows of V logits (the effective MaskCriterion, utils.py:13-26: the mask
 * cancels because nn.CrossEntropyLoss() already reduced to a scalar mean).
 *   logits [R, V] f32 row-major;  target row r = targets[(r / t_inner) * t_so + (r % t_inner) * t_si]
 *   row_loss [R] scratch;  loss: 1 float out;  dlogits: NULL, or [R,V] out = (softmax - onehot) * gscale[0] / R
 *   (dlogits may alias logits).  gscale: device pointer to the upstream scalar gradient, NULL = 1. */
int s2vt_ce_f32(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                float* row_loss, float* loss, float* dlogits, const float* gscale);

/* ------------------------------------------------------------------ Adam
 * torch.optim.Adam step (train.py:89-93,125) over a flat buffer; step_count is the 1-based step index.
 * bf16_copy (nullable): refreshed bf16 shadow of the updated parameters. */
int s2vt_adam_f32(void* stream, float* p, const float* g, float* m, float* v, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step_count, float grad_scale,
                  void* bf16_copy);
/* The same update with the step-dependent scalars in device memory, so that a captured CUDA graph of the whole train step can be
 * replayed: s2vt_adam_prepare does step_dev[0] += 1 and hyper_dev[0..1] = {lr_dev[0] / (1 - beta1^step), 1 / sqrt(1 - beta2^step)};
 * s2vt_adam_f32_dev reads them.  lr lives in device memory as well (a scheduler rewrites it between replays). */
int s2vt_adam_prepare(void* stream, int* step_dev, const float* lr_dev, float beta1, float beta2, float* hyper_dev);
int s2vt_adam_f32_dev(void* stream, float* p, const float* g, float* m, float* v, int64_t n,
                      float beta1, float beta2, float eps, const float* hyper_dev, float grad_scale, void* bf16_copy);

/* ------------------------------------------------------------------ greedy decode, exact fp32
 * The decode loop of S2VT.forward(mode='test'), S2VTModel.py:88-110.
 *   pre2_vid [n_steps, B, 4H]  vid-half pre-activations of word_rnn for the decode steps
 *                              (W_ih[:, E:] output1_t + b_ih + b_hh)
 *   w_cat    [4H, E+H]         [ W_ih[:, :E] | W_hh ] of word_rnn
 *   h2,c2    [B,H]             word_rnn state after the encode stage (updated in place)
 *   tokens   [B, n_steps] i64  out (batch-major, as the reference returns)
 *   ws       >= s2vt_greedy_ws_bytes(B,H,E,V) */
int64_t s2vt_greedy_ws_bytes(int B, int H, int E, int V);
int s2vt_greedy_decode_f32(void* stream, int B, int H, int E, int V, int n_steps, int sos_ix,
                           const float* pre2_vid, const float* w_cat, const float* emb,
                           const float* w_out, const float* b_out,
                           float* h2, float* c2, int64_t* tokens, void* ws);

/* ------------------------------------------------------------------ beam search, exact fp32
 * S2VT.beam_search, S2VTModel.py:149-240, in lock-step over all videos and beams.
 *   state    [4, B, H]   h1,c1,h2,c2 after the encode stage (S2VTModel.py:57-60)
 *   bias1    [4H]        vid_rnn b_ih+b_hh (its input is the zero pad, S2VTModel.py:208-210)
 *   w_hh1    [4H,H];  w_cat2 [4H, E+H+H] = [ W_ih2[:, :E] | W_ih2[:, E:] | W_hh2 ];  bias2 [4H]
 *   len_pen  [max_depth+2] host-computed float(pow(float(n), 0.7)) table (BeamSearchNode.eval)
 *   out_tokens [B, max_depth+1] i64, -1 padded, <sos> first;  out_len [B] i32
 *   topk: the reference expands top-20 (S2VTModel.py:216)
 *   ws >= s2vt_beam_ws_bytes(...) */
int64_t s2vt_beam_ws_bytes(int B, int H, int E, int V, int beam_width, int max_depth, int topk);
int s2vt_beam_search_f32(void* stream, int B, int H, int E, int V, int beam_width, int max_depth, int topk,
                         int sos_ix, int eos_ix,
                         const float* state, const float* bias1, const float* w_hh1,
                         const float* w_cat2, const float* bias2, const float* emb,
                         const float* w_out, const float* b_out, const float* len_pen,
                         int64_t* out_tokens, int32_t* out_len, void* ws);

/* ------------------------------------------------------------------ bf16 training-path helpers
 * bf16 twin of s2vt_embed_gather_f32 (replaces self.embedding(targets), S2VTModel.py:71). */
int s2vt_embed_gather_bf16(void* stream, const void* table_bf16, int E, const int64_t* ids, int64_t ids_ld,
                           int B, int n_t, void* out_bf16, int64_t out_ld);
/* out[n] = sum_m X[m*ld + n] for a bf16 matrix, fp32 accumulation (bias gradients); out2 (nullable) receives the same
 * sums (nn.LSTM's b_ih and b_hh share one gradient). */
int s2vt_colsum_bf16(void* stream, const void* X_bf16, int64_t M, int N, int64_t ld, float* out, float* out2);
/* Mean cross entropy as s2vt_ce_f32, single pass over each row (online max / sum-exp), with the gradient written as bf16
 * (the operand dtype of the backward GEMMs).  row_lse [R] (nullable) stashes each row's log-sum-exp in the forward call;
 * a backward call passes it back with have_lse = 1 and reads every logit exactly once.  row_loss / loss may be NULL. */
int s2vt_ce_bf16(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                 float* row_loss, float* loss, float* row_lse, int have_lse, void* dlogits_bf16, const float* gscale);

/* s2vt_ce_bf16 with the gradient row of logits row r written at element offset omap(r) of dlogits_bf16: reads the batch-major
 * fp32 logits the module API returns and writes dL/dlogits time-major, as the backward GEMMs consume it (no transpose pass). */
int s2vt_ce_bf16_mapped(void* stream, const float* logits, int64_t R, int V, const int64_t* targets, s2vt_rowmap tmap,
                        float* row_loss, float* loss, float* row_lse, int have_lse, void* dlogits_bf16, s2vt_rowmap omap,
                        const float* gscale);

/* ------------------------------------------------------------------ vocab projection fused with the loss statistics
 * logits = A W^T + bias on the tensor cores (A [R,K] bf16, W [V,K] bf16), written ONCE as bf16 [R, ldl]; the GEMM epilogue also
 * reduces, from the fp32 accumulators, each row's online-softmax partials per 256-column tile and picks out the target logit, so
 * the mean cross entropy needs no further pass over the logits:
 *   part_ws  >= s2vt_vocab_ce_ws_bytes(R, V) bytes, ztgt_ws [R] floats: scratch
 *   row_lse [R] out (log-sum-exp per row, kept for the backward call);  row_loss [R] scratch, loss: 1 float out (both nullable)
 * replaces: self.out_linear(...) at S2VTModel.py:80 + nn.CrossEntropyLoss inside MaskCriterion, utils.py:11,22. */
int64_t s2vt_vocab_ce_ws_bytes(int R, int V);
int s2vt_vocab_ce_fwd_bf16(void* stream, int R, int V, int K, const void* A_bf16, int64_t lda, const void* W_bf16, int64_t ldw,
                           const float* bias, void* logits_bf16, int64_t ldl, const int64_t* targets, s2vt_rowmap tmap,
                           void* part_ws, float* ztgt_ws, float* row_lse, float* row_loss, float* loss);
/* In place: bf16 logits -> bf16 dL/dlogits = (softmax - onehot) * gscale[0] / R, using the rows' log-sum-exp from the forward
 * call (replaces: autograd of nn.CrossEntropyLoss under loss.backward(), train.py:124). */
int s2vt_ce_dlogits_inplace_bf16(void* stream, void* logits_bf16, int64_t R, int V, int64_t ld, const float* row_lse,
                                 const int64_t* targets, s2vt_rowmap tmap, const float* gscale);

/* ------------------------------------------------------------------ LSTM recurrence / BPTT for any hidden size (H % 8 == 0)
 * One tcgen05 GEMM per time step (bf16 operands, fp32 accumulation) whose epilogue is the LSTM cell (forward) or the gate-gradient
 * arithmetic (backward); used where the persistent cluster kernels do not apply (H > 512 or H % 128 != 0, e.g. the paper sizing
 * H = 1000).  Time-major layouts, row = t*B + b:
 *   pre [n_pre,B,4H] f32 and gates [T,B,4H] bf16 have their columns INTERLEAVED (4u+g); w_hh_il = W_hh with rows interleaved alike;
 *   dgates [T,B,4H] bf16 comes out in natural gate order (g*H+u), the layout of the time-batched weight-gradient products;
 *   w_hh_t = W_hh^T [H,4H] bf16;  cells [T,B,H] f32;  out [T,B,H] bf16;  ws >= s2vt_lstm_steps_bwd_ws_bytes(B,H) scratch
 *   (running dL/dc, split-K partial sums: the backward product, K = 4H over few tiles, is split until the step fills the machine).
 * replaces: nn.LSTM (S2VTModel.py:19-22,67,77) and its autograd (train.py:124). */
int s2vt_lstm_steps_fwd_bf16(void* stream, int T, int B, int H, int n_pre, const float* pre, const float* bias_il,
                             const void* w_hh_il, void* out, void* gates, float* cells);
int64_t s2vt_lstm_steps_bwd_ws_bytes(int B, int H);
int s2vt_lstm_steps_bwd_bf16(void* stream, int T, int B, int H, int dout_t0, const float* dout, const void* gates,
                             const float* cells, const void* w_hh_t, void* dgates, void* ws);

/* ------------------------------------------------------------------ exact-grade decode on the tensor cores ("x" path)
 * fp32 operands are scaled by a power of two and split into two fp16 planes (hi, lo); a product is three
 * tcgen05.mma.kind::f16 passes (hi*lo + lo*hi + hi*hi) into one fp32 TMEM accumulator: per-term error <= 2^-21, below what an
 * fp32 FMA chain accumulates over K >= 512 terms (csrc/xdec_sm100.cu has the derivation, tests/test_gpu_xdec.py measures it). */
typedef struct s2vt_xdec_cfg {
  int32_t vocab_size, feat_dim, length, dim_hid, dim_embed, sos_ix, eos_ix;
} s2vt_xdec_cfg;

/* C[cmap(m), n] = sum_k A[m*lda + k] B[n*ldb + k] (+ bias[n]) (+ C): fp32 in, fp32 out, split on the fly.
 * ws >= s2vt_xgemm_ws_bytes(M,N,K).  replaces: nn.Linear / addmm (S2VTModel.py:54,80) at fp32-grade accuracy. */
int64_t s2vt_xgemm_ws_bytes(int M, int N, int K);
int s2vt_xgemm_f32(void* stream, int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb,
                   float* C, s2vt_rowmap cmap, const float* bias, int accumulate, void* ws);

/* Debug: %globaltimer stamps of CTA (0,0) of every following tensor-core decode kernel into buf[max_records][8] u64 (device
 * memory; NULL stops).  Returns the number of records handed out since the previous call (tools/trace_xdec.py). */
int s2vt_xdec_set_trace(void* buf, int max_records);

/* Weight preparation for the decode entry points: fp16 (hi, lo) planes of the 13 state_dict tensors (gate rows interleaved,
 * dimensions padded to multiples of 8) and the table EW = embedding . W_ih(word_rnn)[:, :E]^T  [V, 4H].
 *   params: 13 device pointers in state_dict registration order (S2VTModel.py:19-28)
 *   wbuf >= s2vt_xdec_weights_bytes(cfg): caller-owned, valid until the weights change. */
int64_t s2vt_xdec_weights_bytes(s2vt_xdec_cfg cfg);
int s2vt_xdec_prepare(void* stream, s2vt_xdec_cfg cfg, const float* const* params, void* wbuf);

/* S2VT.forward(mode='test'), S2VTModel.py:82-110: feats [B, L, F] f32 -> tokens [B, L-1] i64 (batch-major).
 * feat_linear, both LSTMs (159 + 80 steps), 79 x (word_rnn step, out_linear, argmax); the [B,V] logits are never stored.
 * Work is enqueued on `stream` and on one internal side stream joined back into `stream` before returning. */
int64_t s2vt_xdec_greedy_ws_bytes(s2vt_xdec_cfg cfg, int B);
int s2vt_xdec_greedy(void* stream, s2vt_xdec_cfg cfg, const void* wbuf, int B, const float* feats, int64_t* tokens, void* ws);

/* S2VT.forward(mode='beam_search'), S2VTModel.py:56-61,149-240, lock-step over videos x beams (same outputs as
 * s2vt_beam_search_f32).  beam_width <= 8.  check_every > 0: every that many depths the number of finished videos is copied
 * to pinned host memory; the host looks at the PREVIOUS chunk's count before enqueuing the next chunk and stops when every video has
 * finished (the stream never drains; at most one chunk of depths runs past the end, changing nothing).
 * host_flag: pinned int32[4] for those reads (may be NULL if check_every == 0). */
int64_t s2vt_xdec_beam_ws_bytes(s2vt_xdec_cfg cfg, int B, int beam_width, int max_depth);
int s2vt_xdec_beam(void* stream, s2vt_xdec_cfg cfg, const void* wbuf, int B, const float* feats, int beam_width, int max_depth,
                   int topk, const float* len_pen, int64_t* out_tokens, int32_t* out_len, void* ws, int check_every,
                   int32_t* host_flag);

#ifdef __cplusplus
}
#endif
#endif /* S2VT_B200_H_ */
