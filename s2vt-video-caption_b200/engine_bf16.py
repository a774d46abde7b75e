"""bf16 training engine: the S2VT train step on tensor cores.

Every dense contraction is a TMA-fed tcgen05 GEMM (bf16 operands, fp32 accumulation in TMEM); both recurrences and both
BPTT sweeps run in the persistent cluster kernels; fp32 master weights stay in the nn.Parameters and are mirrored into
bf16 shadows that are rebuilt only when the weights change.  Same mathematics and data flow as model.train_forward_f32 /
train_backward_f32 (the exact path) -- operands are bf16 instead of fp32.
"""
from __future__ import annotations

import ctypes
import os as _os
from typing import Dict, Optional

import torch

from . import lib as L
from . import ops
from .lib import dense, rowmap

BF = torch.bfloat16
# SM budgets for work launched beside two resident sweeps (96 of 148 SMs at B = 64) that must leave room for a kernel launched after it:
# persistent GEMM CTAs (one per SM; used for the embedding half of word_rnn's input product, which runs beside the forward coupling
# kernel) and CTAs of the memory-bound kernels.  With the resident coupling kernel (WAVE_SERVER, the default) the backward's gradient
# products need no budget; the per-chunk coupling path (S2VT_WAVE_SERVER=0) caps them all.
BULK_CTAS = int(_os.environ.get("S2VT_BULK_CTAS", "28"))
BULK_ELT_CTAS = int(_os.environ.get("S2VT_BULK_ELT_CTAS", "104"))


def supported(H: int, E: int, F: int, V: int) -> bool:
    """Shapes the tensor-core path covers: cluster recurrence needs H % 128 == 0 and H <= 512; TMA needs 16-byte rows, i.e. E and F
    multiples of 8 (they are leading dimensions of weights in the checkpoint layout).  The vocabulary size is free: V is a leading
    dimension only of the bf16 logits / dlogits, which are this engine's own buffers and get a row pitch of pad8(V)."""
    return H % 128 == 0 and 128 <= H <= 512 and E % 8 == 0 and F % 8 == 0 and V >= 1


def pad8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------ thin wrappers
def gemm(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, cmap, out_bf16=False, bias=None, accumulate=False, a_off=0, b_off=0, c_off=0,
         short_ctas=False, urgent=False, bulk=False):
    """short_ctas: use the one-tile-per-CTA kernel (products beside a recurrence sweep whose clusters may not be resident yet: a
    persistent CTA keeps its SM for the whole product and the sweep's 16-CTA clusters could not be placed until it retires).
    bulk: a big product beside two RESIDENT sweeps that nothing on the serial chain waits for (weight gradients): the persistent kernel
    on at most BULK_CTAS SMs.  CTAs are dispatched in launch order -- stream and graph-node priorities do not reorder them, measured
    with tools/probe_priority.py -- so a product with more CTAs than free slots would keep every later kernel, i.e. the wave front's
    coupling products, waiting until its last CTA is placed; a capped persistent grid is resident at once and leaves SMs over."""
    lib = L.load()
    dense_c = cmap.inner == 1 and cmap.stride_inner == 0
    bulk = bulk and dense_c and BULK_CTAS > 0
    short_ctas = short_ctas and not bulk
    persistent = not short_ctas and dense_c                                             # the library routes dense outputs there
    tag = "gemm_bf16_%s[%dx%dx%d %s%s%s]" % ("persist" if persistent else "tile", M, N, K, "T" if a_mn else "N", "T" if b_mn else "N",
                                             " bf16out" if out_bf16 else "")
    if short_ctas:
        lib.s2vt_gemm_bf16_set_mode(0, 0)
    elif bulk:
        lib.s2vt_gemm_bf16_set_mode(BULK_CTAS, 1)
    if urgent:                                     # (launch attribute; kept for tools/probe_priority.py -- no measurable effect)
        lib.s2vt_set_launch_priority(1)
    try:
        with ops._timed(tag, 2.0 * M * N * K, 2.0 * (M * K + N * K) + (2.0 if out_bf16 else 4.0) * M * N):
            rc = lib.s2vt_gemm_bf16(L.stream_ptr(A.device), M, N, K, L.ptr(A, a_off), lda, int(a_mn), L.ptr(B, b_off), ldb, int(b_mn),
                                    L.ptr(C, c_off), cmap, int(out_bf16), L.ptr(bias), int(accumulate))
    finally:
        if short_ctas or bulk:
            lib.s2vt_gemm_bf16_set_mode(0, 1)
        if urgent:
            lib.s2vt_set_launch_priority(0)
    L.check(rc, "s2vt_gemm_bf16")


def gemm_gated(M, N, K, A, lda, B, ldb, b_mn, C, ldc, rows, wait, wait_val, done, ready, k0=0, bias=None, accumulate=False,
               max_ctas=20, reverse_m=False, a_off=0, b_off=0, c_off=0):
    """s2vt_gemm_bf16_gated: `rows` = chunk boundaries in rows of this product (0 .. M); wait / done / ready = counter rows, chunk 0 of
    this product being their element k0."""
    arr = (ctypes.c_int * len(rows))(*rows)
    with ops._timed("gemm_bf16_gated[%dx%dx%d N%s]" % (M, N, K, "T" if b_mn else "N"), 2.0 * M * N * K, 2.0 * (M * K + N * K) + 4.0 * M * N):
        rc = L.load().s2vt_gemm_bf16_gated(L.stream_ptr(A.device), M, N, K, L.ptr(A, a_off), lda, L.ptr(B, b_off), ldb, int(b_mn),
                                           L.ptr(C, c_off), ldc, L.ptr(bias), int(accumulate), int(max_ctas), int(reverse_m),
                                           len(rows) - 1, arr, L.ptr(wait, k0), int(wait_val), L.ptr(done, k0), L.ptr(ready, k0))
    L.check(rc, "s2vt_gemm_bf16_gated")


class beside_sweeps:
    """Context for the memory-bound kernels (column sums, Adam) enqueued while two sweeps are resident: grids of at most BULK_ELT_CTAS
    CTAs (s2vt_set_bulk_cta_cap), for the reason given in gemm(bulk=True)."""

    def __init__(self, on: bool = True):
        self.on = on and BULK_ELT_CTAS > 0

    def __enter__(self):
        if self.on:
            L.load().s2vt_set_bulk_cta_cap(BULK_ELT_CTAS)
        return self

    def __exit__(self, *exc):
        if self.on:
            L.load().s2vt_set_bulk_cta_cap(0)


def cast(src: torch.Tensor, rows: int, cols: int, want_t: bool = False, want_plain: bool = True):
    dst = torch.empty(rows, cols, dtype=BF, device=src.device) if want_plain else None
    dst_t = torch.empty(cols, rows, dtype=BF, device=src.device) if want_t else None
    with ops._timed("cast_bf16", 0.0, 6.0 * rows * cols):
        rc = L.load().s2vt_cast_bf16(L.stream_ptr(src.device), L.ptr(src), L.ptr(dst), L.ptr(dst_t), rows, cols)
    L.check(rc, "s2vt_cast_bf16")
    return dst, dst_t


def _sync_args(sync):
    """sync = (bounds, signal, wait, wait_val): chunk boundaries [0, ..., T] and the u32 counter rows (tensors or None)."""
    if sync is None:
        return 0, None, None, None, 0
    bounds, signal, wait, wait_val = sync
    arr = (ctypes.c_int * len(bounds))(*bounds)
    return len(bounds) - 1, arr, L.ptr(signal), L.ptr(wait), int(wait_val)


def lstm_fwd(T, B, H, n_pre, pre, bias_sum, w_bf, out, gates, cells, hT=None, cT=None, h0=None, c0=None, reverse=False,
             pre_off=0, out_off=0, gates_off=0, cells_off=0, tiles_per_cluster=1, sync=None):
    """`*_off`: element offsets, to run a time range [t0, t1) of a longer sweep (chunked launches chained through h0/c0 -> hT/cT).
    tiles_per_cluster=2: two batch tiles share a cluster's resident weights (half the SMs).  sync: wave-front counters (_sync_args)."""
    n_sync, sync_t, signal, wait, wait_val = _sync_args(sync)
    with ops._timed("lstm_fwd_bf16", 2.0 * T * B * H * 4 * H, 4.0 * T * B * 4 * H + 2.0 * T * B * 6 * H):
        rc = L.load().s2vt_lstm_fwd_bf16_sync(L.stream_ptr(out.device), T, B, H, n_pre, L.ptr(pre, pre_off), L.ptr(bias_sum), L.ptr(w_bf),
                                              L.ptr(h0), L.ptr(c0), L.ptr(out, out_off), L.ptr(gates, gates_off), L.ptr(cells, cells_off),
                                              L.ptr(hT), L.ptr(cT), int(reverse), tiles_per_cluster, n_sync, sync_t, signal, wait, wait_val)
    L.check(rc, "s2vt_lstm_fwd_bf16")


def lstm_bwd(T, B, H, dout_t0, dout, gates, cells, w_t_bf, dgates, reverse=False, dout_off=0, gates_off=0, cells_off=0, dgates_off=0,
             dh_in=None, dc_in=None, dh_out=None, dc_out=None, has_prev=False, tiles_per_cluster=1, sync=None):
    """`*_off` (elements) + the state gradients run one time chunk of a longer sweep; sync: wave-front counters (_sync_args)."""
    n_sync, sync_t, signal, wait, wait_val = _sync_args(sync)
    with ops._timed("lstm_bwd_bf16", 2.0 * T * B * H * 4 * H, 4.0 * T * B * 3 * H + 2.0 * T * B * 8 * H):
        rc = L.load().s2vt_lstm_bwd_bf16_sync(L.stream_ptr(dgates.device), T, B, H, dout_t0, L.ptr(dout, dout_off), L.ptr(gates, gates_off),
                                              L.ptr(cells, cells_off), L.ptr(w_t_bf), L.ptr(dgates, dgates_off), int(reverse), L.ptr(dh_in),
                                              L.ptr(dc_in), L.ptr(dh_out), L.ptr(dc_out), int(has_prev), tiles_per_cluster, n_sync, sync_t,
                                              signal, wait, wait_val)
    L.check(rc, "s2vt_lstm_bwd_bf16")


def colsum_bf16(X, M, N, ld, out, out2=None, x_off=0):
    with ops._timed("colsum_bf16", 0.0, 2.0 * M * N):
        rc = L.load().s2vt_colsum_bf16(L.stream_ptr(X.device), L.ptr(X, x_off), M, N, ld, L.ptr(out), L.ptr(out2))
    L.check(rc, "s2vt_colsum_bf16")


def ce_bf16(logits, R, V, targets, t_off, tmap, loss=None, dlogits=None, gscale=None, row_lse=None, have_lse=False, omap=None):
    """omap: row map of the bf16 gradient (default: dense rows in logits order)."""
    row_loss = torch.empty(R, dtype=torch.float32, device=logits.device) if loss is not None else None
    with ops._timed("ce_bf16", 0.0, 4.0 * R * V + (2.0 * R * V if dlogits is not None else 0.0)):
        rc = L.load().s2vt_ce_bf16_mapped(L.stream_ptr(logits.device), L.ptr(logits), R, V, L.ptr(targets, t_off), tmap, L.ptr(row_loss),
                                          L.ptr(loss), L.ptr(row_lse), int(have_lse), L.ptr(dlogits), omap if omap is not None else dense(V),
                                          L.ptr(gscale))
    L.check(rc, "s2vt_ce_bf16_mapped")


# ------------------------------------------------------------------------------------------------ bf16 weight shadows
class ShadowCache:
    """bf16 mirrors of the fp32 master weights (+ W_hh^T for the BPTT kernels and the summed LSTM biases).  Rebuilt when any
    parameter's storage / version changes or an Adam kernel ran (ops.WEIGHT_EPOCH)."""

    def __init__(self):
        self.key = None
        self.t: Dict[str, torch.Tensor] = {}

    def get(self, P: Dict[str, torch.Tensor], adam=None) -> Dict[str, torch.Tensor]:
        """`adam`: the attached FusedAdam, whose kernel already wrote a bf16 copy of every weight it updated."""
        key = (ops.WEIGHT_EPOCH,) + tuple((p.data_ptr(), p._version) for p in P.values())
        if key == self.key:
            return self.t
        fused = adam.shadow_views(list(P.items())) if adam is not None else None
        t = {}
        for name in ("feat_linear.weight", "vid_rnn.weight_ih_l0", "word_rnn.weight_ih_l0", "out_linear.weight", "embedding.weight"):
            w = P[name]
            t[name] = fused[name] if fused is not None else cast(w, w.shape[0], w.shape[1])[0]
        for name in ("vid_rnn.weight_hh_l0", "word_rnn.weight_hh_l0"):
            w = P[name]
            if fused is not None:
                t[name] = fused[name]
                t[name + ".T"] = fused[name + ".T"] if name + ".T" in fused else cast(w, w.shape[0], w.shape[1], want_t=True, want_plain=False)[1]
            else:
                t[name], t[name + ".T"] = cast(w, w.shape[0], w.shape[1], want_t=True)
        for kb, layer in (("b1", "vid_rnn"), ("b2", "word_rnn")):            # (FusedAdam.refresh_derived keeps these current too)
            t[kb] = fused[kb] if fused is not None and kb in fused else \
                ops.add_f32(P[layer + ".bias_ih_l0"], P[layer + ".bias_hh_l0"], torch.empty_like(P[layer + ".bias_ih_l0"]))
        self.key, self.t = key, t
        return t


_CHAIN = {}


def _aux_stream(dev, name: str, priority: int = -1) -> torch.cuda.Stream:
    key = (name, dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _CHAIN:
        _CHAIN[key] = torch.cuda.Stream(device=dev, priority=priority)
    return _CHAIN[key]


def _chain_stream(dev) -> torch.cuda.Stream:
    """High-priority stream for the serial chain of the backward pass (dgrad product -> BPTT sweep -> dgrad product -> sweep):
    whenever a chain kernel and an off-chain kernel are both ready, the chain kernel's CTAs are placed first."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _CHAIN:
        _CHAIN[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _CHAIN[key]


# ------------------------------------------------------------------------------------------------ forward / backward
def vocab_ce_fwd(R, V, K, A, a_off, W, bias, targets_full, t_off, tmap, loss):
    """out_linear fused with the loss statistics: returns (bf16 logits [R,V] with row pitch pad8(V), row log-sum-exp [R]); `loss`
    receives the mean CE."""
    lib = L.load()
    dev = A.device
    logits = torch.empty(R, pad8(V), dtype=BF, device=dev)[:, :V]
    part = torch.empty(int(lib.s2vt_vocab_ce_ws_bytes(R, V)), dtype=torch.uint8, device=dev)
    ztgt = torch.empty(R, device=dev)
    lse = torch.empty(R, device=dev)
    row_loss = torch.empty(R, device=dev)
    with ops._timed("gemm_bf16_persist[%dx%dx%d NN bf16out +CE]" % (R, V, K), 2.0 * R * V * K, 2.0 * (R * K + V * K) + 2.0 * R * V):
        rc = lib.s2vt_vocab_ce_fwd_bf16(L.stream_ptr(dev), R, V, K, L.ptr(A, a_off), K, L.ptr(W), K, L.ptr(bias), L.ptr(logits), logits.stride(0),
                                        L.ptr(targets_full, t_off), tmap, L.ptr(part), L.ptr(ztgt), L.ptr(lse), L.ptr(row_loss), L.ptr(loss))
    L.check(rc, "s2vt_vocab_ce_fwd_bf16")
    return logits, lse


def dlogits_buffer(R: int, V: int, dev) -> torch.Tensor:
    """bf16 [R, V] view with row pitch pad8(V): what train_backward takes as dL/dlogits (TMA rows must be 16-byte aligned)."""
    return torch.empty(R, pad8(V), dtype=BF, device=dev)[:, :V]


def ce_dlogits_inplace(logits_bf, R, V, lse, targets_full, t_off, tmap, gscale):
    with ops._timed("ce_bf16", 0.0, 4.0 * R * V):
        rc = L.load().s2vt_ce_dlogits_inplace_bf16(L.stream_ptr(logits_bf.device), L.ptr(logits_bf), R, V, logits_bf.stride(0), L.ptr(lse), L.ptr(targets_full, t_off),
                                                   tmap, L.ptr(gscale))
    L.check(rc, "s2vt_ce_dlogits_inplace_bf16")
    return logits_bf


WAVEFRONT = _os.environ.get("S2VT_WAVEFRONT", "1") != "0"       # run the two layers' sweeps side by side, a time chunk apart (False: one whole sweep after the other)


MAX_SYNC = 64                # S2VT_MAX_SYNC of the header
# The product coupling the two sweeps: one resident gated launch (s2vt_gemm_bf16_gated) on this many SMs, or ("0") one launch per chunk
# released by stream memory operations.
WAVE_SERVER = _os.environ.get("S2VT_WAVE_SERVER", "1") != "0"
FWD_SERVER_CTAS = int(_os.environ.get("S2VT_FWD_SERVER_CTAS", "24"))
BWD_SERVER_CTAS = int(_os.environ.get("S2VT_BWD_SERVER_CTAS", "20"))


def _time_bounds(Lq: int, T: int, backward: bool = False):
    """Chunk boundaries of the wave front.  The trailing sweep can never be ahead of the leading one by less than a chunk plus its
    coupling tiles, at the start and at the end of the sequence alike, so chunks are short (S2VT_WAVE_CHUNK steps, default 10) and
    shorter still at both ends (3 steps to get the trailing sweep going, 4-step chunks over the last 12 steps: that is the lag with which
    it finishes) -- limited by what a chunk boundary costs the leading sweep (a fence and a barrier) and by MAX_SYNC.  No chunk
    straddles step L (the embedding half of word_rnn's input starts there).  For BPTT the same pattern is laid out from the end of the
    sequence."""
    c = max(2, int(_os.environ.get("S2VT_WAVE_CHUNK", "10")))
    c = max(c, -(-T // (MAX_SYNC - 8)))
    taper = _os.environ.get("S2VT_WAVE_TAPER", "1") != "0" and T >= 4 * c
    pos = [0] + ([3] if taper and c >= 6 else [])
    p = c // 2 + 1
    end = T - 12 if taper else T
    while p < end:
        pos.append(p)
        p += c
    if taper:
        if end - pos[-1] < 3:
            pos.pop()
        pos += [T - 12, T - 8, T - 4]
    elif T - pos[-1] < c // 2 and len(pos) > 1:
        pos.pop()
    pos.append(T)
    cut = (T - Lq) if backward else Lq                       # position of step L in processing order
    pos = sorted(set([x for x in pos if abs(x - cut) >= 3 or x in (0, T)] + [cut]))
    return [T - x for x in pos][::-1] if backward else pos


def _wave_tiles(which: str):
    """(tiles per cluster of the leading sweep, of the trailing sweep).  Two tiles per cluster halve a sweep's SMs (32 instead of 64 at
    B = 64) at 1.2-1.3x its step time; the trailing sweep must use 2 for both sweeps to be co-resident (at most 7 clusters of 16 fit)."""
    a, b = [int(x) for x in _os.environ.get("S2VT_%s_TILES" % which, "1,2").split(",")]
    return a, b


_COUNTERS = {}


def _wave_counters(dev) -> torch.Tensor:
    """[6, MAX_SYNC] u32 (as int32) device counters: rows = forward signal / ready, backward signal / ready (s2vt_lstm_fwd_bf16_sync),
    forward / backward completion scratch of the gated coupling product (s2vt_gemm_bf16_gated)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COUNTERS:
        _COUNTERS[key] = torch.zeros(6, MAX_SYNC, dtype=torch.int32, device=dev)
    return _COUNTERS[key]


def wave_ok(B: int, Lq: int) -> bool:
    return WAVEFRONT and Lq >= 16 and (B + 15) // 16 + (B + 31) // 32 <= 7


def _wavefront_forward(P, S, B, Lq, H, E, T, Bp, pre1, out1, g1, c1, emb_seq, pre2, out2, g2, c2):
    """vid_rnn and word_rnn sweeps as a wave front, each sweep ONE launch (its clusters are placed once, on an empty machine, and the
    weights enter tensor memory once).  The sequence is cut into time chunks: vid_rnn's kernel counts a chunk's finished (CTA, tile)
    pairs in a device counter; the input product of word_rnn for that chunk waits for the count on a side stream (stream memory
    operation, no SM is held), runs on the free SMs and raises the chunk's ready flag, which word_rnn's kernel -- resident from the
    start on its own clusters -- polls before it reads the chunk's pre-activations.  The arithmetic is that of two whole sweeps, so
    the result is bit-identical; the serial chain shrinks from 2 x 159 steps to 159 + one chunk."""
    dev = out1.device
    lib = L.load()
    cur = torch.cuda.current_stream(dev)
    sg, s2 = _aux_stream(dev, "wave_gemm"), _aux_stream(dev, "wave_l2")
    W2 = S["word_rnn.weight_ih_l0"]
    ctr = _wave_counters(dev)
    ctr[0:2].zero_()
    ctr[4].zero_()
    bounds = _time_bounds(Lq, T)
    n_arrive = (H // 32) * ((B + 15) // 16)
    ntl_lead, ntl_trail = _wave_tiles("FWD")
    ev0 = torch.cuda.Event()
    ev0.record(cur)
    R = (Lq - 1) * B
    lstm_fwd(T, B, H, Lq, pre1, S["b1"], S["vid_rnn.weight_hh_l0"], out1, g1, c1, tiles_per_cluster=ntl_lead, sync=(bounds, ctr[0], None, 0))
    ev_g = torch.cuda.Event()

    def trailing_sweep(after_products: bool):
        with torch.cuda.stream(s2):
            s2.wait_event(ev_g if after_products else ev0)
            lstm_fwd(T, B, H, T, pre2, S["b2"], S["word_rnn.weight_hh_l0"], out2, g2, c2, tiles_per_cluster=ntl_trail,
                     sync=(bounds, None, ctr[1], 1))

    def products():
        se = _aux_stream(dev, "wave_emb")
        with torch.cuda.stream(se):
            se.wait_event(ev0)
            # embedding half of word_rnn's input product: no dependence on vid_rnn at all, needed from step L on
            gemm(R, 4 * H, E, emb_seq, E, False, W2, E + H, False, pre2, dense(4 * H), bias=S["b2"], c_off=Lq * B * 4 * H, short_ctas=True,
                 bulk=WAVE_SERVER)
            ev_emb = torch.cuda.Event()
            ev_emb.record(se)
        with torch.cuda.stream(sg):
            sg.wait_event(ev0)
            if WAVE_SERVER:
                kL = bounds.index(Lq)
                # steps < L: pre2 = out1 W^T + b;  steps >= L: added onto the embedding half (which carries the bias)
                gemm_gated(Lq * B, 4 * H, H, out1, H, W2, E + H, False, pre2, 4 * H, [t * B for t in bounds[:kL + 1]], ctr[0], n_arrive,
                           ctr[4], ctr[1], bias=S["b2"], max_ctas=FWD_SERVER_CTAS, b_off=E)
                sg.wait_event(ev_emb)
                gemm_gated((T - Lq) * B, 4 * H, H, out1, H, W2, E + H, False, pre2, 4 * H, [(t - Lq) * B for t in bounds[kL:]], ctr[0],
                           n_arrive, ctr[4], ctr[1], k0=kL, accumulate=True, max_ctas=FWD_SERVER_CTAS, a_off=Lq * B * H, b_off=E,
                           c_off=Lq * B * 4 * H)
            else:
                for k in range(len(bounds) - 1):
                    t0, t1 = bounds[k], bounds[k + 1]
                    L.check(lib.s2vt_stream_wait_value32(sg.cuda_stream, L.ptr(ctr, k), n_arrive), "s2vt_stream_wait_value32")
                    if t0 == Lq:
                        sg.wait_event(ev_emb)
                    gemm((t1 - t0) * B, 4 * H, H, out1, H, False, W2, E + H, False, pre2, dense(4 * H), bias=S["b2"] if t1 <= Lq else None,
                         accumulate=t0 >= Lq, a_off=t0 * B * H, b_off=E, c_off=t0 * B * 4 * H, short_ctas=True)
                    L.check(lib.s2vt_stream_write_value32(sg.cuda_stream, L.ptr(ctr, MAX_SYNC + k), 1), "s2vt_stream_write_value32")
            ev_g.record(sg)

    _run_coupled(("fwd", dev.index, B, Lq, H, E), trailing_sweep, products)
    cur.wait_stream(s2)
    cur.wait_event(ev_g)


_WARM = set()


def _run_coupled(key, trailing_sweep, products):
    """Enqueue a trailing sweep (which spins on device flags) and the products that raise those flags.  Normally the sweep goes first, so
    that it is resident before its first chunk is ready.  The first time a configuration runs, the products go first and the sweep
    starts behind them: the CUDA runtime loads a kernel's code on its first launch and may have to wait for the device to drain to do
    so -- with a spinning kernel resident that would never happen.  (S2VT_WAVEFRONT=safe, or a detected Nsight Compute, keeps this
    order: profilers that serialise kernels need it.)"""
    if key in _WARM and _os.environ.get("S2VT_WAVEFRONT") != "safe" and not ops.serialising_profiler_attached():
        trailing_sweep(False)
        products()
    else:
        products()
        trailing_sweep(True)
        _WARM.add(key)


def _wavefront_backward(S, saved, dl_bf, chain, dout2, dg2, dout1, dg1, events):
    """The two BPTT sweeps as a wave front, the mirror image of _wavefront_forward: word_rnn's sweep (one launch) counts finished chunks,
    latest first; the product giving dL/d(vid_rnn output) for a chunk waits for the count on a side stream, runs on free SMs and raises
    the ready flag that vid_rnn's sweep (one launch, own clusters) polls.  `events` = (dout2 ready, dg2 complete, dout1 complete,
    dg1 complete); on return the chain stream has joined everything."""
    B, Lq, F, H, E, V, T = saved["dims"]
    dev = dl_bf.device
    lib = L.load()
    R = (Lq - 1) * B
    ev_dout2, ev_dg2, ev_dout1, ev_dg1 = events
    sg, s1 = _aux_stream(dev, "wave_bgemm"), _aux_stream(dev, "wave_bl1")
    ntl_lead, ntl_trail = _wave_tiles("BWD")
    bounds = _time_bounds(Lq, T, backward=True)
    n_arrive = (H // 32) * ((B + 15) // 16)
    W2 = S["word_rnn.weight_ih_l0"]
    ev_first = torch.cuda.Event()
    with torch.cuda.stream(chain):
        ctr = _wave_counters(dev)
        ctr[2:4].zero_()
        ctr[5].zero_()
        gemm(R, H, V, dl_bf, dl_bf.stride(0), False, S["out_linear.weight"], H, True, dout2, dense(H), c_off=Lq * B * H)
        ev_dout2.record(chain)
        lstm_bwd(T, B, H, Lq, dout2, saved["g2"], saved["c2"], S["word_rnn.weight_hh_l0.T"], dg2, tiles_per_cluster=ntl_lead,
                 sync=(bounds, ctr[2], None, 0))
        ev_dg2.record(chain)
    def trailing_sweep(after_products: bool):
        with torch.cuda.stream(s1):
            s1.wait_event(ev_dout1 if after_products else ev_dout2)
            lstm_bwd(T, B, H, 0, dout1, saved["g1"], saved["c1"], S["vid_rnn.weight_hh_l0.T"], dg1, tiles_per_cluster=ntl_trail,
                     sync=(bounds, None, ctr[3], 1))
            ev_dg1.record(s1)

    def products():
        with torch.cuda.stream(sg):
            sg.wait_event(ev_dout2)
            if WAVE_SERVER:
                gemm_gated(T * B, H, 4 * H, dg2, 4 * H, W2, E + H, True, dout1, H, [t * B for t in bounds], ctr[2], n_arrive, ctr[5], ctr[3],
                           max_ctas=BWD_SERVER_CTAS, reverse_m=True, b_off=E)
            else:
                for k in range(len(bounds) - 2, -1, -1):
                    t0, t1 = bounds[k], bounds[k + 1]
                    L.check(lib.s2vt_stream_wait_value32(sg.cuda_stream, L.ptr(ctr, 2 * MAX_SYNC + k), n_arrive), "s2vt_stream_wait_value32")
                    gemm((t1 - t0) * B, H, 4 * H, dg2, 4 * H, False, W2, E + H, True, dout1, dense(H), a_off=t0 * B * 4 * H, b_off=E,
                         c_off=t0 * B * H, short_ctas=True)
                    L.check(lib.s2vt_stream_write_value32(sg.cuda_stream, L.ptr(ctr, 3 * MAX_SYNC + k), 1), "s2vt_stream_write_value32")
                    if k == len(bounds) - 2:
                        ev_first.record(sg)
            ev_dout1.record(sg)

    _run_coupled(("bwd", dev.index, B, Lq, H, E), trailing_sweep, products)
    chain.wait_event(ev_dg1)
    chain.wait_event(ev_dout1)
    if WAVE_SERVER:
        # the bulk products of the caller's stream start once word_rnn's sweep has finished its first chunk: both sweeps and the coupling
        # kernel are resident by then, and whatever the bulk CTAs do to the dispatch queue no longer matters to the wave front
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev_dout2)
        L.check(lib.s2vt_stream_wait_value32(cur.cuda_stream, L.ptr(ctr, 2 * MAX_SYNC + len(bounds) - 2), n_arrive), "s2vt_stream_wait_value32")
        return ev_dout2
    return ev_first            # the trailing sweep is under way: from here on the free SMs belong to the weight-gradient products


def train_forward(P, S, feats, targets, stash: bool, batch_major_logits: bool, ce=None):
    """S2VT.forward(mode='train') on tensor cores (S2VTModel.py:48-81).  Returns (fp32 logits, saved); with ce = dict(targets_full,
    t_off, tmap, loss) the vocab projection is fused with the loss and (bf16 time-major logits, saved) is returned, saved['lse']
    holding the rows' log-sum-exp."""
    B, Lq, F = feats.shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    V, E = P["embedding.weight"].shape
    T = 2 * Lq - 1
    dev = feats.device
    Bp = int(L.load().s2vt_lstm_bf16_batch_pad(B))
    # batch-major rows (b, l), bf16: the features' own rounding -- a store that already holds them as bf16 hands them over as they are
    xb = feats.view(B * Lq, F) if feats.dtype == BF else cast(feats, B * Lq, F)[0]
    xproj = torch.empty(Lq * B, H, dtype=BF, device=dev)                         # time-major rows (l, b)
    gemm(B * Lq, H, F, xb, F, False, S["feat_linear.weight"], F, False, xproj, rowmap(Lq, H, B * H), out_bf16=True,
         bias=P["feat_linear.bias"])
    pre1 = torch.empty(Lq * B, 4 * H, device=dev)
    gemm(Lq * B, 4 * H, H, xproj, H, False, S["vid_rnn.weight_ih_l0"], H, False, pre1, dense(4 * H), bias=S["b1"])
    out1 = torch.empty(T * B, H, dtype=BF, device=dev)
    g1 = torch.empty(T * Bp * 4 * H, dtype=BF, device=dev) if stash else None
    c1 = torch.empty(T * Bp * H, device=dev) if stash else None
    emb_seq = torch.empty((Lq - 1) * B, E, dtype=BF, device=dev)
    rc = L.load().s2vt_embed_gather_bf16(L.stream_ptr(dev), L.ptr(S["embedding.weight"]), E, L.ptr(targets), Lq - 1, B, Lq - 1,
                                         L.ptr(emb_seq), E)
    L.check(rc, "s2vt_embed_gather_bf16")
    pre2 = torch.empty(T * B, 4 * H, device=dev)
    out2 = torch.empty(T * B, H, dtype=BF, device=dev)
    g2 = torch.empty(T * Bp * 4 * H, dtype=BF, device=dev) if stash else None
    c2 = torch.empty(T * Bp * H, device=dev) if stash else None
    R = (Lq - 1) * B
    if wave_ok(B, Lq):
        _wavefront_forward(P, S, B, Lq, H, E, T, Bp, pre1, out1, g1, c1, emb_seq, pre2, out2, g2, c2)
    else:
        lstm_fwd(T, B, H, Lq, pre1, S["b1"], S["vid_rnn.weight_hh_l0"], out1, g1, c1)
        gemm(T * B, 4 * H, H, out1, H, False, S["word_rnn.weight_ih_l0"], E + H, False, pre2, dense(4 * H), bias=S["b2"], b_off=E)
        gemm(R, 4 * H, E, emb_seq, E, False, S["word_rnn.weight_ih_l0"], E + H, False, pre2, dense(4 * H), accumulate=True,
             c_off=Lq * B * 4 * H)
        lstm_fwd(T, B, H, T, pre2, S["b2"], S["word_rnn.weight_hh_l0"], out2, g2, c2)
    lse = None
    if ce is not None:
        logits, lse = vocab_ce_fwd(R, V, H, out2, Lq * B * H, S["out_linear.weight"], P["out_linear.bias"], ce["targets_full"], ce["t_off"],
                                   ce["tmap"], ce["loss"])
    else:
        if batch_major_logits:
            logits = torch.empty(B, Lq - 1, V, device=dev)
            cmap = rowmap(B, V, (Lq - 1) * V)
        else:
            logits = torch.empty(R, V, device=dev)
            cmap = dense(V)
        gemm(R, V, H, out2, H, False, S["out_linear.weight"], H, False, logits, cmap, bias=P["out_linear.bias"], a_off=Lq * B * H)
    saved = dict(xb=xb, xproj=xproj, out1=out1, g1=g1, c1=c1, out2=out2, g2=g2, c2=c2, emb_seq=emb_seq, lse=lse,
                 dims=(B, Lq, F, H, E, V, T)) if stash else None
    return logits, saved


def train_backward(P, S, saved, targets, dl_bf: torch.Tensor, need_dfeats: bool, gout: Optional[Dict[str, torch.Tensor]] = None,
                   on_ready=None, early_out_wgrad: bool = False):
    """BPTT for train_forward: dl_bf = dL/dlogits as bf16 [(L-1)B, V] in time-major row order, row pitch a multiple of 8 elements
    (pad8(V): allocate with dlogits_buffer()).  Returns fp32 grads keyed by parameter name (+ 'feats' when requested)."""
    B, Lq, F, H, E, V, T = saved["dims"]
    ldv = dl_bf.stride(0)
    if ldv % 8 or dl_bf.stride(1) != 1:
        raise ValueError("dl_bf must have unit column stride and a row pitch that is a multiple of 8 (engine_bf16.dlogits_buffer)")
    dev = dl_bf.device
    R = (Lq - 1) * B
    out1, out2, xproj = saved["out1"], saved["out2"], saved["xproj"]
    G = {}

    def _new(name, *shape):
        return gout[name] if gout is not None else torch.empty(*shape, device=dev)

    def _ready(bucket):
        if on_ready is not None:
            on_ready(bucket)

    hdec = Lq * B * H
    gW_early = None
    if early_out_wgrad:
        # Data-parallel runs: out_linear's weight gradient (26.6 MB, the first and biggest all-reduce bucket) needs only dlogits and
        # out2, so it is produced FIRST, on the whole machine (80 us), ahead of the serial chain: its all-reduce then starts ~0.3 ms
        # earlier and the chain of five all-reduces ends with the step instead of 0.2-0.35 ms after it.  (Single GPU: no all-reduce to
        # hide, and the product is cheaper beside the sweeps, on SMs that would idle -- the default order below.)
        gW_early = _new("out_linear.weight", V, H)
        gemm(V, H, R, dl_bf, ldv, True, out2, H, True, gW_early, dense(H), b_off=hdec)
        G["out_linear.weight"] = gW_early
        _ready("out_linear:reduce")                 # all-reduce only: out_linear's weights are still to be read by the dgrad product
    # The BPTT sweeps occupy 64 of the 148 SMs and form the critical chain  dl -> dout2 -> sweep(word_rnn) -> dout1 ->
    # sweep(vid_rnn), which runs on a high-priority stream.  Every product that is NOT on that chain (weight / bias / embedding
    # gradients) stays on the caller's stream and runs beside the sweep that follows its inputs, on the other 84 SMs.
    cur = torch.cuda.current_stream(dev)
    chain = _chain_stream(dev)
    chain.wait_stream(cur)
    dout2 = torch.empty(T * B, H, device=dev)                                   # rows < L*B never read (dout_t0 = L)
    dg2 = torch.empty(T * B, 4 * H, dtype=BF, device=dev)
    dout1 = torch.empty(T * B, H, device=dev)
    dg1 = torch.empty(T * B, 4 * H, dtype=BF, device=dev)
    ev_dout2, ev_dg2, ev_dout1, ev_dg1 = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    wave = wave_ok(B, Lq)
    capped = wave and not WAVE_SERVER            # (with the resident coupling kernel the bulk products need no SM budget)
    if wave:
        # (a 13000 x 512 x 5056 weight-gradient product launched now would hold every free SM for 150 us and starve the first chunk's
        # coupling product, i.e. delay the trailing sweep by as much: it starts once that product is through)
        ev_dout2 = _wavefront_backward(S, saved, dl_bf, chain, dout2, dg2, dout1, dg1, (ev_dout2, ev_dg2, ev_dout1, ev_dg1))
    else:
        with torch.cuda.stream(chain):
            gemm(R, H, V, dl_bf, ldv, False, S["out_linear.weight"], H, True, dout2, dense(H), c_off=hdec)
            ev_dout2.record(chain)
            lstm_bwd(T, B, H, Lq, dout2, saved["g2"], saved["c2"], S["word_rnn.weight_hh_l0.T"], dg2)
            ev_dg2.record(chain)
            gemm(T * B, H, 4 * H, dg2, 4 * H, False, S["word_rnn.weight_ih_l0"], E + H, True, dout1, dense(H), b_off=E)
            ev_dout1.record(chain)
            lstm_bwd(T, B, H, 0, dout1, saved["g1"], saved["c1"], S["vid_rnn.weight_hh_l0.T"], dg1)
            ev_dg1.record(chain)
    # Everything below is off the serial chain.  Three independent groups -- out_linear | word_rnn + embedding | vid_rnn + feat_linear --
    # each on its own stream (the caller's, and two side streams): beside the sweeps they share the free SMs, and once the sweeps are
    # through, kernels that are too small to fill the machine alone (column sums, 32-tile products, per-bucket Adam) overlap.
    sB, sC, sD = _aux_stream(dev, "bulk_b", 0), _aux_stream(dev, "bulk_c", 0), _aux_stream(dev, "bulk_d", 0)
    # ---- out_linear:  dW = dl^T h,  db = colsum(dl)   (beside the word_rnn sweep; released together with it)
    cur.wait_event(ev_dout2)
    ev_bulk = torch.cuda.Event()
    ev_bulk.record(cur)
    gb = _new("out_linear.bias", V)
    if gW_early is not None:
        _ready("out_linear")                        # dout2 is done (ev_dout2): the bucket's Adam update may follow its all-reduce
    else:
        gW = _new("out_linear.weight", V, H)
        with beside_sweeps(capped):
            gemm(V, H, R, dl_bf, ldv, True, out2, H, True, gW, dense(H), b_off=hdec, short_ctas=True, bulk=capped)
            G["out_linear.weight"] = gW
            _ready("out_linear")                    # (the bias gradient belongs to the embedding bucket, dp.BUCKETS)
    # ---- word_rnn weight / bias / embedding gradients (beside the vid_rnn sweep)
    gWih2 = _new("word_rnn.weight_ih_l0", 4 * H, E + H)
    gWhh2 = _new("word_rnn.weight_hh_l0", 4 * H, H)
    gb2, gb2b = _new("word_rnn.bias_ih_l0", 4 * H), _new("word_rnn.bias_hh_l0", 4 * H)
    gE = _new("embedding.weight", V, E)
    demb = torch.empty(R, E, device=dev)
    with torch.cuda.stream(sB):
        sB.wait_event(ev_bulk)
        with beside_sweeps(capped):
            colsum_bf16(dl_bf, R, V, ldv, gb)
        G["out_linear.bias"] = gb
        sB.wait_event(ev_dg2)
        with beside_sweeps(capped):
            # embedding first: its bucket is one of the two big ones (V x E), and in a data-parallel run its all-reduce should start early
            gemm(R, E, 4 * H, dg2, 4 * H, False, S["word_rnn.weight_ih_l0"], E + H, True, demb, dense(E), a_off=Lq * B * 4 * H, short_ctas=True,
                 bulk=capped)
            gE.zero_()
            ops.embed_scatter_add_f32(gE, targets, 0, Lq - 1, B, Lq - 1, demb, E)
            G["embedding.weight"] = gE
            _ready("embedding")
            gemm(4 * H, H, T * B, dg2, 4 * H, True, out1, H, True, gWih2, dense(E + H), c_off=E, short_ctas=True, bulk=capped)
            gemm(4 * H, E, R, dg2, 4 * H, True, saved["emb_seq"], E, True, gWih2, dense(E + H), a_off=Lq * B * 4 * H, short_ctas=True,
                 bulk=capped)
            gemm(4 * H, H, (T - 1) * B, dg2, 4 * H, True, out2, H, True, gWhh2, dense(H), a_off=B * 4 * H, short_ctas=True, bulk=capped)
            colsum_bf16(dg2, T * B, 4 * H, 4 * H, gb2, gb2b)
        G.update({"word_rnn.weight_ih_l0": gWih2, "word_rnn.weight_hh_l0": gWhh2, "word_rnn.bias_ih_l0": gb2, "word_rnn.bias_hh_l0": gb2b})
        sB.wait_event(ev_dout1)
        _ready("word_rnn")                                                      # (after the last readers of word_rnn's weights)
    # ---- vid_rnn and feat_linear (the machine is free again): the chain dg1 -> d xproj -> feat_linear on one stream, vid_rnn's own
    # gradients on another
    gWih1 = _new("vid_rnn.weight_ih_l0", 4 * H, H)
    gWhh1 = _new("vid_rnn.weight_hh_l0", 4 * H, H)
    gb1, gb1b = _new("vid_rnn.bias_ih_l0", 4 * H), _new("vid_rnn.bias_hh_l0", 4 * H)
    gWf = _new("feat_linear.weight", H, F)
    gbf = _new("feat_linear.bias", H)
    dxp = torch.empty(B * Lq, H, dtype=BF, device=dev)
    dfeats = torch.empty(B, Lq, F, device=dev) if need_dfeats else None
    ev_dxp = torch.cuda.Event()
    with torch.cuda.stream(sC):
        sC.wait_event(ev_bulk)
        sC.wait_event(ev_dg1)
        # d xproj written back in batch-major row order so that it lines up with the bf16 features
        gemm(Lq * B, H, 4 * H, dg1, 4 * H, False, S["vid_rnn.weight_ih_l0"], H, True, dxp, rowmap(B, H, Lq * H), out_bf16=True)
        ev_dxp.record(sC)
        gemm(H, F, B * Lq, dxp, H, True, saved["xb"], F, True, gWf, dense(F))
        colsum_bf16(dxp, B * Lq, H, H, gbf)
        G.update({"feat_linear.weight": gWf, "feat_linear.bias": gbf})
        if need_dfeats:                                                          # dataloader.py:38 makes feats require grad
            gemm(B * Lq, F, H, dxp, H, False, S["feat_linear.weight"], F, True, dfeats, dense(F))
            G["feats"] = dfeats
    with torch.cuda.stream(sD):
        sD.wait_event(ev_bulk)
        sD.wait_event(ev_dg1)
        gemm(4 * H, H, Lq * B, dg1, 4 * H, True, xproj, H, True, gWih1, dense(H))
        gemm(4 * H, H, (T - 1) * B, dg1, 4 * H, True, out1, H, True, gWhh1, dense(H), a_off=B * 4 * H)
        colsum_bf16(dg1, T * B, 4 * H, 4 * H, gb1, gb1b)
        G.update({"vid_rnn.weight_ih_l0": gWih1, "vid_rnn.weight_hh_l0": gWhh1, "vid_rnn.bias_ih_l0": gb1, "vid_rnn.bias_hh_l0": gb1b})
        sD.wait_event(ev_dxp)
        _ready("vid_rnn")                                                       # (after the last reader of vid_rnn's weights)
    with torch.cuda.stream(sC):
        _ready("feat_linear")
    cur.wait_event(ev_dg1)
    cur.wait_stream(chain)
    cur.wait_stream(sB)
    cur.wait_stream(sC)
    cur.wait_stream(sD)
    return G
