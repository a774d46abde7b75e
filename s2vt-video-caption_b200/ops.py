"""Thin tensor-level wrappers over the C ABI (one Python function per entry point).

Every wrapper takes torch CUDA tensors (+ element offsets where a sub-matrix is addressed), checks
dtype/contiguity, and forwards raw pointers on the current stream.  No arithmetic happens here.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as L
from .lib import RowMap, dense, rowmap


# ---------------------------------------------------------------------------------------------------------
# Optional per-call timing (bench.py's roofline leg): CUDA events on the launching stream around each C call.
_PROFILE = None
WEIGHT_EPOCH = 0          # bumped whenever an Adam kernel rewrites weights behind autograd's back (bf16 shadows key on it)


class profile:
    """with ops.profile() as prof: ...   ->  prof.summary() = {tag: (calls, total_ms, flops, bytes)} after a sync."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _PROFILE
        _PROFILE = self
        return self

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, e0, e1, flops, nbytes in self.records:
            c, ms, f, b = out.get(tag, (0, 0.0, 0.0, 0.0))
            out[tag] = (c + 1, ms + e0.elapsed_time(e1), f + flops, b + nbytes)
        return out


def serialising_profiler_attached() -> bool:
    """True under Nsight Compute, which replays kernels one at a time: kernels that wait for each other on the device (the wave front)
    must then be enqueued in dependency order, and must not sit in one CUDA graph (S2VT_WAVEFRONT=safe, eager steps)."""
    import os
    inj = os.environ.get("CUDA_INJECTION64_PATH", "").lower()
    return "NV_COMPUTE_PROFILER_PERFWORKS_DIR" in os.environ or "nsight-compute" in inj or "libcuda-injection" in inj


MARKS = None             # tools/timeline_step.py: dict(buf=int64 device tensor, names=[...]) -> %globaltimer stamps around every call


def _mark(name):
    m = MARKS
    i = len(m["names"])
    if i < m["buf"].numel():
        m["names"].append(name)
        L.load().s2vt_timestamp(L.stream_ptr(m["buf"].device), L.ptr(m["buf"], i))


class _timed:
    def __init__(self, tag, flops=0.0, nbytes=0.0):
        self.tag, self.flops, self.nbytes = tag, flops, nbytes

    def __enter__(self):
        if MARKS is not None:
            _mark("B " + self.tag)
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if MARKS is not None:
            _mark("E " + self.tag)
        if _PROFILE is not None and exc[0] is None:
            self.e1.record()
            _PROFILE.records.append((self.tag, self.e0, self.e1, self.flops, self.nbytes))


def _f32(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("%s must be a contiguous float32 tensor (got %s, contiguous=%s)" % (name, t.dtype, t.is_contiguous()))


def gemm_f32(M: int, N: int, K: int, A: torch.Tensor, amap: RowMap, a_trans: bool, B: torch.Tensor, bmap: RowMap,
             b_trans: bool, C: torch.Tensor, cmap: RowMap, bias: Optional[torch.Tensor] = None, accumulate: bool = False,
             a_off: int = 0, b_off: int = 0, c_off: int = 0, split_k: int = 1, split_stride: int = 0) -> None:
    _f32(A, "A"); _f32(B, "B"); _f32(C, "C")
    L.require_cuda(A, B, C, bias)
    with _timed("gemm_f32", 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)):
        rc = L.load().s2vt_gemm_f32(L.stream_ptr(A.device), M, N, K, L.ptr(A, a_off), amap, int(a_trans), L.ptr(B, b_off), bmap,
                                    int(b_trans), L.ptr(C, c_off), cmap, L.ptr(bias), int(accumulate), split_k, split_stride)
    L.check(rc, "s2vt_gemm_f32")


def add_f32(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _f32(a, "a"); _f32(b, "b"); _f32(out, "out")
    L.require_cuda(a, b, out)
    L.check(L.load().s2vt_add_f32(L.stream_ptr(a.device), L.ptr(a), L.ptr(b), L.ptr(out), a.numel()), "s2vt_add_f32")
    return out


def bcast_rows_f32(src: torch.Tensor, src_off: int, src_ld: int, B: int, N: int, n_t: int, dst: torch.Tensor) -> torch.Tensor:
    _f32(src, "src"); _f32(dst, "dst")
    L.require_cuda(src, dst)
    L.check(L.load().s2vt_bcast_rows_f32(L.stream_ptr(src.device), L.ptr(src, src_off), src_ld, B, N, n_t, L.ptr(dst)), "s2vt_bcast_rows_f32")
    return dst


def lstm_fwd_f32(T: int, B: int, H: int, n_pre: int, pre: Optional[torch.Tensor], bias_sum: torch.Tensor, w_hh: torch.Tensor,
                 out: torch.Tensor, gates: Optional[torch.Tensor] = None, cells: Optional[torch.Tensor] = None,
                 h0: Optional[torch.Tensor] = None, c0: Optional[torch.Tensor] = None, hT: Optional[torch.Tensor] = None,
                 cT: Optional[torch.Tensor] = None, pre_off: int = 0) -> None:
    _f32(w_hh, "w_hh"); _f32(out, "out")
    L.require_cuda(w_hh, out)
    lib = L.load()
    ws = torch.empty(int(lib.s2vt_lstm_ws_bytes(B, H)), dtype=torch.uint8, device=out.device)
    with _timed("lstm_fwd_f32", 2.0 * T * B * H * 4 * H, 4.0 * T * (4 * H * H + B * 10 * H)):
        rc = lib.s2vt_lstm_fwd_f32(L.stream_ptr(out.device), T, B, H, n_pre, L.ptr(pre, pre_off), L.ptr(bias_sum), L.ptr(w_hh),
                                   L.ptr(h0), L.ptr(c0), L.ptr(out), L.ptr(gates), L.ptr(cells), L.ptr(hT), L.ptr(cT), L.ptr(ws))
    L.check(rc, "s2vt_lstm_fwd_f32")


def lstm_bwd_f32(T: int, B: int, H: int, dout_t0: int, dout: Optional[torch.Tensor], gates: torch.Tensor, cells: torch.Tensor,
                 w_hh: torch.Tensor, dgates: torch.Tensor) -> None:
    lib = L.load()
    L.require_cuda(gates, cells, w_hh, dgates)
    ws = torch.empty(int(lib.s2vt_lstm_ws_bytes(B, H)), dtype=torch.uint8, device=gates.device)
    with _timed("lstm_bwd_f32", 2.0 * T * B * H * 4 * H, 4.0 * T * (4 * H * H + B * 11 * H)):
        rc = lib.s2vt_lstm_bwd_f32(L.stream_ptr(gates.device), T, B, H, dout_t0, L.ptr(dout), L.ptr(gates), L.ptr(cells),
                                   L.ptr(w_hh), L.ptr(dgates), L.ptr(ws))
    L.check(rc, "s2vt_lstm_bwd_f32")


def embed_gather_f32(table: torch.Tensor, ids: torch.Tensor, ids_off: int, ids_ld: int, B: int, n_t: int, out: torch.Tensor,
                     out_ld: int, out_off: int = 0) -> None:
    _f32(table, "table")
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        raise ValueError("ids must be contiguous int64")
    L.require_cuda(table, ids, out)
    rc = L.load().s2vt_embed_gather_f32(L.stream_ptr(out.device), L.ptr(table), table.shape[1], L.ptr(ids, ids_off), ids_ld, B, n_t,
                                        L.ptr(out, out_off), out_ld)
    L.check(rc, "s2vt_embed_gather_f32")


def embed_scatter_add_f32(grad_table: torch.Tensor, ids: torch.Tensor, ids_off: int, ids_ld: int, B: int, n_t: int,
                          src: torch.Tensor, src_ld: int, src_off: int = 0) -> None:
    _f32(grad_table, "grad_table")
    L.require_cuda(grad_table, ids, src)
    rc = L.load().s2vt_embed_scatter_add_f32(L.stream_ptr(src.device), L.ptr(grad_table), grad_table.shape[1], L.ptr(ids, ids_off),
                                             ids_ld, B, n_t, L.ptr(src, src_off), src_ld)
    L.check(rc, "s2vt_embed_scatter_add_f32")


def colsum_f32(X: torch.Tensor, M: int, N: int, ld: int, out: torch.Tensor, accumulate: bool = False, x_off: int = 0) -> None:
    L.require_cuda(X, out)
    rc = L.load().s2vt_colsum_f32(L.stream_ptr(X.device), L.ptr(X, x_off), M, N, ld, L.ptr(out), int(accumulate))
    L.check(rc, "s2vt_colsum_f32")


def ce_f32(logits: torch.Tensor, R: int, V: int, targets: torch.Tensor, t_off: int, tmap: RowMap, loss: torch.Tensor,
           dlogits: Optional[torch.Tensor] = None, gscale: Optional[torch.Tensor] = None) -> None:
    _f32(logits, "logits")
    if targets.dtype != torch.int64 or not targets.is_contiguous():
        raise ValueError("targets must be contiguous int64")
    L.require_cuda(logits, targets, loss, dlogits, gscale)
    row_loss = torch.empty(R, dtype=torch.float32, device=logits.device)
    with _timed("ce_f32", 0.0, 4.0 * R * V * (2 if dlogits is not None else 1)):
        rc = L.load().s2vt_ce_f32(L.stream_ptr(logits.device), L.ptr(logits), R, V, L.ptr(targets, t_off), tmap, L.ptr(row_loss),
                                  L.ptr(loss), L.ptr(dlogits), L.ptr(gscale))
    L.check(rc, "s2vt_ce_f32")


def adam_f32(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, beta1: float, beta2: float,
             eps: float, step: int, grad_scale: float = 1.0, bf16_copy: Optional[torch.Tensor] = None) -> None:
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1
    L.require_cuda(p, g, m, v)
    with _timed("adam_f32", 0.0, 28.0 * p.numel()):
        rc = L.load().s2vt_adam_f32(L.stream_ptr(p.device), L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, beta1, beta2, eps,
                                    step, grad_scale, L.ptr(bf16_copy))
    L.check(rc, "s2vt_adam_f32")


def adam_prepare(step_dev: torch.Tensor, lr_dev: torch.Tensor, beta1: float, beta2: float, hyper_dev: torch.Tensor) -> None:
    rc = L.load().s2vt_adam_prepare(L.stream_ptr(step_dev.device), L.ptr(step_dev), L.ptr(lr_dev), beta1, beta2, L.ptr(hyper_dev))
    L.check(rc, "s2vt_adam_prepare")


def adam_f32_dev(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, beta1: float, beta2: float, eps: float,
                 hyper_dev: torch.Tensor, grad_scale: float = 1.0, bf16_copy: Optional[torch.Tensor] = None) -> None:
    """adam_f32 with this step's scalars read from device memory (adam_prepare): replayable inside a CUDA graph."""
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1
    L.require_cuda(p, g, m, v, hyper_dev)
    with _timed("adam_f32", 0.0, 28.0 * p.numel()):
        rc = L.load().s2vt_adam_f32_dev(L.stream_ptr(p.device), L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), beta1, beta2, eps,
                                        L.ptr(hyper_dev), grad_scale, L.ptr(bf16_copy))
    L.check(rc, "s2vt_adam_f32_dev")


def greedy_decode_f32(B: int, H: int, E: int, V: int, n_steps: int, sos_ix: int, pre2_vid: torch.Tensor, pre_off: int,
                      w_cat: torch.Tensor, emb: torch.Tensor, w_out: torch.Tensor, b_out: torch.Tensor, h2: torch.Tensor,
                      c2: torch.Tensor, tokens: torch.Tensor) -> None:
    lib = L.load()
    L.require_cuda(pre2_vid, w_cat, emb, w_out, b_out, h2, c2, tokens)
    ws = torch.empty(int(lib.s2vt_greedy_ws_bytes(B, H, E, V)), dtype=torch.uint8, device=h2.device)
    rc = lib.s2vt_greedy_decode_f32(L.stream_ptr(h2.device), B, H, E, V, n_steps, sos_ix, L.ptr(pre2_vid, pre_off), L.ptr(w_cat),
                                    L.ptr(emb), L.ptr(w_out), L.ptr(b_out), L.ptr(h2), L.ptr(c2), L.ptr(tokens), L.ptr(ws))
    L.check(rc, "s2vt_greedy_decode_f32")


def beam_search_f32(B: int, H: int, E: int, V: int, beam_width: int, max_depth: int, topk: int, sos_ix: int, eos_ix: int,
                    state: torch.Tensor, bias1: torch.Tensor, w_hh1: torch.Tensor, w_cat2: torch.Tensor, bias2: torch.Tensor,
                    emb: torch.Tensor, w_out: torch.Tensor, b_out: torch.Tensor, len_pen: torch.Tensor, out_tokens: torch.Tensor,
                    out_len: torch.Tensor) -> None:
    lib = L.load()
    L.require_cuda(state, bias1, w_hh1, w_cat2, bias2, emb, w_out, b_out, len_pen, out_tokens, out_len)
    ws = torch.empty(int(lib.s2vt_beam_ws_bytes(B, H, E, V, beam_width, max_depth, topk)), dtype=torch.uint8, device=state.device)
    rc = lib.s2vt_beam_search_f32(L.stream_ptr(state.device), B, H, E, V, beam_width, max_depth, topk, sos_ix, eos_ix, L.ptr(state),
                                  L.ptr(bias1), L.ptr(w_hh1), L.ptr(w_cat2), L.ptr(bias2), L.ptr(emb), L.ptr(w_out), L.ptr(b_out),
                                  L.ptr(len_pen), L.ptr(out_tokens), L.ptr(out_len), L.ptr(ws))
    L.check(rc, "s2vt_beam_search_f32")


# ---------------------------------------------------------------------------------------------------------
# exact-grade tensor-core path (csrc/xdec_sm100.cu)
def xgemm_f32(M: int, N: int, K: int, A: torch.Tensor, lda: int, B: torch.Tensor, ldb: int, C: torch.Tensor, cmap: RowMap,
              bias: Optional[torch.Tensor] = None, accumulate: bool = False) -> None:
    """C = A . B^T (+ bias) (+ C) with fp32 operands split into fp16 (hi, lo) planes on the fly; tcgen05, fp32-grade accuracy."""
    _f32(A, "A"); _f32(B, "B"); _f32(C, "C")
    L.require_cuda(A, B, C, bias)
    lib = L.load()
    ws = torch.empty(int(lib.s2vt_xgemm_ws_bytes(M, N, K)), dtype=torch.uint8, device=A.device)
    with _timed("xgemm_f32", 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)):
        rc = lib.s2vt_xgemm_f32(L.stream_ptr(A.device), M, N, K, L.ptr(A), lda, L.ptr(B), ldb, L.ptr(C), cmap, L.ptr(bias),
                                int(accumulate), L.ptr(ws))
    L.check(rc, "s2vt_xgemm_f32")


def xdec_cfg(V: int, F: int, Lq: int, H: int, E: int, sos: int, eos: int) -> "L.XdecCfg":
    return L.XdecCfg(int(V), int(F), int(Lq), int(H), int(E), int(sos), int(eos))


def xdec_prepare(cfg, params) -> torch.Tensor:
    """fp16-plane copies of the 13 weights + the embedding-product table; returns the caller-owned weight buffer."""
    import ctypes as C
    lib = L.load()
    L.require_cuda(*params)
    for p in params:
        _f32(p, "parameter")
    dev = params[0].device
    wbuf = torch.empty(int(lib.s2vt_xdec_weights_bytes(cfg)), dtype=torch.uint8, device=dev)
    arr = (C.c_void_p * 13)(*[p.data_ptr() for p in params])
    with _timed("xdec_prepare"):
        rc = lib.s2vt_xdec_prepare(L.stream_ptr(dev), cfg, arr, L.ptr(wbuf))
    L.check(rc, "s2vt_xdec_prepare")
    return wbuf


def xdec_greedy(cfg, wbuf: torch.Tensor, feats: torch.Tensor, tokens: torch.Tensor, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    _f32(feats, "feats")
    L.require_cuda(wbuf, feats, tokens)
    lib = L.load()
    B = feats.shape[0]
    need = int(lib.s2vt_xdec_greedy_ws_bytes(cfg, B))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=feats.device)
    flops = 2.0 * B * (cfg.length * cfg.feat_dim * cfg.dim_hid + (5 * cfg.length - 3) * 4 * cfg.dim_hid * cfg.dim_hid
                       + (cfg.length - 1) * cfg.dim_hid * cfg.vocab_size)
    with _timed("xdec_greedy", flops):
        rc = lib.s2vt_xdec_greedy(L.stream_ptr(feats.device), cfg, L.ptr(wbuf), B, L.ptr(feats), L.ptr(tokens), L.ptr(ws))
    L.check(rc, "s2vt_xdec_greedy")
    return ws


def xdec_beam(cfg, wbuf: torch.Tensor, feats: torch.Tensor, beam_width: int, max_depth: int, topk: int, len_pen: torch.Tensor,
              out_tokens: torch.Tensor, out_len: torch.Tensor, ws: Optional[torch.Tensor] = None, check_every: int = 0,
              host_flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    _f32(feats, "feats")
    L.require_cuda(wbuf, feats, len_pen, out_tokens, out_len)
    lib = L.load()
    B = feats.shape[0]
    need = int(lib.s2vt_xdec_beam_ws_bytes(cfg, B, beam_width, max_depth))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=feats.device)
    with _timed("xdec_beam"):
        rc = lib.s2vt_xdec_beam(L.stream_ptr(feats.device), cfg, L.ptr(wbuf), B, L.ptr(feats), beam_width, max_depth, topk,
                                L.ptr(len_pen), L.ptr(out_tokens), L.ptr(out_len), L.ptr(ws), int(check_every),
                                host_flag.data_ptr() if host_flag is not None else None)
    L.check(rc, "s2vt_xdec_beam")
    return ws


__all__ = [n for n in dir() if n.endswith("_f32")] + ["RowMap", "dense", "rowmap"]
