"""Drop-in Att_Baseline (attention_baseline.py:9-105): the reference's encoder-decoder "with temporal attention", same
constructor / forward(feats, targets, mode) / 22-tensor state_dict, with the arithmetic executed by libs2vt_b200.so.

    model  = Att_Baseline(vocab_size, dim_feat, length, dim_hid=512, dim_embed=512).cuda()
    logits = model(feats, targets=targets[:, :-1], mode='train')      # [B, L-1, V] float32, differentiable
    tokens = model(feats, mode='test')                                  # int64 [B, L]

What the reference actually computes is kept, quirks included: `attention()` applies softmax over a singleton dimension
(attention_baseline.py:52-54), so every frame weight is exactly 1 and the context is sum_l enc_outputs[:, l] -- one [B, 2H]
vector for all decode steps, independent of the decoder state.  Consequently the context's input product is computed once
and broadcast over the steps, and the three att_* layers receive exactly-zero gradients (as autograd gives the reference).
The reverse encoder direction runs on the same recurrence kernel over time-reversed rows; the reversal lives in the GEMM
row maps, nothing is copied.  Dropout > 0 is not supported (the reference's defaults are 0).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import nn

from . import engine_bf16 as EB
from . import lib as L
from . import ops
from .lib import dense, require_cuda, rowmap
from .model import _LinearParams

BF = torch.bfloat16

ATT_PARAM_ORDER = (
    "encoder.weight_ih_l0", "encoder.weight_hh_l0", "encoder.bias_ih_l0", "encoder.bias_hh_l0",
    "encoder.weight_ih_l0_reverse", "encoder.weight_hh_l0_reverse", "encoder.bias_ih_l0_reverse", "encoder.bias_hh_l0_reverse",
    "decoder.weight_ih_l0", "decoder.weight_hh_l0", "decoder.bias_ih_l0", "decoder.bias_hh_l0",
    "feat_linear.weight", "feat_linear.bias", "embedding.weight", "out_linear.weight", "out_linear.bias",
    "att_enc.weight", "att_enc.bias", "att_prev_hid.weight", "att_prev_hid.bias", "att_apply.weight",
)
_ATT_ZERO_GRAD = ATT_PARAM_ORDER[17:]


class _LSTMParamsN(nn.Module):
    """The tensors of a 1-layer nn.LSTM (optionally bidirectional) under nn.LSTM's names and init order
    (attention_baseline.py:23-24)."""

    def __init__(self, input_size: int, hidden_size: int, bidirectional: bool = False):
        super().__init__()
        self.input_size, self.hidden_size, self.bidirectional = input_size, hidden_size, bidirectional
        for sfx in ("", "_reverse") if bidirectional else ("",):
            self.register_parameter("weight_ih_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size, input_size)))
            self.register_parameter("weight_hh_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size, hidden_size)))
            self.register_parameter("bias_ih_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size)))
            self.register_parameter("bias_hh_l0" + sfx, nn.Parameter(torch.empty(4 * hidden_size)))
        k = 1.0 / math.sqrt(hidden_size) if hidden_size > 0 else 0.0
        for p in self.parameters():
            nn.init.uniform_(p, -k, k)

    def extra_repr(self):
        return "%d, %d, batch_first=True, bidirectional=%s (sm_100a kernels)" % (self.input_size, self.hidden_size, self.bidirectional)


class _LinearNoBias(nn.Module):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class _EmbeddingPad0(nn.Module):
    """nn.Embedding(V, E, padding_idx=0): N(0,1) init with row 0 zeroed; row 0 never receives gradient."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.num_embeddings, self.embedding_dim, self.padding_idx = num_embeddings, embedding_dim, 0
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        nn.init.normal_(self.weight)
        with torch.no_grad():
            self.weight[0].fill_(0)


def _sum_bias(P, pre, sfx=""):
    a, b = P[pre + ".bias_ih_l0" + sfx], P[pre + ".bias_hh_l0" + sfx]
    return ops.add_f32(a, b, torch.empty_like(a))


def _encode(P, feats, stash: bool):
    """feat_linear + bidirectional encoder (attention_baseline.py:63-65) and the context's decoder pre-activation.
    Rows are time-major; the reverse direction's tensors are in PROCESSING order (step s <-> frame L-1-s)."""
    B, L, F = feats.shape
    H = P["encoder.weight_hh_l0"].shape[1]
    E = P["embedding.weight"].shape[1]
    dev = feats.device
    xproj = torch.empty(L * B, H, device=dev)
    ops.gemm_f32(L * B, H, F, feats, rowmap(B, F, L * F), False, P["feat_linear.weight"], dense(F), False, xproj, dense(H),
                 bias=P["feat_linear.bias"])
    enc = {}
    for sfx in ("", "_reverse"):
        pre = torch.empty(L * B, 4 * H, device=dev)
        b = _sum_bias(P, "encoder", sfx)
        if sfx == "":
            ops.gemm_f32(L * B, 4 * H, H, xproj, dense(H), False, P["encoder.weight_ih_l0"], dense(H), False, pre, dense(4 * H), bias=b)
        else:                                                        # row (t,b) lands at processing step L-1-t
            ops.gemm_f32(L * B, 4 * H, H, xproj, dense(H), False, P["encoder.weight_ih_l0_reverse"], dense(H), False, pre,
                         rowmap(B, -B * 4 * H, 4 * H), bias=b, c_off=(L - 1) * B * 4 * H)
        out = torch.empty(L * B, H, device=dev)
        g = torch.empty(L * B, 4 * H, device=dev) if stash else None
        c = torch.empty(L * B, H, device=dev) if stash else None
        ops.lstm_fwd_f32(L, B, H, L, pre, b, P["encoder.weight_hh_l0" + sfx], out, g, c)
        ctx = torch.empty(B, H, device=dev)                          # sum over frames: the "attention" context half
        ops.colsum_f32(out, L, B * H, B * H, ctx)
        enc[sfx] = (out, g, c, ctx)
    Wd = P["decoder.weight_ih_l0"]
    ctx_pre = torch.empty(B, 4 * H, device=dev)
    ops.gemm_f32(B, 4 * H, H, enc[""][3], dense(H), False, Wd, dense(E + 2 * H), False, ctx_pre, dense(4 * H),
                 bias=_sum_bias(P, "decoder"), b_off=E)
    ops.gemm_f32(B, 4 * H, H, enc["_reverse"][3], dense(H), False, Wd, dense(E + 2 * H), False, ctx_pre, dense(4 * H),
                 accumulate=True, b_off=E + H)
    return xproj, enc, ctx_pre


def att_forward_f32(P: Dict[str, torch.Tensor], feats, targets, stash: bool):
    """Att_Baseline.forward(mode='train'), attention_baseline.py:69-84 -> (logits [B,L-1,V], saved)."""
    B, L, F = feats.shape
    H = P["encoder.weight_hh_l0"].shape[1]
    V, E = P["embedding.weight"].shape
    dev = feats.device
    R = (L - 1) * B
    xproj, enc, ctx_pre = _encode(P, feats, stash)
    emb_seq = torch.empty(R, E, device=dev)
    ops.embed_gather_f32(P["embedding.weight"], targets, 0, L - 1, B, L - 1, emb_seq, E)
    pre_d = torch.empty(R, 4 * H, device=dev)
    ops.bcast_rows_f32(ctx_pre, 0, 4 * H, B, 4 * H, L - 1, pre_d)
    ops.gemm_f32(R, 4 * H, E, emb_seq, dense(E), False, P["decoder.weight_ih_l0"], dense(E + 2 * H), False, pre_d, dense(4 * H),
                 accumulate=True)
    out_d = torch.empty(R, H, device=dev)
    g_d = torch.empty(R, 4 * H, device=dev) if stash else None
    c_d = torch.empty(R, H, device=dev) if stash else None
    ops.lstm_fwd_f32(L - 1, B, H, L - 1, pre_d, None, P["decoder.weight_hh_l0"], out_d, g_d, c_d)
    logits = torch.empty(B, L - 1, V, device=dev)
    ops.gemm_f32(R, V, H, out_d, dense(H), False, P["out_linear.weight"], dense(H), False, logits, rowmap(B, V, (L - 1) * V),
                 bias=P["out_linear.bias"])
    saved = dict(xproj=xproj, enc=enc, emb_seq=emb_seq, out_d=out_d, g_d=g_d, c_d=c_d, dims=(B, L, F, H, E, V)) if stash else None
    return logits, saved


def att_backward_f32(P, saved, feats, targets, dl, need_dfeats: bool):
    """BPTT for att_forward_f32; dl = dL/dlogits [B,L-1,V].  Returns ({name: grad}, dfeats or None)."""
    B, L, F, H, E, V = saved["dims"]
    dev = dl.device
    R = (L - 1) * B
    out_d, enc, xproj = saved["out_d"], saved["enc"], saved["xproj"]
    G = {}
    new = lambda *s: torch.empty(*s, device=dev)                                    # noqa: E731
    # ---- out_linear (dl rows are batch-major (b,t); h rows time-major (t,b))
    G["out_linear.weight"] = new(V, H)
    ops.gemm_f32(V, H, R, dl, dense(V), True, out_d, rowmap(L - 1, H, B * H), True, G["out_linear.weight"], dense(H))
    G["out_linear.bias"] = new(V)
    ops.colsum_f32(dl, R, V, V, G["out_linear.bias"])
    dout_d = new(R, H)
    ops.gemm_f32(R, H, V, dl, rowmap(B, V, (L - 1) * V), False, P["out_linear.weight"], dense(H), True, dout_d, dense(H))
    # ---- decoder
    dg_d = new(R, 4 * H)
    ops.lstm_bwd_f32(L - 1, B, H, 0, dout_d, saved["g_d"], saved["c_d"], P["decoder.weight_hh_l0"], dg_d)
    Wd = P["decoder.weight_ih_l0"]
    dctx_pre = new(B, 4 * H)                                                         # the context is shared by every step
    ops.colsum_f32(dg_d, L - 1, B * 4 * H, B * 4 * H, dctx_pre)
    gWd = new(4 * H, E + 2 * H)
    ops.gemm_f32(4 * H, E, R, dg_d, dense(4 * H), True, saved["emb_seq"], dense(E), True, gWd, dense(E + 2 * H))
    ops.gemm_f32(4 * H, H, B, dctx_pre, dense(4 * H), True, enc[""][3], dense(H), True, gWd, dense(E + 2 * H), c_off=E)
    ops.gemm_f32(4 * H, H, B, dctx_pre, dense(4 * H), True, enc["_reverse"][3], dense(H), True, gWd, dense(E + 2 * H), c_off=E + H)
    G["decoder.weight_ih_l0"] = gWd
    G["decoder.weight_hh_l0"] = new(4 * H, H)
    if L > 2:
        ops.gemm_f32(4 * H, H, (L - 2) * B, dg_d, dense(4 * H), True, out_d, dense(H), True, G["decoder.weight_hh_l0"], dense(H),
                     a_off=B * 4 * H)
    else:
        G["decoder.weight_hh_l0"].zero_()
    for k in ("decoder.bias_ih_l0", "decoder.bias_hh_l0"):
        G[k] = new(4 * H)
        ops.colsum_f32(dg_d, R, 4 * H, 4 * H, G[k])
    demb = new(R, E)
    ops.gemm_f32(R, E, 4 * H, dg_d, dense(4 * H), False, Wd, dense(E + 2 * H), True, demb, dense(E))
    gE = new(V, E)
    gE.zero_()
    ops.embed_scatter_add_f32(gE, targets, 0, L - 1, B, L - 1, demb, E)
    gE[0].zero_()                                                                    # padding_idx=0
    G["embedding.weight"] = gE
    # ---- encoder: d enc_outputs[:, l] = d context for every frame
    dxp = new(L * B, H)
    for sfx, col in (("", E), ("_reverse", E + H)):
        out, g, c, _ = enc[sfx]
        dctx = new(B, H)
        ops.gemm_f32(B, H, 4 * H, dctx_pre, dense(4 * H), False, Wd, dense(E + 2 * H), True, dctx, dense(H), b_off=col)
        dout = new(L * B, H)
        ops.bcast_rows_f32(dctx, 0, H, B, H, L, dout)
        dg = new(L * B, 4 * H)
        ops.lstm_bwd_f32(L, B, H, 0, dout, g, c, P["encoder.weight_hh_l0" + sfx], dg)
        gWih, gWhh = new(4 * H, H), new(4 * H, H)
        Wih = P["encoder.weight_ih_l0" + sfx]
        if sfx == "":
            ops.gemm_f32(4 * H, H, L * B, dg, dense(4 * H), True, xproj, dense(H), True, gWih, dense(H))
            ops.gemm_f32(L * B, H, 4 * H, dg, dense(4 * H), False, Wih, dense(H), True, dxp, dense(H))
        else:                                                                        # processing step s <-> frame L-1-s
            ops.gemm_f32(4 * H, H, L * B, dg, dense(4 * H), True, xproj, rowmap(B, -B * H, H), True, gWih, dense(H),
                         b_off=(L - 1) * B * H)
            ops.gemm_f32(L * B, H, 4 * H, dg, dense(4 * H), False, Wih, dense(H), True, dxp, rowmap(B, -B * H, H),
                         accumulate=True, c_off=(L - 1) * B * H)
        if L > 1:
            ops.gemm_f32(4 * H, H, (L - 1) * B, dg, dense(4 * H), True, out, dense(H), True, gWhh, dense(H), a_off=B * 4 * H)
        else:
            gWhh.zero_()
        G["encoder.weight_ih_l0" + sfx], G["encoder.weight_hh_l0" + sfx] = gWih, gWhh
        for k in ("encoder.bias_ih_l0" + sfx, "encoder.bias_hh_l0" + sfx):
            G[k] = new(4 * H)
            ops.colsum_f32(dg, L * B, 4 * H, 4 * H, G[k])
    # ---- feat_linear
    G["feat_linear.weight"] = new(H, F)
    ops.gemm_f32(H, F, L * B, dxp, dense(H), True, feats, rowmap(B, F, L * F), True, G["feat_linear.weight"], dense(F))
    G["feat_linear.bias"] = new(H)
    ops.colsum_f32(dxp, L * B, H, H, G["feat_linear.bias"])
    dfeats = None
    if need_dfeats:
        dfeats = new(B, L, F)
        ops.gemm_f32(L * B, F, H, dxp, dense(H), False, P["feat_linear.weight"], dense(F), True, dfeats, rowmap(B, F, L * F))
    for k in _ATT_ZERO_GRAD:                                                         # softmax over a singleton: zero gradient
        G[k] = torch.zeros_like(P[k])
    return G, dfeats


# --------------------------------------------------------------------------- tensor-core (bf16) engine
class _AttShadows:
    """bf16 mirrors of the weights (+ W_hh^T for the BPTT kernels, summed LSTM biases); rebuilt when a weight changes."""

    def __init__(self):
        self.key, self.t = None, {}

    def get(self, P):
        key = (ops.WEIGHT_EPOCH,) + tuple((p.data_ptr(), p._version) for p in P.values())
        if key == self.key:
            return self.t
        t = {}
        for name in ("feat_linear.weight", "encoder.weight_ih_l0", "encoder.weight_ih_l0_reverse", "decoder.weight_ih_l0", "out_linear.weight",
                     "embedding.weight"):
            t[name] = EB.cast(P[name], P[name].shape[0], P[name].shape[1])[0]
        for name in ("encoder.weight_hh_l0", "encoder.weight_hh_l0_reverse", "decoder.weight_hh_l0"):
            t[name], t[name + ".T"] = EB.cast(P[name], P[name].shape[0], P[name].shape[1], want_t=True)
        t["b"] = _sum_bias(P, "encoder")
        t["b_reverse"] = _sum_bias(P, "encoder", "_reverse")
        t["b_dec"] = _sum_bias(P, "decoder")
        self.key, self.t = key, t
        return t


def att_bf16_supported(H, E, F, V):
    return EB.supported(H, E, F, V)


def _side_by_side(B: int) -> bool:
    """The two encoder directions are independent: run them on two streams.  At most 7 clusters of 16 CTAs fit on the part, so the
    second direction takes two batch tiles per cluster (half the SMs at ~1.25x the step time) and both sweeps are resident at once."""
    return EB.WAVEFRONT and (B + 15) // 16 + (B + 31) // 32 <= 7


def att_forward_bf16(P, S, feats, targets, stash: bool, ce=None):
    """Att_Baseline.forward(mode='train') on tensor cores: every contraction a tcgen05 GEMM, the three recurrences (encoder forward,
    encoder reverse via the direction flag, decoder) in the persistent cluster kernel.  Returns (fp32 logits [B,L-1,V], saved); with
    ce = dict(targets_full, loss) the vocab projection is fused with the loss (time-major bf16 logits, saved['lse'])."""
    B, Lq, F = feats.shape
    H = P["encoder.weight_hh_l0"].shape[1]
    V, E = P["embedding.weight"].shape
    dev = feats.device
    R = (Lq - 1) * B
    Bp = int(L.load().s2vt_lstm_bf16_batch_pad(B))
    xb, _ = EB.cast(feats, B * Lq, F)                                              # batch-major rows (b, l)
    xproj = torch.empty(Lq * B, H, dtype=BF, device=dev)                           # time-major rows (l, b)
    EB.gemm(B * Lq, H, F, xb, F, False, S["feat_linear.weight"], F, False, xproj, rowmap(Lq, H, B * H), out_bf16=True,
            bias=P["feat_linear.bias"])
    enc = {}
    cur = torch.cuda.current_stream(dev)
    side = _side_by_side(B)
    s_rev = EB._aux_stream(dev, "att_rev") if side else cur
    bufs = {}
    for sfx in ("", "_reverse"):                                                   # (allocated on the caller's stream, used on both)
        bufs[sfx] = (torch.empty(Lq * B, 4 * H, device=dev), torch.empty(Lq * B, H, dtype=BF, device=dev),
                     torch.empty(Lq * Bp * 4 * H, dtype=BF, device=dev) if stash else None,
                     torch.empty(Lq * Bp * H, device=dev) if stash else None, torch.empty(B, H, device=dev),
                     torch.empty(B, H, dtype=BF, device=dev))
    ev0 = torch.cuda.Event()
    ev0.record(cur)
    for sfx, stream in (("", cur), ("_reverse", s_rev)):
        pre, out, g, c, ctx, ctx_bf = bufs[sfx]                                    # out: time order for both directions
        with torch.cuda.stream(stream):
            if stream is not cur:
                stream.wait_event(ev0)
            EB.gemm(Lq * B, 4 * H, H, xproj, H, False, S["encoder.weight_ih_l0" + sfx], H, False, pre, dense(4 * H), bias=S["b" + sfx])
            EB.lstm_fwd(Lq, B, H, Lq, pre, S["b" + sfx], S["encoder.weight_hh_l0" + sfx], out, g, c, reverse=(sfx != ""),
                        tiles_per_cluster=2 if (side and sfx != "") else 1)
            EB.colsum_bf16(out, Lq, B * H, B * H, ctx)                             # sum over frames (attention weights are all 1)
            rc = L.load().s2vt_cast_bf16(L.stream_ptr(dev), L.ptr(ctx), L.ptr(ctx_bf), None, B, H)
            L.check(rc, "s2vt_cast_bf16")
        enc[sfx] = (out, g, c, ctx_bf)
    if side:
        cur.wait_stream(s_rev)
    Wd = S["decoder.weight_ih_l0"]
    ctx_pre = torch.empty(B, 4 * H, device=dev)
    EB.gemm(B, 4 * H, H, enc[""][3], H, False, Wd, E + 2 * H, False, ctx_pre, dense(4 * H), bias=S["b_dec"], b_off=E)
    EB.gemm(B, 4 * H, H, enc["_reverse"][3], H, False, Wd, E + 2 * H, False, ctx_pre, dense(4 * H), accumulate=True, b_off=E + H)
    emb_seq = torch.empty(R, E, dtype=BF, device=dev)
    rc = L.load().s2vt_embed_gather_bf16(L.stream_ptr(dev), L.ptr(S["embedding.weight"]), E, L.ptr(targets), Lq - 1, B, Lq - 1, L.ptr(emb_seq), E)
    L.check(rc, "s2vt_embed_gather_bf16")
    pre_d = torch.empty(R, 4 * H, device=dev)
    ops.bcast_rows_f32(ctx_pre, 0, 4 * H, B, 4 * H, Lq - 1, pre_d)
    EB.gemm(R, 4 * H, E, emb_seq, E, False, Wd, E + 2 * H, False, pre_d, dense(4 * H), accumulate=True)
    out_d = torch.empty(R, H, dtype=BF, device=dev)
    g_d = torch.empty((Lq - 1) * Bp * 4 * H, dtype=BF, device=dev) if stash else None
    c_d = torch.empty((Lq - 1) * Bp * H, device=dev) if stash else None
    EB.lstm_fwd(Lq - 1, B, H, Lq - 1, pre_d, S["b_dec"], S["decoder.weight_hh_l0"], out_d, g_d, c_d)
    lse = None
    if ce is not None:
        # row (t,b) of the time-major logits is scored against targets_full[b, t+1]
        logits, lse = EB.vocab_ce_fwd(R, V, H, out_d, 0, S["out_linear.weight"], P["out_linear.bias"], ce["targets_full"], 1,
                                      rowmap(B, 1, ce["targets_full"].shape[1]), ce["loss"])
    else:
        logits = torch.empty(B, Lq - 1, V, device=dev)
        EB.gemm(R, V, H, out_d, H, False, S["out_linear.weight"], H, False, logits, rowmap(B, V, (Lq - 1) * V), bias=P["out_linear.bias"])
    saved = dict(xb=xb, xproj=xproj, enc=enc, emb_seq=emb_seq, out_d=out_d, g_d=g_d, c_d=c_d, lse=lse, dims=(B, Lq, F, H, E, V)) if stash else None
    return logits, saved


def att_backward_bf16(P, S, saved, targets, dl_bf, need_dfeats: bool):
    """BPTT for att_forward_bf16; dl_bf = dL/dlogits as bf16 [(L-1)B, V] in time-major row order."""
    B, Lq, F, H, E, V = saved["dims"]
    dev = dl_bf.device
    R = (Lq - 1) * B
    out_d, enc, xproj, xb = saved["out_d"], saved["enc"], saved["xproj"], saved["xb"]
    G = {}
    new = lambda *s: torch.empty(*s, device=dev)                                    # noqa: E731
    G["out_linear.weight"] = new(V, H)
    ldv = dl_bf.stride(0)                                                            # pad8(V): EB.dlogits_buffer
    EB.gemm(V, H, R, dl_bf, ldv, True, out_d, H, True, G["out_linear.weight"], dense(H))
    G["out_linear.bias"] = new(V)
    EB.colsum_bf16(dl_bf, R, V, ldv, G["out_linear.bias"])
    dout_d = new(R, H)
    EB.gemm(R, H, V, dl_bf, ldv, False, S["out_linear.weight"], H, True, dout_d, dense(H))
    dg_d = torch.empty(R, 4 * H, dtype=BF, device=dev)
    EB.lstm_bwd(Lq - 1, B, H, 0, dout_d, saved["g_d"], saved["c_d"], S["decoder.weight_hh_l0.T"], dg_d)
    Wd = S["decoder.weight_ih_l0"]
    dctx_pre = new(B, 4 * H)                                                         # the context feeds every decode step
    EB.colsum_bf16(dg_d, Lq - 1, B * 4 * H, B * 4 * H, dctx_pre)
    dctx_pre_bf = EB.cast(dctx_pre, B, 4 * H)[0]
    gWd = new(4 * H, E + 2 * H)
    EB.gemm(4 * H, E, R, dg_d, 4 * H, True, saved["emb_seq"], E, True, gWd, dense(E + 2 * H))
    EB.gemm(4 * H, H, B, dctx_pre_bf, 4 * H, True, enc[""][3], H, True, gWd, dense(E + 2 * H), c_off=E)
    EB.gemm(4 * H, H, B, dctx_pre_bf, 4 * H, True, enc["_reverse"][3], H, True, gWd, dense(E + 2 * H), c_off=E + H)
    G["decoder.weight_ih_l0"] = gWd
    G["decoder.weight_hh_l0"] = new(4 * H, H)
    EB.gemm(4 * H, H, (Lq - 2) * B, dg_d, 4 * H, True, out_d, H, True, G["decoder.weight_hh_l0"], dense(H), a_off=B * 4 * H)
    G["decoder.bias_ih_l0"], G["decoder.bias_hh_l0"] = new(4 * H), new(4 * H)
    EB.colsum_bf16(dg_d, R, 4 * H, 4 * H, G["decoder.bias_ih_l0"], G["decoder.bias_hh_l0"])
    demb = new(R, E)
    EB.gemm(R, E, 4 * H, dg_d, 4 * H, False, Wd, E + 2 * H, True, demb, dense(E))
    gE = new(V, E)
    gE.zero_()
    ops.embed_scatter_add_f32(gE, targets, 0, Lq - 1, B, Lq - 1, demb, E)
    gE[0].zero_()                                                                    # padding_idx=0
    G["embedding.weight"] = gE
    # ---- encoder: d enc_outputs[:, l] = d context for every frame; each direction contributes to d xproj
    gWf, gbf = new(H, F), new(H)
    dfeats = new(B, Lq, F) if need_dfeats else None
    # the two directions' BPTT sweeps side by side (second one on two tiles per cluster, see _side_by_side); the products that
    # accumulate into shared outputs (feat_linear, dfeats) follow on the caller's stream
    cur = torch.cuda.current_stream(dev)
    side = _side_by_side(B)
    s_rev = EB._aux_stream(dev, "att_rev") if side else cur
    pre_bufs = {sfx: (new(B, H), new(Lq * B, H), torch.empty(Lq * B, 4 * H, dtype=BF, device=dev)) for sfx in ("", "_reverse")}
    ev0 = torch.cuda.Event()
    ev0.record(cur)
    for sfx, col, stream in (("", E, cur), ("_reverse", E + H, s_rev)):
        out, g, c, _ = enc[sfx]
        dctx, dout, dg = pre_bufs[sfx]
        with torch.cuda.stream(stream):
            if stream is not cur:
                stream.wait_event(ev0)
            EB.gemm(B, H, 4 * H, dctx_pre_bf, 4 * H, False, Wd, E + 2 * H, True, dctx, dense(H), b_off=col)
            ops.bcast_rows_f32(dctx, 0, H, B, H, Lq, dout)
            EB.lstm_bwd(Lq, B, H, 0, dout, g, c, S["encoder.weight_hh_l0" + sfx + ".T"], dg, reverse=(sfx != ""),
                        tiles_per_cluster=2 if (side and sfx != "") else 1)
    if side:
        cur.wait_stream(s_rev)
    for i, (sfx, col) in enumerate((("", E), ("_reverse", E + H))):
        out, g, c, _ = enc[sfx]
        rev = sfx != ""
        dg = pre_bufs[sfx][2]
        gWih, gWhh = new(4 * H, H), new(4 * H, H)
        EB.gemm(4 * H, H, Lq * B, dg, 4 * H, True, xproj, H, True, gWih, dense(H))
        # previous hidden state in PROCESSING order: time t-1 for the forward direction, t+1 for the reverse one
        EB.gemm(4 * H, H, (Lq - 1) * B, dg, 4 * H, True, out, H, True, gWhh, dense(H), a_off=0 if rev else B * 4 * H, b_off=B * H if rev else 0)
        G["encoder.weight_ih_l0" + sfx], G["encoder.weight_hh_l0" + sfx] = gWih, gWhh
        G["encoder.bias_ih_l0" + sfx], G["encoder.bias_hh_l0" + sfx] = new(4 * H), new(4 * H)
        EB.colsum_bf16(dg, Lq * B, 4 * H, 4 * H, G["encoder.bias_ih_l0" + sfx], G["encoder.bias_hh_l0" + sfx])
        dxp = torch.empty(B * Lq, H, dtype=BF, device=dev)                            # batch-major rows, lines up with xb
        EB.gemm(Lq * B, H, 4 * H, dg, 4 * H, False, S["encoder.weight_ih_l0" + sfx], H, True, dxp, rowmap(B, H, Lq * H), out_bf16=True)
        EB.gemm(H, F, B * Lq, dxp, H, True, xb, F, True, gWf, dense(F), accumulate=(i > 0))
        if i == 0:
            EB.colsum_bf16(dxp, B * Lq, H, H, gbf)
        else:
            part = new(H)
            EB.colsum_bf16(dxp, B * Lq, H, H, part)
            ops.add_f32(gbf, part, gbf)
        if need_dfeats:
            EB.gemm(B * Lq, F, H, dxp, H, False, S["feat_linear.weight"], F, True, dfeats, dense(F), accumulate=(i > 0))
    G["feat_linear.weight"], G["feat_linear.bias"] = gWf, gbf
    for k in _ATT_ZERO_GRAD:
        G[k] = torch.zeros_like(P[k])
    return G, dfeats


class _AttTrainBf16Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, feats, targets, *params):
        P = dict(zip(ATT_PARAM_ORDER, params))
        need = any(ctx.needs_input_grad)
        ctx.S = module._shadow.get(P)
        logits, saved = att_forward_bf16(P, ctx.S, feats, targets, stash=need)
        ctx.saved, ctx.P, ctx.targets = saved, P, targets
        return logits

    @staticmethod
    def backward(ctx, dl):
        B, Lm1, V = dl.shape
        dl_tm = EB.dlogits_buffer(Lm1 * B, V, dl.device)                                # layout glue for the caller-supplied gradient
        dl_tm.view(Lm1, B, V).copy_(dl.transpose(0, 1))
        G, dfeats = att_backward_bf16(ctx.P, ctx.S, ctx.saved, ctx.targets, dl_tm, ctx.needs_input_grad[1])
        ctx.saved = None
        return (None, dfeats, None) + tuple(G[k] for k in ATT_PARAM_ORDER)


class _AttLossBf16Fn(torch.autograd.Function):
    """Fused forward + MaskCriterion (utils.py:13-26) for Att_Baseline: the vocab projection's epilogue reduces the loss statistics,
    the logits are written once as bf16 and turned into dL/dlogits in place (same kernels as S2VT.forward_loss)."""

    @staticmethod
    def forward(ctx, module, feats, targets_full, *params):
        P = dict(zip(ATT_PARAM_ORDER, params))
        need = any(ctx.needs_input_grad)
        ctx.S = module._shadow.get(P)
        Lq = feats.shape[1]
        tin = targets_full[:, :Lq - 1].contiguous()
        loss = torch.empty((), device=feats.device)
        logits, saved = att_forward_bf16(P, ctx.S, feats, tin, stash=need, ce=dict(targets_full=targets_full, loss=loss))
        ctx.saved, ctx.P, ctx.tin, ctx.tfull, ctx.logits = saved, P, tin, targets_full, logits
        return loss

    @staticmethod
    def backward(ctx, gloss):
        B, Lq, F, H, E, V = ctx.saved["dims"]
        g = gloss.contiguous().to(torch.float32)
        dl = EB.ce_dlogits_inplace(ctx.logits, (Lq - 1) * B, V, ctx.saved["lse"], ctx.tfull, 1, rowmap(B, 1, ctx.tfull.shape[1]), g)
        G, dfeats = att_backward_bf16(ctx.P, ctx.S, ctx.saved, ctx.tin, dl, ctx.needs_input_grad[1])
        ctx.saved = ctx.logits = None
        return (None, dfeats, None) + tuple(G[k] for k in ATT_PARAM_ORDER)


class _AttTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, targets, *params):
        P = dict(zip(ATT_PARAM_ORDER, params))
        need = any(ctx.needs_input_grad)
        logits, saved = att_forward_f32(P, feats, targets, stash=need)
        ctx.saved, ctx.P, ctx.feats, ctx.targets = saved, P, feats, targets
        return logits

    @staticmethod
    def backward(ctx, dl):
        G, dfeats = att_backward_f32(ctx.P, ctx.saved, ctx.feats, ctx.targets, dl.contiguous(), ctx.needs_input_grad[0])
        ctx.saved = None
        return (dfeats, None) + tuple(G[k] for k in ATT_PARAM_ORDER)


class Att_Baseline(nn.Module):
    def __init__(self, vocab_size, dim_feat, length, dim_hid=500, dim_embed=500, feat_dropout=0, out_dropout=0, sos_ix=3, eos_ix=4,
                 train_precision: str = "auto"):
        super().__init__()
        if train_precision not in ("auto", "bf16", "fp32"):
            raise ValueError("train_precision must be 'auto', 'bf16' or 'fp32'")
        self.train_precision = train_precision
        self._shadow = _AttShadows()
        if feat_dropout or out_dropout:
            raise NotImplementedError("dropout > 0 is not supported on the sm_100a path (the reference defaults are 0)")
        self.dim_feat = dim_feat
        self.length = length
        self.dim_hid = dim_hid
        self.dim_embed = dim_embed
        self.sos_ix = sos_ix
        self.eos_ix = eos_ix
        self.vocab_size = vocab_size
        # registration order == attention_baseline.py:23-33, so a seeded construction draws the reference's weights
        self.encoder = _LSTMParamsN(dim_hid, dim_hid, bidirectional=True)
        self.decoder = _LSTMParamsN(dim_hid * 2 + dim_embed, dim_hid)
        self.feat_linear = _LinearParams(dim_feat, dim_hid)
        self.embedding = _EmbeddingPad0(vocab_size, dim_embed)
        self.out_linear = _LinearParams(dim_hid, vocab_size)
        self.att_enc = _LinearParams(dim_hid * 2, dim_hid)
        self.att_prev_hid = _LinearParams(dim_hid, dim_hid)
        self.att_apply = _LinearNoBias(dim_hid, 1)

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_shadow"] = _AttShadows()
        return st

    def _use_bf16(self) -> bool:
        ok = att_bf16_supported(self.dim_hid, self.dim_embed, self.dim_feat, self.vocab_size) and self.length >= 3
        if self.train_precision == "bf16" and not ok:
            raise NotImplementedError("train_precision='bf16' needs dim_hid % 128 == 0, dim_hid <= 512 and dim_embed, dim_feat, "
                                      "vocab_size multiples of 8; use train_precision='fp32' for other shapes")
        return ok and self.train_precision != "fp32"

    def _params(self) -> Dict[str, torch.Tensor]:
        sd = dict(self.named_parameters())
        return {k: sd[k] for k in ATT_PARAM_ORDER}

    # ---- data-parallel trainer hooks (dp.DataParallelTrainer): one gradient bucket, private bf16 weight copies
    DP_BUCKETS = (("all", ATT_PARAM_ORDER),)

    def _needs_derived_shadows(self) -> bool:
        return False

    def forward_loss(self, feats, targets, mask=None):
        """criterion(model(feats, targets[:, :-1], 'train'), targets, mask) (train.py:120-122 with Att_Baseline, attention_baseline.py:
        59-84) as one scalar, without handing [B,L-1,V] logits to autograd.  Tensor-core path only (the exact path goes through
        forward(mode='train') + MaskCriterion)."""
        self._check(feats)
        if targets.dim() != 2 or targets.shape[1] < self.length:
            raise RuntimeError("targets must be [B, >= length] (got %s)" % (tuple(targets.shape),))
        P = self._params()
        if not self._use_bf16():
            from .criterion import MaskCriterion
            m = mask if mask is not None else torch.ones_like(targets, dtype=torch.float32)
            return MaskCriterion()(self(feats, targets=targets[:, :-1], mode="train"), targets[:, :self.length], m[:, :self.length])
        return _AttLossBf16Fn.apply(self, feats.contiguous(), targets[:, :self.length].contiguous().to(torch.int64), *[P[k] for k in ATT_PARAM_ORDER])

    def _check(self, feats):
        if feats.dim() != 3 or feats.shape[1] != self.length or feats.shape[2] != self.dim_feat:
            raise ValueError("feats must be [B, %d, %d] (got %s)" % (self.length, self.dim_feat, tuple(feats.shape)))
        require_cuda(feats, self.embedding.weight)
        if feats.dtype != torch.float32:
            raise ValueError("feats must be float32")

    def forward(self, feats, targets=None, mode='train'):
        self._check(feats)
        feats = feats.contiguous()
        if mode == 'train':
            if targets is None or targets.dim() != 2 or targets.shape[0] != feats.shape[0] or targets.shape[1] < self.length - 1:
                raise RuntimeError("mode='train' needs targets [B, >= length-1] (attention_baseline.py:73-75)")
            t = targets[:, :self.length - 1].contiguous().to(torch.int64)           # the loop only reads columns < L-1
            P = self._params()
            if self._use_bf16():
                return _AttTrainBf16Fn.apply(self, feats, t, *[P[k] for k in ATT_PARAM_ORDER])
            return _AttTrainFn.apply(feats, t, *[P[k] for k in ATT_PARAM_ORDER])
        elif mode == 'test':
            with torch.no_grad():
                return self._greedy(feats.detach())
        raise ValueError("unknown mode %r" % (mode,))

    def _greedy(self, feats):
        """attention_baseline.py:85-105: L greedy steps from <sos>, zero initial state, constant context."""
        P = {k: v.detach() for k, v in self._params().items()}
        B, L, _ = feats.shape
        H, E, V = self.dim_hid, self.dim_embed, self.vocab_size
        dev = feats.device
        _, _, ctx_pre = _encode(P, feats, False)
        pre = torch.empty(L * B, 4 * H, device=dev)
        ops.bcast_rows_f32(ctx_pre, 0, 4 * H, B, 4 * H, L, pre)
        w_cat = torch.cat([P["decoder.weight_ih_l0"][:, :E], P["decoder.weight_hh_l0"]], dim=1).contiguous()
        h = torch.zeros(B, H, device=dev)
        c = torch.zeros(B, H, device=dev)
        tokens = torch.empty(B, L, dtype=torch.int64, device=dev)
        ops.greedy_decode_f32(B, H, E, V, L, int(self.sos_ix), pre, 0, w_cat, P["embedding.weight"], P["out_linear.weight"],
                              P["out_linear.bias"], h, c, tokens)
        return tokens
