"""bf16 training engine for hidden sizes the cluster kernels do not cover (H > 512 or H % 128 != 0; e.g. the reference's constructor
default dim_hid = 500, S2VTModel.py:11, and the paper sizing H = 1000 / E = 500 of BASELINE configs[3]).

Same data flow as engine_bf16 (every dense product a TMA-fed tcgen05 GEMM, fp32 master weights mirrored as bf16), but the two
recurrences and their BPTT run one kernel launch per time step (csrc/lstm_step_bf16_sm100.cu: the step's GEMM with the LSTM cell /
gate-gradient arithmetic as its epilogue) instead of the persistent cluster sweeps, and the two layers run one after the other.  At
the batch sizes this path is meant for (256 per GPU) a step's launch is shared by hundreds of videos; DataParallelTrainer captures
the whole step -- ~650 launches -- into one CUDA graph.

Needs H % 8 == 0 and F % 8 == 0 (leading dimensions of checkpoint-layout weights); E and V are free (padded private copies).
Exposes the interface of engine_bf16 that model.py uses: ShadowCache, train_forward, train_backward.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import engine_bf16 as EB
from . import lib as L
from . import ops
from .lib import dense, rowmap

BF = torch.bfloat16


def supported(H: int, E: int, F: int, V: int) -> bool:
    return H % 8 == 0 and H >= 8 and F % 8 == 0 and E >= 1 and V >= 1


def _il(w: torch.Tensor) -> torch.Tensor:
    """gate rows g*H+u -> 4u+g (a thread of the step kernel's epilogue then owns all four gates of its units)"""
    H4 = w.shape[0]
    H = H4 // 4
    return w.reshape(4, H, -1).transpose(0, 1).reshape(H4, -1) if w.dim() == 2 else w.reshape(4, H).t().reshape(H4)


def _pad_cols(w: torch.Tensor, Kp: int) -> torch.Tensor:
    if w.shape[1] == Kp:
        return w.contiguous()
    out = torch.zeros(w.shape[0], Kp, dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w
    return out


class ShadowCache:
    """bf16 copies of the weights in the layouts this engine's kernels read; rebuilt when a weight changes (ops.WEIGHT_EPOCH or a
    parameter's version / storage).  Private buffers: the optimizer may update the fp32 masters while backward still reads these."""

    def __init__(self):
        self.key = None
        self.t: Dict[str, torch.Tensor] = {}

    def get(self, P: Dict[str, torch.Tensor], adam=None) -> Dict[str, torch.Tensor]:
        key = (ops.WEIGHT_EPOCH,) + tuple((p.data_ptr(), p._version) for p in P.values())
        if key == self.key:
            return self.t
        with torch.no_grad():
            H = P["vid_rnn.weight_hh_l0"].shape[1]
            V, E = P["embedding.weight"].shape
            Ep = EB.pad8(E)
            w1i, w1h = P["vid_rnn.weight_ih_l0"], P["vid_rnn.weight_hh_l0"]
            w2i, w2h = P["word_rnn.weight_ih_l0"], P["word_rnn.weight_hh_l0"]
            t = {
                "feat": P["feat_linear.weight"].to(BF).contiguous(),
                "ih1": w1i.to(BF).contiguous(), "ih1_il": _il(w1i).to(BF).contiguous(),
                "hh1_il": _il(w1h).to(BF).contiguous(), "hh1_t": w1h.t().to(BF).contiguous(),
                "ih2v": w2i[:, E:].to(BF).contiguous(), "ih2v_il": _il(w2i[:, E:]).to(BF).contiguous(),
                "ih2e": _pad_cols(w2i[:, :E].to(BF), Ep), "ih2e_il": _pad_cols(_il(w2i[:, :E]).to(BF), Ep),
                "hh2_il": _il(w2h).to(BF).contiguous(), "hh2_t": w2h.t().to(BF).contiguous(),
                "out": P["out_linear.weight"].to(BF).contiguous(), "emb": _pad_cols(P["embedding.weight"].to(BF), Ep),
                "b1_il": _il(P["vid_rnn.bias_ih_l0"] + P["vid_rnn.bias_hh_l0"]).contiguous(),
                "b2_il": _il(P["word_rnn.bias_ih_l0"] + P["word_rnn.bias_hh_l0"]).contiguous(),
            }
        self.key, self.t = key, t
        return t


def lstm_steps_fwd(T, B, H, n_pre, pre, bias_il, w_hh_il, out, gates, cells):
    with ops._timed("lstm_steps_fwd_bf16", 2.0 * T * B * H * 4 * H, 4.0 * T * B * 5 * H + 2.0 * T * (B * 5 * H + 4 * H * H)):
        rc = L.load().s2vt_lstm_steps_fwd_bf16(L.stream_ptr(out.device), T, B, H, n_pre, L.ptr(pre), L.ptr(bias_il), L.ptr(w_hh_il), L.ptr(out),
                                               L.ptr(gates), L.ptr(cells))
    L.check(rc, "s2vt_lstm_steps_fwd_bf16")


def lstm_steps_bwd(T, B, H, dout_t0, dout, gates, cells, w_hh_t, dgates):
    lib = L.load()
    ws = torch.empty(int(lib.s2vt_lstm_steps_bwd_ws_bytes(B, H)), dtype=torch.uint8, device=dgates.device)
    with ops._timed("lstm_steps_bwd_bf16", 2.0 * T * B * H * 4 * H, 4.0 * T * B * 5 * H + 2.0 * T * (B * 12 * H + 4 * H * H)):
        rc = lib.s2vt_lstm_steps_bwd_bf16(L.stream_ptr(dgates.device), T, B, H, dout_t0, L.ptr(dout), L.ptr(gates), L.ptr(cells),
                                               L.ptr(w_hh_t), L.ptr(dgates), L.ptr(ws))
    L.check(rc, "s2vt_lstm_steps_bwd_bf16")


def train_forward(P, S, feats, targets, stash: bool, batch_major_logits: bool, ce=None):
    """S2VT.forward(mode='train') (S2VTModel.py:48-81); same contract as engine_bf16.train_forward."""
    B, Lq, F = feats.shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    V, E = P["embedding.weight"].shape
    Ep = EB.pad8(E)
    T = 2 * Lq - 1
    R = (Lq - 1) * B
    dev = feats.device
    xb = feats.view(B * Lq, F) if feats.dtype == BF else EB.cast(feats, B * Lq, F)[0]
    xproj = torch.empty(Lq * B, H, dtype=BF, device=dev)                         # time-major rows (l, b)
    EB.gemm(B * Lq, H, F, xb, F, False, S["feat"], F, False, xproj, rowmap(Lq, H, B * H), out_bf16=True, bias=P["feat_linear.bias"])
    pre1 = torch.empty(Lq * B, 4 * H, device=dev)                                # gate columns interleaved (4u+g), as are b*_il
    EB.gemm(Lq * B, 4 * H, H, xproj, H, False, S["ih1_il"], H, False, pre1, dense(4 * H), bias=S["b1_il"])
    out1 = torch.empty(T * B, H, dtype=BF, device=dev)
    g1 = torch.empty(T * B, 4 * H, dtype=BF, device=dev) if stash else None
    c1 = torch.empty(T * B, H, device=dev)
    lstm_steps_fwd(T, B, H, Lq, pre1, S["b1_il"], S["hh1_il"], out1, g1, c1)
    emb_seq = torch.empty(R, Ep, dtype=BF, device=dev)
    rc = L.load().s2vt_embed_gather_bf16(L.stream_ptr(dev), L.ptr(S["emb"]), Ep, L.ptr(targets), Lq - 1, B, Lq - 1, L.ptr(emb_seq), Ep)
    L.check(rc, "s2vt_embed_gather_bf16")
    pre2 = torch.empty(T * B, 4 * H, device=dev)
    EB.gemm(T * B, 4 * H, H, out1, H, False, S["ih2v_il"], H, False, pre2, dense(4 * H), bias=S["b2_il"])
    EB.gemm(R, 4 * H, Ep, emb_seq, Ep, False, S["ih2e_il"], Ep, False, pre2, dense(4 * H), accumulate=True, c_off=Lq * B * 4 * H)
    out2 = torch.empty(T * B, H, dtype=BF, device=dev)
    g2 = torch.empty(T * B, 4 * H, dtype=BF, device=dev) if stash else None
    c2 = torch.empty(T * B, H, device=dev)
    lstm_steps_fwd(T, B, H, T, pre2, S["b2_il"], S["hh2_il"], out2, g2, c2)
    lse = None
    if ce is not None:
        logits, lse = EB.vocab_ce_fwd(R, V, H, out2, Lq * B * H, S["out"], P["out_linear.bias"], ce["targets_full"], ce["t_off"], ce["tmap"],
                                      ce["loss"])
    else:
        if batch_major_logits:
            logits = torch.empty(B, Lq - 1, V, device=dev)
            cmap = rowmap(B, V, (Lq - 1) * V)
        else:
            logits = torch.empty(R, V, device=dev)
            cmap = dense(V)
        EB.gemm(R, V, H, out2, H, False, S["out"], H, False, logits, cmap, bias=P["out_linear.bias"], a_off=Lq * B * H)
    saved = dict(xb=xb, xproj=xproj, out1=out1, g1=g1, c1=c1, out2=out2, g2=g2, c2=c2, emb_seq=emb_seq, lse=lse,
                 dims=(B, Lq, F, H, E, V, T)) if stash else None
    return logits, saved


def train_backward(P, S, saved, targets, dl_bf: torch.Tensor, need_dfeats: bool, gout: Optional[Dict[str, torch.Tensor]] = None,
                   on_ready=None, early_out_wgrad: bool = False):
    """BPTT for train_forward; same contract as engine_bf16.train_backward (dl_bf: bf16 [(L-1)B, V] time-major, row pitch % 8 == 0).
    The bf16 weight copies are private to this engine, so a bucket is released (all-reduce + Adam may touch its fp32 masters) as soon
    as its gradients are complete."""
    B, Lq, F, H, E, V, T = saved["dims"]
    Ep = EB.pad8(E)
    dev = dl_bf.device
    R = (Lq - 1) * B
    ldv = dl_bf.stride(0)
    if ldv % 8 or dl_bf.stride(1) != 1:
        raise ValueError("dl_bf must have unit column stride and a row pitch that is a multiple of 8 (engine_bf16.dlogits_buffer)")
    out1, out2, xproj = saved["out1"], saved["out2"], saved["xproj"]
    G = {}

    def _new(name, *shape):
        return gout[name] if gout is not None else torch.empty(*shape, device=dev)

    def _ready(bucket):
        if on_ready is not None:
            on_ready(bucket)

    hdec = Lq * B * H
    # ---- out_linear and the gradient entering word_rnn
    dout2 = torch.empty(T * B, H, device=dev)                                    # rows < L*B never read (dout_t0 = L)
    EB.gemm(R, H, V, dl_bf, ldv, False, S["out"], H, True, dout2, dense(H), c_off=hdec)
    gW = _new("out_linear.weight", V, H)
    EB.gemm(V, H, R, dl_bf, ldv, True, out2, H, True, gW, dense(H), b_off=hdec)
    gb = _new("out_linear.bias", V)
    EB.colsum_bf16(dl_bf, R, V, ldv, gb)
    G["out_linear.weight"], G["out_linear.bias"] = gW, gb
    _ready("out_linear")
    # ---- word_rnn
    dg2 = torch.empty(T * B, 4 * H, dtype=BF, device=dev)                         # natural gate order (g*H + u)
    lstm_steps_bwd(T, B, H, Lq, dout2, saved["g2"], saved["c2"], S["hh2_t"], dg2)
    demb = torch.empty(R, E, device=dev)
    EB.gemm(R, E, 4 * H, dg2, 4 * H, False, S["ih2e"], Ep, True, demb, dense(E), a_off=Lq * B * 4 * H)
    gE = _new("embedding.weight", V, E)
    gE.zero_()
    ops.embed_scatter_add_f32(gE, targets, 0, Lq - 1, B, Lq - 1, demb, E)
    G["embedding.weight"] = gE
    _ready("embedding")
    gWih2 = _new("word_rnn.weight_ih_l0", 4 * H, E + H)
    gWhh2 = _new("word_rnn.weight_hh_l0", 4 * H, H)
    gb2, gb2b = _new("word_rnn.bias_ih_l0", 4 * H), _new("word_rnn.bias_hh_l0", 4 * H)
    EB.gemm(4 * H, H, T * B, dg2, 4 * H, True, out1, H, True, gWih2, dense(E + H), c_off=E)
    EB.gemm(4 * H, E, R, dg2, 4 * H, True, saved["emb_seq"], Ep, True, gWih2, dense(E + H), a_off=Lq * B * 4 * H)
    EB.gemm(4 * H, H, (T - 1) * B, dg2, 4 * H, True, out2, H, True, gWhh2, dense(H), a_off=B * 4 * H)
    EB.colsum_bf16(dg2, T * B, 4 * H, 4 * H, gb2, gb2b)
    G.update({"word_rnn.weight_ih_l0": gWih2, "word_rnn.weight_hh_l0": gWhh2, "word_rnn.bias_ih_l0": gb2, "word_rnn.bias_hh_l0": gb2b})
    dout1 = torch.empty(T * B, H, device=dev)
    EB.gemm(T * B, H, 4 * H, dg2, 4 * H, False, S["ih2v"], H, True, dout1, dense(H))
    _ready("word_rnn")
    # ---- vid_rnn
    dg1 = torch.empty(T * B, 4 * H, dtype=BF, device=dev)
    lstm_steps_bwd(T, B, H, 0, dout1, saved["g1"], saved["c1"], S["hh1_t"], dg1)
    gWih1 = _new("vid_rnn.weight_ih_l0", 4 * H, H)
    gWhh1 = _new("vid_rnn.weight_hh_l0", 4 * H, H)
    gb1, gb1b = _new("vid_rnn.bias_ih_l0", 4 * H), _new("vid_rnn.bias_hh_l0", 4 * H)
    EB.gemm(4 * H, H, Lq * B, dg1, 4 * H, True, xproj, H, True, gWih1, dense(H))
    EB.gemm(4 * H, H, (T - 1) * B, dg1, 4 * H, True, out1, H, True, gWhh1, dense(H), a_off=B * 4 * H)
    EB.colsum_bf16(dg1, T * B, 4 * H, 4 * H, gb1, gb1b)
    G.update({"vid_rnn.weight_ih_l0": gWih1, "vid_rnn.weight_hh_l0": gWhh1, "vid_rnn.bias_ih_l0": gb1, "vid_rnn.bias_hh_l0": gb1b})
    # d xproj written back in batch-major row order so that it lines up with the bf16 features
    dxp = torch.empty(B * Lq, H, dtype=BF, device=dev)
    EB.gemm(Lq * B, H, 4 * H, dg1, 4 * H, False, S["ih1"], H, True, dxp, rowmap(B, H, Lq * H), out_bf16=True)
    _ready("vid_rnn")
    # ---- feat_linear
    gWf = _new("feat_linear.weight", H, F)
    gbf = _new("feat_linear.bias", H)
    EB.gemm(H, F, B * Lq, dxp, H, True, saved["xb"], F, True, gWf, dense(F))
    EB.colsum_bf16(dxp, B * Lq, H, H, gbf)
    G.update({"feat_linear.weight": gWf, "feat_linear.bias": gbf})
    if need_dfeats:                                                              # dataloader.py:38 makes feats require grad
        dfeats = torch.empty(B, Lq, F, device=dev)
        EB.gemm(B * Lq, F, H, dxp, H, False, S["feat"], F, True, dfeats, dense(F))
        G["feats"] = dfeats
    _ready("feat_linear")
    return G
