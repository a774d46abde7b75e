"""Drop-in S2VT module: the reference's constructor / forward / state_dict contract
(S2VTModel.py:10-110, 149-240) with the arithmetic executed by libs2vt_b200.so.

    model = S2VT(vocab_size, feat_dim, length, dim_hid=512, dim_embed=512).cuda()
    logits = model(feats, targets=targets[:, :-1], mode='train')      # train.py:120
    tokens = model(feats, mode='test')                                  # eval.py:52
    sents  = model(feats, mode='beam_search')                           # eval.py:88

Only what the reference's Opt() permits is supported (LSTM, one layer, unidirectional, dropout 0);
anything else raises NotImplementedError -- there is no cuDNN / CPU fallback on this path.
"""
from __future__ import annotations

import math
import weakref
from typing import Dict, List, Optional

import torch
from torch import nn

from . import engine_bf16 as EB
from . import engine_step as ES
from . import ops
from .lib import dense, rowmap, require_cuda

PARAM_ORDER = (
    "vid_rnn.weight_ih_l0", "vid_rnn.weight_hh_l0", "vid_rnn.bias_ih_l0", "vid_rnn.bias_hh_l0",
    "word_rnn.weight_ih_l0", "word_rnn.weight_hh_l0", "word_rnn.bias_ih_l0", "word_rnn.bias_hh_l0",
    "feat_linear.weight", "feat_linear.bias", "out_linear.weight", "out_linear.bias",
    "embedding.weight",
)


# --------------------------------------------------------------------------- parameter holders
class _LSTMParams(nn.Module):
    """Holds the four tensors of a 1-layer unidirectional nn.LSTM under the same names
    (weight_ih_l0 [4H,I], weight_hh_l0 [4H,H], bias_ih_l0, bias_hh_l0; gate rows i,f,g,o) and with
    nn.LSTM's init (U(+-1/sqrt(H)) drawn in registration order), S2VTModel.py:19-22."""

    def __init__(self, input_size: int, hidden_size: int):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.weight_ih_l0 = nn.Parameter(torch.empty(4 * hidden_size, input_size))
        self.weight_hh_l0 = nn.Parameter(torch.empty(4 * hidden_size, hidden_size))
        self.bias_ih_l0 = nn.Parameter(torch.empty(4 * hidden_size))
        self.bias_hh_l0 = nn.Parameter(torch.empty(4 * hidden_size))
        k = 1.0 / math.sqrt(hidden_size) if hidden_size > 0 else 0.0
        for p in self.parameters():
            nn.init.uniform_(p, -k, k)

    def extra_repr(self):
        return "%d, %d, batch_first=True (sm_100a kernels)" % (self.input_size, self.hidden_size)


class _LinearParams(nn.Module):
    """weight [out,in] + bias [out] with nn.Linear's default init, S2VTModel.py:26-27."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_features) if in_features > 0 else 0.0
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return "in_features=%d, out_features=%d (sm_100a kernels)" % (self.in_features, self.out_features)


class _EmbeddingParams(nn.Module):
    """weight [V,E] ~ N(0,1) like nn.Embedding, S2VTModel.py:28 (dense gradient)."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        nn.init.normal_(self.weight)

    def extra_repr(self):
        return "%d, %d (sm_100a kernels)" % (self.num_embeddings, self.embedding_dim)


# --------------------------------------------------------------------------- fp32 engine
def _bias_sums(P):
    dev = P["vid_rnn.bias_ih_l0"].device
    b1 = ops.add_f32(P["vid_rnn.bias_ih_l0"], P["vid_rnn.bias_hh_l0"], torch.empty_like(P["vid_rnn.bias_ih_l0"], device=dev))
    b2 = ops.add_f32(P["word_rnn.bias_ih_l0"], P["word_rnn.bias_hh_l0"], torch.empty_like(P["word_rnn.bias_ih_l0"], device=dev))
    return b1, b2


def _encode_vid_f32(P, feats, T: int, stash: bool, b1):
    """feat_linear + vid_rnn over T steps (S2VTModel.py:52-67; T=L for beam search, S2VTModel.py:57)."""
    B, L, F = feats.shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    dev = feats.device
    xproj = torch.empty(L * B, H, device=dev)
    # time-major rows (t,b) read from the batch-major [B,L,F] input
    ops.gemm_f32(L * B, H, F, feats, rowmap(B, F, L * F), False, P["feat_linear.weight"], dense(F), False, xproj, dense(H),
                 bias=P["feat_linear.bias"])
    pre1 = torch.empty(L * B, 4 * H, device=dev)
    ops.gemm_f32(L * B, 4 * H, H, xproj, dense(H), False, P["vid_rnn.weight_ih_l0"], dense(H), False, pre1, dense(4 * H), bias=b1)
    out1 = torch.empty(T * B, H, device=dev)
    g1 = torch.empty(T * B, 4 * H, device=dev) if stash else None
    c1 = torch.empty(T * B, H, device=dev) if stash else None
    hT = torch.empty(B, H, device=dev)
    cT = torch.empty(B, H, device=dev)
    ops.lstm_fwd_f32(T, B, H, min(L, T), pre1, b1, P["vid_rnn.weight_hh_l0"], out1, g1, c1, hT=hT, cT=cT)
    return xproj, out1, g1, c1, hT, cT


def _word_pre_vid_f32(P, out1, rows: int, b2, E: int, H: int):
    """vid half of word_rnn's input product: out1 . W_ih[:, E:]^T + (b_ih + b_hh)  (S2VTModel.py:75: embedding columns first)."""
    pre2 = torch.empty(rows, 4 * H, device=out1.device)
    ops.gemm_f32(rows, 4 * H, H, out1, dense(H), False, P["word_rnn.weight_ih_l0"], dense(E + H), False, pre2, dense(4 * H),
                 bias=b2, b_off=E)
    return pre2


def train_forward_f32(P: Dict[str, torch.Tensor], feats: torch.Tensor, targets: torch.Tensor, stash: bool, batch_major_logits: bool):
    """S2VT.forward(mode='train') in exact fp32.  Returns (logits, saved)."""
    B, L, F = feats.shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    V, E = P["embedding.weight"].shape
    T = 2 * L - 1
    dev = feats.device
    b1, b2 = _bias_sums(P)
    xproj, out1, g1, c1, _, _ = _encode_vid_f32(P, feats, T, stash, b1)
    emb_seq = torch.empty((L - 1) * B, E, device=dev)
    ops.embed_gather_f32(P["embedding.weight"], targets, 0, L - 1, B, L - 1, emb_seq, E)
    pre2 = _word_pre_vid_f32(P, out1, T * B, b2, E, H)
    # embedding half only exists for the decode steps t >= L (pad_embed is zero before, S2VTModel.py:72-73)
    ops.gemm_f32((L - 1) * B, 4 * H, E, emb_seq, dense(E), False, P["word_rnn.weight_ih_l0"], dense(E + H), False, pre2,
                 dense(4 * H), accumulate=True, c_off=L * B * 4 * H)
    out2 = torch.empty(T * B, H, device=dev)
    g2 = torch.empty(T * B, 4 * H, device=dev) if stash else None
    c2 = torch.empty(T * B, H, device=dev) if stash else None
    ops.lstm_fwd_f32(T, B, H, T, pre2, b2, P["word_rnn.weight_hh_l0"], out2, g2, c2)
    R = (L - 1) * B
    if batch_major_logits:
        logits = torch.empty(B, L - 1, V, device=dev)
        cmap = rowmap(B, V, (L - 1) * V)                 # row (t,b) -> [b, t, :]
    else:
        logits = torch.empty(R, V, device=dev)
        cmap = dense(V)
    ops.gemm_f32(R, V, H, out2, dense(H), False, P["out_linear.weight"], dense(H), False, logits, cmap, bias=P["out_linear.bias"],
                 a_off=L * B * H)
    saved = dict(xproj=xproj, out1=out1, g1=g1, c1=c1, out2=out2, g2=g2, c2=c2, emb_seq=emb_seq, dims=(B, L, F, H, E, V, T)) if stash else None
    return logits, saved


def train_backward_f32(P, saved, feats, targets, dl: torch.Tensor, dl_batch_major: bool, need_dfeats: bool,
                       gout: Optional[Dict[str, torch.Tensor]] = None, on_ready=None):
    """BPTT for train_forward_f32: dl = dL/dlogits ([B,L-1,V] batch-major or [(L-1)B,V] time-major).
    Returns grads in PARAM_ORDER (+ dfeats or None).  `gout` (name -> preallocated tensor) makes the kernels write
    straight into e.g. the optimizer's flat gradient buffer; `on_ready(bucket)` is called as soon as all kernels
    producing a bucket of dp.BUCKETS have been enqueued (the data-parallel all-reduce hooks in here)."""
    B, L, F, H, E, V, T = saved["dims"]
    dev = dl.device
    R = (L - 1) * B
    out1, out2 = saved["out1"], saved["out2"]
    G = {}

    def _new(name, *shape):
        return gout[name] if gout is not None else torch.empty(*shape, device=dev)

    def _ready(bucket):
        if on_ready is not None:
            on_ready(bucket)

    hdec_off = L * B * H
    # ---- out_linear
    gW = _new("out_linear.weight", V, H)
    if dl_batch_major:
        # K runs over batch-major rows k = b*(L-1)+t; the matching h row is (L+t)*B + b
        ops.gemm_f32(V, H, R, dl, dense(V), True, out2, rowmap(L - 1, H, B * H), True, gW, dense(H), b_off=hdec_off)
        amap_rows = rowmap(B, V, (L - 1) * V)
    else:
        ops.gemm_f32(V, H, R, dl, dense(V), True, out2, dense(H), True, gW, dense(H), b_off=hdec_off)
        amap_rows = dense(V)
    G["out_linear.weight"] = gW
    gb = _new("out_linear.bias", V)
    ops.colsum_f32(dl, R, V, V, gb)
    G["out_linear.bias"] = gb
    dout2 = torch.empty(T * B, H, device=dev)          # rows < L*B are never read (dout_t0 = L)
    ops.gemm_f32(R, H, V, dl, amap_rows, False, P["out_linear.weight"], dense(H), True, dout2, dense(H), c_off=hdec_off)
    _ready("out_linear")                               # on_ready may update the bucket's weights: only after their last reader
    # ---- word_rnn
    dg2 = torch.empty(T * B, 4 * H, device=dev)
    ops.lstm_bwd_f32(T, B, H, L, dout2, saved["g2"], saved["c2"], P["word_rnn.weight_hh_l0"], dg2)
    gWih2 = _new("word_rnn.weight_ih_l0", 4 * H, E + H)
    ops.gemm_f32(4 * H, H, T * B, dg2, dense(4 * H), True, out1, dense(H), True, gWih2, dense(E + H), c_off=E)
    ops.gemm_f32(4 * H, E, R, dg2, dense(4 * H), True, saved["emb_seq"], dense(E), True, gWih2, dense(E + H), a_off=L * B * 4 * H)
    G["word_rnn.weight_ih_l0"] = gWih2
    gWhh2 = _new("word_rnn.weight_hh_l0", 4 * H, H)
    ops.gemm_f32(4 * H, H, (T - 1) * B, dg2, dense(4 * H), True, out2, dense(H), True, gWhh2, dense(H), a_off=B * 4 * H)
    G["word_rnn.weight_hh_l0"] = gWhh2
    gb2 = _new("word_rnn.bias_ih_l0", 4 * H)
    ops.colsum_f32(dg2, T * B, 4 * H, 4 * H, gb2)
    G["word_rnn.bias_ih_l0"] = gb2
    gb2b = _new("word_rnn.bias_hh_l0", 4 * H)
    ops.colsum_f32(dg2, T * B, 4 * H, 4 * H, gb2b)
    G["word_rnn.bias_hh_l0"] = gb2b
    # d input2 = dg2 . W_ih2: vid half -> dL/d output1, embedding half -> dense embedding grad
    dout1 = torch.empty(T * B, H, device=dev)
    ops.gemm_f32(T * B, H, 4 * H, dg2, dense(4 * H), False, P["word_rnn.weight_ih_l0"], dense(E + H), True, dout1, dense(H), b_off=E)
    demb = torch.empty(R, E, device=dev)
    ops.gemm_f32(R, E, 4 * H, dg2, dense(4 * H), False, P["word_rnn.weight_ih_l0"], dense(E + H), True, demb, dense(E),
                 a_off=L * B * 4 * H)
    _ready("word_rnn")
    gE = _new("embedding.weight", V, E)
    gE.zero_()
    ops.embed_scatter_add_f32(gE, targets, 0, L - 1, B, L - 1, demb, E)
    G["embedding.weight"] = gE
    _ready("embedding")
    # ---- vid_rnn
    dg1 = torch.empty(T * B, 4 * H, device=dev)
    ops.lstm_bwd_f32(T, B, H, 0, dout1, saved["g1"], saved["c1"], P["vid_rnn.weight_hh_l0"], dg1)
    gWih1 = _new("vid_rnn.weight_ih_l0", 4 * H, H)
    ops.gemm_f32(4 * H, H, L * B, dg1, dense(4 * H), True, saved["xproj"], dense(H), True, gWih1, dense(H))
    G["vid_rnn.weight_ih_l0"] = gWih1
    gWhh1 = _new("vid_rnn.weight_hh_l0", 4 * H, H)
    ops.gemm_f32(4 * H, H, (T - 1) * B, dg1, dense(4 * H), True, out1, dense(H), True, gWhh1, dense(H), a_off=B * 4 * H)
    G["vid_rnn.weight_hh_l0"] = gWhh1
    gb1 = _new("vid_rnn.bias_ih_l0", 4 * H)
    ops.colsum_f32(dg1, T * B, 4 * H, 4 * H, gb1)
    G["vid_rnn.bias_ih_l0"] = gb1
    gb1b = _new("vid_rnn.bias_hh_l0", 4 * H)
    ops.colsum_f32(dg1, T * B, 4 * H, 4 * H, gb1b)
    G["vid_rnn.bias_hh_l0"] = gb1b
    # ---- feat_linear
    dxp = torch.empty(L * B, H, device=dev)
    ops.gemm_f32(L * B, H, 4 * H, dg1, dense(4 * H), False, P["vid_rnn.weight_ih_l0"], dense(H), True, dxp, dense(H))
    _ready("vid_rnn")
    gWf = _new("feat_linear.weight", H, F)
    ops.gemm_f32(H, F, L * B, dxp, dense(H), True, feats, rowmap(B, F, L * F), True, gWf, dense(F))
    G["feat_linear.weight"] = gWf
    gbf = _new("feat_linear.bias", H)
    ops.colsum_f32(dxp, L * B, H, H, gbf)
    G["feat_linear.bias"] = gbf
    dfeats = None
    if need_dfeats:                                     # dataloader.py:38 makes feats require grad
        dfeats = torch.empty(B, L, F, device=dev)
        ops.gemm_f32(L * B, F, H, dxp, dense(H), False, P["feat_linear.weight"], dense(F), True, dfeats, rowmap(B, F, L * F))
    _ready("feat_linear")
    return [G[k] for k in PARAM_ORDER], dfeats


def _direct_grad_targets(module):
    """When an optimizer attached its flat gradient buffer (FusedAdam.attach) and no gradient is pending, backward
    writes into those views and installs them as p.grad itself (autograd then has nothing to accumulate)."""
    views = getattr(module, "_grad_views", None)
    if views is None:
        return None, None
    params = dict(module.named_parameters())
    for k in PARAM_ORDER:
        p = params[k]
        if p.grad is not None or views[k].device != p.device or not p.requires_grad:
            return None, None
    return views, getattr(module, "_on_bucket_ready", None)


def _finish_grads(module, direct, grads):
    if direct is None:
        return tuple(grads)
    params = dict(module.named_parameters())
    for k in PARAM_ORDER:
        params[k].grad = direct[k]
    return (None,) * len(PARAM_ORDER)


class _TrainLogitsFn(torch.autograd.Function):
    """forward(mode='train') -> materialised fp32 logits [B,L-1,V], differentiable (train.py:120-124)."""

    @staticmethod
    def forward(ctx, module, feats, targets, *params):
        P = dict(zip(PARAM_ORDER, params))
        need = any(ctx.needs_input_grad)
        logits, saved = train_forward_f32(P, feats, targets, stash=need, batch_major_logits=True)
        ctx.saved, ctx.P, ctx.feats, ctx.targets, ctx.module = saved, P, feats, targets, module
        return logits

    @staticmethod
    def backward(ctx, dl):
        dl = dl.contiguous()
        direct, cb = _direct_grad_targets(ctx.module)
        grads, dfeats = train_backward_f32(ctx.P, ctx.saved, ctx.feats, ctx.targets, dl, True, ctx.needs_input_grad[1],
                                           gout=direct, on_ready=cb)
        ctx.saved = None
        return (None, dfeats, None) + _finish_grads(ctx.module, direct, grads)


class _TrainLossFn(torch.autograd.Function):
    """Fused forward + MaskCriterion (utils.py:13-26) without handing the logits to autograd."""

    @staticmethod
    def forward(ctx, module, feats, targets_full, *params):
        P = dict(zip(PARAM_ORDER, params))
        ctx.module = module
        B, L, _ = feats.shape
        V = P["embedding.weight"].shape[0]
        need = any(ctx.needs_input_grad)
        tin = targets_full[:, :-1].contiguous()
        ctx.bf16 = module._use_bf16()
        loss = torch.empty((), device=feats.device)
        # row (t,b) of the time-major logits is scored against targets_full[b, t+1]
        if ctx.bf16:
            ctx.eng, ctx.S = module._engine_and_shadows(P)
            # vocab projection + loss statistics in one kernel: the logits are written once, as bf16
            logits, saved = ctx.eng.train_forward(P, ctx.S, feats, tin, stash=need, batch_major_logits=False,
                                                  ce=dict(targets_full=targets_full, t_off=1, tmap=rowmap(B, 1, L), loss=loss))
        else:
            logits, saved = train_forward_f32(P, feats, tin, stash=need, batch_major_logits=False)
            ops.ce_f32(logits, (L - 1) * B, V, targets_full, 1, rowmap(B, 1, L), loss)
        ctx.saved, ctx.P, ctx.feats, ctx.tin, ctx.tfull, ctx.logits = saved, P, feats, tin, targets_full, logits
        return loss

    @staticmethod
    def backward(ctx, gloss):
        B, L, _ = ctx.feats.shape
        V = ctx.P["embedding.weight"].shape[0]
        logits = ctx.logits
        g = gloss.contiguous().to(torch.float32)
        direct, cb = _direct_grad_targets(ctx.module)
        if ctx.bf16:
            dl = EB.ce_dlogits_inplace(logits, (L - 1) * B, V, ctx.saved["lse"], ctx.tfull, 1, rowmap(B, 1, L), g)
            G = ctx.eng.train_backward(ctx.P, ctx.S, ctx.saved, ctx.tin, dl, ctx.needs_input_grad[1], gout=direct, on_ready=cb,
                                       early_out_wgrad=cb is not None and getattr(ctx.module, "_dp_world", 1) > 1)
            grads, dfeats = [G[k] for k in PARAM_ORDER], G.get("feats")
        else:
            scratch = torch.empty((), device=logits.device)
            ops.ce_f32(logits, (L - 1) * B, V, ctx.tfull, 1, rowmap(B, 1, L), scratch, dlogits=logits, gscale=g)
            grads, dfeats = train_backward_f32(ctx.P, ctx.saved, ctx.feats, ctx.tin, logits, False, ctx.needs_input_grad[1],
                                               gout=direct, on_ready=cb)
        ctx.saved = ctx.logits = None
        return (None, dfeats, None) + _finish_grads(ctx.module, direct, grads)


# forward(mode='train') hands fp32 logits to the caller; when the caller's loss is this package's MaskCriterion, its backward
# looks the producer up here (by the logits' address) and leaves dL/dlogits for it as time-major bf16 -- what the backward GEMMs
# read -- instead of a 263 MB fp32 tensor that would have to be transposed and cast again.
_LOGITS_PRODUCERS: Dict[int, "weakref.ref"] = {}


def lookup_logits_producer(logits2d: torch.Tensor):
    ref = _LOGITS_PRODUCERS.get(logits2d.data_ptr())
    ctx = ref() if ref is not None else None
    if ctx is None or getattr(ctx, "logits_numel", -1) != logits2d.numel():
        return None
    return ctx


class _TrainLogitsBf16Fn(torch.autograd.Function):
    """forward(mode='train') on tensor cores -> materialised fp32 logits [B,L-1,V] (the API contract); backward receives
    dL/dlogits from autograd (e.g. from MaskCriterion), casts it to bf16 in time-major order and runs the bf16 BPTT."""

    @staticmethod
    def forward(ctx, module, feats, targets, *params):
        P = dict(zip(PARAM_ORDER, params))
        need = any(ctx.needs_input_grad)
        ctx.eng, ctx.S = module._engine_and_shadows(P)
        logits, saved = ctx.eng.train_forward(P, ctx.S, feats, targets, stash=need, batch_major_logits=True)
        ctx.saved, ctx.P, ctx.targets, ctx.module = saved, P, targets, module
        ctx.dl_bf16, ctx.logits_numel = None, logits.numel()
        if need:
            if len(_LOGITS_PRODUCERS) > 64:
                for k in [k for k, r in _LOGITS_PRODUCERS.items() if r() is None]:
                    del _LOGITS_PRODUCERS[k]
            _LOGITS_PRODUCERS[logits.data_ptr()] = weakref.ref(ctx)
        return logits

    @staticmethod
    def backward(ctx, dl):
        B, Lm1, V = dl.shape
        fused = ctx.dl_bf16                     # left here by MaskCriterion's backward (time-major bf16), or None
        ctx.dl_bf16 = None
        if fused is not None and all(st == 0 for st in dl.stride()):
            dl_tm = fused                       # `dl` is MaskCriterion's zero-stride placeholder: nothing else contributed
        else:
            # a caller-supplied gradient: [B, L-1, V] f32 -> time-major [(L-1)B, V] bf16
            dl_tm = EB.dlogits_buffer(Lm1 * B, V, dl.device)
            dl_tm.view(Lm1, B, V).copy_(dl.transpose(0, 1))
            if fused is not None:
                dl_tm += fused
        direct, cb = _direct_grad_targets(ctx.module)
        G = ctx.eng.train_backward(ctx.P, ctx.S, ctx.saved, ctx.targets, dl_tm, ctx.needs_input_grad[1], gout=direct, on_ready=cb)
        ctx.saved = None
        return (None, G.get("feats"), None) + _finish_grads(ctx.module, direct, [G[k] for k in PARAM_ORDER])


# --------------------------------------------------------------------------- the module
class S2VT(nn.Module):
    def __init__(self, vocab_size, feat_dim, length, dim_hid=500, dim_embed=500, feat_dropout=0, rnn_dropout=0,
                 out_dropout=0, num_layers=1, bidirectional=False, rnn_type='lstm', sos_ix=3, eos_ix=4,
                 train_precision: str = "auto", decode_precision: str = "x"):
        super().__init__()
        if str(rnn_type).lower() != 'lstm':
            raise NotImplementedError("only rnn_type='lstm' is supported (the reference warns against GRU, train.py:35)")
        if num_layers != 1 or bidirectional:
            raise NotImplementedError("only num_layers=1, bidirectional=False is supported (train.py:33-34)")
        if feat_dropout or rnn_dropout or out_dropout:
            raise NotImplementedError("dropout > 0 is not supported on the sm_100a path (Opt() uses 0, train.py:30-32)")
        # construction order == the reference's, so a seeded construction draws identical weights
        self.vid_rnn = _LSTMParams(dim_hid, dim_hid)
        self.word_rnn = _LSTMParams(dim_hid + dim_embed, dim_hid)
        self.feat_linear = _LinearParams(feat_dim, dim_hid)
        self.out_linear = _LinearParams(dim_hid, vocab_size)
        self.embedding = _EmbeddingParams(vocab_size, dim_embed)
        self.feat_dim = feat_dim
        self.length = length
        self.dim_hid = dim_hid
        self.dim_embed = dim_embed
        self.sos_ix = sos_ix
        self.eos_ix = eos_ix
        self.vocab_size = vocab_size
        self.rnn_type = rnn_type
        self.train_precision = train_precision
        self.decode_precision = decode_precision
        self.beam_topk = 20                     # S2VTModel.py:216
        self._grad_views = None                 # set by FusedAdam.attach(): backward writes gradients in place
        self._on_bucket_ready = None
        self._adam_shadow = None                # set by FusedAdam.attach(): its kernel keeps bf16 copies of the weights current
        self._shadow = EB.ShadowCache()         # bf16 mirrors of the weights (derived, rebuilt lazily, never saved)
        self._xdec = None                       # (key, cfg, weight planes, workspace) of the tensor-core decode path
        self._shadow_step = None                # engine_step.ShadowCache (shapes outside the cluster kernels' range)
        self.beam_check_every = 3               # beam search: the host looks for "every video finished" every this many depths (0 = never)

    def __getstate__(self):
        st = self.__dict__.copy()               # whole-module pickles (train.py:167) carry parameters only
        st["_grad_views"], st["_on_bucket_ready"], st["_adam_shadow"], st["_shadow"] = None, None, None, EB.ShadowCache()
        st["_xdec"] = None
        st["_shadow_step"] = None
        return st

    def __setstate__(self, st):
        st.setdefault("_xdec", None)
        st.setdefault("beam_check_every", 3)
        st.setdefault("_shadow_step", None)
        self.__dict__.update(st)

    # ---- tensor-core decode path ("x": fp32 operands as fp16 hi/lo planes, csrc/xdec_sm100.cu)
    def _use_xdec(self, beam_width=None) -> bool:
        if self.decode_precision == "fp32":
            return False
        if self.decode_precision != "x":
            raise ValueError("decode_precision must be 'x' (tcgen05, fp32-grade) or 'fp32' (CUDA-core FMA)")
        return beam_width is None or beam_width <= 8        # wider beams than the fused top-k keeps: FMA path

    def _xdec_state(self):
        P = self._params()
        key = (ops.WEIGHT_EPOCH, int(self.sos_ix), int(self.eos_ix)) + tuple((p.data_ptr(), p._version) for p in P.values())
        if self._xdec is None or self._xdec["key"] != key:
            cfg = ops.xdec_cfg(self.vocab_size, self.feat_dim, self.length, self.dim_hid, self.dim_embed, self.sos_ix, self.eos_ix)
            wbuf = ops.xdec_prepare(cfg, [P[k].detach() for k in PARAM_ORDER])
            self._xdec = dict(key=key, cfg=cfg, wbuf=wbuf, ws=None, flag=None, pen={})
        return self._xdec

    def _bf16_engine(self):
        """The tensor-core training engine for this shape: engine_bf16 (persistent cluster recurrences: H % 128 == 0, H <= 512, E and F
        multiples of 8), else engine_step (one launch per time step: any H % 8 == 0, F % 8 == 0, e.g. H = 500 / 1000), else None."""
        dims = (self.dim_hid, self.dim_embed, self.feat_dim, self.vocab_size)
        if EB.supported(*dims):
            return EB
        # ('auto' leaves toy shapes, H < 128, on the exact path; train_precision='bf16' takes any shape the step engine supports)
        if ES.supported(*dims) and (self.train_precision == "bf16" or self.dim_hid >= 128):
            return ES
        return None

    def _needs_derived_shadows(self) -> bool:
        return self._bf16_engine() is EB         # engine_step casts private copies at the head of every step (dp.DataParallelTrainer)

    def _engine_and_shadows(self, P):
        eng = self._bf16_engine()
        if eng is EB:
            return EB, self._shadow.get(P, getattr(self, '_adam_shadow', None))
        if getattr(self, "_shadow_step", None) is None:
            self._shadow_step = ES.ShadowCache()
        return ES, self._shadow_step.get(P)

    def _use_bf16(self) -> bool:
        """train_precision: 'bf16' = tensor cores (raises if the shapes are unsupported), 'fp32' = exact CUDA-core path,
        'auto' = bf16 whenever the shapes allow it."""
        ok = self._bf16_engine() is not None
        if self.train_precision == "bf16":
            if not ok:
                raise NotImplementedError("train_precision='bf16' needs dim_hid and feat_dim to be multiples of 8 (any dim_embed, any "
                                          "vocab_size); use train_precision='fp32' for other shapes")
            return True
        if self.train_precision == "fp32":
            return False
        if self.train_precision == "auto":
            return ok
        raise ValueError("train_precision must be 'auto', 'bf16' or 'fp32'")

    # ---- helpers
    def _params(self) -> Dict[str, torch.Tensor]:
        sd = dict(self.named_parameters())
        return {k: sd[k] for k in PARAM_ORDER}

    def _check_inputs(self, feats: torch.Tensor, allow_bf16: bool = False):
        if feats.dim() != 3 or feats.shape[1] != self.length or feats.shape[2] != self.feat_dim:
            raise ValueError("feats must be [B, %d, %d] (got %s)" % (self.length, self.feat_dim, tuple(feats.shape)))
        require_cuda(feats, self.embedding.weight)
        if feats.dtype == torch.bfloat16 and allow_bf16 and self._use_bf16() and not feats.requires_grad:
            return                  # pre-rounded features (data.DeviceFeatureStore(dtype=bfloat16)): the tensor-core path skips its cast
        if feats.dtype != torch.float32:
            raise ValueError("feats must be float32 (bfloat16 only for forward_loss on the bf16 training path, without requires_grad)")
        if str(self.rnn_type).lower() != 'lstm':
            raise NotImplementedError("only rnn_type='lstm' is supported")

    def forward(self, feats, targets=None, mode='train', beam_width=3, max_beam_depth=30):
        """
        :param feats: [B, L, feat_dim] float32 (CUDA)
        :param targets: [B, L-1] int64 (mode='train')
        :param mode: train: fixed-length teacher forcing -> logits [B, L-1, V]
                     test: greedy -> int64 [B, L-1]   beam_search: list[B] of list[Tensor], <sos> first
        """
        self._check_inputs(feats)
        feats = feats.contiguous()
        if mode == 'train':
            if targets is None:
                raise ValueError("mode='train' needs targets [B, L-1]")
            if targets.dim() != 2 or targets.shape[0] != feats.shape[0] or targets.shape[1] != self.length - 1:
                raise RuntimeError("targets must be [B, length-1] = [%d, %d] (got %s); the reference fails in torch.cat "
                                   "(S2VTModel.py:73-75)" % (feats.shape[0], self.length - 1, tuple(targets.shape)))
            targets = targets.contiguous().to(torch.int64)
            P = self._params()
            fn = _TrainLogitsBf16Fn if self._use_bf16() else _TrainLogitsFn
            return fn.apply(self, feats, targets, *[P[k] for k in PARAM_ORDER])
        elif mode == 'test':
            with torch.no_grad():
                return self._greedy(feats.detach())
        elif mode == 'beam_search':
            with torch.no_grad():
                toks, lens = self.beam_search_ids(feats.detach(), beam_width=beam_width, max_beam_depth=max_beam_depth)
            return self._ids_to_reference_lists(toks, lens)
        raise ValueError("unknown mode %r" % (mode,))

    def forward_loss(self, feats, targets, mask=None):
        """Fused train forward + MaskCriterion: returns the scalar the reference computes with
        criterion(model(feats, targets[:, :-1], 'train'), targets, mask) (train.py:120-122) without
        materialising [B,L-1,V] logits for autograd.  `mask` is accepted for signature parity; the
        reference's criterion is independent of it (utils.py:19-26)."""
        self._check_inputs(feats, allow_bf16=True)
        if targets.dim() != 2 or targets.shape[1] != self.length:
            raise RuntimeError("targets must be [B, length] (got %s)" % (tuple(targets.shape),))
        P = self._params()
        return _TrainLossFn.apply(self, feats.contiguous(), targets.contiguous().to(torch.int64), *[P[k] for k in PARAM_ORDER])

    # ---- pretrained word vectors (S2VTModel.py:112-147)
    def load_glove_weights(self, glove_path, glove_dim, ix2word, word2embed=None):
        """Re-initialise the embedding table from GloVe vectors: xavier-normal rows for words GloVe does not know, the GloVe vector
        otherwise (S2VTModel.py:129-147; the table stays trainable, nn.Embedding.from_pretrained(freeze=False)).  `word2embed`: an
        already-extracted {word: vector} dict or a JSON path (the reference caches one at ./data/word2embed.json); None parses
        `glove_path` ("word v1 v2 ..." per line) for the words of `ix2word`."""
        import json
        assert glove_dim == self.dim_embed
        if word2embed is None:
            vocab = set(ix2word.values())
            word2embed = {}
            with open(glove_path, encoding="utf-8") as f:
                for line in f:
                    parts = line.rstrip().split(" ")
                    if parts[0] in vocab:
                        word2embed[parts[0]] = [float(x) for x in parts[1:]]
        elif isinstance(word2embed, str):
            with open(word2embed, encoding="utf-8") as fp:
                word2embed = json.load(fp)
        dev = self.embedding.weight.device
        weights = torch.zeros([self.vocab_size, glove_dim], dtype=torch.float, device=dev)
        torch.nn.init.xavier_normal_(weights)
        for ix, word in ix2word.items():
            if word in word2embed:
                weights[int(ix)] = torch.tensor(word2embed[word], dtype=torch.float, device=dev)
        with torch.no_grad():
            self.embedding.weight.copy_(weights)
        return len(word2embed)

    # ---- greedy (S2VTModel.py:82-110)
    def _greedy(self, feats):
        B, L, _ = feats.shape
        if self._use_xdec():
            X = self._xdec_state()
            tokens = torch.empty(B, L - 1, dtype=torch.int64, device=feats.device)
            X["ws"] = ops.xdec_greedy(X["cfg"], X["wbuf"], feats, tokens, X["ws"])
            return tokens
        P = {k: v.detach() for k, v in self._params().items()}
        H, E, V = self.dim_hid, self.dim_embed, self.vocab_size
        T = 2 * L - 1
        dev = feats.device
        b1, b2 = _bias_sums(P)
        _, out1, _, _, _, _ = _encode_vid_f32(P, feats, T, False, b1)
        pre2 = _word_pre_vid_f32(P, out1, T * B, b2, E, H)
        enc2 = torch.empty(L * B, H, device=dev)
        h2 = torch.empty(B, H, device=dev)
        c2 = torch.empty(B, H, device=dev)
        ops.lstm_fwd_f32(L, B, H, L, pre2, b2, P["word_rnn.weight_hh_l0"], enc2, hT=h2, cT=c2)
        w_cat = torch.cat([P["word_rnn.weight_ih_l0"][:, :E], P["word_rnn.weight_hh_l0"]], dim=1).contiguous()
        tokens = torch.empty(B, L - 1, dtype=torch.int64, device=dev)
        ops.greedy_decode_f32(B, H, E, V, L - 1, int(self.sos_ix), pre2, L * B * 4 * H, w_cat, P["embedding.weight"],
                              P["out_linear.weight"], P["out_linear.bias"], h2, c2, tokens)
        return tokens

    # ---- beam search (S2VTModel.py:56-61, 149-240)
    def beam_search_ids(self, feats, beam_width=3, max_beam_depth=30):
        """Beam search returning (tokens int64 [B, max_depth+1] padded with -1, lengths int32 [B]); row b holds
        what the reference returns as sentences[b], <sos> first."""
        self._check_inputs(feats)
        feats = feats.contiguous()
        P = {k: v.detach() for k, v in self._params().items()}
        B, L, _ = feats.shape
        H, E, V = self.dim_hid, self.dim_embed, self.vocab_size
        dev = feats.device
        if V < self.beam_topk:
            raise RuntimeError("beam search expands topk(%d) (S2VTModel.py:216): vocab_size must be >= %d" % (self.beam_topk, self.beam_topk))
        if self._use_xdec(beam_width):
            X = self._xdec_state()
            pen = X["pen"].get(max_beam_depth)
            if pen is None:         # BeamSearchNode.eval: logp / pow(float(leng), 0.7) -- divisor computed in Python double precision
                pen = torch.tensor([0.0] + [pow(float(n), 0.7) for n in range(1, max_beam_depth + 3)], dtype=torch.float32).to(dev)
                X["pen"][max_beam_depth] = pen
            if X["flag"] is None:
                X["flag"] = torch.zeros(4, dtype=torch.int32).pin_memory()
            toks = torch.empty(B, max_beam_depth + 1, dtype=torch.int64, device=dev)
            lens = torch.empty(B, dtype=torch.int32, device=dev)
            X["ws"] = ops.xdec_beam(X["cfg"], X["wbuf"], feats, int(beam_width), int(max_beam_depth), self.beam_topk, pen, toks, lens,
                                    X["ws"], int(self.beam_check_every), X["flag"])
            return toks, lens
        b1, b2 = _bias_sums(P)
        _, out1, _, _, h1, c1 = _encode_vid_f32(P, feats, L, False, b1)
        pre2 = _word_pre_vid_f32(P, out1, L * B, b2, E, H)
        enc2 = torch.empty(L * B, H, device=dev)
        state = torch.empty(4, B, H, device=dev)
        state[0].copy_(h1); state[1].copy_(c1)
        ops.lstm_fwd_f32(L, B, H, L, pre2, b2, P["word_rnn.weight_hh_l0"], enc2, hT=state[2], cT=state[3])
        w_cat2 = torch.cat([P["word_rnn.weight_ih_l0"], P["word_rnn.weight_hh_l0"]], dim=1).contiguous()
        # BeamSearchNode.eval: logp / pow(float(leng), 0.7) -- divisor computed in Python double precision
        pen = torch.tensor([0.0] + [pow(float(n), 0.7) for n in range(1, max_beam_depth + 3)], dtype=torch.float32).to(dev)
        toks = torch.empty(B, max_beam_depth + 1, dtype=torch.int64, device=dev)
        lens = torch.empty(B, dtype=torch.int32, device=dev)
        ops.beam_search_f32(B, H, E, V, int(beam_width), int(max_beam_depth), self.beam_topk, int(self.sos_ix), int(self.eos_ix),
                            state, b1, P["vid_rnn.weight_hh_l0"], w_cat2, b2, P["embedding.weight"], P["out_linear.weight"],
                            P["out_linear.bias"], pen, toks, lens)
        return toks, lens

    @staticmethod
    def _ids_to_reference_lists(toks: torch.Tensor, lens: torch.Tensor) -> List[List[torch.Tensor]]:
        """Shape the result like the reference: list[B] of list[Tensor]; element 0 is a [1,1] LongTensor holding
        <sos>, the others are 0-dim int64 tensors (S2VTModel.py:177,220,231-238); eval.py:91 only calls .item()."""
        host = toks.cpu()
        n = lens.cpu().tolist()
        out = []
        for b in range(host.shape[0]):
            row = host[b]
            out.append([row[0:1].view(1, 1)] + [row[j] for j in range(1, n[b])])
        return out


S2VTModel = S2VT   # BASELINE.json's wording; the reference class is S2VTModel.S2VT
