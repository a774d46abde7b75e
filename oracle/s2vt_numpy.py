"""CPU oracle for the S2VT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a numpy (fp32) restatement of the reference's algorithm for the one path
this repo accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``s2vt-video-caption_b200``) never does.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *unmodified reference module itself*, run in
the build container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the oracle against
those files on every run.

The arithmetic the reference delegates to PyTorch (third-party, not under
/root/reference; torch 2.11.0 in this image, the reference pins no version) is restated
from its published semantics:
  nn.LSTM   : gates = W_ih x + b_ih + W_hh h + b_hh, row blocks ordered i, f, g, o;
              c' = sigmoid(f) * c + sigmoid(i) * tanh(g);  h' = sigmoid(o) * tanh(c')
  nn.Linear : y = x W^T + b
  nn.CrossEntropyLoss(reduction='mean') : mean_r( logsumexp(z_r) - z_r[target_r] )
  optim.Adam: m,v moment estimates with bias correction, eps outside the sqrt.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32

PARAM_NAMES = (
    "vid_rnn.weight_ih_l0", "vid_rnn.weight_hh_l0", "vid_rnn.bias_ih_l0", "vid_rnn.bias_hh_l0",
    "word_rnn.weight_ih_l0", "word_rnn.weight_hh_l0", "word_rnn.bias_ih_l0", "word_rnn.bias_hh_l0",
    "feat_linear.weight", "feat_linear.bias", "out_linear.weight", "out_linear.bias",
    "embedding.weight",
)


def param_shapes(V: int, F: int, H: int, E: int) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the 13 state_dict tensors (S2VTModel.py:19-28)."""
    return {
        "vid_rnn.weight_ih_l0": (4 * H, H), "vid_rnn.weight_hh_l0": (4 * H, H),
        "vid_rnn.bias_ih_l0": (4 * H,), "vid_rnn.bias_hh_l0": (4 * H,),
        "word_rnn.weight_ih_l0": (4 * H, E + H), "word_rnn.weight_hh_l0": (4 * H, H),
        "word_rnn.bias_ih_l0": (4 * H,), "word_rnn.bias_hh_l0": (4 * H,),
        "feat_linear.weight": (H, F), "feat_linear.bias": (H,),
        "out_linear.weight": (V, H), "out_linear.bias": (V,),
        "embedding.weight": (V, E),
    }


def synth_params(V: int, F: int, H: int, E: int, seed: int = 0, out_scale: float = 1.0,
                 eos_ix: int = 4, eos_bias: float = 0.0) -> Dict[str, np.ndarray]:
    """Deterministic random-init weights with the reference's init *distributions*
    (torch defaults: LSTM U(+-1/sqrt(H)), Linear U(+-1/sqrt(fan_in)), Embedding N(0,1)),
    drawn from numpy's PCG64 so the same tensors can be rebuilt on the GPU box without
    shipping 83 MB fixtures.  ``out_scale``/``eos_bias`` make the logits peaky so that
    <eos> paths are exercised."""
    rng = np.random.default_rng(seed)
    P = {}
    for name, shp in param_shapes(V, F, H, E).items():
        if name == "embedding.weight":
            P[name] = rng.standard_normal(shp, dtype=np.float32)
        else:
            if name.startswith("feat_linear"):
                k = 1.0 / math.sqrt(F)
            else:
                k = 1.0 / math.sqrt(H)
            P[name] = ((rng.random(shp, dtype=np.float32) * 2.0 - 1.0) * k).astype(F32)
    if out_scale != 1.0:
        P["out_linear.weight"] = (P["out_linear.weight"] * F32(out_scale)).astype(F32)
    if eos_bias != 0.0:
        P["out_linear.bias"] = P["out_linear.bias"].copy()
        P["out_linear.bias"][eos_ix] += F32(eos_bias)
    return P


def synth_batch(B: int, L: int, F: int, V: int, seed: int = 1234, real_tokens: int = 28,
                sos_ix: int = 3, eos_ix: int = 4):
    """Synthetic MSVD-shaped batch: feats ~ N(0,1) f32 [B,L,F];
    captions = <sos>, real_tokens-2 words in [5,V), <eos>, then <pad>=0 up to L."""
    rng = np.random.default_rng(seed)
    feats = rng.standard_normal((B, L, F), dtype=np.float32)
    real = min(real_tokens, L)
    targets = np.zeros((B, L), dtype=np.int64)
    targets[:, 0] = sos_ix
    if real > 2:
        targets[:, 1:real - 1] = rng.integers(5, V, size=(B, real - 2))
    targets[:, real - 1] = eos_ix
    mask = np.zeros((B, L), dtype=np.float32)
    mask[:, :real] = 1.0
    return feats, targets, mask


# --------------------------------------------------------------------------- LSTM cell
def _sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def lstm_cell(pre: np.ndarray, h: np.ndarray, c: np.ndarray, W_hh: np.ndarray):
    """One nn.LSTM step given the input-side pre-activation ``pre = W_ih x + b_ih + b_hh``.
    Returns (h', c', (i, f, g, o, tanh(c')))."""
    H = h.shape[1]
    gates = (pre + h @ W_hh.T).astype(F32)
    i = _sigmoid(gates[:, 0 * H:1 * H])
    f = _sigmoid(gates[:, 1 * H:2 * H])
    g = np.tanh(gates[:, 2 * H:3 * H], dtype=F32)
    o = _sigmoid(gates[:, 3 * H:4 * H])
    c2 = (f * c + i * g).astype(F32)
    tc = np.tanh(c2, dtype=F32)
    h2 = (o * tc).astype(F32)
    return h2, c2, (i, f, g, o, tc)


def _run_lstm(pre_seq: np.ndarray, W_hh: np.ndarray, h0=None, c0=None, keep=False):
    """pre_seq [T,B,4H] -> outputs [T,B,H], final (h,c), optional per-step stash."""
    T, B, G = pre_seq.shape
    H = G // 4
    h = np.zeros((B, H), F32) if h0 is None else h0.astype(F32)
    c = np.zeros((B, H), F32) if c0 is None else c0.astype(F32)
    outs = np.empty((T, B, H), F32)
    stash = []
    for t in range(T):
        c_prev = c
        h, c, acts = lstm_cell(pre_seq[t], h, c, W_hh)
        outs[t] = h
        if keep:
            stash.append((acts, c_prev))
    return outs, (h, c), stash


# --------------------------------------------------------------------------- forward
def _project_feats(P, feats):
    """feat_drop (p=0) + feat_linear, S2VTModel.py:52-54.  Returns time-major [L,B,H]."""
    B, L, Fd = feats.shape
    x = feats.reshape(B * L, Fd).astype(F32) @ P["feat_linear.weight"].T + P["feat_linear.bias"]
    return np.ascontiguousarray(x.reshape(B, L, -1).transpose(1, 0, 2)).astype(F32)


def _vid_pre(P, xproj, T):
    """Input-side pre-activations of vid_rnn over T steps; steps >= L see the zero pad that
    S2VTModel.py:64-65 appends *after* the projection, i.e. bias only."""
    L, B, H = xproj.shape
    b = (P["vid_rnn.bias_ih_l0"] + P["vid_rnn.bias_hh_l0"]).astype(F32)
    pre = np.empty((T, B, 4 * H), F32)
    pre[:L] = xproj @ P["vid_rnn.weight_ih_l0"].T + b
    pre[L:] = b
    return pre


def _word_pre(P, out1, emb_seq, E):
    """Input-side pre-activations of word_rnn for input2 = [pad_embed || output1]
    (S2VTModel.py:72-75; embedding columns come FIRST).  ``emb_seq`` is [T,B,E] or None for
    an all-zero embedding half."""
    W = P["word_rnn.weight_ih_l0"]
    b = (P["word_rnn.bias_ih_l0"] + P["word_rnn.bias_hh_l0"]).astype(F32)
    pre = out1 @ W[:, E:].T + b
    if emb_seq is not None:
        pre = pre + emb_seq @ W[:, :E].T
    return pre.astype(F32)


def forward_train(P: Dict[str, np.ndarray], feats: np.ndarray, targets: np.ndarray, keep: bool = False):
    """S2VT.forward(mode='train'), S2VTModel.py:48-81.
    feats f32 [B,L,F]; targets i64 [B,L-1]  ->  logits f32 [B,L-1,V] (and a cache for backward)."""
    B, L, _ = feats.shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    E = P["embedding.weight"].shape[1]
    T = 2 * L - 1
    if targets.shape != (B, L - 1):
        raise ValueError("targets must be [B, L-1]")  # reference fails in torch.cat, S2VTModel.py:73-75
    xproj = _project_feats(P, feats)
    pre1 = _vid_pre(P, xproj, T)
    out1, _, st1 = _run_lstm(pre1, P["vid_rnn.weight_hh_l0"], keep=keep)
    emb = P["embedding.weight"][targets]                       # [B,L-1,E]  S2VTModel.py:71
    emb_seq = np.zeros((T, B, E), F32)
    emb_seq[L:] = emb.transpose(1, 0, 2)
    pre2 = _word_pre(P, out1, emb_seq, E)
    out2, _, st2 = _run_lstm(pre2, P["word_rnn.weight_hh_l0"], keep=keep)
    hdec = out2[L:]                                            # [L-1,B,H]  S2VTModel.py:78
    logits = hdec @ P["out_linear.weight"].T + P["out_linear.bias"]
    logits = np.ascontiguousarray(logits.transpose(1, 0, 2)).astype(F32)
    if not keep:
        return logits
    cache = dict(feats=feats, targets=targets, xproj=xproj, out1=out1, out2=out2, st1=st1, st2=st2,
                 emb_seq=emb_seq, B=B, L=L, H=H, E=E, T=T)
    return logits, cache


def log_softmax(z: np.ndarray) -> np.ndarray:
    m = z.max(axis=-1, keepdims=True)
    s = z - m
    return (s - np.log(np.exp(s, dtype=F32).sum(axis=-1, keepdims=True, dtype=F32))).astype(F32)


def mask_criterion(logits: np.ndarray, target: np.ndarray, mask: np.ndarray) -> np.float32:
    """MaskCriterion.forward, utils.py:13-26.  nn.CrossEntropyLoss() reduces to a scalar mean
    BEFORE the mask is applied, so sum(loss*mask)/sum(mask) == loss: the mask cancels and the
    result is the plain mean CE over all B*(L-1) positions, <pad> targets included."""
    B, Lm1, V = logits.shape
    tgt = target[:, 1:].reshape(-1)
    lp = log_softmax(logits.reshape(B * Lm1, V).astype(F32))
    loss = F32(-lp[np.arange(B * Lm1), tgt].astype(np.float64).mean())
    m = mask[:, 1:].reshape(-1).astype(F32)
    return F32(np.sum(loss * m, dtype=F32) / np.sum(m, dtype=F32))


def dlogits_of_loss(logits: np.ndarray, target: np.ndarray) -> np.ndarray:
    """d loss / d logits for the (unmasked) mean CE above."""
    B, Lm1, V = logits.shape
    p = np.exp(log_softmax(logits.reshape(B * Lm1, V)), dtype=F32)
    p[np.arange(B * Lm1), target[:, 1:].reshape(-1)] -= F32(1.0)
    return (p / F32(B * Lm1)).reshape(B, Lm1, V).astype(F32)


# --------------------------------------------------------------------------- backward
def _lstm_backward(dout: np.ndarray, stash, W_hh: np.ndarray):
    """BPTT through one LSTM layer from zero initial state.
    dout [T,B,H] = dL/d output_t.  Returns dgates [T,B,4H] (pre-activation grads)."""
    T, B, H = dout.shape
    dg = np.empty((T, B, 4 * H), F32)
    dh_rec = np.zeros((B, H), F32)
    dc = np.zeros((B, H), F32)
    for t in range(T - 1, -1, -1):
        (i, f, g, o, tc), c_prev = stash[t]
        dh = dout[t] + dh_rec
        do = dh * tc
        dc = dc + dh * o * (F32(1) - tc * tc)
        di = dc * g
        df = dc * c_prev
        dgg = dc * i
        dg[t, :, 0 * H:1 * H] = di * i * (F32(1) - i)
        dg[t, :, 1 * H:2 * H] = df * f * (F32(1) - f)
        dg[t, :, 2 * H:3 * H] = dgg * (F32(1) - g * g)
        dg[t, :, 3 * H:4 * H] = do * o * (F32(1) - o)
        dc = dc * f
        dh_rec = dg[t] @ W_hh
    return dg


def backward(P, cache, dlogits: np.ndarray) -> Dict[str, np.ndarray]:
    """Gradients of all 13 parameters (+ 'feats') given dL/dlogits [B,L-1,V]; the numpy twin of
    what autograd does for train.py:124."""
    B, L, H, E, T = cache["B"], cache["L"], cache["H"], cache["E"], cache["T"]
    out1, out2 = cache["out1"], cache["out2"]
    G = {}
    dl = dlogits.transpose(1, 0, 2).reshape((L - 1) * B, -1).astype(F32)      # time-major rows
    hdec = out2[L:].reshape((L - 1) * B, H)
    G["out_linear.weight"] = dl.T @ hdec
    G["out_linear.bias"] = dl.sum(0)
    dout2 = np.zeros((T, B, H), F32)
    dout2[L:] = (dl @ P["out_linear.weight"]).reshape(L - 1, B, H)
    dg2 = _lstm_backward(dout2, cache["st2"], P["word_rnn.weight_hh_l0"])
    hprev2 = np.concatenate([np.zeros((1, B, H), F32), out2[:-1]], 0)
    in2 = np.concatenate([cache["emb_seq"], out1], 2)                       # [T,B,E+H]
    dg2f = dg2.reshape(T * B, 4 * H)
    G["word_rnn.weight_ih_l0"] = dg2f.T @ in2.reshape(T * B, E + H)
    G["word_rnn.weight_hh_l0"] = dg2f.T @ hprev2.reshape(T * B, H)
    G["word_rnn.bias_ih_l0"] = dg2f.sum(0)
    G["word_rnn.bias_hh_l0"] = dg2f.sum(0)
    din2 = (dg2f @ P["word_rnn.weight_ih_l0"]).reshape(T, B, E + H)
    demb = din2[L:, :, :E]                                                  # [L-1,B,E]
    gE = np.zeros_like(P["embedding.weight"])
    np.add.at(gE, cache["targets"].T.reshape(-1), demb.reshape(-1, E))
    G["embedding.weight"] = gE
    dout1 = np.ascontiguousarray(din2[:, :, E:])
    dg1 = _lstm_backward(dout1, cache["st1"], P["vid_rnn.weight_hh_l0"])
    hprev1 = np.concatenate([np.zeros((1, B, H), F32), out1[:-1]], 0)
    dg1f = dg1.reshape(T * B, 4 * H)
    xp = cache["xproj"].reshape(L * B, H)
    G["vid_rnn.weight_ih_l0"] = dg1f[:L * B].T @ xp
    G["vid_rnn.weight_hh_l0"] = dg1f.T @ hprev1.reshape(T * B, H)
    G["vid_rnn.bias_ih_l0"] = dg1f.sum(0)
    G["vid_rnn.bias_hh_l0"] = dg1f.sum(0)
    dxp = (dg1f[:L * B] @ P["vid_rnn.weight_ih_l0"]).reshape(L, B, H).transpose(1, 0, 2).reshape(B * L, H)
    X = cache["feats"].reshape(B * L, -1)
    G["feat_linear.weight"] = dxp.T @ X
    G["feat_linear.bias"] = dxp.sum(0)
    G["feats"] = (dxp @ P["feat_linear.weight"]).reshape(cache["feats"].shape)
    return {k: v.astype(F32) for k, v in G.items()}


def loss_and_grads(P, feats, targets_full, mask):
    """The train-step arithmetic of train.py:120-124: returns (loss, logits, grads)."""
    logits, cache = forward_train(P, feats, targets_full[:, :-1], keep=True)
    loss = mask_criterion(logits, targets_full, mask)
    grads = backward(P, cache, dlogits_of_loss(logits, targets_full))
    return loss, logits, grads


def adam_step(P, G, state, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (train.py:89-93: lr=1e-4, defaults otherwise, no weight decay)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for k in PARAM_NAMES:
        m = state.setdefault("m." + k, np.zeros_like(P[k]))
        v = state.setdefault("v." + k, np.zeros_like(P[k]))
        g = G[k]
        m[...] = b1 * m + (1 - b1) * g
        v[...] = b2 * v + (1 - b2) * g * g
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = np.sqrt(v) / math.sqrt(bc2) + eps
        P[k] = (P[k] - (lr / bc1) * (m / denom)).astype(F32)
    return P


# --------------------------------------------------------------------------- greedy
def greedy(P, feats: np.ndarray, sos_ix: int = 3) -> np.ndarray:
    """S2VT.forward(mode='test'), S2VTModel.py:82-110 -> int64 [B, L-1]; no early stop at <eos>."""
    B, L, _ = feats.shape
    E = P["embedding.weight"].shape[1]
    T = 2 * L - 1
    xproj = _project_feats(P, feats)
    out1, _, _ = _run_lstm(_vid_pre(P, xproj, T), P["vid_rnn.weight_hh_l0"])
    pre2_vid = _word_pre(P, out1, None, E)                      # vid half + biases for all T steps
    _, (h2, c2), _ = _run_lstm(pre2_vid[:L], P["word_rnn.weight_hh_l0"])
    Wemb = P["word_rnn.weight_ih_l0"][:, :E]
    tok = np.full((B,), sos_ix, np.int64)
    pred = np.empty((B, L - 1), np.int64)
    for k in range(L - 1):
        pre = (pre2_vid[L + k] + P["embedding.weight"][tok] @ Wemb.T).astype(F32)
        h2, c2, _ = lstm_cell(pre, h2, c2, P["word_rnn.weight_hh_l0"])
        logits = h2 @ P["out_linear.weight"].T + P["out_linear.bias"]
        tok = logits.argmax(1)                                   # ties -> lowest index, like torch
        pred[:, k] = tok
    return pred


def greedy_margins(P, feats, sos_ix: int = 3) -> np.ndarray:
    """Top-1 minus top-2 logit at every greedy step, for picking fixtures whose argmax is robust
    to fp32 summation-order noise."""
    B, L, _ = feats.shape
    pred = greedy(P, feats, sos_ix)
    tg = np.concatenate([np.full((B, 1), sos_ix, np.int64), pred[:, :-1]], 1)
    lg = forward_train(P, feats, tg)
    s = np.sort(lg, axis=-1)
    return (s[..., -1] - s[..., -2]).astype(F32)


# --------------------------------------------------------------------------- beam search
def _len_penalty(n: int, alpha: float = 0.7) -> np.float32:
    """BeamSearchNode.eval, S2VTModel.py:261-269: logp / pow(float(leng), alpha) with the divisor
    computed in Python double precision, then applied to an fp32 tensor."""
    return F32(pow(float(n), alpha))


def beam_encode(P, feats):
    """Encode stage of mode='beam_search', S2VTModel.py:56-60: vid_rnn over the L real frames only,
    word_rnn over [0 || output1].  Returns per-video (h1,c1,h2,c2)."""
    E = P["embedding.weight"].shape[1]
    xproj = _project_feats(P, feats)
    L = xproj.shape[0]
    out1, (h1, c1), _ = _run_lstm(_vid_pre(P, xproj, L), P["vid_rnn.weight_hh_l0"])
    _, (h2, c2), _ = _run_lstm(_word_pre(P, out1, None, E), P["word_rnn.weight_hh_l0"])
    return h1, c1, h2, c2


def beam_search(P, feats, beam_width: int = 3, max_depth: int = 30, sos_ix: int = 3, eos_ix: int = 4,
                topk: int = 20, return_scores: bool = False):
    """S2VT.beam_search, S2VTModel.py:149-240, restated in lock-step form.

    Per video the reference keeps a PriorityQueue keyed by -(logp_last / len**0.7) where logp_last is
    the log-prob of the node's OWN token only (not cumulative, S2VTModel.py:220).  Each depth it pops
    the <= beam_width best entries, clears the queue, re-queues finished (<eos>) entries unchanged and
    expands every other entry into its top-20 next tokens (S2VTModel.py:216-223).  It stops when the
    queue holds <= beam_width entries (S2VTModel.py:227) or after max_depth rounds, and returns the
    token chain of the best queue entry, <sos> included (S2VTModel.py:231-238).
    Exact key ties follow heapq internals in the reference; they are measure-zero for distinct fp32
    log-probs and the golden vectors are checked to be tie-free.
    """
    B = feats.shape[0]
    V, E = P["embedding.weight"].shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    h1, c1, h2, c2 = beam_encode(P, feats)
    b1 = (P["vid_rnn.bias_ih_l0"] + P["vid_rnn.bias_hh_l0"]).astype(F32)[None]
    b2 = (P["word_rnn.bias_ih_l0"] + P["word_rnn.bias_hh_l0"]).astype(F32)[None]
    W2 = P["word_rnn.weight_ih_l0"]
    k = min(topk, V)
    sentences, scores = [], []
    for b in range(B):
        # hyp = (key, tokens, state, finished)
        queue = [(F32(-0.0), [sos_ix], (h1[b:b + 1], c1[b:b + 1], h2[b:b + 1], c2[b:b + 1]), False)]
        for _depth in range(max_depth):
            queue.sort(key=lambda q: q[0])                      # stable; smallest key = best
            beam, queue = queue[:beam_width], []
            for key, toks, st, fin in beam:
                if fin:
                    queue.append((key, toks, st, fin))
                    continue
                a1, d1, a2, d2 = st
                # vid_rnn on a zero input (S2VTModel.py:208-210): pre-activation = biases only
                a1, d1, _ = lstm_cell(b1, a1, d1, P["vid_rnn.weight_hh_l0"])
                x2 = np.concatenate([P["embedding.weight"][toks[-1]][None], a1], 1)
                a2, d2, _ = lstm_cell((x2 @ W2.T + b2).astype(F32), a2, d2, P["word_rnn.weight_hh_l0"])
                lp = log_softmax((a2 @ P["out_linear.weight"].T + P["out_linear.bias"])[0])
                top = np.argsort(-lp, kind="stable")[:k]
                n = len(toks) + 1
                pen = _len_penalty(n)
                for tok in sorted(top.tolist()):                 # reference inserts in index order
                    queue.append((F32(-(lp[tok] / pen)), toks + [int(tok)], (a1, d1, a2, d2),
                                  int(tok) == eos_ix))
            if len(queue) <= beam_width:
                break
        best = min(queue, key=lambda q: q[0])
        sentences.append(best[1])
        scores.append(best[0])
    if return_scores:
        return sentences, scores
    return sentences


def beam_tie_free(P, feats, **kw) -> bool:
    """True when no two queue keys that matter for selection are exactly equal (fixtures only)."""
    sents, scores = beam_search(P, feats, return_scores=True, **kw)
    return all(np.isfinite(s) for s in scores)
