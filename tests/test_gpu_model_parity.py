"""Module-level parity of the drop-in S2VT (CUDA, through the C ABI) against
  (a) golden vectors dumped from the unmodified reference module (tests/golden/*.npz), and
  (b) the numpy oracle on fresh seeded inputs.
Tolerances (fp32 exact mode): logits atol 2e-5 * max(1,|logit|max); loss rtol 1e-5; gradients
1e-4 * |grad|max; greedy / beam token ids bit-exact.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest
import torch

from conftest import beam_rows_to_lists, golden_inputs, load_golden
from oracle import s2vt_numpy as O

pytestmark = pytest.mark.gpu

import s2vt_b200  # noqa: E402

SAMPLE_STRIDE = 997


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    s2vt_b200.load()
    return torch.device("cuda:0")


def build_model(c, P, dev, **kw):
    kw.setdefault("train_precision", "fp32")          # this file checks the exact path; bf16 parity lives in test_gpu_bf16.py
    m = s2vt_b200.S2VT(c["V"], c["F"], c["L"], dim_hid=c["H"], dim_embed=c["E"], sos_ix=3, eos_ix=4, **kw)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()}, strict=True)
    return m.to(dev)


def _sample(a):
    return a.reshape(-1)[::SAMPLE_STRIDE]


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_train_logits_loss_grads_vs_reference_golden(dev, name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev)
    crit = s2vt_b200.MaskCriterion()
    tf = torch.from_numpy(feats).to(dev).requires_grad_(True)
    tt = torch.from_numpy(targets).to(dev)
    tm = torch.from_numpy(mask).to(dev)
    logits = model(tf, targets=tt[:, :-1], mode="train")
    assert logits.shape == (c["B"], c["L"] - 1, c["V"]) and logits.dtype == torch.float32
    loss = crit(logits, tt, tm)
    loss.backward()
    scale = max(1.0, float(np.abs(g["logits"]).max()))
    assert np.abs(logits.detach().cpu().numpy() - g["logits"]).max() <= 2e-5 * scale
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    grads["feats"] = tf.grad.cpu().numpy()
    for k, gv in grads.items():
        ref = g["grad/" + k]
        assert np.abs(gv - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, k


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_fused_forward_loss_matches(dev, name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev)
    tf = torch.from_numpy(feats).to(dev)
    tt = torch.from_numpy(targets).to(dev)
    loss = model.forward_loss(tf, tt, torch.from_numpy(mask).to(dev))
    (loss * 1.0).backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k, p in model.named_parameters():
        ref = g["grad/" + k]
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, k


@pytest.mark.parametrize("dec", ["x", "fp32"])          # tcgen05 fp16-split path (default) and the CUDA-core FFMA path
@pytest.mark.parametrize("name", ["tiny", "mid", "msvd", "msvd_peaky", "paper"])
def test_greedy_tokens_bit_exact(dev, name, dec):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev, decode_precision=dec).eval()
    with torch.no_grad():
        pred = model(torch.from_numpy(feats).to(dev), mode="test")
    assert pred.dtype == torch.int64 and tuple(pred.shape) == (c["B"], c["L"] - 1)
    assert np.array_equal(pred.cpu().numpy(), g["greedy"])


@pytest.mark.parametrize("dec", ["x", "fp32"])
@pytest.mark.parametrize("name", ["tiny", "mid", "msvd", "msvd_peaky"])
def test_beam_tokens_bit_exact(dev, name, dec):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev, decode_precision=dec).eval()
    tf = torch.from_numpy(feats).to(dev)
    for key in [k for k in g if k.startswith("beam")]:
        bw = int(key[4:])
        ref = beam_rows_to_lists(g[key])
        with torch.no_grad():
            sents = model(tf[:len(ref)], mode="beam_search", beam_width=bw, max_beam_depth=30)
        got = [[int(x.item()) for x in s] for s in sents]
        assert got == ref, (name, key)
        assert tuple(sents[0][0].shape) == (1, 1) and sents[0][0].dtype == torch.int64     # reference return type


@pytest.mark.parametrize("name", ["msvd", "paper"])
def test_train_msvd_shape_vs_reference_golden(dev, name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev)
    tf = torch.from_numpy(feats).to(dev).requires_grad_(True)
    tt = torch.from_numpy(targets).to(dev)
    logits = model(tf, targets=tt[:, :-1], mode="train")
    loss = s2vt_b200.MaskCriterion()(logits, tt, torch.from_numpy(mask).to(dev))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.abs(_sample(logits.detach().cpu().numpy()) - g["logits_sample"]).max() <= 2e-5
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    grads["feats"] = tf.grad.cpu().numpy()
    for k, gv in grads.items():
        ref = g["grad_sample/" + k]
        assert np.abs(_sample(gv) - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-9, k
        n = np.linalg.norm(gv.astype(np.float64))
        assert abs(n - float(g["grad_norm/" + k])) <= 1e-4 * float(g["grad_norm/" + k]), k


def test_fresh_inputs_vs_oracle(dev):
    """Seeds the goldens never saw: CUDA path vs the numpy oracle (ragged dims, E != H, B not a tile multiple)."""
    V, F, H, E, Lq, B = 77, 36, 20, 28, 5, 7
    P = O.synth_params(V, F, H, E, seed=101, out_scale=20.0, eos_bias=1.5)
    feats, targets, mask = O.synth_batch(B, Lq, F, V, seed=202, real_tokens=4)
    c = dict(V=V, F=F, H=H, E=E, L=Lq, B=B)
    model = build_model(c, P, dev)
    tf = torch.from_numpy(feats).to(dev)
    tt = torch.from_numpy(targets).to(dev)
    loss_ref, logits_ref, grads_ref = O.loss_and_grads(P, feats, targets, mask)
    logits = model(tf, targets=tt[:, :-1], mode="train")
    loss = s2vt_b200.MaskCriterion()(logits, tt, torch.from_numpy(mask).to(dev))
    loss.backward()
    assert np.abs(logits.detach().cpu().numpy() - logits_ref).max() <= 2e-5 * max(1.0, np.abs(logits_ref).max())
    assert abs(loss.item() - loss_ref) <= 1e-5 * abs(loss_ref)
    for k, p in model.named_parameters():
        assert np.abs(p.grad.cpu().numpy() - grads_ref[k]).max() <= 1e-4 * np.abs(grads_ref[k]).max() + 1e-7, k
    with torch.no_grad():
        assert np.array_equal(model(tf, mode="test").cpu().numpy(), O.greedy(P, feats))
        for bw in (1, 2, 4):
            toks, lens = model.beam_search_ids(tf, beam_width=bw, max_beam_depth=9)
            got = [toks[b, :lens[b]].tolist() for b in range(B)]
            assert got == O.beam_search(P, feats, beam_width=bw, max_depth=9), bw


def test_cross_mode_invariants(dev):
    """(i) teacher-forcing the greedy output reproduces it as argmax of the train logits;
    (ii) beam_width=1 returns <sos> + a prefix of the greedy tokens."""
    g = load_golden("mid")
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev).eval()
    tf = torch.from_numpy(feats).to(dev)
    with torch.no_grad():
        pred = model(tf, mode="test")
        tg = torch.cat([torch.full((c["B"], 1), 3, dtype=torch.int64, device=dev), pred[:, :-1]], 1)
        lg = model(tf, targets=tg, mode="train")
        assert torch.equal(lg.argmax(-1), pred)
        toks, lens = model.beam_search_ids(tf, beam_width=1, max_beam_depth=8)
    for b in range(c["B"]):
        body = toks[b, 1:lens[b]].tolist()
        assert body == pred[b, :len(body)].tolist()


def test_fused_adam_matches_torch_adam(dev):
    g = load_golden("tiny")
    P, feats, targets, mask, c = golden_inputs(g)
    model = build_model(c, P, dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-4)
    tf = torch.from_numpy(feats).to(dev)
    tt = torch.from_numpy(targets).to(dev)
    loss = model.forward_loss(tf, tt)
    opt.zero_grad()
    loss.backward()
    opt.step()
    for k, p in model.named_parameters():
        upd_ref = g["adam1/" + k] - P[k]
        upd = p.detach().cpu().numpy() - P[k]
        nz = np.abs(g["grad/" + k]) > 1e-6
        assert np.abs(upd - upd_ref)[nz].max(initial=0.0) <= 2e-6, k


def test_api_errors(dev):
    with pytest.raises(NotImplementedError):
        s2vt_b200.S2VT(40, 24, 6, rnn_type="gru")
    with pytest.raises(NotImplementedError):
        s2vt_b200.S2VT(40, 24, 6, bidirectional=True)
    m = s2vt_b200.S2VT(40, 24, 6, 16, 12).to(dev)
    f = torch.randn(2, 6, 24, device=dev)
    with pytest.raises(RuntimeError):
        m(f, targets=torch.zeros(2, 6, dtype=torch.int64, device=dev), mode="train")      # [B,L] instead of [B,L-1]
    with pytest.raises(RuntimeError):
        m.cpu()(f.cpu(), mode="test")                                                     # no CPU fallback
