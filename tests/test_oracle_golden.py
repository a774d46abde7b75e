"""Pin the numpy oracle against golden vectors dumped from the unmodified reference module
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import beam_rows_to_lists, golden_inputs, load_golden
from oracle import s2vt_numpy as O

SAMPLE_STRIDE = 997


def _sample(a):
    return a.reshape(-1)[::SAMPLE_STRIDE]


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_train_forward_loss_grads_full(name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    loss, logits, grads = O.loss_and_grads(P, feats, targets, mask)
    scale = np.abs(g["logits"]).max()
    assert np.abs(logits - g["logits"]).max() <= 2e-5 * max(1.0, scale)
    assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"])
    for k in list(O.PARAM_NAMES) + ["feats"]:
        ref = g["grad/" + k]
        tol = 1e-4 * np.abs(ref).max() + 1e-7
        assert np.abs(grads[k] - ref).max() <= tol, k


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_adam_step(name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    G = {k: g["grad/" + k] for k in O.PARAM_NAMES}
    P2 = O.adam_step({k: v.copy() for k, v in P.items()}, G, {}, lr=1e-4)
    for k in O.PARAM_NAMES:
        # first Adam step moves every weight by ~lr; compare the update, not the weight
        upd_ref = g["adam1/" + k] - P[k]
        upd = P2[k] - P[k]
        nz = np.abs(G[k]) > 1e-6          # tiny grads: m/(sqrt(v)+eps) is eps-sensitive, skip
        assert np.abs(upd - upd_ref)[nz].max(initial=0.0) <= 2e-6, k


@pytest.mark.parametrize("name", ["tiny", "mid", "msvd", "msvd_peaky", "paper"])
def test_greedy_tokens(name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    pred = O.greedy(P, feats)
    assert pred.dtype == np.int64 and pred.shape == (c["B"], c["L"] - 1)
    assert np.array_equal(pred, g["greedy"])


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_beam_small(name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    for key in [k for k in g if k.startswith("beam")]:
        bw = int(key[4:])
        ref = beam_rows_to_lists(g[key])
        got = O.beam_search(P, feats[:len(ref)], beam_width=bw, max_depth=30)
        assert got == ref, (key, got, ref)
        assert all(s[0] == 3 for s in got)        # <sos> is element 0 (S2VTModel.py:231-238)


@pytest.mark.parametrize("name", ["msvd", "msvd_peaky"])
def test_beam_msvd_shape(name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    for key in [k for k in g if k.startswith("beam")]:
        bw = int(key[4:])
        ref = beam_rows_to_lists(g[key])
        got = O.beam_search(P, feats[:len(ref)], beam_width=bw, max_depth=30)
        assert got == ref, key


@pytest.mark.parametrize("name", ["msvd", "paper"])
def test_train_msvd_shape_samples(name):
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    loss, logits, grads = O.loss_and_grads(P, feats, targets, mask)
    assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"])
    assert np.abs(_sample(logits) - g["logits_sample"]).max() <= 2e-5
    for k in list(O.PARAM_NAMES) + ["feats"]:
        ref = g["grad_sample/" + k]
        assert np.abs(_sample(grads[k]) - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-9, k
        n = np.linalg.norm(grads[k].astype(np.float64))
        assert abs(n - g["grad_norm/" + k]) <= 1e-4 * g["grad_norm/" + k], k


def test_loss_mask_cancels():
    """MaskCriterion does not mask (utils.py:19-26): the value is independent of the mask."""
    g = load_golden("tiny")
    P, feats, targets, mask, c = golden_inputs(g)
    a = O.mask_criterion(g["logits"], targets, mask)
    m2 = np.ones_like(mask)
    b = O.mask_criterion(g["logits"], targets, m2)
    assert abs(a - b) < 1e-6 and abs(a - g["loss"]) < 1e-5


def test_beam_width_one_is_greedy_prefix():
    g = load_golden("mid")
    P, feats, targets, mask, c = golden_inputs(g)
    gr = O.greedy(P, feats)
    bs = O.beam_search(P, feats, beam_width=1, max_depth=8)
    for b, s in enumerate(bs):
        body = s[1:]
        assert body == gr[b, :len(body)].tolist()


def test_targets_shape_error():
    g = load_golden("tiny")
    P, feats, targets, mask, c = golden_inputs(g)
    with pytest.raises(ValueError):
        O.forward_train(P, feats, targets)        # [B,L] instead of [B,L-1]


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_torch_port_matches_golden(name):
    """The CPU-baseline port (oracle/torch_port.py) reproduces the reference's logits, loss, grads and greedy ids."""
    import torch
    from oracle.torch_port import S2VTCpuPort
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    m = S2VTCpuPort(c["V"], c["F"], c["L"], c["H"], c["E"])
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()}, strict=True)
    tf = torch.from_numpy(feats).requires_grad_(True)
    tt, tm = torch.from_numpy(targets), torch.from_numpy(mask)
    logits = m.train_logits(tf, tt[:, :-1])
    loss = m.criterion(logits, tt, tm)
    loss.backward()
    assert np.abs(logits.detach().numpy() - g["logits"]).max() <= 2e-5 * max(1.0, np.abs(g["logits"]).max())
    assert abs(loss.item() - g["loss"]) <= 1e-5 * abs(g["loss"])
    for k, p in m.named_parameters():
        ref = g["grad/" + k]
        assert np.abs(p.grad.numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, k
    assert np.array_equal(m.greedy(tf.detach()).numpy(), g["greedy"])
