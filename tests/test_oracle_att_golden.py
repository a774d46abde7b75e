"""Pin the Att_Baseline numpy oracle (oracle/att_numpy.py) against golden vectors dumped from the unmodified reference
class (tests/golden/make_golden_att.py).  CPU only."""
import numpy as np
import pytest

from conftest import att_golden_inputs, load_golden
from oracle import att_numpy as A
from oracle import s2vt_numpy as O

STRIDE = 997


def _loss_and_grads(P, feats, targets, mask):
    logits, cache = A.forward_train(P, feats, targets[:, :-1], keep=True)
    loss = O.mask_criterion(logits, targets, mask)
    return loss, logits, A.backward(P, cache, O.dlogits_of_loss(logits, targets))


@pytest.mark.parametrize("name", ["att_tiny", "att_mid"])
def test_att_train_full(name):
    g = load_golden(name)
    P, feats, targets, mask, c = att_golden_inputs(g)
    loss, logits, grads = _loss_and_grads(P, feats, targets, mask)
    assert logits.shape == (c["B"], c["L"] - 1, c["V"])
    assert np.abs(logits - g["logits"]).max() <= 2e-5 * max(1.0, np.abs(g["logits"]).max())
    assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"])
    for k in list(A.PARAM_NAMES) + ["feats"]:
        ref = g["grad/" + k]
        assert np.abs(grads[k] - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, k
    for k in ("att_enc.weight", "att_enc.bias", "att_prev_hid.weight", "att_prev_hid.bias", "att_apply.weight"):
        assert not g["grad/" + k].any()                 # softmax over a singleton: the reference's att layers get zero gradient
    assert not g["grad/embedding.weight"][0].any()      # padding_idx=0


def test_att_train_msvd_samples():
    g = load_golden("att_msvd")
    P, feats, targets, mask, c = att_golden_inputs(g)
    loss, logits, grads = _loss_and_grads(P, feats, targets, mask)
    ref = g["logits_sample"]
    assert np.abs(logits.reshape(-1)[::STRIDE] - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"])
    for k in list(A.PARAM_NAMES) + ["feats"]:
        ref = g["grad_sample/" + k]
        tol = 2e-4 * max(np.abs(ref).max(), g["grad_norm/" + k] / np.sqrt(grads[k].size)) + 1e-9
        assert np.abs(grads[k].reshape(-1)[::STRIDE] - ref).max() <= tol, k
        assert abs(np.linalg.norm(grads[k].astype(np.float64)) - g["grad_norm/" + k]) <= 1e-4 * g["grad_norm/" + k] + 1e-12, k


@pytest.mark.parametrize("name", ["att_tiny", "att_mid", "att_msvd"])
def test_att_greedy(name):
    g = load_golden(name)
    P, feats, targets, mask, c = att_golden_inputs(g)
    pred, _ = A.greedy(P, feats)
    assert pred.dtype == np.int64 and pred.shape == (c["B"], c["L"])      # L steps, attention_baseline.py:93
    assert np.array_equal(pred, g["greedy"])
