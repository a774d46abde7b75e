"""Golden vectors for Att_Baseline from the UNMODIFIED reference class (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_att.py

Imports Att_Baseline from /root/reference/attention_baseline.py and MaskCriterion from /root/reference/utils.py, runs them
on CPU fp32 with weights/inputs from oracle.att_numpy.synth_params / oracle.s2vt_numpy.synth_batch and writes
tests/golden/att_*.npz (logits, loss, all 22 gradients + feats.grad, greedy ids).  The MSVD-shaped case stores strided samples
only; its weights are rebuilt from the seed.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from attention_baseline import Att_Baseline  # noqa: E402  (reference)
from utils import MaskCriterion  # noqa: E402  (reference)

from oracle import att_numpy as A  # noqa: E402
from oracle import s2vt_numpy as O  # noqa: E402

CASES = {
    "att_tiny": dict(V=40, F=24, H=16, E=12, L=6, B=3, real=5, wseed=41, dseed=42, out_scale=20.0, ctx_scale=0.3, full=True),
    "att_mid": dict(V=300, F=160, H=128, E=96, L=12, B=5, real=9, wseed=51, dseed=52, out_scale=8.0, ctx_scale=0.2, full=True),
    "att_msvd": dict(V=13000, F=4096, H=512, E=512, L=80, B=4, real=28, wseed=61, dseed=62, out_scale=1.0, ctx_scale=1.0, full=False),
}
SAMPLE_STRIDE = 997


def sample(a):
    return np.ascontiguousarray(a.reshape(-1)[::SAMPLE_STRIDE])


def run_case(name, c):
    t0 = time.time()
    P = A.synth_params(c["V"], c["F"], c["H"], c["E"], seed=c["wseed"], out_scale=c["out_scale"], ctx_scale=c["ctx_scale"])
    feats, targets, mask = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"], real_tokens=c["real"])
    m = Att_Baseline(c["V"], c["F"], c["L"], dim_hid=c["H"], dim_embed=c["E"], sos_ix=3, eos_ix=4)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()}, strict=True)
    assert tuple(m.state_dict().keys()) == A.PARAM_NAMES
    tf = torch.from_numpy(feats).requires_grad_(True)
    tt = torch.from_numpy(targets)
    m.train()
    logits = m(tf, targets=tt[:, :-1], mode="train")
    loss = MaskCriterion()(logits, tt, torch.from_numpy(mask))
    loss.backward()
    grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in m.named_parameters()}
    grads["feats"] = tf.grad.numpy()
    m.eval()
    with torch.no_grad():
        greedy = m(tf.detach(), mode="test").numpy()
    _, margins = A.greedy(P, feats)
    out = dict(loss=np.float32(loss.item()), greedy=greedy.astype(np.int64), greedy_min_margin=margins.min())
    for k in ("V", "F", "H", "E", "L", "B", "real", "wseed", "dseed", "out_scale", "ctx_scale"):
        out["cfg_" + k] = np.asarray(c[k])
    lg = logits.detach().numpy()
    if c["full"]:
        out["logits"] = lg
        out["feats"], out["targets"], out["mask"] = feats, targets, mask
        for k, v in P.items():
            out["param/" + k] = v
        for k, v in grads.items():
            out["grad/" + k] = v
    else:
        out["logits_sample"] = sample(lg)
        for k, v in grads.items():
            out["grad_sample/" + k] = sample(v)
            out["grad_norm/" + k] = np.float64(np.linalg.norm(v.astype(np.float64)))
            if v.size <= 20000:
                out["grad_full/" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-9s loss=%.6f min greedy margin=%.3e  %.1fs" % (name, out["loss"], out["greedy_min_margin"], time.time() - t0), flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n, CASES[n])
