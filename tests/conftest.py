import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s missing" % name)
    z = np.load(path)
    return {k: z[k] for k in z.files}


def golden_cfg(g):
    return {k[4:]: g[k].item() for k in g if k.startswith("cfg_")}


def golden_inputs(g):
    """Rebuild (params, feats, targets, mask) for a fixture: stored tensors for the small cases,
    regenerated from the recorded seeds for the MSVD-shaped ones."""
    from oracle import s2vt_numpy as O
    c = golden_cfg(g)
    if "feats" in g:
        P = {k[6:]: g[k] for k in g if k.startswith("param/")}
        return P, g["feats"], g["targets"], g["mask"], c
    P = O.synth_params(c["V"], c["F"], c["H"], c["E"], seed=c["wseed"], out_scale=c["out_scale"],
                       eos_bias=c["eos_bias"])
    feats, targets, mask = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"], real_tokens=c["real"])
    return P, feats, targets, mask, c


def beam_rows_to_lists(rows):
    return [[int(t) for t in r if t >= 0] for r in rows]


def att_golden_inputs(g):
    """(params, feats, targets, mask, cfg) of an Att_Baseline fixture (tests/golden/make_golden_att.py)."""
    from oracle import att_numpy as A
    from oracle import s2vt_numpy as O
    c = golden_cfg(g)
    if "feats" in g:
        P = {k[6:]: g[k] for k in g if k.startswith("param/")}
        return P, g["feats"], g["targets"], g["mask"], c
    P = A.synth_params(c["V"], c["F"], c["H"], c["E"], seed=c["wseed"], out_scale=c["out_scale"], ctx_scale=c["ctx_scale"])
    feats, targets, mask = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"], real_tokens=c["real"])
    return P, feats, targets, mask, c
