"""Generate golden vectors from the UNMODIFIED reference module (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports S2VT / MaskCriterion from /root/reference (read-only), runs them on CPU fp32 at fixed
seeds, and writes small .npz fixtures next to this file.  The fixtures travel to the GPU box;
/root/reference does not.  Weights and inputs come from oracle.s2vt_numpy.synth_params/
synth_batch (numpy PCG64), so the MSVD-shaped cases store only outputs (tokens, loss, strided
samples of logits and gradients) and the 83 MB weight set is rebuilt from its seed.
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

from S2VTModel import S2VT  # noqa: E402  (reference)
from utils import MaskCriterion  # noqa: E402  (reference)

from oracle import s2vt_numpy as O  # noqa: E402

CASES = {
    # name: dims, batch, seeds, peaky-logit knobs, what to store
    "tiny": dict(V=40, F=24, H=16, E=12, L=6, B=3, real=5, wseed=11, dseed=12, out_scale=30.0, eos_bias=2.0,
                 full=True, beams=(1, 3, 5), beam_videos=3),
    "mid": dict(V=300, F=160, H=128, E=96, L=12, B=5, real=9, wseed=21, dseed=22, out_scale=12.0, eos_bias=1.0,
                full=True, beams=(3, 5), beam_videos=5),
    "msvd": dict(V=13000, F=4096, H=512, E=512, L=80, B=8, real=28, wseed=0, dseed=1234, out_scale=1.0,
                 eos_bias=0.0, full=False, beams=(3, 5), beam_videos=2),
    "msvd_peaky": dict(V=13000, F=4096, H=512, E=512, L=80, B=4, real=28, wseed=5, dseed=77, out_scale=40.0,
                       eos_bias=3.0, full=False, beams=(3,), beam_videos=2),
    # BASELINE configs[1], the benched configuration: batch 64, plus a 5-step Adam trajectory over 5 different batches
    "c2": dict(V=13000, F=4096, H=512, E=512, L=80, B=64, real=28, wseed=0, dseed=4000, out_scale=1.0,
               eos_bias=0.0, full=False, beams=(), beam_videos=0, adam_steps=5),
    "paper": dict(V=5000, F=2048, H=1000, E=500, L=80, B=2, real=28, wseed=31, dseed=32, out_scale=1.0,
                  eos_bias=0.0, full=False, beams=(), beam_videos=0),
}

SAMPLE_STRIDE = 997


def sample(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a.reshape(-1)[::SAMPLE_STRIDE])


def build_reference(c, P):
    m = S2VT(c["V"], c["F"], c["L"], dim_hid=c["H"], dim_embed=c["E"], sos_ix=3, eos_ix=4)
    sd = {k: torch.from_numpy(v.copy()) for k, v in P.items()}
    m.load_state_dict(sd, strict=True)
    return m


def run_case(name, c):
    t0 = time.time()
    P = O.synth_params(c["V"], c["F"], c["H"], c["E"], seed=c["wseed"], out_scale=c["out_scale"],
                       eos_bias=c["eos_bias"])
    feats, targets, mask = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"], real_tokens=c["real"])
    model = build_reference(c, P)
    crit = MaskCriterion()
    tf = torch.from_numpy(feats).requires_grad_(True)      # dataloader.py:38 -> feats require grad
    tt = torch.from_numpy(targets)
    tm = torch.from_numpy(mask)
    model.train()
    logits = model(tf, targets=tt[:, :-1], mode="train")    # train.py:120
    loss = crit(logits, tt, tm)                              # train.py:122
    loss.backward()                                          # train.py:124
    grads = {k: p.grad.numpy() for k, p in model.named_parameters()}
    grads["feats"] = tf.grad.numpy()
    model.eval()
    with torch.no_grad():
        greedy = model(tf.detach(), mode="test").numpy()     # eval.py:52
    out = dict(loss=np.float32(loss.item()), greedy=greedy.astype(np.int64),
               greedy_min_margin=O.greedy_margins(P, feats).min())
    meta = {k: c[k] for k in ("V", "F", "H", "E", "L", "B", "real", "wseed", "dseed", "out_scale", "eos_bias")}
    for k, v in meta.items():
        out["cfg_" + k] = np.asarray(v)
    lg = logits.detach().numpy()
    if c["full"]:
        out["logits"] = lg
        for k, v in P.items():
            out["param/" + k] = v
        out["feats"], out["targets"], out["mask"] = feats, targets, mask
        for k, v in grads.items():
            out["grad/" + k] = v
    else:
        out["logits_sample"] = sample(lg)
        for k, v in grads.items():
            out["grad_sample/" + k] = sample(v)
            out["grad_norm/" + k] = np.float64(np.linalg.norm(v.astype(np.float64)))
            if v.size <= 20000:                              # biases: keep the whole vector
                out["grad_full/" + k] = v
    # one Adam step (train.py:125) on the full-tensor cases: updated params
    if c["full"]:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        opt.step()
        for k, p in model.named_parameters():
            out["adam1/" + k] = p.detach().numpy().copy()
        model = build_reference(c, P)                        # restore weights for beam search
        model.eval()
    # N-step training trajectory (train.py:116-127 with Adam lr=1e-4): loss before every step, sampled weights after the last
    if c.get("adam_steps"):
        model = build_reference(c, P)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        losses = []
        for i in range(c["adam_steps"]):
            f_i, t_i, m_i = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"] + i, real_tokens=c["real"])
            opt.zero_grad()
            lg_i = model(torch.from_numpy(f_i).requires_grad_(True), targets=torch.from_numpy(t_i)[:, :-1], mode="train")
            l_i = crit(lg_i, torch.from_numpy(t_i), torch.from_numpy(m_i))
            l_i.backward()
            opt.step()
            losses.append(l_i.item())
        out["traj_loss"] = np.asarray(losses, np.float64)
        for k, p in model.named_parameters():
            v = p.detach().numpy()
            out["traj_param_sample/" + k] = sample(v)
            out["traj_delta_norm/" + k] = np.float64(np.linalg.norm((v - P[k]).astype(np.float64)))
        out["cfg_adam_steps"] = np.asarray(c["adam_steps"])
        model = build_reference(c, P)
        model.eval()
    for bw in c["beams"]:
        nv = c["beam_videos"]
        with torch.no_grad():
            sents = model(tf.detach()[:nv], mode="beam_search", beam_width=bw, max_beam_depth=30)  # eval.py:88
        flat = np.full((nv, 31), -1, np.int64)
        for i, s in enumerate(sents):
            ids = [int(x.item()) for x in s]
            flat[i, :len(ids)] = ids
        out["beam%d" % bw] = flat
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-11s loss=%.6f min greedy margin=%.3e  %.1fs" % (name, out["loss"], out["greedy_min_margin"],
                                                             time.time() - t0), flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    which = sys.argv[1:] or list(CASES)
    for n in which:
        run_case(n, CASES[n])
