#!/usr/bin/env python
"""Where the bf16 training path's logit error comes from (runs on CPU; no GPU, no reference import needed).

The MSVD-shaped golden case (tests/golden/msvd.npz: B=8, V=13000, H=E=512) is re-run in plain torch fp32 with bf16 rounding
switched on one place at a time -- exactly the places ANY bf16 tensor-core design has to round:
   W   weights and input features rounded to bf16 (operands of every product), fp32 accumulation and state
   +h  the hidden state h_t rounded to bf16 where it is an operand (recurrent product, next layer's input, vocab projection)
   +x  the other activations that are operands (feat_linear output, embeddings) rounded to bf16
   +z  logits stored as bf16 (the fused-loss path keeps them in bf16; forward(mode='train') returns fp32)
and the result is compared with the reference's fp32 logits (golden `logits_sample`).  sigma_z is the standard deviation of the
reference logits.  The GPU kernels add tanh.approx / ex2.approx on top (MUFU, ~2^-11 relative) -- measured on the GPU by
tests/test_gpu_bf16.py::test_bf16_train_step_vs_reference_golden, which prints the realised error next to this floor.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from conftest import golden_inputs, load_golden


def r(x, on):
    return x.to(torch.bfloat16).to(torch.float32) if on else x


def lstm(pre, w_hh, round_h):
    """pre [T,B,4H] (input-side pre-activations incl. biases) -> out [T,B,H]; gate order i,f,g,o"""
    T, B, G = pre.shape
    H = G // 4
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    out = []
    for t in range(T):
        g = pre[t] + r(h, round_h) @ w_hh.t()
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out)


def forward(P, feats, targets_in, rw, rh, rx, rz):
    B, L, F = feats.shape
    H = P["vid_rnn.weight_hh_l0"].shape[1]
    E = P["embedding.weight"].shape[1]
    W = {k: (r(v, rw) if v.dim() == 2 else v) for k, v in P.items()}
    x = r(feats, rw).reshape(B * L, F) @ W["feat_linear.weight"].t() + P["feat_linear.bias"]
    x = x.reshape(B, L, H).transpose(0, 1)                                       # [L,B,H]
    b1 = P["vid_rnn.bias_ih_l0"] + P["vid_rnn.bias_hh_l0"]
    b2 = P["word_rnn.bias_ih_l0"] + P["word_rnn.bias_hh_l0"]
    T = 2 * L - 1
    pre1 = torch.zeros(T, B, 4 * H) + b1
    pre1[:L] += r(x, rx) @ W["vid_rnn.weight_ih_l0"].t()
    out1 = lstm(pre1, W["vid_rnn.weight_hh_l0"], rh)
    pre2 = r(out1, rh) @ W["word_rnn.weight_ih_l0"][:, E:].t() + b2
    emb = W["embedding.weight"][targets_in].transpose(0, 1)                      # [L-1,B,E]  (table rounded with the weights)
    pre2[L:] += r(emb, rx) @ W["word_rnn.weight_ih_l0"][:, :E].t()
    out2 = lstm(pre2, W["word_rnn.weight_hh_l0"], rh)
    z = r(out2[L:], rh) @ W["out_linear.weight"].t() + P["out_linear.bias"]       # [L-1,B,V]
    return r(z, rz).transpose(0, 1).contiguous()                                  # [B,L-1,V]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    g = load_golden("msvd")
    Pn, feats, targets, mask, c = golden_inputs(g)
    P = {k: torch.from_numpy(v.copy()) for k, v in Pn.items()}
    f, t = torch.from_numpy(feats), torch.from_numpy(targets)[:, :-1]
    ref = g["logits_sample"].astype(np.float64)
    sig = ref.std()
    print("reference logits: |z|max %.3f  sigma_z %.4f  (msvd golden, %d samples, stride 997)" % (np.abs(ref).max(), sig, ref.size))
    print("%-28s %12s %12s %10s" % ("rounding switched on", "max |err|", "rms err", "max/sigma"))
    for tag, flags in (("none (fp32 restatement)", (0, 0, 0, 0)), ("W", (1, 0, 0, 0)), ("W +h", (1, 1, 0, 0)), ("W +h +x", (1, 1, 1, 0)),
                       ("W +h +x +z", (1, 1, 1, 1)), ("h only", (0, 1, 0, 0)), ("z only", (0, 0, 0, 1))):
        with torch.no_grad():
            z = forward(P, f, t, *[bool(v) for v in flags]).numpy().reshape(-1)[::997].astype(np.float64)
        e = np.abs(z - ref)
        print("%-28s %12.3e %12.3e %10.3f" % (tag, e.max(), np.sqrt((e ** 2).mean()), e.max() / sig))


if __name__ == "__main__":
    main()
