#!/usr/bin/env python
"""Small end-to-end exercise of the round-2 kernels for compute-sanitizer (memcheck): tensor-core decode (greedy + beam) at ragged
shapes, the step engine (forward + BPTT with split-K) and the fused beam bookkeeping.

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import s2vt_b200

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (V, F, H, E, L, B) in ((77, 40, 24, 28, 5, 7), (300, 64, 72, 40, 6, 133)):
    m = s2vt_b200.S2VT(V, F, L, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    x = torch.randn(B, L, F, device=dev)
    t = torch.randint(0, V, (B, L), device=dev)
    with torch.no_grad():
        g = m(x, mode="test")
        toks, lens = m.beam_search_ids(x, beam_width=3, max_beam_depth=6)
    loss = m.forward_loss(x, t)
    loss.backward()
    torch.cuda.synchronize()
    print("ok", V, F, H, E, L, B, float(loss), g.shape, toks.shape)
assert s2vt_b200.load().s2vt_device_error_flag(None) == 0
print("done")
