#!/usr/bin/env python
"""One greedy and one beam-5 call of the tensor-core decode path at the MSVD shape, for ncu (launch list / --set full).

    ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file out.csv python tools/profile_decode.py --batch 256
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import s2vt_b200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--mode", default="both", choices=["greedy", "beam", "both"])
ap.add_argument("--depth", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = s2vt_b200.S2VT(13000, 4096, 80, dim_hid=512, dim_embed=512).to(dev).eval()
x = torch.randn(args.batch, 80, 4096, device=dev)
with torch.no_grad():
    if args.mode in ("greedy", "both"):
        t = m(x, mode="test")
    if args.mode in ("beam", "both"):
        b = m.beam_search_ids(x, beam_width=5, max_beam_depth=args.depth)
torch.cuda.synchronize()
assert s2vt_b200.load().s2vt_device_error_flag(None) == 0
print("ok")
