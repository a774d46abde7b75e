#!/usr/bin/env python
"""In-kernel phase timeline of the tensor-core decode step kernels (CTA (0,0) of every launch, %globaltimer).

    python tools/trace_xdec.py --batch 512 [--mode greedy|beam]
Prints, per kernel class (grid shape is not recorded: classes are told apart by their position in the call), the mean of
  gap   = start - previous kernel's epilogue end          launch  = start -> weights requested (prologue)
  dep   = waiting for the predecessor (griddepcontrol)    fill    = first tile landed after the dependency resolved
  stream= first -> last tile landed                       drain   = last tile -> accumulators complete
  epi   = epilogue
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import s2vt_b200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--mode", default="greedy", choices=["greedy", "beam"])
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = s2vt_b200.S2VT(13000, 4096, 80, dim_hid=512, dim_embed=512).to(dev).eval()
x = torch.randn(args.batch, 80, 4096, device=dev)
lib = s2vt_b200.load()


def run():
    with torch.no_grad():
        if args.mode == "greedy":
            m(x, mode="test")
        else:
            m.beam_check_every = 0
            m.beam_search_ids(x, beam_width=5, max_beam_depth=30)


run(); run()
torch.cuda.synchronize()
N = 4096
buf = torch.zeros(N, 8, dtype=torch.int64, device=dev)
lib.s2vt_xdec_set_trace(buf.data_ptr(), N)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
n = lib.s2vt_xdec_set_trace(None, 0)
t = buf[:n].cpu().numpy().astype(np.float64) / 1e3          # us
print("%d xgemm launches, call %.3f ms" % (n, e0.elapsed_time(e1)))
L = 80
if args.mode == "greedy":
    # launch order inside s2vt_xdec_greedy: 2 store GEMMs, then per chunk of 16: 16 vid steps, 1 store, <=16 word steps; then 79 x (word step, vocab)
    kinds = ["store", "store"]
    T = 2 * L - 1
    for t0 in range(0, T, 16):
        t1 = min(T, t0 + 16)
        kinds += ["vid_step"] * (t1 - t0) + ["store_pre2"] + ["word_step"] * max(0, min(t1, L) - t0)
    for k in range(L - 1):
        kinds += ["dec_step", "vocab_argmax"]
else:
    kinds = ["store", "store"]
    for t0 in range(0, L, 16):
        t1 = min(L, t0 + 16)
        kinds += ["vid_step"] * (t1 - t0) + ["store_pre2"] + ["word_step"] * (t1 - t0)
    for d in range(30):
        kinds += ["beam_vid", "beam_word", "vocab_beam"]
assert len(kinds) == n, (len(kinds), n)
rows = {}
for i, k in enumerate(kinds):
    r = t[i]
    rows.setdefault(k, []).append([r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6] - r[0]])
print("%-14s %5s %8s %8s %8s %8s %8s %8s %8s" % ("kernel", "n", "launch", "dep", "fill", "stream", "drain", "epi", "total"))
for k, v in rows.items():
    a = np.median(np.array(v), axis=0)
    print("%-14s %5d %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f" % ((k, len(v)) + tuple(a)))
# spacing of consecutive decode steps on the side stream
if args.mode == "greedy":
    idx = [i for i, k in enumerate(kinds) if k == "dec_step"]
    print("dec_step period (start to start): %.2f us" % np.median(np.diff(t[idx, 0])))
    idx = [i for i, k in enumerate(kinds) if k == "vid_step"]
    print("vid_step period: %.2f us" % np.median(np.diff(t[idx, 0])))
else:
    idx = [i for i, k in enumerate(kinds) if k == "beam_vid"]
    print("beam depth period: %.2f us" % np.median(np.diff(t[idx, 0])))
