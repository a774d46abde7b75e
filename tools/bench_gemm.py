#!/usr/bin/env python
"""Times every bf16 GEMM shape of the S2VT train step (B=64, MSVD shape) through the C ABI, for the persistent kernel and the
one-tile-per-CTA kernel.  Operands rotate over enough copies to exceed the 126 MB L2.  Prints one JSON line per shape.

    python tools/bench_gemm.py [--iters 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import s2vt_b200
from s2vt_b200 import lib as L
from s2vt_b200.lib import dense

SHAPES = [  # name, M, N, K, a_mn, b_mn, out_bf16, bias
    ("feat_linear fwd", 5120, 512, 4096, 0, 0, 1, 1),
    ("vid_rnn pre", 5120, 2048, 512, 0, 0, 0, 1),
    ("word_rnn pre (vid half)", 10176, 2048, 512, 0, 0, 0, 1),
    ("word_rnn pre (emb half)", 5056, 2048, 512, 0, 0, 0, 0),
    ("out_linear fwd (logits)", 5056, 13000, 512, 0, 0, 0, 1),
    ("out_linear dgrad", 5056, 512, 13000, 0, 1, 0, 0),
    ("out_linear wgrad", 13000, 512, 5056, 1, 1, 0, 0),
    ("word_rnn dgrad (vid half)", 10176, 512, 2048, 0, 1, 0, 0),
    ("word_rnn wgrad W_ih vid", 2048, 512, 10176, 1, 1, 0, 0),
    ("word_rnn wgrad W_ih emb", 2048, 512, 5056, 1, 1, 0, 0),
    ("word_rnn wgrad W_hh", 2048, 512, 10112, 1, 1, 0, 0),
    ("embedding dgrad", 5056, 512, 2048, 0, 1, 0, 0),
    ("vid_rnn wgrad W_ih", 2048, 512, 5120, 1, 1, 0, 0),
    ("vid_rnn dgrad", 5120, 512, 2048, 0, 1, 1, 0),
    ("feat_linear wgrad", 512, 4096, 5120, 1, 1, 0, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="", help="substring filter on the shape name")
    ap.add_argument("--persistent-only", action="store_true")
    args = ap.parse_args()
    lib = s2vt_b200.load()
    dev = torch.device("cuda:0")
    tot = {0: 0.0, 1: 0.0}
    for name, M, N, K, a_mn, b_mn, obf, has_bias in SHAPES:
        if args.only and args.only not in name:
            continue
        nbytes = 2 * (M * K + N * K) + (2 if obf else 4) * M * N
        ncopy = max(2, min(8, int(300e6 // nbytes) + 1))
        As = [torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16() for _ in range(ncopy)]
        Bs = [torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16() for _ in range(ncopy)]
        Cs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32) for _ in range(ncopy)]
        bias = torch.randn(N, device=dev) if has_bias else None
        res = {}
        for persistent in ((1,) if args.persistent_only else (1, 0)):
            lib.s2vt_gemm_bf16_set_mode(0, persistent)

            def run(i):
                A, B, C = As[i % ncopy], Bs[i % ncopy], Cs[i % ncopy]
                rc = lib.s2vt_gemm_bf16(L.stream_ptr(dev), M, N, K, L.ptr(A), M if a_mn else K, a_mn, L.ptr(B), N if b_mn else K, b_mn,
                                        L.ptr(C), dense(N), obf, L.ptr(bias), 0)
                L.check(rc, "gemm")
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(6_000_000)          # ~3 ms of GPU idle time: the host queues every launch before the first one runs
            e0.record()
            for i in range(args.iters):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / args.iters
            res[persistent] = us
            tot[persistent] += us
        res.setdefault(0, float("nan"))
        lib.s2vt_gemm_bf16_set_mode(0, 1)
        fl = 2.0 * M * N * K
        print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "layout": "%s%s" % ("T" if a_mn else "N", "T" if b_mn else "N"),
                          "persistent_us": round(res[1], 1), "persistent_tflops": round(fl / res[1] / 1e6, 1),
                          "tile_per_cta_us": round(res[0], 1), "tile_per_cta_tflops": round(fl / res[0] / 1e6, 1)}), flush=True)
    print(json.dumps({"total_us": {"persistent": round(tot[1], 1), "tile_per_cta": round(tot[0], 1)},
                      "flag": lib.s2vt_device_error_flag(L.stream_ptr(dev))}))


if __name__ == "__main__":
    main()
