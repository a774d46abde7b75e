#!/usr/bin/env python
"""Secondary metrics of BASELINE.json: greedy and beam-search captions/s (exact fp32 decode path) on one GPU, or sharded
over the ranks of a torchrun launch (videos are independent: no collective on the path, only a final gather of ids).

    python tools/bench_decode.py [--videos 1970] [--beam 5] [--batch 64] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import s2vt_b200
from s2vt_b200.dp import shard_range

CFG = dict(V=13000, F=4096, H=512, E=512, L=80)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=1970)
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--decode", default="x", choices=["x", "fp32"], help="x: tcgen05 fp16-split path, fp32: CUDA-core FFMA path")
    ap.add_argument("--beam-batch", type=int, default=230)
    ap.add_argument("--cpu-videos", type=int, default=0, help="also time the CPU port's greedy decode on this many videos")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                  # (at VERSION the banner ignores NCCL_DEBUG_FILE and lands on stdout)
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = s2vt_b200.S2VT(CFG["V"], CFG["F"], CFG["L"], dim_hid=CFG["H"], dim_embed=CFG["E"], decode_precision=args.decode).to(dev).eval()
    lo, hi = shard_range(args.videos, rank, world)
    n = hi - lo
    g = torch.Generator().manual_seed(77 + rank)
    feats = torch.randn(n, CFG["L"], CFG["F"], generator=g).to(dev)

    enq = []

    def timed(fn):
        fn()                                     # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h0 = time.perf_counter()
            fn()
            enq.append((time.perf_counter() - h0) * 1e3)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def greedy_all():
        with torch.no_grad():
            return [model(feats[i:i + args.batch], mode="test") for i in range(0, n, args.batch)]

    def beam_all():
        with torch.no_grad():
            return [model.beam_search_ids(feats[i:i + 256], beam_width=args.beam, max_beam_depth=30) for i in range(0, n, args.beam_batch)]

    ms_g = timed(greedy_all)
    enq_g = float(np.median(enq)); enq.clear()
    ms_b = timed(beam_all)
    enq_b = float(np.median(enq))
    out = {"n_gpus": world, "videos": args.videos, "greedy_captions_per_s": round(args.videos / (ms_g / 1e3), 1), "greedy_ms": round(ms_g, 2), "greedy_host_enqueue_ms": round(enq_g, 2), "beam_host_enqueue_ms": round(enq_b, 2),
           "beam_width": args.beam, "beam_captions_per_s": round(args.videos / (ms_b / 1e3), 1), "beam_ms": round(ms_b, 2),
           "precision": "fp32-grade (token ids bit-identical to the reference, tests/test_gpu_model_parity.py)", "decode_path": args.decode,
           "greedy_batch": args.batch, "beam_batch": args.beam_batch, "max_beam_depth": 30}
    if args.cpu_videos and rank == 0:
        from oracle.torch_port import S2VTCpuPort
        torch.set_num_threads(os.cpu_count() or 1)
        m = S2VTCpuPort(CFG["V"], CFG["F"], CFG["L"], CFG["H"], CFG["E"])
        x = torch.randn(args.cpu_videos, CFG["L"], CFG["F"])
        m.greedy(x[:2])
        t0 = time.perf_counter()
        m.greedy(x)
        out["cpu_port_greedy_captions_per_s"] = round(args.cpu_videos / (time.perf_counter() - t0), 1)
        out["cpu_cores"] = os.cpu_count()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
