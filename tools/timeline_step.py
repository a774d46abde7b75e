#!/usr/bin/env python
"""Timeline of one train step of the bench workload: start / end of every C-ABI call (CUDA events on the launching stream,
relative to the step's first call), plus host enqueue time per step against device time per step.

    python tools/timeline_step.py [--batch 64] > gpurun_out/timeline.txt
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import s2vt_b200  # noqa: E402
from s2vt_b200 import ops  # noqa: E402
from s2vt_b200.dp import DataParallelTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    C = bench.CFG
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:                                   # under torchrun: the data-parallel step (rank 0 prints)
        import torch.distributed as dist
        os.environ.setdefault("NCCL_MIN_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
        if rank != 0:
            sys.stdout = open(os.devnull, "w")
    torch.manual_seed(0)
    model = s2vt_b200.S2VT(C["V"], C["F"], C["L"], dim_hid=C["H"], dim_embed=C["E"], train_precision="bf16").to(dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-4)
    trainer = DataParallelTrainer(model, opt)
    batches = [bench.synth_batch(args.batch, 1234 + i, device=dev) for i in range(4)]
    for i in range(5):
        trainer.step(*batches[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        trainer.step(*batches[i % 4])
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("host enqueue %.3f ms/step, wall incl. device %.3f ms/step" % (1e3 * t_enq / args.steps, 1e3 * t_all / args.steps))
    print("device error flag:", s2vt_b200.load().s2vt_device_error_flag(None))
    if trainer.use_graph and trainer._graphs:
        import collections
        import ctypes
        ent = next(iter(trainer._graphs.values()))
        if ent[4] is not None:
            pr = (ctypes.c_int * 1024)()
            n = ctypes.c_int(0)
            s2vt_b200.load().s2vt_graph_kernel_priorities(ent[0].raw_cuda_graph(), pr, 1024, ctypes.byref(n))
            print("kernel-node priorities of the captured step (priority: nodes):", dict(collections.Counter(pr[:min(n.value, 1024)])))
    if trainer.use_graph:
        # device-side timeline of a REPLAYED step: %globaltimer stamps captured into a fresh graph around every call
        buf = torch.zeros(512, dtype=torch.int64, device=dev)
        ops.MARKS = dict(buf=buf, names=[])
        f, t, m = bench.synth_batch(args.batch, 999, device=dev)
        trainer.step(f, t, m)                       # capture (+ first replay) with the stamps in
        names = list(ops.MARKS["names"])
        ops.MARKS = None
        for _ in range(3):
            trainer.step(f, t, m)
        torch.cuda.synchronize()
        ts = buf.cpu().tolist()[:len(names)]
        t0g = min(ts)
        open_ = {}
        print("graph replay: step span %.1f us (with %d stamp kernels inside)" % ((max(ts) - t0g) / 1e3, len(names)))
        for nm, tv in zip(names, ts):
            kind, tag = nm[0], nm[2:]
            if kind == "B":
                open_.setdefault(tag, []).append(tv)
            else:
                b = open_[tag].pop(0)
                print("%9.1f %9.1f %8.1f  %s" % ((b - t0g) / 1e3, (tv - t0g) / 1e3, (tv - b) / 1e3, tag))
        print()
    with ops.profile() as prof:
        trainer.step(*batches[0])
        torch.cuda.synchronize()
    recs = prof.records
    first = recs[0][1]
    rows = []
    for tag, e0, e1, flops, nbytes in recs:
        rows.append((first.elapsed_time(e0) * 1e3, first.elapsed_time(e1) * 1e3, tag))
    end = max(r[1] for r in rows)
    print("step span %.1f us (instrumented)" % end)
    for a, b, tag in rows:
        print("%9.1f %9.1f %8.1f  %s" % (a, b, b - a, tag))


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)
