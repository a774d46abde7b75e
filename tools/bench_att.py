#!/usr/bin/env python
"""BASELINE.json configs[4]: Att_Baseline (attention_baseline.py) training throughput on one GPU or data-parallel under torchrun.
One step = zero_grad + forward + MaskCriterion + backward + (gradient all-reduce) + Adam, batch 64/GPU, MSVD shape (80 x 4096
features, V = 13000, H = E = 512), synthetic data, random-init weights.  Default: DataParallelTrainer (fused forward_loss, one
all-reduce bucket, CUDA-graph replay); --api times the unchanged loop body (module forward + MaskCriterion + backward + step).
Prints one JSON line.

    python tools/bench_att.py [--steps 20] [--precision bf16|fp32] [--api]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_att.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import s2vt_b200
from bench import CFG, synth_batch


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--api", action="store_true")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--gpus", type=int, default=None, help="(informational: the launcher decides the world size)")
    args = ap.parse_args(argv)
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
    model = s2vt_b200.Att_Baseline(CFG["V"], CFG["F"], CFG["L"], dim_hid=CFG["H"], dim_embed=CFG["E"], train_precision=args.precision).to(dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-4)
    crit = s2vt_b200.MaskCriterion()
    flat_g = None
    batches = [synth_batch(CFG["B"], 99 + rank * 10 + i, device=dev) for i in range(3)]

    from s2vt_b200.dp import DataParallelTrainer
    trainer = None
    if not args.api and args.precision != "fp32":
        trainer = DataParallelTrainer(model, opt)
        for f, t, _ in batches:
            trainer.register_inputs(f, t)

    def step(i):
        f, t, m = batches[i % 3]
        if trainer is not None:
            return trainer.step(f, t, m)
        opt.zero_grad(set_to_none=True)
        loss = crit(model(f, targets=t[:, :-1], mode="train"), t, m)
        loss.backward()
        if world > 1:
            for p in model.parameters():
                dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
        opt.step()
        return loss

    for i in range(max(args.warmup, 6)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    if rank == 0:
        print(json.dumps({"metric": "Att_Baseline train videos/sec", "value": round(world * CFG["B"] / (ms / 1e3), 1), "unit": "videos/s",
                          "ms_per_step": round(ms, 3), "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 6), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
                          "config": {"workload": "BASELINE configs[4]: attention_baseline.py encoder-decoder, batch %d/GPU, MSVD shape 80x4096, "
                                                 "V=13000, H=E=512, random init" % CFG["B"], "global_batch": world * CFG["B"], "parallelism": "dp%d" % world},
                          "precision": args.precision, "batch_per_gpu": CFG["B"], "loss": float(loss.item()),
                          "path": "DataParallelTrainer (forward_loss, graph replays %d)" % trainer.replays if trainer is not None else "module + MaskCriterion + FusedAdam.step, eager",
                          "launches_total": int(s2vt_b200.launch_count())}))
    if world > 1:
        if trainer is not None:
            trainer.release_graphs()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
