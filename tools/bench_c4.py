#!/usr/bin/env python
"""BASELINE.json configs[3]: S2VT at the paper's sizing (ResNet152 2048-d features, hidden 1000, embedding 500 = the constructor
default, S2VTModel.py:11), batch 256 per GPU, one train step = forward_loss + backward + FusedAdam.  H = 1000 is outside the cluster
recurrence's range (H % 128 == 0, H <= 512): train_precision 'auto' / 'bf16' runs it on the per-step tensor-core engine
(engine_step.py, csrc/lstm_step_bf16_sm100.cu), 'fp32' on the exact CUDA-core path.  Prints one JSON line with a per-kernel breakdown.

    python tools/bench_c4.py [--steps 10] [--batch 256] [--precision auto|bf16|fp32]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import s2vt_b200
from s2vt_b200.dp import DataParallelTrainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    V, F, L, H, E, B = 13000, 2048, 80, 1000, 500, args.batch
    torch.manual_seed(0)
    model = s2vt_b200.S2VT(V, F, L, dim_hid=H, dim_embed=E, train_precision=args.precision).to(dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-4)
    trainer = DataParallelTrainer(model, opt)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(B, L, F, generator=g).to(dev)
    targets = torch.zeros(B, L, dtype=torch.int64)
    targets[:, 0] = 3
    targets[:, 1:27] = torch.randint(5, V, (B, 26), generator=g)
    targets[:, 27] = 4
    targets = targets.to(dev)
    trainer.register_inputs(feats, targets)
    for _ in range(max(args.warmup, 3)):
        trainer.step(feats, targets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = trainer.step(feats, targets)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    from s2vt_b200 import ops
    with ops.profile() as prof:
        trainer.step(feats, targets)
    kern = {k: {"calls": c, "ms": round(t, 3), "tflops": round(f / (t * 1e-3) / 1e12, 1) if f and t else None}
            for k, (c, t, f, b) in sorted(prof.summary().items(), key=lambda kv: -kv[1][1])[:12]}
    eng = model._bf16_engine().__name__.split(".")[-1] if model._use_bf16() else "fp32 exact (CUDA cores)"
    print(json.dumps({"metric": "S2VT train videos/sec, paper sizing (configs[3])", "value": round(B / (ms / 1e3), 1), "ms_per_step": round(ms, 2),
                      "n_gpus": 1, "precision": args.precision, "engine": eng, "graph_replays": trainer.replays, "batch_per_gpu": B,
                      "kernels_one_eager_step": kern,
                      "dims": {"V": V, "F": F, "L": L, "H": H, "E": E}, "loss": float(loss.item()),
                      "device_error_flag": int(s2vt_b200.load().s2vt_device_error_flag(None))}))


if __name__ == "__main__":
    main()
