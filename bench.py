#!/usr/bin/env python
"""bench.py -- train videos/s of the S2VT hot path on N B200s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = the reference's train-loop body (train.py:116-127): zero_grad + forward(mode='train') + MaskCriterion +
backward + (gradient all-reduce) + Adam step on one batch of 64 synthetic MSVD-shaped videos per GPU
(80 x 4096 fp32 features, 28 real tokens padded to 80, V = 13000, H = E = 512, random-init weights).
Rank 0 prints ONE JSON line (see the keys below); `--impl reference` times the CPU port of the reference's own
PyTorch path (oracle/torch_port.py) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

# workload = BASELINE.json configs[1] / SURVEY.md section 8(d)
CFG = dict(V=13000, F=4096, H=512, E=512, L=80, B=64, real_tokens=28)
FLOP_PER_VIDEO_TRAIN = 7.827e9        # algorithmic work, structural zeros skipped
FLOP_PER_VIDEO_FWD = 2.721e9


# profiler tag -> kernel name in the ncu launch list (profiles/*_ncu_launch_list_summary.txt)
KERNEL_OF_TAG = {"gemm_bf16_persist": "gemm_bf16_persist_kernel", "gemm_bf16_tile": "gemm_bf16_kernel", "lstm_fwd_bf16": "lstm_fwd_cluster_kernel",
                 "lstm_bwd_bf16": "lstm_bwd_cluster_kernel", "adam_f32": "adam_kernel", "colsum_bf16": "colsum_bf16_v8_kernel",
                 "ce_bf16": "ce_dlogits_inplace_kernel", "cast_bf16": "cast_bf16_kernel",
                 "gemm_bf16_gated": "gemm_bf16_persist_kernel (gated: resident beside the sweeps, duration includes waiting for them)"}


def load_traffic():
    """DRAM bytes per launch of each kernel from the newest committed `ncu --set full` capture (profiles/*_dram_traffic.json)."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        cands = sorted(f for f in os.listdir(pdir) if f.endswith("_dram_traffic.json"))
        return (json.load(open(os.path.join(pdir, cands[-1]))), cands[-1]) if cands else ({}, None)
    except Exception:
        return {}, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def synth_batch(B, seed, device=None, pinned=False):
    """feats ~ N(0,1); captions <sos>=3, 26 words in [5,V), <eos>=4, <pad>=0 to length 80."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, CFG["L"], CFG["F"], generator=g)
    targets = torch.zeros(B, CFG["L"], dtype=torch.int64)
    r = CFG["real_tokens"]
    targets[:, 0] = 3
    targets[:, 1:r - 1] = torch.randint(5, CFG["V"], (B, r - 2), generator=g)
    targets[:, r - 1] = 4
    mask = torch.zeros(B, CFG["L"])
    mask[:, :r] = 1
    if pinned:
        return feats.pin_memory(), targets.pin_memory(), mask.pin_memory()
    if device is not None:
        return feats.to(device), targets.to(device), mask.to(device)
    return feats, targets, mask


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line).  nvidia-smi needs a few hundred
    milliseconds to start and the timed region lasts tens of milliseconds, so the sampler is started before the warm-up, samples every
    20 ms with nvidia-smi's own timestamps, and stop(t0, t1) keeps the samples that fall inside the timed window [t0, t1] (host
    clock); if none does (a very short region), the samples within half a second of it are used and `window` says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        try:
            self.t.join(timeout=1)
        except Exception:
            pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                rows.append((self._epoch(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
            except ValueError:
                continue
        window = "all"
        if t0 is not None and t1 is not None and rows and all(r[0] is not None for r in rows):
            inside = [r for r in rows if t0 - 0.02 <= r[0] <= t1 + 0.02]
            if inside:
                rows, window = inside, "timed region"
            else:
                rows, window = [r for r in rows if t0 - 0.5 <= r[0] <= t1 + 0.5] or rows, "within 0.5 s of the timed region"
        sm, mx, reasons = [r[1] for r in rows], [r[2] for r in rows], set()
        for r in rows:
            for nm, v in zip(names, r[3]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_reference_run(steps, warmup, B, threads=None):
    """Times the CPU port of the reference's train step (oracle/torch_port.py) on the host cores."""
    from oracle.torch_port import S2VTCpuPort
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = S2VTCpuPort(CFG["V"], CFG["F"], CFG["L"], CFG["H"], CFG["E"])
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    feats, targets, mask = synth_batch(B, 1234)
    for _ in range(warmup):
        m.train_step(opt, feats, targets, mask)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        m.train_step(opt, feats, targets, mask)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts))
    return dict(value=B / (ms / 1e3), ms_per_step=ms, cores=threads, B=B)


def run_reference(args, rank):
    if rank != 0:
        return
    B = 16                                   # bounded sample: Opt().batch_size videos per step (train.py:36)
    steps, warmup = min(args.steps, 8), min(args.warmup, 2)
    r = cpu_reference_run(steps, max(1, warmup), B)
    sample = "%d train steps of %d videos (same shapes as the workload; batch bounded so the run ends in minutes)" % (steps, B)
    line = {
        "impl": "reference", "metric": "train videos/sec", "value": r["value"], "unit": "videos/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "S2VT train step, MSVD shape 80x4096, V=13000, H=E=512, batch %d on CPU" % B, "batch_per_step": B},
        "cpu_baseline": {"value": r["value"], "unit": "videos/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("S2VT_BENCH_PRECISION", "auto"))
    ap.add_argument("--batch", type=int, default=CFG["B"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import s2vt_b200
    from s2vt_b200 import ops
    from s2vt_b200.dp import DataParallelTrainer

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback on the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_MIN_CTAS", "16")          # the 20-27 MB gradient buckets are bandwidth-bound: measured +4% at 8 GPUs
        if not os.environ.get("S2VT_KEEP_NCCL_DEBUG"):
            os.environ.pop("NCCL_DEBUG", None)        # NCCL's version banner goes to stdout; rank 0 must print ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    numa_node = None
    if world > 1 and os.environ.get("S2VT_NUMA_BIND", "1") != "0":
        from s2vt_b200.dp import bind_to_local_numa
        numa_node = bind_to_local_numa(local_rank)          # before any pinned host buffer exists (end-to-end loop feeds from them)
    s2vt_b200.load()
    peaks = load_peaks()
    B = args.batch
    precision = args.precision
    if precision == "auto":
        precision = "bf16" if getattr(s2vt_b200, "BF16_TRAIN_READY", False) else "fp32"

    torch.manual_seed(0)                      # identical replicas on every rank
    model = s2vt_b200.S2VT(CFG["V"], CFG["F"], CFG["L"], dim_hid=CFG["H"], dim_embed=CFG["E"], train_precision=precision).to(dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-4)
    trainer = DataParallelTrainer(model, opt)

    # device-resident inputs, rotated so that consecutive steps do not reuse L2 (4 x 84 MB > 126 MB L2)
    n_rot = 4
    batches = [synth_batch(B, 1234 + rank * 100 + i, device=dev) for i in range(n_rot)]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(i):
        f, t, m = batches[i % n_rot]
        return trainer.step(f, t, m)

    # (with CUDA graphs: two eager steps, then one capture per rotating input buffer -- all of it before the timed region)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(max(args.warmup, n_rot + 2) if trainer.use_graph else args.warmup):
        step(i)
    sync_all()
    launches0 = s2vt_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = s2vt_b200.launch_count() - launches0
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step / 1e3)
    final_loss = float(loss.item())

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs + D2H of the loss, every step.
    # Next step's inputs are prefetched on a copy stream while this step computes (each copy is inside the timed region).
    host = [synth_batch(B, 4321 + rank * 100 + i, pinned=True) for i in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.empty((), pin_memory=True)

    def run_e2e(host):
        dbuf = [(torch.empty_like(host[0][0], device=dev), torch.empty_like(host[0][1], device=dev)) for _ in range(2)]

        def e2e_loop(n):
            evs = [None, None]

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    dbuf[i % 2][0].copy_(host[i % 2][0], non_blocking=True)
                    dbuf[i % 2][1].copy_(host[i % 2][1], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(copy_stream); evs[i % 2] = ev
            prefetch(0)
            for i in range(n):
                torch.cuda.current_stream().wait_event(evs[i % 2])
                if i + 1 < n:
                    copy_stream.wait_stream(torch.cuda.current_stream()) if i >= 1 else None
                    prefetch(i + 1)
                l = trainer.step(dbuf[i % 2][0], dbuf[i % 2][1])
                loss_host.copy_(l.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_loop(2)
        sync_all()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        sync_all()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        return world * B / (t2.item() / args.steps / 1e3), host[0][0].numel() * host[0][0].element_size() + host[0][1].numel() * 8

    e2e_value, h2d = run_e2e(host)
    # the same loop fed from a pinned bf16 feature store (data.DeviceFeatureStore(dtype=bfloat16) semantics: features rounded once at
    # load time to what the tensor-core path rounds them to at every step; identical loss and gradients) -- half the H2D bytes
    e2e_bf16 = None
    if precision == "bf16" and trainer.use_graph and trainer.max_graphs >= n_rot + 4:
        host_bf = [(f.to(torch.bfloat16).pin_memory(), t, m) for f, t, m in host]
        v, nb = run_e2e(host_bf)
        e2e_bf16 = {"value": round(v, 2), "unit": "videos/s", "h2d_bytes_per_step": int(nb), "d2h_bytes_per_step": 4,
                    "note": "features held as bf16 in pinned host memory (rounded once at load); `e2e` is the float32-input figure"}
    clocks = sampler.stop(t_wall0, time.time())                # the window spans the timed regions (device-resident and end-to-end)

    # ---- per-kernel-family timing of one extra instrumented step (CUDA events on the launching stream)
    with ops.profile() as prof:
        for i in range(2):
            step(i)
    detail = prof.summary()
    summ, gemm_detail = {}, []
    for tag, (calls, ms, flops, nbytes) in detail.items():       # per-shape GEMM tags fold into one family line
        base = tag.split("[")[0]
        c0, m0, f0, b0 = summ.get(base, (0, 0.0, 0.0, 0.0))
        summ[base] = (c0 + calls, m0 + ms, f0 + flops, b0 + nbytes)
        if "[" in tag:
            gemm_detail.append({"shape": tag[tag.index("[") + 1:-1], "us": round(1e3 * ms / calls, 1),
                                "tflops": round(flops / calls / (ms / calls * 1e-3) / 1e12, 1)})
    gemm_detail.sort(key=lambda d: -d["us"])
    kernels = []
    tot_ms = sum(v[1] for v in summ.values()) or 1.0
    for tag, (calls, ms, flops, nbytes) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
        ms1 = ms / 2.0
        kernels.append({"op": tag, "calls_per_step": calls // 2, "ms_per_step": round(ms1, 4), "share": round(ms / tot_ms, 4),
                        "tflops": round(flops / 2.0 / (ms1 * 1e-3) / 1e12, 3) if flops else None,
                        "gbs": round(nbytes / 2.0 / (ms1 * 1e-3) / 1e9, 1) if nbytes else None})
    traffic, traffic_src = load_traffic()

    def roof(k):
        kern = KERNEL_OF_TAG.get(k["op"], k["op"])
        tr = traffic.get(kern, {}).get("dram_bytes_per_launch")
        base = {"kernel": kern, "launches_per_step": k["calls_per_step"], "us_per_launch": round(1e3 * k["ms_per_step"] / max(1, k["calls_per_step"]), 2),
                "share_of_step": k["share"], "traffic": None if tr is None else round(tr), "traffic_source": traffic_src if tr is not None else None}
        if k["tflops"] is not None:
            base.update({"bound": "tensor", "achieved": k["tflops"], "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                         "frac": round(k["tflops"] / peaks["tf_sust"], 5), "peak_source": peaks["src"] + " (sustained bf16, kernel timed inside the step)"})
            if k["op"].startswith("lstm_"):
                # a chain of T = 159 dependent steps: what bounds it is the latency of one step (tensor-memory MMA -> activations -> DSMEM
                # exchange across the 16-CTA cluster); the tensor fraction is reported for information
                base["us_per_timestep"] = round(base["us_per_launch"] / (2 * CFG["L"] - 1), 3)
                # on-chip traffic of one step, all clusters of the sweep: every CTA's MMAs read its 4H/CS x H bf16 weight slice (tensor
                # memory) and the 16-column h / dgates operand (shared memory); the exchange moves B x H bf16 through distributed
                # shared memory to each of the CS = H/32 CTAs of a cluster
                H_, B_ = CFG["H"], B
                step_s = base["us_per_timestep"] * 1e-6
                base["onchip_operand_gbs"] = round((4 * H_ * H_ * 2 * ((B_ + 15) // 16) + B_ * H_ * 2 * (H_ // 32)) / step_s / 1e9, 1)
                base["dsmem_exchange_gbs"] = round(B_ * H_ * 2 * (H_ // 32) / step_s / 1e9, 1)
                base["note"] = ("serial recurrence (159 dependent steps, north-star target < 5 us/step at batch 64); two sweeps run side by "
                                "side as a wave front, so a launch's duration includes the trailing sweep's wait for the leading one")
        else:
            base.update({"bound": "hbm", "achieved": k["gbs"], "peak": peaks["hbm"], "unit": "GB/s", "frac": round(k["gbs"] / peaks["hbm"], 5),
                         "peak_source": peaks["src"]})
        return base
    roofline = roof(kernels[0]) if kernels else None            # the dominant kernel (largest share of the instrumented step)
    roofline_all = [roof(k) for k in kernels]

    line = {
        "metric": "train videos/sec", "value": round(value, 2), "unit": "videos/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: S2VT train step, batch %d/GPU, MSVD shape 80x4096 fp32 feats, 28-token captions "
                               "padded to 80, V=13000, H=E=512, random init" % B,
                   "global_batch": world * B, "parallelism": "dp%d" % world, "precision": precision,
                   "l2_policy": "inputs rotate over 4 device-resident batches (336 MB > 126 MB L2)",
                   "launch": "CUDA graph replay of the whole step (one graph per input buffer)" if trainer.use_graph and trainer._graphs else "eager",
                   "host_numa_node_rank0": numa_node},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": round(e2e_value, 2), "unit": "videos/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
        "e2e_bf16_store": e2e_bf16,
        "roofline": roofline, "roofline_all": roofline_all, "kernels": kernels, "gemm_detail": gemm_detail,
        "step_tflops": round(world * B * FLOP_PER_VIDEO_TRAIN / (ms_step * 1e-3) / 1e12, 3), "loss": final_loss,
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        r = cpu_reference_run(4, 1, 16)
        line["cpu_baseline"] = {"value": round(r["value"], 2), "unit": "videos/s", "cores": r["cores"], "kind": "port",
                                "sample": "4 train steps of 16 videos (Opt().batch_size), same shapes, oracle/torch_port.py on the host cores"}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        wd = threading.Timer(60.0, lambda: os._exit(0))       # the line is out: whatever teardown does, leave within a minute
        wd.daemon = True
        wd.start()
        # Every rank has finished its work once it passes this barrier.  The process then leaves without tearing the NCCL communicator
        # down: destroy_process_group() with captured graphs that hold NCCL kernels still alive was observed to hang (2-GPU run), and
        # nothing remains to be flushed but the standard streams.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
