#!/usr/bin/env python
"""bench.py -- train videos/s of the S2VT hot path on N B200s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = the reference's train-loop body (train.py:116-127): zero_grad + forward(mode='train') + MaskCriterion +
backward + (gradient all-reduce) + Adam step on one batch of 64 synthetic MSVD-shaped videos per GPU
(80 x 4096 fp32 features, 28 real tokens padded to 80, V = 13000, H = E = 512, random-init weights).
Rank 0 prints ONE JSON line.  Besides the base contract's keys it carries
  decode        greedy / beam-5 captions/s of the tensor-core decode path on every rank's own videos (weak scaling), us per decode step
  api_path      the UNCHANGED reference loop body (model(..., 'train') -> MaskCriterion -> backward -> optimizer.step()) through the drop-in
  sustained     the timed loop again over >= 2000 steps (clocks settle at their sustained value)
  gpu_incumbent (N=1) the unmodified reference module on this GPU through stock PyTorch (cuDNN RNN + cuBLAS): train + greedy + beam
  dp_check      (N>1) max - min over ranks of a checksum of the weights after the timed loop (0 = replicas identical)
`--impl reference` times the reference's own CPU path on the host cores instead: the unmodified modules from oracle/_ref when that
directory travelled with the repository (kind "reference"), else the port in oracle/torch_port.py (kind "port").
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
def reference_bundle(device="cpu"):
    """(model, criterion, kind): the unmodified reference S2VT + MaskCriterion from oracle/_ref (kind 'reference'), else the torch
    port of the same library calls (kind 'port').  Test / baseline infrastructure: never on the product path."""
    torch.manual_seed(0)
    try:
        from oracle import build_ref
        ref = build_ref.import_reference()
    except Exception:
        ref = None
    if ref is not None:
        S2VT, MaskCriterion, _ = ref
        m = S2VT(CFG["V"], CFG["F"], CFG["L"], dim_hid=CFG["H"], dim_embed=CFG["E"], sos_ix=3, eos_ix=4).to(device)
        return m, MaskCriterion(), "reference"
    from oracle.torch_port import S2VTCpuPort
    m = S2VTCpuPort(CFG["V"], CFG["F"], CFG["L"], CFG["H"], CFG["E"]).to(device)

    class _Shim(torch.nn.Module):            # same call surface as the reference module
        def __init__(self, port):
            super().__init__()
            self.port = port

        def forward(self, feats, targets=None, mode="train", beam_width=3, max_beam_depth=30):
            if mode == "train":
                return self.port.train_logits(feats, targets)
            if mode == "test":
                return self.port.greedy(feats)
            raise NotImplementedError("the port has no beam search")
    return _Shim(m), (lambda lg, t, mk: S2VTCpuPort.criterion(lg, t, mk)), "port"


def reference_train_step(model, crit, opt, feats, targets, mask):
    """train.py:116-127, verbatim order: zero_grad, forward, criterion, backward, step, loss.item()"""
    opt.zero_grad()
    feats = feats.detach().requires_grad_(True)                                  # dataloader.py:38
    loss = crit(model(feats, targets=targets[:, :-1], mode="train"), targets, mask)
    loss.backward()
    opt.step()
    return loss.item()


def cpu_reference_run(steps, warmup, B, threads=None, decode=False):
    """Times the reference's own CPU path (config C1, batch 8) on the host cores."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    m, crit, kind = reference_bundle("cpu")
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    feats, targets, mask = synth_batch(B, 1234)
    for _ in range(warmup):
        reference_train_step(m, crit, opt, feats, targets, mask)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        reference_train_step(m, crit, opt, feats, targets, mask)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts))
    out = dict(value=B / (ms / 1e3), ms_per_step=ms, cores=threads, B=B, kind=kind)
    if decode:
        m.eval()
        with torch.no_grad():
            m(feats, mode="test")
            t0 = time.perf_counter()
            m(feats, mode="test")
            out["greedy_captions_s"] = B / (time.perf_counter() - t0)
            out["us_per_lstm_timestep"] = None
            if kind == "reference":
                t0 = time.perf_counter()
                m(feats[:1], mode="beam_search", beam_width=5, max_beam_depth=30)
                out["beam5_captions_s"] = 1.0 / (time.perf_counter() - t0)
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    B = 8                                    # config C1 (batch 8), the reference's own CPU-runnable case
    steps, warmup = min(args.steps, 8), min(max(1, args.warmup), 2)
    r = cpu_reference_run(steps, warmup, B, decode=True)
    sample = "%d train steps of %d videos (BASELINE config C1 shapes: 80x4096 feats, V=13000, H=E=512), 1 greedy pass of %d videos%s" % (
        steps, B, B, ", beam-5 on 1 video" if "beam5_captions_s" in r else "")
    line = {
        "impl": "reference", "metric": "train videos/sec", "value": r["value"], "unit": "videos/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "S2VT train step (train.py:116-127), MSVD shape 80x4096, V=13000, H=E=512, batch %d on the host CPU" % B,
                   "batch_per_step": B, "torch_threads": r["cores"]},
        "cpu_baseline": {"value": r["value"], "unit": "videos/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "decode": {"greedy_captions_s": r.get("greedy_captions_s"), "beam5_captions_s": r.get("beam5_captions_s")},
        "e2e": {"value": r["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_incumbent_run(dev, B):
    """SURVEY 2.1 / BASELINE.md section 3 "second incumbent": the same unmodified module on this B200 through stock PyTorch
    (cuDNN RNN + cuBLAS), at the benched batch.  Three numeric settings: torch defaults (fp32 matmul, cuDNN may use TF32),
    TF32 everywhere, and bf16 autocast (what a user would switch on to use the tensor cores)."""
    out = {"impl": None, "batch": B}
    feats, targets, mask = synth_batch(B, 1234, device=dev)
    m, crit, kind = reference_bundle(dev)
    out["impl"] = "%s module .cuda(): cuDNN %s RNN + cuBLAS, torch %s" % (kind, torch.backends.cudnn.version(), torch.__version__)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)

    def timed_train(n, autocast):
        def one():
            if autocast:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    opt.zero_grad()
                    f = feats.detach().requires_grad_(True)
                    loss = crit(m(f, targets=targets[:, :-1], mode="train").float(), targets, mask)
                loss.backward()
                opt.step()
                return loss.item()
            return reference_train_step(m, crit, opt, feats, targets, mask)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            one()
        e1.record()
        torch.cuda.synchronize()
        return B / (e0.elapsed_time(e1) / n / 1e3)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        out["train_videos_s_default"] = round(timed_train(10, False), 1)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        out["train_videos_s_tf32"] = round(timed_train(10, False), 1)
        out["train_videos_s_bf16_autocast"] = round(timed_train(10, True), 1)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    m.eval()
    with torch.no_grad():
        for nb in (B, 512):
            x = torch.randn(nb, CFG["L"], CFG["F"], device=dev)
            m(x, mode="test")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m(x, mode="test")
            torch.cuda.synchronize()
            out["greedy_captions_s_batch%d" % nb] = round(nb / (time.perf_counter() - t0), 1)
        if kind == "reference":
            t0 = time.perf_counter()
            m(feats[:1], mode="beam_search", beam_width=5, max_beam_depth=30)     # a Python loop over nodes with .item() syncs: ~1 min per video
            torch.cuda.synchronize()
            out["beam5_captions_s"] = round(1.0 / (time.perf_counter() - t0), 4)
    del m, opt
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------- decode (greedy / beam) on this rank's videos
def decode_run(s2vt_b200, dev, rank, world, sync_all, peaks):
    """BASELINE metric, second half: greedy and beam-5 captions/s.  Every rank decodes its own synthetic videos (weak scaling; no
    collective on the path): 1024 videos greedy in batches of 512, 460 videos beam-5 in batches of 230, features resident in HBM;
    `e2e` adds the H2D copy of the features from pinned host memory and the D2H read of the token ids."""
    torch.manual_seed(0)
    model = s2vt_b200.S2VT(CFG["V"], CFG["F"], CFG["L"], dim_hid=CFG["H"], dim_embed=CFG["E"]).to(dev).eval()
    GB, NG, BB, NB = 512, 1024, 230, 460      # beam: 230 videos x 5 beams = 1150 slots = 9 row tiles -> the step kernels are one wave (144 CTAs)
    g = torch.Generator().manual_seed(99 + rank)
    host = torch.randn(NG, CFG["L"], CFG["F"], generator=g).pin_memory()
    feats = host.to(dev)

    def greedy_all(src, h2d):
        outs = []
        for i in range(0, NG, GB):
            x = src[i:i + GB].to(dev, non_blocking=True) if h2d else src[i:i + GB]
            outs.append(model(x, mode="test"))
        return [o.cpu() for o in outs] if h2d else outs

    def beam_all(src, h2d):
        outs = []
        for i in range(0, NB, BB):
            x = src[i:i + BB].to(dev, non_blocking=True) if h2d else src[i:i + BB]
            outs.append(model.beam_search_ids(x, beam_width=5, max_beam_depth=30))
        return [(a.cpu(), b.cpu()) for a, b in outs] if h2d else outs

    def timed(fn, *a):
        fn(*a)
        sync_all()
        n0 = s2vt_b200.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*a)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), s2vt_b200.launch_count() - n0
    with torch.no_grad():
        ms_g, n_g = timed(greedy_all, feats, False)
        ms_b, n_b = timed(beam_all, feats, False)
        ms_ge, _ = timed(greedy_all, host, True)
        ms_be, _ = timed(beam_all, host, True)
    L_ = CFG["L"]
    try:
        dtraffic = json.load(open(os.path.join(ROOT, "profiles", "r02g_decode_dram_traffic.json")))
    except Exception:
        dtraffic = None
    gcap = world * NG / (ms_g / 1e3)
    # algorithmic work: SURVEY 8(d), 2.721 GFLOP per greedy caption; every fp32-grade product is three fp16 tensor-core passes
    tf = gcap / world * FLOP_PER_VIDEO_FWD / 1e12
    steps_serial = (2 * L_ - 1) + (L_ - 1)          # vid_rnn steps, then the decode steps (word_rnn's encode steps run beside vid_rnn)
    return {
        "greedy_captions_s": round(gcap, 1), "beam5_captions_s": round(world * NB / (ms_b / 1e3), 1),
        "greedy_e2e_captions_s": round(world * NG / (ms_ge / 1e3), 1), "beam5_e2e_captions_s": round(world * NB / (ms_be / 1e3), 1),
        "greedy_ms_per_batch": round(ms_g / (NG / GB), 3), "beam5_ms_per_batch": round(ms_b / (NB / BB), 3),
        "us_per_decode_step": round(1e3 * ms_g / (NG / GB) / steps_serial, 2),
        "videos_per_rank": {"greedy": NG, "beam": NB}, "batch": {"greedy": GB, "beam": BB}, "beam_width": 5, "max_beam_depth": 30,
        "gpu_launches": {"greedy": int(n_g), "beam": int(n_b)},
        "precision": "fp32-grade: fp32 operands as fp16 hi/lo planes, 3 tcgen05 passes, split fp32 TMEM accumulators; token ids "
                     "bit-identical to the reference on every golden (tests/test_gpu_model_parity.py)",
        "roofline": {"bound": "tensor", "kernel": "xgemm_kernel (all epilogues; whole greedy call)", "achieved": round(tf, 1),
                     "achieved_mma": round(3 * tf, 1), "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": round(tf / peaks["tf_sust"], 4),
                     "frac_mma": round(3 * tf / peaks["tf_sust"], 4),
                     "traffic": None if dtraffic is None else {k: v.get("dram_bytes_per_launch") for k, v in dtraffic.items() if isinstance(v, dict)},
                     "traffic_source": None if dtraffic is None else "profiles/r02g_decode_dram_traffic.json (ncu --set full, per launch)",
                     "note": "achieved = 2.721 GFLOP per caption (SURVEY 8d) / time; achieved_mma counts the three fp16 passes per product. "
                             "A chain of 238 dependent step kernels + 79 vocab products per batch: in-kernel phase times in profiles/r02_trace_*.txt"},
    }


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
    ap.add_argument("--quick", action="store_true", help="tuning runs: only the device-resident timed loop and the end-to-end loop")
    ap.add_argument("--model", default="s2vt", choices=["s2vt", "att"],
                    help="att: BASELINE configs[4], the attention_baseline.py encoder-decoder (tools/bench_att.py; train videos/s only)")
    args = ap.parse_args()
    if args.model == "att" and args.impl == "ours":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_att
        return bench_att.main(["--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--precision", args.precision] if args.precision in ("bf16", "fp32") else []))

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
        # NCCL logs to stdout by default and rank 0's stdout must carry ONE JSON line: send its log to stderr.  At level VERSION NCCL
        # ignores NCCL_DEBUG_FILE (the banner "NCCL version ..." goes to stdout regardless), so that level is raised to WARN; on top of
        # that, stdout is pointed at stderr while the communicator comes up.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
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
    for f_, t_, _ in batches:
        trainer.register_inputs(f_, t_)          # long-lived buffers: captured and replayed in place (no copy into the trainer's own)

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

    # ---- the same loop over >= 2000 steps: seconds of continuous load, clocks at their sustained value
    n_sus = max(2000, args.steps) if not args.quick else args.steps
    sync_all()
    t_sus0 = time.time()
    e0.record()
    for i in range(n_sus):
        step(i)
    e1.record()
    sync_all()
    t_sus1 = time.time()
    ts_ = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
    sustained = {"steps": n_sus, "ms_per_step": round(ts_.item() / n_sus, 4), "value": round(world * B / (ts_.item() / n_sus / 1e3), 2),
                 "unit": "videos/s", "seconds": round(ts_.item() / 1e3, 2)}

    # ---- data-parallel sanity: every rank must hold bit-identical weights after the same number of averaged updates
    dp_check = None
    if world > 1:
        cs = opt._flat["p"].double().sum().reshape(1)
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        vals = [float(x.item()) for x in allcs]
        dp_check = {"weight_checksum_max_minus_min": max(vals) - min(vals), "ranks": world, "steps_checked": int(opt._flat["step"])}
        # the gradient all-reduce averages rank gradients of per-rank MEAN losses: right iff mean_r loss(batch_r) == loss(all batches)
        f0, t0_, _ = batches[0]
        with torch.no_grad():
            l_here = model.forward_loss(f0, t0_).reshape(1).double()
        l_all = [torch.zeros_like(l_here) for _ in range(world)]
        dist.all_gather(l_all, l_here)
        fg = [torch.empty_like(f0) for _ in range(world)] if rank == 0 else None
        tg = [torch.empty_like(t0_) for _ in range(world)] if rank == 0 else None
        dist.gather(f0, fg, dst=0)
        dist.gather(t0_, tg, dst=0)
        if rank == 0:
            with torch.no_grad():
                l_cat = float(model.forward_loss(torch.cat(fg), torch.cat(tg)).item())
            l_mean = float(torch.stack(l_all).mean().item())
            dp_check.update({"loss_concat_batch": l_cat, "loss_mean_of_ranks": l_mean, "rel_diff": abs(l_cat - l_mean) / abs(l_cat)})
            del fg, tg

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs + D2H of the loss, every step.
    # Next step's inputs are prefetched on a copy stream while this step computes (each copy is inside the timed region).
    host = [synth_batch(B, 4321 + rank * 100 + i, pinned=True) for i in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.empty((), pin_memory=True)

    def run_e2e(host):
        dbuf = [(torch.empty_like(host[0][0], device=dev), torch.empty_like(host[0][1], device=dev)) for _ in range(2)]
        for f_, t_ in dbuf:
            trainer.register_inputs(f_, t_)

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

    e2e_f32_value, h2d_f32 = run_e2e(host)
    # The end-to-end headline feeds the loop from a pinned bf16 feature store -- data.DeviceFeatureStore(dtype=bfloat16) semantics: the
    # features are rounded ONCE at load time to exactly what the tensor-core path rounds them to at every step (identical loss and
    # gradients, tests/test_gpu_bf16.py::test_bf16_feature_store_batches_match_float32_batches), which halves the H2D bytes; the
    # float32-host-buffer figure is kept beside it as `e2e_f32_host`.
    e2e_value, h2d, e2e_src = e2e_f32_value, h2d_f32, "float32 host buffers"
    if precision == "bf16":
        trainer.max_graphs = max(trainer.max_graphs, n_rot + 4)
        host_bf = [(f.to(torch.bfloat16).pin_memory(), t, m) for f, t, m in host]
        e2e_value, h2d = run_e2e(host_bf)
        e2e_src = "bf16 feature store in pinned host memory (features rounded once at load time)"
    e2e_f32 = {"value": round(e2e_f32_value, 2), "unit": "videos/s", "h2d_bytes_per_step": int(h2d_f32), "d2h_bytes_per_step": 4}
    clocks = sampler.stop(t_wall0, time.time())                # the window spans the timed regions (device-resident and end-to-end)

    # ---- the UNCHANGED reference loop body through the drop-in (train.py:116-127): module forward -> MaskCriterion -> backward -> step
    crit = s2vt_b200.MaskCriterion()

    def api_step(i):
        f, t, m = batches[i % n_rot]
        opt.zero_grad()
        loss_ = crit(model(f, targets=t[:, :-1], mode="train"), t, m)
        loss_.backward()
        opt.step()
        return loss_
    for i in range(3):
        api_step(i)
    sync_all()
    n_api = min(args.steps, 20) if not args.quick else 2
    e0.record()
    for i in range(n_api):
        api_step(i)
    e1.record()
    sync_all()
    ta = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    api_path = {"value": round(world * B / (ta.item() / n_api / 1e3), 2), "unit": "videos/s", "ms_per_step": round(ta.item() / n_api, 4),
                "ratio_to_value": round(world * B / (ta.item() / n_api / 1e3) / value, 4),
                "what": "model(feats, targets[:, :-1], 'train') -> MaskCriterion -> backward -> FusedAdam.step(), eager, fp32 logits "
                        "[B,79,V] materialised as the API promises" + ("; no gradient all-reduce on this path (single-process loop body)" if world > 1 else "")}

    # ---- the same unchanged body, captured once into a CUDA graph by the caller (s2vt_b200.GraphedLoopBody) and replayed
    api_graphed = None
    if precision == "bf16" and not args.quick:
        def body(f, t, m):
            opt.zero_grad()
            loss_ = crit(model(f, targets=t[:, :-1], mode="train"), t, m)
            loss_.backward()
            opt.step()
            return loss_
        gstep = s2vt_b200.GraphedLoopBody(body, batches[0], optimizer=opt)
        for i in range(3):
            gstep(*batches[i % n_rot])
        sync_all()
        e0.record()
        for i in range(n_api):
            lg_ = gstep(*batches[i % n_rot])
        e1.record()
        sync_all()
        tg_ = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tg_, op=dist.ReduceOp.MAX)
        api_graphed = {"value": round(world * B / (tg_.item() / n_api / 1e3), 2), "unit": "videos/s", "ms_per_step": round(tg_.item() / n_api, 4),
                       "ratio_to_value": round(world * B / (tg_.item() / n_api / 1e3) / value, 4), "loss": float(lg_.item()),
                       "what": "the same body (zero_grad, module forward, MaskCriterion, backward, FusedAdam.step) captured once with "
                               "s2vt_b200.GraphedLoopBody and replayed; every step copies its batch (84 MB) into the captured buffers" +
                               ("; no gradient all-reduce on this path (single-process loop body)" if world > 1 else "")}
        del gstep

    # ---- decode: greedy / beam captions per second on this rank's videos
    decode = decode_run(s2vt_b200, dev, rank, world, sync_all, peaks) if not args.quick else None

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
                # headline for a latency chain: microseconds per time step against the north star's "< 5 us at batch 64"
                base.update({"bound": "latency", "tensor_tflops": base["achieved"], "tensor_frac": base["frac"],
                             "achieved": base["us_per_timestep"], "peak": 5.0, "unit": "us/timestep (lower is better)",
                             "frac": round(5.0 / max(1e-9, base["us_per_timestep"]), 4),
                             "peak_source": "north-star target: < 5 us per recurrent step at batch 64 (frac = target / achieved)"})
                base["note"] = ("serial recurrence (159 dependent steps); two sweeps run side by side as a wave front, so a launch's "
                                "duration includes the trailing sweep's wait for the leading one; tensor_tflops / tensor_frac for information")
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
        "e2e": {"value": round(e2e_value, 2), "unit": "videos/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "source": e2e_src},
        "e2e_f32_host": e2e_f32, "sustained": sustained, "api_path": api_path, "api_path_graphed": api_graphed, "decode": decode, "dp_check": dp_check,
        "roofline": roofline, "roofline_all": roofline_all, "kernels": kernels, "gemm_detail": gemm_detail,
        "step_tflops": round(world * B * FLOP_PER_VIDEO_TRAIN / (ms_step * 1e-3) / 1e12, 3), "loss": final_loss,
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline and not args.quick:
        line["gpu_incumbent"] = gpu_incumbent_run(dev, B)
        r = cpu_reference_run(4, 1, 8, decode=True)
        line["cpu_baseline"] = {"value": round(r["value"], 2), "unit": "videos/s", "cores": r["cores"], "kind": r["kind"],
                                "greedy_captions_s": round(r["greedy_captions_s"], 2),
                                "beam5_captions_s": round(r["beam5_captions_s"], 4) if "beam5_captions_s" in r else None,
                                "sample": "4 train steps + 1 greedy pass of 8 videos (BASELINE config C1) + beam-5 on 1 video, the %s on the "
                                          "host cores" % ("unmodified reference modules (oracle/_ref)" if r["kind"] == "reference" else "torch port")}
    else:
        line["cpu_baseline"] = None
        line["gpu_incumbent"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        wd = threading.Timer(60.0, lambda: os._exit(0))       # the line is out: whatever teardown does, leave within a minute
        wd.daemon = True
        wd.start()
        # Orderly teardown: the captured step graphs hold NCCL kernels, and destroying the communicator underneath them hung in
        # round 1 -- release them first (DataParallelTrainer.release_graphs), then destroy the process group.
        dist.barrier()
        torch.cuda.synchronize()
        trainer.release_graphs()
        sys.stdout.flush()
        sys.stderr.flush()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
