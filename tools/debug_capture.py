#!/usr/bin/env python
"""Debug aid: finds the call that invalidates a CUDA graph capture of the train step.

Every stream / event operation and every C-ABI call made while the step is being captured is followed by a
cudaStreamIsCapturing query on the capture stream; the first call after which the capture is no longer active is printed.

    python tools/debug_capture.py [--case small|bench]
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import s2vt_b200  # noqa: E402
from s2vt_b200 import lib as L  # noqa: E402
from s2vt_b200.dp import DataParallelTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="small")
    ap.add_argument("--interleave", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            pass
    if rt is None:
        import glob
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        rt = ctypes.CDLL(cands[0])
    state = {"stream": None, "bad": False, "log": []}

    def status():
        if state["stream"] is None:
            return -1
        st = ctypes.c_int(0)
        rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(state["stream"]), ctypes.byref(st))
        return st.value if rc == 0 else 100 + rc

    def check(what):
        if state["stream"] is None or state["bad"]:
            return
        s = status()
        state["log"].append((what, s))
        if s != 1:
            state["bad"] = True
            print("capture status %d right after: %s" % (s, what))
            print("last calls:", state["log"][-8:])
            traceback.print_stack(limit=12)

    def wrap_method(cls, name):
        orig = getattr(cls, name)

        def f(self, *a, **k):
            try:
                return orig(self, *a, **k)
            finally:
                check("%s.%s(%s)" % (cls.__name__, name, ", ".join(type(x).__name__ + ":" + hex(getattr(x, "cuda_stream", 0) or 0) for x in a)))
        setattr(cls, name, f)

    wrap_method(torch.cuda.Stream, "wait_event")
    wrap_method(torch.cuda.Stream, "wait_stream")
    wrap_method(torch.cuda.Stream, "record_event")
    wrap_method(torch.cuda.Event, "record")
    lib = L.load()
    for fname in list(L.SIGNATURES):
        fn = getattr(lib, fname, None)
        if fn is None or fname in ("s2vt_last_error",):
            continue

        def mk(fn, fname):
            def g(*a):
                r = fn(*a)
                check(fname)
                return r
            return g
        setattr(lib, fname, mk(fn, fname))

    if args.case == "small":
        V, F, Lq, H, E, B = 520, 64, 10, 128, 64, 24
    else:
        V, F, Lq, H, E, B = 1000, 256, 80, 512, 512, 64
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    models = []
    for _ in range(2):
        torch.manual_seed(11)
        models.append(s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev))
    opt_a = s2vt_b200.FusedAdam(models[0].parameters(), lr=1e-3)
    trainer = DataParallelTrainer(models[0], opt_a)
    opt_b = s2vt_b200.FusedAdam(models[1].parameters(), lr=1e-3)
    opt_b.attach(models[1])
    orig_capture = trainer._capture

    def cap(key, f, t, m):
        # torch.cuda.graph switches to its own capture stream: find it from inside via a hook on _step_eager
        orig_eager = trainer._step_eager

        def eager(*a, **k):
            state["stream"] = torch.cuda.current_stream(dev).cuda_stream
            print("capture stream", hex(state["stream"]), "status", status())
            try:
                return orig_eager(*a, **k)
            finally:
                print("end of captured step: status", status(), "calls checked", len(state["log"]))
                state["stream"] = None
        trainer._step_eager = eager
        try:
            return orig_capture(key, f, t, m)
        finally:
            trainer._step_eager = orig_eager
    trainer._capture = cap
    for i in range(4):
        try:
            la = trainer.step(feats, targets)
        except Exception as e:                                   # noqa: BLE001
            print("step %d raised: %s" % (i, str(e).splitlines()[0]))
            break
        if args.interleave:
            opt_b.zero_grad(set_to_none=True)
            lb = models[1].forward_loss(feats, targets)
            lb.backward()
            opt_b.step()
            print("step", i, float(la.item()), float(lb.item()))
        else:
            print("step", i, float(la.item()))
    print("graphs:", len(trainer._graphs))


if __name__ == "__main__":
    main()
