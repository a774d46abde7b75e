#!/usr/bin/env python
"""Probe: does a small high-priority product get SM slots ahead of the queued CTAs of a big low-priority product?
Three launch modes: eager streams, torch graph replay (no node priorities), own executable graph with
cudaGraphInstantiateFlagUseNodePriority.  Prints the small product's start / end relative to the big one (%globaltimer stamps).

    python tools/probe_priority.py
"""
from __future__ import annotations

import collections
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import s2vt_b200  # noqa: E402
from s2vt_b200 import engine_bf16 as EB  # noqa: E402
from s2vt_b200 import lib as L  # noqa: E402
from s2vt_b200.lib import dense  # noqa: E402

BF = torch.bfloat16


def main():
    dev = torch.device("cuda", 0)
    lib = s2vt_b200.load()
    V, H, R = 13000, 512, 5056
    dl = torch.randn(R, V, device=dev).to(BF)
    out2 = torch.randn(R, H, device=dev).to(BF)
    gW = torch.empty(V, H, device=dev)
    dg = torch.randn(1280, 4 * H, device=dev).to(BF)
    W2 = torch.randn(4 * H, 2 * H, device=dev).to(BF)
    do = torch.empty(1280, H, device=dev)
    tiny = torch.randn(1024, device=dev)
    tiny2 = torch.empty(1024, dtype=BF, device=dev)
    stamps = torch.zeros(8, dtype=torch.int64, device=dev)
    hi = torch.cuda.Stream(device=dev, priority=-5)
    lo = torch.cuda.Stream(device=dev, priority=0)

    def stamp(i):
        lib.s2vt_timestamp(L.stream_ptr(dev), L.ptr(stamps, i))

    def body(kslice, urgent):
        cur = torch.cuda.current_stream(dev)
        ev0 = torch.cuda.Event()
        ev0.record(cur)
        with torch.cuda.stream(lo):
            lo.wait_event(ev0)
            stamp(0)
            EB.BULK_KSLICE = kslice
            EB.gemm(V, H, R, dl, V, True, out2, H, True, gW, dense(H), short_ctas=True, bulk=kslice > 0)
            stamp(1)
            e1 = torch.cuda.Event()
            e1.record(lo)
        with torch.cuda.stream(hi):
            hi.wait_event(ev0)
            for _ in range(12):                                   # ~30 us of tiny launches: the big product is under way by then
                lib.s2vt_cast_bf16(L.stream_ptr(dev), L.ptr(tiny), L.ptr(tiny2), None, 1, 1024)
            stamp(2)
            EB.gemm(1280, H, 4 * H, dg, 4 * H, False, W2, 2 * H, True, do, dense(H), b_off=H, short_ctas=True, urgent=urgent)
            stamp(3)
            e2 = torch.cuda.Event()
            e2.record(hi)
        cur.wait_event(e1)
        cur.wait_event(e2)

    def report(tag):
        torch.cuda.synchronize()
        t = stamps.cpu().tolist()
        print("%-46s big %6.1f us | small starts at %6.1f, takes %6.1f us" % (tag, (t[1] - t[0]) / 1e3, (t[2] - t[0]) / 1e3, (t[3] - t[2]) / 1e3))

    for kslice in (0, 16):
        for urgent in (False, True):
            for _ in range(3):
                body(kslice, urgent)
            report("eager   kslice=%d urgent=%d" % (kslice, urgent))
            for own in (False, True):
                g = torch.cuda.CUDAGraph(keep_graph=True)
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    body(kslice, urgent)
                if own:
                    pr = (ctypes.c_int * 256)()
                    n = ctypes.c_int(0)
                    rc = lib.s2vt_graph_kernel_priorities(g.raw_cuda_graph(), pr, 256, ctypes.byref(n))
                    if rc != 0:
                        print("priorities query failed:", L.last_error())
                    prio = dict(collections.Counter(pr[:n.value]))
                    ex = ctypes.c_void_p()
                    L.check(lib.s2vt_graph_instantiate(g.raw_cuda_graph(), 1, ctypes.byref(ex)), "instantiate")
                    for _ in range(3):
                        L.check(lib.s2vt_graph_launch(ex.value, L.stream_ptr(dev)), "launch")
                    report("graph+UseNodePriority kslice=%d urgent=%d %s" % (kslice, urgent, prio))
                    lib.s2vt_graph_exec_destroy(ex.value)
                else:
                    g.instantiate()
                    for _ in range(3):
                        g.replay()
                    report("graph (torch replay) kslice=%d urgent=%d" % (kslice, urgent))


if __name__ == "__main__":
    main()
