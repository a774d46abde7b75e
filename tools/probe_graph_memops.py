"""Probe: are cuStreamWaitValue32 / cuStreamWriteValue32 capturable into a CUDA graph on this driver, across forked streams?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2vt_b200
from s2vt_b200 import lib as L

lib = L.load()
dev = torch.device("cuda", 0)
ctr = torch.zeros(4, dtype=torch.int32, device=dev)
x = torch.zeros(1024, device=dev)
s1 = torch.cuda.Stream()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        ctr.zero_()
        ev = torch.cuda.Event(); ev.record(cur)
        with torch.cuda.stream(s1):
            s1.wait_event(ev)
            L.check(lib.s2vt_stream_wait_value32(s1.cuda_stream, L.ptr(ctr, 0), 1), "wait")
            x.add_(1.0)
            ev2 = torch.cuda.Event(); ev2.record(s1)
        y = x * 2
        L.check(lib.s2vt_stream_write_value32(cur.cuda_stream, L.ptr(ctr, 0), 1), "write")
        cur.wait_event(ev2)
        z = x + 0
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print("capture+replay ok: x[0] =", x[0].item(), "z[0] =", z[0].item(), "ctr =", ctr.tolist())
except Exception as e:
    print("FAILED:", type(e).__name__, e)
