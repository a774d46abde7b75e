"""Debug tool (not a test): per-step clock64 breakdown of the cluster LSTM kernel, CTA 0, steps 16..47."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2vt_b200
from s2vt_b200 import lib as L

dev = torch.device("cuda:0")
lib = L.load()
T, B, H = 159, 64, 512
g = torch.Generator().manual_seed(1)
wb = ((torch.rand(4 * H, H, generator=g) * 2 - 1) / H ** 0.5).to(dev).bfloat16()
bias = torch.zeros(4 * H, device=dev)
pre = torch.randn(T, B, 4 * H, generator=g).to(dev)
out = torch.empty(T, B, H, device=dev, dtype=torch.bfloat16)
gates = torch.empty(T, B, 4 * H, device=dev, dtype=torch.bfloat16)
cells = torch.empty(T, B, H, device=dev)
raw = C.CDLL(L.LIB_PATH)
raw.s2vt_debug_trace_enable(1)
VARIANTS = {"W in TMEM, full stash": (L.ptr(gates), L.ptr(cells), 0), "W in TMEM, no stash": (None, None, 0),
            "W in smem, full stash": (L.ptr(gates), L.ptr(cells), 8)}
for variant, (gp, cp, flags) in VARIANTS.items():
    raw.s2vt_debug_set_flags(flags)
    for _ in range(3):
        rc = lib.s2vt_lstm_fwd_bf16(L.stream_ptr(dev), T, B, H, T, L.ptr(pre), L.ptr(bias), L.ptr(wb), None, None, L.ptr(out), gp, cp, None, None)
        L.check(rc, "lstm")
    buf = (C.c_longlong * 256)()
    raw.s2vt_debug_trace_read(buf, 256)
    import numpy as np
    a = np.array(list(buf), dtype=np.int64).reshape(32, 8)
    names = ["h_full wait done", "mma issued", "epi: mma_done seen", "phase1 done", "bar1 passed", "phase2 h packed", "st.async issued", "stash stores issued"]
    d = np.diff(a, axis=1)
    print("==", variant)
    print("step period (cycles):", np.diff(a[:, 0])[:12])
    for i in range(7):
        print("%-22s -> %-22s median %6.0f cycles" % (names[i], names[i + 1], np.median(d[4:28, i])))
    print("st.async issued -> next h_full wait done: median %6.0f cycles" % np.median(a[5:28, 0] - a[4:27, 6]))
