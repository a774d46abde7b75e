"""Parity of the tensor-core (bf16, tcgen05) kernels against fp32/fp64 PyTorch expressions of the same op.
Tolerances are stated per test; the inputs of each reference are the SAME bf16-rounded operands the kernel sees,
so the remaining error is accumulation order + the bf16 rounding of h between steps.  Run with -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import s2vt_b200  # noqa: E402
from s2vt_b200 import lib as L  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    s2vt_b200.load()
    return torch.device("cuda:0")


def lstm_ref(pre, bias, w_hh, T, n_pre, h0=None, c0=None, round_h=True):
    """fp64 LSTM over T steps from input-side pre-activations; h is rounded to bf16 before each recurrent product
    (as the kernel feeds it to the tensor cores) when round_h."""
    B = pre.shape[1] if pre is not None else h0.shape[0]
    H = w_hh.shape[1]
    W = w_hh.double()
    h = torch.zeros(B, H, dtype=torch.float64, device=w_hh.device) if h0 is None else h0.double()
    c = torch.zeros(B, H, dtype=torch.float64, device=w_hh.device) if c0 is None else c0.double()
    outs, gates, cells = [], [], []
    for t in range(T):
        hin = h.float().bfloat16().double() if round_h else h
        x = (pre[t].double() if t < n_pre else bias.double()[None]) + hin @ W.T
        i, f, g, o = x[:, :H].sigmoid(), x[:, H:2 * H].sigmoid(), x[:, 2 * H:3 * H].tanh(), x[:, 3 * H:].sigmoid()
        c = f * c + i * g
        h = o * c.tanh()
        outs.append(h); gates.append(torch.cat([i, f, g, o], 1)); cells.append(c)
    return torch.stack(outs), torch.stack(gates), torch.stack(cells), h, c


def run_lstm_bf16(dev, T, B, H, n_pre, with_state=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    k = 1.0 / H ** 0.5
    w = ((torch.rand(4 * H, H, generator=g) * 2 - 1) * k).to(dev)
    wb = w.bfloat16()
    bias = ((torch.rand(4 * H, generator=g) * 2 - 1) * k).to(dev)
    pre = (torch.randn(max(n_pre, 1), B, 4 * H, generator=g) * 0.7).to(dev)
    h0 = (torch.randn(B, H, generator=g) * 0.3).to(dev) if with_state else None
    c0 = (torch.randn(B, H, generator=g) * 0.5).to(dev) if with_state else None
    lib = L.load()
    Bp = int(lib.s2vt_lstm_bf16_batch_pad(B))
    out = torch.full((T, B, H), float("nan"), device=dev, dtype=torch.bfloat16)
    gates = torch.full((T * Bp * 4 * H,), float("nan"), device=dev, dtype=torch.bfloat16)
    cells = torch.full((T * Bp * H,), float("nan"), device=dev)
    hT = torch.empty(B, H, device=dev); cT = torch.empty(B, H, device=dev)
    rc = lib.s2vt_lstm_fwd_bf16(L.stream_ptr(dev), T, B, H, n_pre, L.ptr(pre), L.ptr(bias), L.ptr(wb), L.ptr(h0), L.ptr(c0),
                                L.ptr(out), L.ptr(gates), L.ptr(cells), L.ptr(hT), L.ptr(cT))
    L.check(rc, "s2vt_lstm_fwd_bf16")
    flag = lib.s2vt_device_error_flag(L.stream_ptr(dev))
    assert flag == 0, "device error flag %d" % flag
    # undo the kernel-private stash layout: gates [T][nbt][CS][16][32][4], cells [T][nbt][CS][16][32]
    nbt, CS = Bp // 16, H // 32
    gates = gates.view(T, nbt, CS, 16, 32, 4).permute(0, 1, 3, 5, 2, 4).reshape(T, Bp, 4 * H)[:, :B]
    cells = cells.view(T, nbt, CS, 16, 32).permute(0, 1, 3, 2, 4).reshape(T, Bp, H)[:, :B]
    ro, rg, rc_, rh, rcT = lstm_ref(pre, bias, wb.float(), T, n_pre, h0, c0)
    return (out, gates, cells, hT, cT), (ro, rg, rc_, rh, rcT)


@pytest.mark.parametrize("T,B,H,n_pre", [(1, 16, 64, 1), (3, 16, 128, 3), (7, 5, 256, 4), (12, 64, 512, 12), (159, 64, 512, 80), (9, 37, 512, 0)])
def test_lstm_fwd_bf16_cluster(dev, T, B, H, n_pre):
    (out, gates, cells, hT, cT), (ro, rg, rc_, rh, rcT) = run_lstm_bf16(dev, T, B, H, n_pre, seed=T + B + H)
    # h in (-1,1): bf16 output rounding 4e-3, fast-math activations ~1e-6, accumulated drift over T steps
    tol = 8e-3 + 2e-4 * T
    assert torch.isfinite(out.float()).all()
    assert (out.double() - ro).abs().max().item() < tol
    assert (gates.double() - rg).abs().max().item() < tol
    assert (cells.double() - rc_).abs().max().item() < 2 * tol
    assert (hT.double() - rh).abs().max().item() < tol
    assert (cT.double() - rcT).abs().max().item() < 2 * tol


def test_lstm_fwd_bf16_initial_state(dev):
    (out, gates, cells, hT, cT), (ro, rg, rc_, rh, rcT) = run_lstm_bf16(dev, 6, 20, 512, 6, with_state=True, seed=99)
    assert (out.double() - ro).abs().max().item() < 1e-2
    assert (cT.double() - rcT).abs().max().item() < 2e-2


def test_lstm_fwd_bf16_speed(dev):
    """Report us per timestep at the headline shape (B=64, H=512, T=159); asserts only a loose upper bound."""
    T, B, H = 159, 64, 512
    g = torch.Generator().manual_seed(1)
    wb = ((torch.rand(4 * H, H, generator=g) * 2 - 1) / H ** 0.5).to(dev).bfloat16()
    bias = torch.zeros(4 * H, device=dev)
    pre = torch.randn(T, B, 4 * H, generator=g).to(dev)
    out = torch.empty(T, B, H, device=dev, dtype=torch.bfloat16)
    gates = torch.empty(T, B, 4 * H, device=dev, dtype=torch.bfloat16)
    cells = torch.empty(T, B, H, device=dev)
    lib = L.load()
    assert int(lib.s2vt_lstm_bf16_batch_pad(B)) == B

    def run():
        rc = lib.s2vt_lstm_fwd_bf16(L.stream_ptr(dev), T, B, H, T, L.ptr(pre), L.ptr(bias), L.ptr(wb), None, None,
                                    L.ptr(out), L.ptr(gates), L.ptr(cells), None, None)
        L.check(rc, "s2vt_lstm_fwd_bf16")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us_step = e0.elapsed_time(e1) * 1e3 / 10 / T
    print("\nlstm_fwd_bf16: %.3f us per timestep (B=64, H=512, T=159)" % us_step)
    assert us_step < 20.0
