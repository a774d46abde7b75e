"""Parity of the tensor-core (bf16, tcgen05) kernels against fp32/fp64 PyTorch expressions of the same op.
Tolerances are stated per test; the inputs of each reference are the SAME bf16-rounded operands the kernel sees,
so the remaining error is accumulation order + the bf16 rounding of h between steps.  Run with -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import s2vt_b200  # noqa: E402
from s2vt_b200 import lib as L  # noqa: E402
from s2vt_b200 import engine_bf16 as EB  # noqa: E402


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


@pytest.mark.parametrize("T,B,H,n_pre,state", [(5, 64, 512, 5, False), (159, 64, 512, 80, False), (9, 37, 512, 0, False), (6, 48, 256, 6, True),
                                                 (4, 16, 512, 4, True)])
def test_lstm_fwd_bf16_two_tiles_per_cluster(dev, T, B, H, n_pre, state):
    """Two batch tiles sharing one cluster's resident weights (half the SMs per sweep) run the same arithmetic per column:
    every output must be bit-identical to the one-tile-per-cluster launch (ragged batches leave the second tile partly / fully idle)."""
    lib = L.load()
    one, _ = run_lstm_bf16(dev, T, B, H, n_pre, with_state=state, seed=7)
    lib.s2vt_lstm_bf16_set_tiles_per_cluster(2)
    try:
        two, _ = run_lstm_bf16(dev, T, B, H, n_pre, with_state=state, seed=7)
    finally:
        lib.s2vt_lstm_bf16_set_tiles_per_cluster(1)
    for a, b, name in zip(one, two, ("out", "gates", "cells", "hT", "cT")):
        assert torch.equal(a, b), name


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
    lib.s2vt_lstm_bf16_set_tiles_per_cluster(2)
    try:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
    finally:
        lib.s2vt_lstm_bf16_set_tiles_per_cluster(1)
    print("lstm_fwd_bf16, two tiles per cluster (32 SMs): %.3f us per timestep" % (e0.elapsed_time(e1) * 1e3 / 10 / T))


def lstm_bwd_ref(dout, gates, cells, w_hh, dout_t0):
    """fp64 BPTT from the stash (gates [T,B,4H] post-activation, cells [T,B,H]); dgates are rounded to bf16 before the
    recurrent product, as the kernel feeds them to the tensor cores."""
    T, B, H4 = gates.shape
    H = H4 // 4
    W = w_hh.double()
    dg = torch.zeros(T, B, H4, dtype=torch.float64, device=gates.device)
    dh_rec = torch.zeros(B, H, dtype=torch.float64, device=gates.device)
    dc = torch.zeros(B, H, dtype=torch.float64, device=gates.device)
    for t in range(T - 1, -1, -1):
        i, f, g, o = [gates[t, :, k * H:(k + 1) * H].double() for k in range(4)]
        c_t = cells[t].double()
        c_prev = cells[t - 1].double() if t > 0 else torch.zeros_like(c_t)
        dh = dh_rec + (dout[t].double() if t >= dout_t0 else 0.0)
        tc = c_t.tanh()
        dcv = dc + dh * o * (1 - tc * tc)
        dg[t] = torch.cat([dcv * g * i * (1 - i), dcv * c_prev * f * (1 - f), dcv * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
        dc = dcv * f
        dh_rec = dg[t].float().bfloat16().double() @ W
    return dg


@pytest.mark.parametrize("T,B,H,dout_t0", [(1, 16, 128, 0), (4, 16, 128, 0), (6, 21, 256, 2), (12, 64, 512, 0), (159, 64, 512, 80)])
def test_lstm_bwd_bf16_cluster(dev, T, B, H, dout_t0):
    g = torch.Generator().manual_seed(T * 3 + B + H)
    lib = L.load()
    Bp = int(lib.s2vt_lstm_bf16_batch_pad(B))
    nbt, CS = Bp // 16, H // 32
    k = 1.0 / H ** 0.5
    w = ((torch.rand(4 * H, H, generator=g) * 2 - 1) * k).to(dev).bfloat16()
    wt = w.T.contiguous()
    # a plausible stash: gates in (0,1)/(-1,1), cells ~ N(0,1); the kernel only needs them to be self-consistent inputs
    gates = torch.rand(T, Bp, 4 * H, generator=g).to(dev)
    gates[:, :, 2 * H:3 * H] = gates[:, :, 2 * H:3 * H] * 2 - 1
    gates = gates.bfloat16()
    cells = torch.randn(T, Bp, H, generator=g).to(dev)
    dout = (torch.randn(T, B, H, generator=g) * 0.1).to(dev)
    # private layouts: gates [T][nbt][CS][16][32][4], cells [T][nbt][CS][16][32]
    gates_p = gates.view(T, nbt, 16, 4, CS, 32).permute(0, 1, 4, 2, 5, 3).contiguous()
    cells_p = cells.view(T, nbt, 16, CS, 32).permute(0, 1, 3, 2, 4).contiguous()
    dg = torch.full((T, B, 4 * H), float("nan"), device=dev, dtype=torch.bfloat16)
    rc = lib.s2vt_lstm_bwd_bf16(L.stream_ptr(dev), T, B, H, dout_t0, L.ptr(dout), L.ptr(gates_p), L.ptr(cells_p), L.ptr(wt), L.ptr(dg))
    L.check(rc, "s2vt_lstm_bwd_bf16")
    flag = lib.s2vt_device_error_flag(L.stream_ptr(dev))
    assert flag == 0, "device error flag %d" % flag
    ref = lstm_bwd_ref(dout, gates[:, :B].float(), cells[:, :B], w.float(), dout_t0)
    assert torch.isfinite(dg.float()).all()
    scale = ref.abs().max().item()
    err = (dg.double() - ref).abs().max().item()
    assert err < (1e-2 + 1e-3 * T) * scale, (err, scale)


def _bwd_inputs(dev, T, B, H, seed):
    g = torch.Generator().manual_seed(seed)
    lib = L.load()
    Bp = int(lib.s2vt_lstm_bf16_batch_pad(B))
    nbt, CS = Bp // 16, H // 32
    w = ((torch.rand(4 * H, H, generator=g) * 2 - 1) / H ** 0.5).to(dev).bfloat16()
    gates = torch.rand(T, Bp, 4 * H, generator=g).to(dev)
    gates[:, :, 2 * H:3 * H] = gates[:, :, 2 * H:3 * H] * 2 - 1
    gates = gates.bfloat16()
    cells = torch.randn(T, Bp, H, generator=g).to(dev)
    dout = (torch.randn(T, B, H, generator=g) * 0.1).to(dev)
    gates_p = gates.view(T, nbt, 16, 4, CS, 32).permute(0, 1, 4, 2, 5, 3).contiguous()
    cells_p = cells.view(T, nbt, 16, CS, 32).permute(0, 1, 3, 2, 4).contiguous()
    return w.T.contiguous(), gates_p, cells_p, dout, Bp


@pytest.mark.parametrize("T,B,H,dout_t0,cuts,ntl", [(12, 64, 512, 0, (5,), 1), (12, 64, 512, 0, (), 2), (159, 64, 512, 80, (40, 80, 120), 1),
                                                     (159, 64, 512, 80, (40, 80, 120), 2), (9, 37, 256, 3, (1, 8), 2), (7, 16, 128, 0, (3,), 2),
                                                     (6, 48, 512, 0, (2, 4), 2)])
def test_lstm_bwd_bf16_chunks_and_tiles(dev, T, B, H, dout_t0, cuts, ntl):
    """A sweep cut into time chunks (latest first, chained through dh / dc) and / or run with two batch tiles per cluster gives the
    whole-sweep result: identical bits for the tile variant, fp32-reassociation noise (the carried dh is summed in another order)
    for the chunked one."""
    lib = L.load()
    wt, gates_p, cells_p, dout, Bp = _bwd_inputs(dev, T, B, H, seed=T + B + H)
    whole = torch.full((T, B, 4 * H), float("nan"), device=dev, dtype=torch.bfloat16)
    L.check(lib.s2vt_lstm_bwd_bf16(L.stream_ptr(dev), T, B, H, dout_t0, L.ptr(dout), L.ptr(gates_p), L.ptr(cells_p), L.ptr(wt), L.ptr(whole)), "bwd")
    got = torch.full((T, B, 4 * H), float("nan"), device=dev, dtype=torch.bfloat16)
    bounds = [0] + list(cuts) + [T]
    dh, dc = None, None
    for k in range(len(bounds) - 2, -1, -1):
        t0, t1 = bounds[k], bounds[k + 1]
        dh_o = torch.full((B, H), float("nan"), device=dev) if t0 > 0 else None
        dc_o = torch.full((B, H), float("nan"), device=dev) if t0 > 0 else None
        EB.lstm_bwd(t1 - t0, B, H, max(0, dout_t0 - t0), dout, gates_p, cells_p, wt, got, dout_off=t0 * B * H, gates_off=t0 * Bp * 4 * H,
                    cells_off=t0 * Bp * H, dgates_off=t0 * B * 4 * H, dh_in=dh, dc_in=dc, dh_out=dh_o, dc_out=dc_o, has_prev=t0 > 0,
                    tiles_per_cluster=ntl)
        dh, dc = dh_o, dc_o
    assert lib.s2vt_device_error_flag(L.stream_ptr(dev)) == 0
    assert torch.isfinite(got.float()).all()
    if not cuts:
        assert torch.equal(got, whole)
    else:
        scale = whole.float().abs().max().item()
        assert (got.float() - whole.float()).abs().max().item() <= 2e-2 * scale
        assert (got.float() - whole.float()).abs().mean().item() <= 1e-3 * scale


def test_lstm_bwd_bf16_speed(dev):
    T, B, H = 159, 64, 512
    g = torch.Generator().manual_seed(2)
    lib = L.load()
    wt = ((torch.rand(H, 4 * H, generator=g) * 2 - 1) / H ** 0.5).to(dev).bfloat16()
    gates = torch.rand(T * B * 4 * H, generator=g).to(dev).bfloat16()
    cells = torch.randn(T * B * H, generator=g).to(dev)
    dout = (torch.randn(T, B, H, generator=g) * 0.1).to(dev)
    dg = torch.empty(T, B, 4 * H, device=dev, dtype=torch.bfloat16)

    def run():
        L.check(lib.s2vt_lstm_bwd_bf16(L.stream_ptr(dev), T, B, H, 0, L.ptr(dout), L.ptr(gates), L.ptr(cells), L.ptr(wt), L.ptr(dg)), "bwd")
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
    print("\nlstm_bwd_bf16: %.3f us per timestep (B=64, H=512, T=159)" % us_step)
    assert us_step < 20.0

    def run2():
        L.check(lib.s2vt_lstm_bwd_bf16_chunk(L.stream_ptr(dev), T, B, H, 0, L.ptr(dout), L.ptr(gates), L.ptr(cells), L.ptr(wt), L.ptr(dg), 0,
                                             None, None, None, None, 0, 2), "bwd")
    for _ in range(3):
        run2()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run2()
    e1.record()
    torch.cuda.synchronize()
    print("lstm_bwd_bf16, two tiles per cluster (32 SMs): %.3f us per timestep" % (e0.elapsed_time(e1) * 1e3 / 10 / T))


def test_wavefront_matches_sequential_sweeps(dev):
    """The wave front (two single-launch sweeps coupled through device counters, chunked input products between them) computes what
    the sequential order computes: forward quantities are bit-identical (same kernels, same operands); gradients below the coupling
    product differ only by the summation order of its K-split."""
    torch.manual_seed(5)
    V, F, H, E, Lq, B = 1000, 256, 512, 512, 80, 64
    model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    feats = torch.randn(B, Lq, F, device=dev)
    targets = torch.randint(0, V, (B, Lq), device=dev)
    assert EB.wave_ok(B, Lq)
    res = {}
    old = EB.WAVEFRONT
    try:
        for mode in (True, False, True):
            EB.WAVEFRONT = mode
            model.zero_grad(set_to_none=True)
            loss = model.forward_loss(feats, targets)
            loss.backward()
            torch.cuda.synchronize()
            assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
            res[mode] = (loss.item(), {k: p.grad.clone() for k, p in model.named_parameters()})
    finally:
        EB.WAVEFRONT = old
    assert abs(res[True][0] - res[False][0]) <= 1e-6 * abs(res[False][0])
    for k in res[True][1]:
        a, b = res[True][1][k].double(), res[False][1][k].double()
        # not bit-identical: word_rnn's pre-activations are summed in another order (embedding half first), and bf16 roundings of
        # activations / dgates flip on 1e-7 differences
        assert (a - b).norm().item() <= 1e-2 * b.norm().item(), k


def test_cuda_graph_step_matches_eager(dev):
    """DataParallelTrainer replays a captured CUDA graph of the whole step from its third step on: weights after 5 steps must equal
    those of an all-eager trainer (same kernels in the same order; Adam's step count and lr live on the device), and an lr change
    between replays must take effect."""
    from s2vt_b200.dp import DataParallelTrainer
    V, F, H, E, Lq, B = 1000, 256, 512, 512, 80, 64
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    out = {}
    for graph in (False, True):
        torch.manual_seed(3)
        model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
        opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-3)
        tr = DataParallelTrainer(model, opt, cuda_graph=graph)
        losses = []
        for i in range(6):
            if i == 4:
                opt.param_groups[0]["lr"] = 5e-4
            losses.append(float(tr.step(feats, targets).item()))
        torch.cuda.synchronize()
        assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
        assert bool(tr._graphs) == graph
        out[graph] = (losses, {k: p.detach().clone() for k, p in model.named_parameters()}, opt._flat["step"], int(opt._flat["step_dev"].item()))
    assert out[True][2] == out[False][2] == 6 and out[True][3] == out[False][3] == 6
    for a, b in zip(out[True][0], out[False][0]):
        assert abs(a - b) <= 1e-4 * abs(b), (out[True][0], out[False][0])
    assert out[False][0][-1] < out[False][0][0]
    for k in out[True][1]:
        a, b = out[True][1][k].double(), out[False][1][k].double()
        assert (a - b).norm().item() <= 1e-3 * b.norm().item(), k


# ------------------------------------------------------------------ whole train step on tensor cores vs the reference goldens
def _bf16_model(name, dev):
    from conftest import golden_inputs, load_golden
    g = load_golden(name)
    P, feats, targets, mask, c = golden_inputs(g)
    m = s2vt_b200.S2VT(c["V"], c["F"], c["L"], dim_hid=c["H"], dim_embed=c["E"], train_precision="bf16")
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()}, strict=True)
    return g, m.to(dev), torch.from_numpy(feats).to(dev), torch.from_numpy(targets).to(dev), torch.from_numpy(mask).to(dev), c


def _report(tag, got, ref):
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    return "%s: max|err| %.3e  (|ref|max %.3e, rel-to-max %.3e)" % (tag, err.max(), np.abs(ref).max(), err.max() / max(1e-30, np.abs(ref).max()))


def test_bf16_train_step_vs_reference_golden(dev):
    """Tolerances of the bf16 mode (BASELINE north star: loss / logits within rtol 1e-3 at bf16): loss rtol 1e-3;
    logits |err| <= 1e-3 * max(1, |logit|max) + 1e-2 * sigma_logit; gradients: relative L2 error <= 5e-2 per tensor,
    norms within 2e-2."""
    g, model, tf, tt, tm, c = _bf16_model("msvd", dev)
    tf = tf.requires_grad_(True)
    logits = model(tf, targets=tt[:, :-1], mode="train")
    loss = s2vt_b200.MaskCriterion()(logits, tt, tm)
    loss.backward()
    assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
    ls = logits.detach().cpu().numpy().reshape(-1)[::997]
    ref = g["logits_sample"]
    print("\n" + _report("logits", ls, ref), " sigma_ref %.3e" % ref.std())
    print("loss %.6f vs %.6f (rel %.2e)" % (loss.item(), float(g["loss"]), abs(loss.item() - float(g["loss"])) / float(g["loss"])))
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    assert np.abs(ls - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()) + 1e-2 * ref.std()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    grads["feats"] = tf.grad.cpu().numpy()
    for k, gv in grads.items():
        if "grad_full/" + k in g:
            rs, gs = g["grad_full/" + k].reshape(-1), gv.reshape(-1)
        else:
            rs, gs = g["grad_sample/" + k], gv.reshape(-1)[::997]
        rel = np.linalg.norm(gs.astype(np.float64) - rs) / max(1e-30, np.linalg.norm(rs.astype(np.float64)))
        n = np.linalg.norm(gv.astype(np.float64))
        nrel = abs(n - float(g["grad_norm/" + k])) / float(g["grad_norm/" + k])
        print("grad %-24s rel-L2 err (sampled) %.3e   norm rel err %.3e" % (k, rel, nrel))
        assert rel <= 1e-2, k                 # observed 4e-3 .. 6e-3
        assert nrel <= 1e-2, k
    # The error floor of ANY design whose products take bf16 operands, measured on the CPU by rounding operands only
    # (tools/bf16_error_budget.py, profiles/r02_bf16_error_budget.txt): weights + inputs 1.71e-3, + h_t 1.88e-3, + other activations
    # 1.92e-3 (= 0.016 sigma_z).  SURVEY 8(c)'s "1e-3 sigma" (1.2e-4) is below that floor; what the kernels may add on top
    # (tanh.approx, ex2.approx, summation order) is held to 20 % of it.
    assert np.abs(ls - ref).max() <= 1.2 * 1.915e-3


def test_bf16_fused_loss_matches_api_path(dev):
    g, model, tf, tt, tm, c = _bf16_model("msvd", dev)
    loss = model.forward_loss(tf, tt, tm)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    for k, p in model.named_parameters():
        n = np.linalg.norm(p.grad.cpu().numpy().astype(np.float64))
        assert abs(n - float(g["grad_norm/" + k])) <= 2e-2 * float(g["grad_norm/" + k]), k


def test_bf16_training_reduces_loss(dev):
    """A few fused-Adam steps on one batch: the loss must fall monotonically (end-to-end sanity of fwd+bwd+update)."""
    g, model, tf, tt, tm, c = _bf16_model("msvd", dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-3)
    opt.attach(model)
    losses = []
    for _ in range(5):
        opt.zero_grad(set_to_none=True)
        loss = model.forward_loss(tf, tt, tm)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("\nlosses", losses)
    assert all(b < a for a, b in zip(losses, losses[1:]))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_bucketwise_adam_matches_whole_buffer_step(dev, precision):
    """DataParallelTrainer updates each bucket right behind its gradients (beside the BPTT sweeps); three steps of it must
    leave the same weights as forward_loss / backward / FusedAdam.step() on the whole flat buffer."""
    from s2vt_b200.dp import DataParallelTrainer
    V, F, Lq, H, E, B = 520, 64, 10, 128, 64, 24
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    models = []
    for _ in range(2):
        torch.manual_seed(11)
        models.append(s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision=precision).to(dev))
    opt_a = s2vt_b200.FusedAdam(models[0].parameters(), lr=1e-3)
    trainer = DataParallelTrainer(models[0], opt_a)
    opt_b = s2vt_b200.FusedAdam(models[1].parameters(), lr=1e-3)
    opt_b.attach(models[1])
    for _ in range(3):
        la = trainer.step(feats, targets)
        opt_b.zero_grad(set_to_none=True)
        lb = models[1].forward_loss(feats, targets)
        lb.backward()
        opt_b.step()
        assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())
    for (k, a), (_, b) in zip(models[0].state_dict().items(), models[1].state_dict().items()):
        # identical kernels; only the atomic accumulation order of split-K / scatter-add partial sums may differ
        assert (a - b).abs().max().item() <= 2e-5 * max(1e-3, b.abs().max().item()), k


def test_bf16_feature_store_batches_match_float32_batches(dev):
    """Features handed over as bfloat16 (data.DeviceFeatureStore(dtype=bfloat16): rounded once at load time) give the loss and gradients
    of the float32 features: the tensor-core path's first act on float32 features is that very rounding.  Decode and requires_grad
    inputs still insist on float32."""
    V, F, H, E, Lq, B = 520, 64, 128, 64, 10, 24
    g = torch.Generator().manual_seed(23)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    torch.manual_seed(4)
    model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    out = []
    for x in (feats, feats.to(torch.bfloat16)):
        model.zero_grad(set_to_none=True)
        loss = model.forward_loss(x, targets)
        loss.backward()
        out.append((loss.item(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[0][0])
    for k in out[0][1]:
        a, b = out[1][1][k].double(), out[0][1][k].double()
        assert (a - b).norm().item() <= 1e-5 * max(1e-30, b.norm().item()), k     # (split-K / atomic summation order only)
    with pytest.raises(ValueError):
        model(feats.to(torch.bfloat16), mode="test")
    with pytest.raises(ValueError):
        model.forward_loss(feats.to(torch.bfloat16).requires_grad_(True), targets)


# ------------------------------------------------------------------ the benched configuration (BASELINE configs[1]) vs the reference
def test_c2_batch64_trajectory_through_trainer_with_graph_replay(dev):
    """tests/golden/c2.npz: the unmodified reference at B=64 (loss, sampled gradients, and a 5-step Adam trajectory over five
    different batches).  Here the same five batches go through DataParallelTrainer.step() as fresh tensors, i.e. through the
    trainer's static input buffers and, from the third step on, CUDA-graph replay -- the path bench.py times."""
    from conftest import golden_cfg, load_golden
    from oracle import s2vt_numpy as O
    from s2vt_b200.dp import DataParallelTrainer
    g = load_golden("c2")
    c = golden_cfg(g)
    P = O.synth_params(c["V"], c["F"], c["H"], c["E"], seed=c["wseed"], out_scale=c["out_scale"], eos_bias=c["eos_bias"])
    model = s2vt_b200.S2VT(c["V"], c["F"], c["L"], dim_hid=c["H"], dim_embed=c["E"], train_precision="bf16")
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()}, strict=True)
    model = model.to(dev)
    # (1) loss and gradients of the first batch, fused path
    f0, t0, m0 = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"], real_tokens=c["real"])
    loss = model.forward_loss(torch.from_numpy(f0).to(dev), torch.from_numpy(t0).to(dev))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    for k, p in model.named_parameters():
        gv = p.grad.detach().cpu().numpy()
        rs = g["grad_full/" + k].reshape(-1) if "grad_full/" + k in g else g["grad_sample/" + k]
        gs = gv.reshape(-1) if "grad_full/" + k in g else gv.reshape(-1)[::997]
        rel = np.linalg.norm(gs.astype(np.float64) - rs) / max(1e-30, np.linalg.norm(rs.astype(np.float64)))
        assert rel <= 1e-2, (k, rel)
        p.grad = None
    del loss        # (its autograd graph pins the parameters' AccumulateGrad nodes to this stream; a later graph capture must not meet them)
    # (2) five training steps, fresh tensors every step
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-4)
    tr = DataParallelTrainer(model, opt, cuda_graph=True)
    losses = []
    for i in range(int(c["adam_steps"])):
        f_i, t_i, _ = O.synth_batch(c["B"], c["L"], c["F"], c["V"], seed=c["dseed"] + i, real_tokens=c["real"])
        losses.append(float(tr.step(torch.from_numpy(f_i).to(dev), torch.from_numpy(t_i).to(dev)).item()))
    torch.cuda.synchronize()
    assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
    assert tr.replays >= 3 and len(tr._graphs) == 1
    print("\nc2 trajectory loss: ours", ["%.5f" % x for x in losses], " reference", ["%.5f" % x for x in g["traj_loss"]])
    for a, b in zip(losses, g["traj_loss"]):
        assert abs(a - b) <= 1e-3 * abs(b), (losses, g["traj_loss"])
    worst = 0.0
    for k, p in model.named_parameters():
        pv = p.detach().cpu().numpy()
        d_ours = pv.reshape(-1)[::997].astype(np.float64) - P[k].reshape(-1)[::997]
        d_ref = g["traj_param_sample/" + k].astype(np.float64) - P[k].reshape(-1)[::997]
        rel = np.linalg.norm(d_ours - d_ref) / max(1e-30, np.linalg.norm(d_ref))
        nrm = np.linalg.norm((pv - P[k]).astype(np.float64))
        nrel = abs(nrm - float(g["traj_delta_norm/" + k])) / float(g["traj_delta_norm/" + k])
        print("5-step update %-24s rel-L2 err of the sampled update %.3e   update-norm rel err %.3e" % (k, rel, nrel))
        worst = max(worst, rel)
        # Adam's first steps move every weight by ~lr * sign-like(g): an element whose tiny gradient changes sign under bf16 rounding
        # moves the other way, so the update agrees in norm much more tightly than element by element
        assert rel <= 0.25, (k, rel)
        assert nrel <= 2e-2, (k, nrel)


def test_fresh_batches_replay_one_graph_per_shape(dev):
    """What fit() does: a loader yields freshly allocated batches (two shapes: full and tail).  The trainer must replay, not
    re-capture per pointer (round-1 defect: graphs were keyed on data_ptr and never replayed under fit())."""
    from s2vt_b200.dp import DataParallelTrainer
    V, F, H, E, Lq = 136, 64, 128, 64, 6
    torch.manual_seed(5)
    model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-3)
    tr = DataParallelTrainer(model, opt, cuda_graph=True)
    ref_model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    ref_model.load_state_dict(model.state_dict())
    ref_opt = s2vt_b200.FusedAdam(ref_model.parameters(), lr=1e-3)
    ref_tr = DataParallelTrainer(ref_model, ref_opt, cuda_graph=False)
    g = torch.Generator().manual_seed(9)
    data = []
    for i in range(50):
        B = 5 if i % 10 == 9 else 8                       # a tail batch now and then
        data.append((torch.randn(B, Lq, F, generator=g), torch.randint(0, V, (B, Lq), generator=g)))
    # (one trainer after the other: the weight epoch that guards the bf16 shadows is process-wide, so interleaving two optimizers
    # would -- correctly, but uselessly for this test -- force every step down the eager path)
    ref_losses = [float(ref_tr.step(f.to(dev), t.to(dev)).item()) for f, t in data]
    losses = [float(tr.step(f.to(dev), t.to(dev)).item()) for f, t in data]          # fresh device tensors every step
    for i, (a, b) in enumerate(zip(losses, ref_losses)):
        assert abs(a - b) <= 2e-3 * abs(b), (i, a, b)
    assert len(tr._graphs) <= 2 and tr.replays >= 45, (len(tr._graphs), tr.replays)
    tr.check_device_errors()
    # zero-copy route: a loader that fills the trainer's own buffers
    fb, tb = tr.input_buffers((8, Lq, F), (8, Lq))
    fb.normal_(); tb.random_(0, V)
    n = tr.replays
    tr.step(fb, tb)
    assert tr.replays == n + 1 and len(tr._graphs) <= 2


def test_frozen_parameter_still_trains_the_others(dev):
    """with one tensor frozen backward falls back to ordinary autograd gradients; the trainer must bring them into
    the flat buffer instead of stepping on stale zeros."""
    from s2vt_b200.dp import DataParallelTrainer
    V, F, H, E, Lq, B = 136, 64, 128, 64, 6, 4
    torch.manual_seed(6)
    model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    model.embedding.weight.requires_grad_(False)
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-3)
    tr = DataParallelTrainer(model, opt, cuda_graph=False)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    l0 = float(tr.step(feats, targets).item())
    for _ in range(5):
        l1 = float(tr.step(feats, targets).item())
    assert l1 < l0
    assert torch.equal(model.embedding.weight, before["embedding.weight"])
    assert not torch.equal(model.out_linear.weight, before["out_linear.weight"])


def test_fused_adam_state_dict_roundtrip(dev):
    V, F, H, E, Lq, B = 136, 64, 128, 64, 6, 4
    torch.manual_seed(7)
    model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E).to(dev)
    opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    for _ in range(3):
        opt.zero_grad()
        model.forward_loss(feats, targets).backward()
        opt.step()
    sd = opt.state_dict()
    assert sd["fused"]["step"] == 3 and float(sd["fused"]["v"].abs().sum()) > 0
    model2 = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E).to(dev)
    model2.load_state_dict(model.state_dict())
    opt2 = s2vt_b200.FusedAdam(model2.parameters(), lr=1e-3)
    opt2.load_state_dict(sd)
    for o, m in ((opt, model), (opt2, model2)):
        o.zero_grad()
        m.forward_loss(feats, targets).backward()
        o.step()
    for (k, a), (_, b) in zip(model.named_parameters(), model2.named_parameters()):
        assert torch.allclose(a, b, rtol=0, atol=1e-7), k
    with pytest.raises(ValueError):
        s2vt_b200.FusedAdam([{"params": [model.embedding.weight]}, {"params": [model.out_linear.weight], "lr": 1e-2}])


@pytest.mark.parametrize("V", [1001, 203])
def test_vocab_not_a_multiple_of_8_stays_on_tensor_cores(dev, V):
    """A real vocabulary (prepare_captions.py:9-24) is not a multiple of 8: the bf16 path must take it (round-1 cliff: V % 8 != 0 fell
    back to the CUDA-core path).  Loss and gradients against the exact fp32 path of the same module, both API routes."""
    F, H, E, Lq, B = 64, 128, 64, 6, 5
    torch.manual_seed(8)
    mb = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    mf = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="fp32").to(dev)
    mf.load_state_dict(mb.state_dict())
    assert mb._use_bf16()
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    mask = torch.ones(B, Lq, device=dev)
    lf = mf.forward_loss(feats, targets)
    lf.backward()
    ref = {k: p.grad.clone() for k, p in mf.named_parameters()}
    for route in ("fused", "api"):
        for p in mb.parameters():
            p.grad = None
        if route == "fused":
            lb = mb.forward_loss(feats, targets)
        else:
            lb = s2vt_b200.MaskCriterion()(mb(feats, targets=targets[:, :-1], mode="train"), targets, mask)
        lb.backward()
        assert abs(lb.item() - lf.item()) <= 2e-3 * abs(lf.item()), (route, lb.item(), lf.item())
        for k, p in mb.named_parameters():
            rel = (p.grad - ref[k]).norm().item() / max(1e-30, ref[k].norm().item())
            assert rel <= 3e-2, (route, k, rel)
    assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0


# ------------------------------------------------------------------ shapes outside the cluster kernels' range (engine_step)
@pytest.mark.parametrize("dims", [(210, 64, 72, 20, 6, 9), (300, 48, 200, 100, 7, 130)])
def test_step_engine_matches_exact_path(dev, dims):
    """H not a multiple of 128 (and E, V not multiples of 8): the per-step tensor-core recurrence + BPTT against the exact fp32 path of
    the same module, fused and API routes; then five optimizer steps through DataParallelTrainer with graph replay."""
    from s2vt_b200 import engine_step as ES
    from s2vt_b200.dp import DataParallelTrainer
    V, F, H, E, Lq, B = dims
    torch.manual_seed(12)
    mb = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
    mf = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="fp32").to(dev)
    mf.load_state_dict(mb.state_dict())
    assert mb._bf16_engine() is ES
    g = torch.Generator().manual_seed(4)
    feats = torch.randn(B, Lq, F, generator=g).to(dev)
    targets = torch.randint(0, V, (B, Lq), generator=g).to(dev)
    mask = torch.ones(B, Lq, device=dev)
    ff = feats.clone().requires_grad_(True)
    lf = mf.forward_loss(ff, targets)
    lf.backward()
    ref = {k: p.grad.clone() for k, p in mf.named_parameters()}
    ref["feats"] = ff.grad.clone()
    for route in ("fused", "api"):
        for p in mb.parameters():
            p.grad = None
        fb = feats.clone().requires_grad_(True)
        if route == "fused":
            lb = mb.forward_loss(fb, targets)
        else:
            lb = s2vt_b200.MaskCriterion()(mb(fb, targets=targets[:, :-1], mode="train"), targets, mask)
        lb.backward()
        assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
        assert abs(lb.item() - lf.item()) <= 2e-3 * abs(lf.item()), (route, lb.item(), lf.item())
        got = {k: p.grad for k, p in mb.named_parameters()}
        got["feats"] = fb.grad
        for k, gv in got.items():
            rel = (gv - ref[k]).norm().item() / max(1e-30, ref[k].norm().item())
            print("step engine %-5s %-24s rel-L2 %.3e" % (route, k, rel))
            assert rel <= 3e-2, (route, k, rel)
    del lb, lf
    for p in mb.parameters():
        p.grad = None
    opt = s2vt_b200.FusedAdam(mb.parameters(), lr=1e-3)
    tr = DataParallelTrainer(mb, opt, cuda_graph=True)
    losses = [float(tr.step(feats, targets).item()) for _ in range(6)]
    assert tr.replays >= 3 and losses[-1] < losses[0], (tr.replays, losses)
    tr.check_device_errors()


def test_step_engine_paper_sizing_vs_reference_golden(dev):
    """BASELINE configs[3] sizing (H = 1000, E = 500, F = 2048): loss and sampled gradients of the unmodified reference (paper.npz)."""
    from s2vt_b200 import engine_step as ES
    g, model, tf, tt, tm, c = _bf16_model("paper", dev)
    assert model._bf16_engine() is ES
    tf = tf.requires_grad_(True)
    logits = model(tf, targets=tt[:, :-1], mode="train")
    loss = s2vt_b200.MaskCriterion()(logits, tt, tm)
    loss.backward()
    assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
    ls = logits.detach().cpu().numpy().reshape(-1)[::997]
    ref = g["logits_sample"]
    print("\n" + _report("logits (paper sizing)", ls, ref), " sigma_ref %.3e" % ref.std())
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    assert np.abs(ls - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()) + 2e-2 * ref.std()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    grads["feats"] = tf.grad.cpu().numpy()
    for k, gv in grads.items():
        if "grad_full/" + k in g:
            rs, gs = g["grad_full/" + k].reshape(-1), gv.reshape(-1)
        else:
            rs, gs = g["grad_sample/" + k], gv.reshape(-1)[::997]
        rel = np.linalg.norm(gs.astype(np.float64) - rs) / max(1e-30, np.linalg.norm(rs.astype(np.float64)))
        print("grad %-24s rel-L2 err (sampled) %.3e" % (k, rel))
        assert rel <= 2e-2, k


def test_graphed_loop_body_matches_eager_loop_body(dev):
    """The reference's unchanged loop body (zero_grad, module forward, MaskCriterion, backward, optimizer.step) captured with
    GraphedLoopBody: same losses and weights as running the body eagerly, lr changes between replays take effect."""
    V, F, H, E, Lq, B = 203, 64, 128, 64, 8, 6
    g = torch.Generator().manual_seed(31)
    data = [(torch.randn(B, Lq, F, generator=g).to(dev), torch.randint(0, V, (B, Lq), generator=g).to(dev), torch.ones(B, Lq, device=dev))
            for _ in range(3)]
    out = {}
    for graphed in (False, True):
        torch.manual_seed(32)
        model = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E, train_precision="bf16").to(dev)
        opt = s2vt_b200.FusedAdam(model.parameters(), lr=1e-3)
        opt.attach(model)
        crit = s2vt_b200.MaskCriterion()

        def body(f, t, m):
            opt.zero_grad()
            loss = crit(model(f, targets=t[:, :-1], mode="train"), t, m)
            loss.backward()
            opt.step()
            return loss
        losses = []
        if graphed:
            step = s2vt_b200.GraphedLoopBody(body, data[0], optimizer=opt, warmup=2)
            n_warm = 2
        else:
            step, n_warm = body, 0
            for _ in range(2):                       # the warm-up steps GraphedLoopBody runs on its example batch
                float(body(*data[0]).item())
        for i in range(6):
            if i == 3:
                opt.param_groups[0]["lr"] = 3e-4
            losses.append(float(step(*data[i % 3]).item()))
        torch.cuda.synchronize()
        assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0
        out[graphed] = (losses, {k: p.detach().clone() for k, p in model.named_parameters()}, opt._flat["step"])
    assert out[True][2] == out[False][2] == 8
    for a, b in zip(out[True][0], out[False][0]):
        assert abs(a - b) <= 2e-3 * abs(b), (out[True][0], out[False][0])
    for k in out[True][1]:
        a, b = out[True][1][k].double(), out[False][1][k].double()
        assert (a - b).norm().item() <= 2e-3 * max(1e-30, b.norm().item()), k
