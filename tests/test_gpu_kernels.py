"""Unit parity of each CUDA kernel against a plain PyTorch expression (fp64 math on the same inputs).
All calls go through the C ABI (ctypes).  Needs a B200: run with -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import s2vt_b200  # noqa: E402
from s2vt_b200 import lib as L  # noqa: E402
from s2vt_b200 import ops  # noqa: E402
from s2vt_b200.lib import dense, rowmap  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    s2vt_b200.load()
    return torch.device("cuda:0")


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(1e-30, b.double().abs().max().item())


# ------------------------------------------------------------------ fp32 GEMM
@pytest.mark.parametrize("M,N,K", [(1, 1, 4), (64, 64, 64), (237, 40, 16), (300, 2048, 512), (5, 13000, 512), (513, 130, 1000)])
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_f32(dev, M, N, K, a_trans, b_trans):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(dev)
    Bm = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = (A.double() @ Bm.double().T + bias.double()).float()
    As = A.T.contiguous() if a_trans else A
    Bs = Bm.T.contiguous() if b_trans else Bm
    C = torch.full((M, N), float("nan"), device=dev)
    ops.gemm_f32(M, N, K, As, dense(M if a_trans else K), a_trans, Bs, dense(N if b_trans else K), b_trans, C, dense(N), bias=bias)
    assert _rel(C, ref) < 2e-6
    # accumulate on top
    ops.gemm_f32(M, N, K, As, dense(M if a_trans else K), a_trans, Bs, dense(N if b_trans else K), b_trans, C, dense(N), accumulate=True)
    ref2 = (2 * (A.double() @ Bm.double().T) + bias.double()).float()
    assert _rel(C, ref2) < 2e-6


def test_gemm_f32_rowmaps_and_splitk(dev):
    B_, L_, F_, H_ = 5, 7, 24, 16
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(B_, L_, F_, generator=g).to(dev)
    W = torch.randn(H_, F_, generator=g).to(dev)
    out = torch.empty(L_ * B_, H_, device=dev)
    ops.gemm_f32(L_ * B_, H_, F_, feats, rowmap(B_, F_, L_ * F_), False, W, dense(F_), False, out, dense(H_))
    ref = (feats.double() @ W.double().T).transpose(0, 1).reshape(L_ * B_, H_).float()
    assert _rel(out, ref) < 2e-6
    # write time-major rows into a batch-major output
    out_bm = torch.empty(B_, L_, H_, device=dev)
    ops.gemm_f32(L_ * B_, H_, F_, feats, rowmap(B_, F_, L_ * F_), False, W, dense(F_), False, out_bm, rowmap(B_, H_, L_ * H_))
    assert _rel(out_bm, (feats.double() @ W.double().T).float()) < 2e-6
    # split-K partial planes sum to the product
    M, N, K, S = 33, 70, 512, 4
    A = torch.randn(M, K, generator=g).to(dev); Bm = torch.randn(N, K, generator=g).to(dev)
    part = torch.zeros(S, M, N, device=dev)
    ops.gemm_f32(M, N, K, A, dense(K), False, Bm, dense(K), False, part, dense(N), split_k=S, split_stride=M * N)
    assert _rel(part.sum(0), (A.double() @ Bm.double().T).float()) < 2e-6


# ------------------------------------------------------------------ bf16 tcgen05 GEMM
def _gemm_bf16(dev, M, N, K, a_mn, b_mn, out_bf16=False, bias=True, accumulate=False, seed=0):
    g = torch.Generator().manual_seed(seed + M + N + K)
    A = torch.randn(M, K, generator=g).to(dev).bfloat16()
    Bm = torch.randn(N, K, generator=g).to(dev).bfloat16()
    bv = torch.randn(N, generator=g).to(dev) if bias else None
    ref = A.double() @ Bm.double().T
    if bias:
        ref = ref + bv.double()
    As = A.T.contiguous() if a_mn else A
    Bs = Bm.T.contiguous() if b_mn else Bm
    C0 = torch.randn(M, N, generator=g).to(dev)
    C = C0.clone() if not out_bf16 else torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    if accumulate:
        ref = ref + C0.double()
    rc = L.load().s2vt_gemm_bf16(L.stream_ptr(dev), M, N, K, L.ptr(As), (M if a_mn else K), int(a_mn), L.ptr(Bs), (N if b_mn else K),
                                 int(b_mn), L.ptr(C), dense(N), int(out_bf16), L.ptr(bv), int(accumulate))
    L.check(rc, "s2vt_gemm_bf16")
    flag = L.load().s2vt_device_error_flag(L.stream_ptr(dev))
    assert flag == 0, "device error flag %d" % flag
    return C, ref


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 512), (256, 384, 4096), (5120, 512, 4096), (200, 136, 72), (5056, 13000, 512)])
def test_gemm_bf16_kmajor(dev, M, N, K):
    C, ref = _gemm_bf16(dev, M, N, K, False, False)
    err = (C.double() - ref).abs().max().item()
    assert err < 2e-3 * (K ** 0.5), err          # exact products, fp32 accumulation: error ~ 1e-6 * sqrt(K) * |terms|


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (2048, 512, 10176), (520, 264, 200)])
def test_gemm_bf16_mn_major(dev, M, N, K, a_mn, b_mn):
    C, ref = _gemm_bf16(dev, M, N, K, a_mn, b_mn)
    err = (C.double() - ref).abs().max().item()
    assert err < 2e-3 * (K ** 0.5), err


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(5056, 512, 13000, False, True), (13000, 512, 5056, True, True), (1000, 520, 2048, False, True),
                                             (136, 260, 1104, True, False)])
def test_gemm_bf16_persistent_splitk_and_cta_cap(dev, M, N, K, a_mn, b_mn):
    """The persistent kernel (dense C) with split-K reduce-add, ragged edges, and a capped grid (the setting used while the
    recurrence clusters hold part of the machine); the one-tile-per-CTA kernel must agree with it."""
    lib = L.load()
    outs = []
    for max_ctas, persistent in ((0, 1), (37, 1), (0, 0)):
        lib.s2vt_gemm_bf16_set_mode(max_ctas, persistent)
        try:
            C, ref = _gemm_bf16(dev, M, N, K, a_mn, b_mn)
        finally:
            lib.s2vt_gemm_bf16_set_mode(0, 1)
        assert (C.double() - ref).abs().max().item() < 2e-3 * (K ** 0.5)
        outs.append(C)
    assert (outs[0] - outs[2]).abs().max().item() < 1e-3 * (K ** 0.5)
    assert (outs[0] - outs[1]).abs().max().item() < 1e-3 * (K ** 0.5)


def test_gemm_bf16_persistent_back_to_back_launches(dev):
    """Many launches reuse the self-resetting scheduler slots; results must stay correct launch after launch."""
    for i in range(40):
        C, ref = _gemm_bf16(dev, 384 + 8 * (i % 3), 520, 192, False, False, seed=i)
        assert (C.double() - ref).abs().max().item() < 2e-3 * (192 ** 0.5), i
    C, ref = _gemm_bf16(dev, 700, 328, 640, False, False, out_bf16=True)
    assert (C.double() - ref).abs().max().item() < 0.02 * ref.abs().max().item()


@pytest.mark.parametrize("b_mn,reverse_m,accumulate", [(False, False, False), (True, True, False), (False, False, True), (True, False, True)])
def test_gemm_bf16_gated_matches_plain_product(dev, b_mn, reverse_m, accumulate):
    """s2vt_gemm_bf16_gated with every chunk already released: the plain product, and every chunk's ready flag raised exactly once."""
    from s2vt_b200 import engine_bf16 as EB
    M, N, K = 1000, 512, 320                              # 8 row tiles (the last one partial), chunk bounds off the tile grid
    g = torch.Generator().manual_seed(17)
    A = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    Bm = torch.randn(N, K, generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    rows = [0, 64, 200, 456, 999, 1000]
    n = len(rows) - 1
    ctr = torch.zeros(3, EB.MAX_SYNC, dtype=torch.int32, device=dev)
    ctr[0, :n] = 7
    C0 = torch.randn(M, N, generator=g).to(dev)
    C = C0.clone() if accumulate else torch.full((M, N), float("nan"), device=dev)
    Bs = Bm.T.contiguous() if b_mn else Bm
    EB.gemm_gated(M, N, K, A, K, Bs, N if b_mn else K, b_mn, C, N, rows, ctr[0], 7, ctr[1], ctr[2], bias=None if accumulate else bias,
                  accumulate=accumulate, max_ctas=5, reverse_m=reverse_m)
    torch.cuda.synchronize()
    ref = A.double() @ Bm.double().T + (C0.double() if accumulate else bias.double())
    assert _rel(C, ref.float()) < 1e-5
    assert ctr[2, :n].tolist() == [1] * n and ctr[2, n:].abs().sum().item() == 0
    assert L.load().s2vt_device_error_flag(L.stream_ptr(dev)) == 0


def test_gemm_bf16_gated_waits_for_its_chunks(dev):
    """The gated product is launched FIRST and spins; its chunks are released one by one from another stream (last chunk first, as a
    backward sweep does) behind copies that fill the corresponding rows of A -- a tile read before its release would see zeros."""
    from s2vt_b200 import engine_bf16 as EB
    M, N, K = 1536, 256, 512
    g = torch.Generator().manual_seed(19)
    A_src = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    Bm = torch.randn(N, K, generator=g).to(dev).to(torch.bfloat16)
    A = torch.zeros_like(A_src)
    rows = [0, 300, 700, 1100, 1536]
    n = len(rows) - 1
    ctr = torch.zeros(3, EB.MAX_SYNC, dtype=torch.int32, device=dev)
    C = torch.full((M, N), float("nan"), device=dev)
    lib = L.load()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):                          # first launches load kernel code, which can wait for the device to drain:
        torch.cuda._sleep(1000)                            # do them before a spinning kernel is resident
        A[:8].copy_(A_src[:8])
        A[:8].zero_()
        L.check(lib.s2vt_stream_write_value32(side.cuda_stream, L.ptr(ctr[0], EB.MAX_SYNC - 1), 0), "s2vt_stream_write_value32")
    torch.cuda.synchronize()
    EB.gemm_gated(M, N, K, A, K, Bm, K, False, C, N, rows, ctr[0], 1, ctr[1], ctr[2], max_ctas=4, reverse_m=True)
    with torch.cuda.stream(side):
        for k in range(n - 1, -1, -1):
            torch.cuda._sleep(200000)                      # ~0.1 ms between releases
            A[rows[k]:rows[k + 1]].copy_(A_src[rows[k]:rows[k + 1]])
            L.check(lib.s2vt_stream_write_value32(side.cuda_stream, L.ptr(ctr[0], k), 1), "s2vt_stream_write_value32")
    torch.cuda.synchronize()
    assert _rel(C, (A_src.double() @ Bm.double().T).float()) < 1e-5
    assert ctr[2, :n].tolist() == [1] * n
    assert lib.s2vt_device_error_flag(L.stream_ptr(dev)) == 0


def test_gemm_bf16_epilogues(dev):
    C, ref = _gemm_bf16(dev, 300, 200, 256, False, False, out_bf16=True)
    assert (C.double() - ref).abs().max().item() < 0.02 * ref.abs().max().item()
    C, ref = _gemm_bf16(dev, 300, 200, 256, False, False, bias=False, accumulate=True)
    assert (C.double() - ref).abs().max().item() < 2e-3 * 16


# ------------------------------------------------------------------ small kernels
def test_ce_and_colsum_and_embed(dev):
    g = torch.Generator().manual_seed(5)
    R, V = 37, 1300
    z = (torch.randn(R, V, generator=g) * 3).to(dev)
    t = torch.randint(0, V, (R,), generator=g).to(dev)
    loss = torch.empty((), device=dev)
    dl = torch.empty_like(z)
    gs = torch.tensor(0.5, device=dev)
    ops.ce_f32(z, R, V, t, 0, dense(1), loss, dlogits=dl, gscale=gs)
    zr = z.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(zr, t)
    (ref * 0.5).backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert (dl.double() - zr.grad).abs().max().item() < 1e-7
    cs = torch.empty(V, device=dev)
    ops.colsum_f32(z, R, V, V, cs)
    assert _rel(cs, z.double().sum(0).float()) < 1e-5
    E, Bn, nt = 12, 3, 5
    tab = torch.randn(50, E, generator=g).to(dev)
    ids = torch.randint(0, 50, (Bn, nt + 1), generator=g).to(dev)
    out = torch.empty(nt * Bn, E, device=dev)
    ops.embed_gather_f32(tab, ids, 0, nt + 1, Bn, nt, out, E)
    assert torch.equal(out.view(nt, Bn, E), tab[ids[:, :nt]].transpose(0, 1))
    gt = torch.zeros(50, E, device=dev)
    ops.embed_scatter_add_f32(gt, ids, 0, nt + 1, Bn, nt, out, E)
    ref_g = torch.zeros(50, E, device=dev, dtype=torch.float64)
    ref_g.index_add_(0, ids[:, :nt].T.reshape(-1), out.double())
    assert _rel(gt, ref_g.float()) < 1e-6


def test_adam_matches_torch(dev):
    g = torch.Generator().manual_seed(9)
    p0 = torch.randn(1001, generator=g).to(dev)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-4)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(1001, generator=g).to(dev)
        ref_p.grad = gr.clone()
        opt.step()
        ops.adam_f32(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, step)
        assert (p - ref_p.data).abs().max().item() < 1e-7


def test_cast_bf16(dev):
    x = torch.randn(70, 130, device=dev)
    d = torch.empty(70, 130, dtype=torch.bfloat16, device=dev)
    dt = torch.empty(130, 70, dtype=torch.bfloat16, device=dev)
    rc = L.load().s2vt_cast_bf16(L.stream_ptr(dev), L.ptr(x), L.ptr(d), L.ptr(dt), 70, 130)
    L.check(rc, "cast")
    assert torch.equal(d, x.bfloat16()) and torch.equal(dt, x.bfloat16().T.contiguous())
    d2 = torch.empty(70 * 130, dtype=torch.bfloat16, device=dev)
    L.check(L.load().s2vt_cast_bf16(L.stream_ptr(dev), L.ptr(x), L.ptr(d2), None, 70, 130), "cast")
    assert torch.equal(d2.view(70, 130), x.bfloat16())


@pytest.mark.parametrize("R,V,K", [(300, 1000, 128), (5056, 13000, 512), (129, 520, 64)])
def test_vocab_projection_fused_with_ce(dev, R, V, K):
    """s2vt_vocab_ce_fwd_bf16 / s2vt_ce_dlogits_inplace_bf16 against a plain fp32 torch reference of the same op
    (bf16 operands, fp32 accumulation): loss rtol 1e-4, lse atol 1e-4, logits to bf16 resolution, dlogits to bf16 resolution."""
    from s2vt_b200.lib import rowmap
    g = torch.Generator().manual_seed(R + V)
    lib = L.load()
    A = (torch.randn(R, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(V, K, generator=g) * (2.0 / K ** 0.5)).to(dev).bfloat16()
    bias = torch.randn(V, generator=g).to(dev)
    Bq, Lq = 4, R // 4 + 1                                   # targets addressed through a row map, like targets_full[b, t+1]
    tg_full = torch.randint(0, V, (Bq, Lq + 1), generator=g).to(dev)
    tmap = rowmap(Bq, 1, Lq + 1)                             # row r -> tg_full[r % Bq, r // Bq + 1]
    rows = torch.arange(R, device=dev)
    tgt = tg_full[rows % Bq, rows // Bq + 1]
    z = A.float() @ W.float().T + bias
    lse_ref = torch.logsumexp(z.double(), dim=1)
    loss_ref = (lse_ref - z.double()[rows, tgt]).mean().item()
    logits = torch.empty(R, V, device=dev, dtype=torch.bfloat16)
    part = torch.empty(int(lib.s2vt_vocab_ce_ws_bytes(R, V)), dtype=torch.uint8, device=dev)
    ztgt, lse, row_loss = (torch.empty(R, device=dev) for _ in range(3))
    loss = torch.empty((), device=dev)
    rc = lib.s2vt_vocab_ce_fwd_bf16(L.stream_ptr(dev), R, V, K, L.ptr(A), K, L.ptr(W), K, L.ptr(bias), L.ptr(logits), V, L.ptr(tg_full, 1), tmap,
                                    L.ptr(part), L.ptr(ztgt), L.ptr(lse), L.ptr(row_loss), L.ptr(loss))
    L.check(rc, "s2vt_vocab_ce_fwd_bf16")
    assert lib.s2vt_device_error_flag(L.stream_ptr(dev)) == 0
    assert abs(loss.item() - loss_ref) <= 1e-4 * abs(loss_ref)
    assert (lse.double() - lse_ref).abs().max().item() <= 1e-4 * max(1.0, lse_ref.abs().max().item())
    assert (logits.float() - z).abs().max().item() <= 2 ** -8 * z.abs().max().item() + 1e-6
    gs = torch.tensor(0.7, device=dev)
    rc = lib.s2vt_ce_dlogits_inplace_bf16(L.stream_ptr(dev), L.ptr(logits), R, V, V, L.ptr(lse), L.ptr(tg_full, 1), tmap, L.ptr(gs))
    L.check(rc, "s2vt_ce_dlogits_inplace_bf16")
    prob = torch.softmax(z.double(), dim=1)
    dref = prob.clone()
    dref[rows, tgt] -= 1.0
    dref *= 0.7 / R
    err = (logits.double() - dref).abs()
    # the probabilities come from bf16-rounded logits (relative error <= |z| * 2^-9 each), the result is rounded to bf16 (2^-9)
    tol = 2 ** -8 * (1.0 + z.double().abs()) * prob * (0.7 / R) + 2 ** -8 * dref.abs() + 1e-12
    assert (err <= tol).all(), (err - tol).max().item()
