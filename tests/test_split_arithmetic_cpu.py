"""CPU restatement of the fp16 hi/lo operand split of the tensor-core decode path (csrc/xdec_sm100.cu: scale_from_bits, split_kernel,
the three-pass product) and the bounds DESIGN.md section 5.6 states for it.  numpy only: documents the arithmetic the GPU tests
(tests/test_gpu_xdec.py) then measure on the hardware, where the tensor core's truncating accumulate is the remaining error source."""
import numpy as np

from s2vt_b200 import engine_step as ES
from s2vt_b200 import engine_bf16 as EB
import torch


def pow2_scale(absmax: np.float32) -> np.float32:
    """scale_from_bits: the power of two that puts max|x| into [2^14, 2^15)"""
    bits = np.float32(absmax).view(np.uint32)
    e = int((bits >> 23) & 0xFF) - 127
    if (bits & 0x7F800000) in (0, 0x7F800000):
        return np.float32(1.0)
    return np.float32(2.0 ** max(-100, min(100, 14 - e)))


def split(x: np.ndarray):
    s = pow2_scale(np.abs(x).max())
    xs = (x * s).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi, lo, s


def test_scale_puts_the_maximum_into_2_14_2_15():
    rng = np.random.default_rng(0)
    for mag in (1e-6, 3e-3, 0.04, 1.0, 37.0, 6.5e4, 3e9):
        x = (rng.standard_normal(1000) * mag).astype(np.float32)
        _, _, s = split(x)
        m = np.abs(x).max() * s
        assert 2 ** 14 <= m < 2 ** 15
        assert np.log2(s) == np.round(np.log2(s))            # a power of two: scaling and un-scaling are exact


def test_hi_plus_lo_represents_x_to_2_pow_minus_23():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(200000) * np.exp(3 * rng.standard_normal(200000))).astype(np.float32)    # ~6 decades
    hi, lo, s = split(x)
    xs = x.astype(np.float64) * float(s)
    err = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - xs)
    # relative 2^-23 where the residual is a normal fp16 (|x s| >= 2^-3), absolute 2^-25 below that (subnormal residual)
    assert np.all(err <= np.maximum(2.0 ** -23 * np.abs(xs), 2.0 ** -25))
    assert not np.isinf(hi.astype(np.float32)).any()


def test_three_pass_product_error_bound():
    """sum_k (a_hi b_lo + a_lo b_hi + a_hi b_hi) against the fp64 dot product: the dropped a_lo b_lo term and the plane rounding
    give <= 2^-21 relative per term, i.e. <= 2^-21 * sum |a_k b_k| in total (no accumulation error here: the sums run in fp64)."""
    rng = np.random.default_rng(2)
    K = 512
    a = np.tanh(rng.standard_normal((64, K))).astype(np.float32)                      # hidden states
    b = ((rng.random((96, K)) * 2 - 1) / np.sqrt(K)).astype(np.float32)               # nn.Linear-style weights
    ah, al, sa = split(a)
    bh, bl, sb = split(b)
    A_h, A_l, B_h, B_l = (t.astype(np.float64) for t in (ah, al, bh, bl))
    got = (A_h @ B_l.T + A_l @ B_h.T + A_h @ B_h.T) / (float(sa) * float(sb))
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    bound = 2.0 ** -21 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T)
    assert np.all(np.abs(got - ref) <= bound)
    # every fp16 x fp16 product is exact in fp32 (22 significant bits at most)
    p = ah[:4, :8].astype(np.float32)[:, None, :] * bh[:4, :8].astype(np.float32)[None, :, :]
    assert np.array_equal(p.astype(np.float64), ah[:4, :8].astype(np.float64)[:, None, :] * bh[:4, :8].astype(np.float64)[None, :, :])


def test_gate_interleave_and_padding_helpers():
    H, K = 6, 5
    w = torch.arange(4 * H * K, dtype=torch.float32).reshape(4 * H, K)
    wi = ES._il(w)
    for g in range(4):
        for u in range(H):
            assert torch.equal(wi[4 * u + g], w[g * H + u])
    b = torch.arange(4 * H, dtype=torch.float32)
    bi = ES._il(b)
    assert all(bi[4 * u + g] == b[g * H + u] for g in range(4) for u in range(H))
    p = ES._pad_cols(w[:, :3].contiguous(), 8)
    assert p.shape == (4 * H, 8) and torch.equal(p[:, :3], w[:, :3]) and not p[:, 3:].any()
    assert [EB.pad8(n) for n in (1, 8, 9, 13000, 13001)] == [8, 8, 16, 13000, 13008]
    assert ES.supported(1000, 500, 2048, 13001) and not ES.supported(1001, 500, 2048, 13000) and not ES.supported(1000, 500, 2050, 13000)
    assert EB.supported(512, 512, 4096, 13001) and not EB.supported(1000, 512, 4096, 13000)
