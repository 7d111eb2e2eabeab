"""A numpy model of the fp32-parity dense layer's arithmetic (csrc/gemm_tc.cu, MODE_F16X3): every operand scaled by a
power of two picked from (a bound of) its amax, split into fp16 planes hi = fp16(x 2^s), lo = fp16((x 2^s - hi) 2^11),
products hi*hi in one truncating tensor-core accumulator and hi*lo + lo*hi in a second one, promotion to round-to-nearest
registers every 128 reduction elements, scales undone at the end.  It documents why the design is safe: the error is
fp32-class (the parity bar is 1e-5 per tensor) independent of K, for gradient-sized and heavy-tailed data, and with an
amax BOUND that is loose by many binades -- which is what lets producer kernels write planes before they know the true
maximum (hid <= 2 max|PQ|, |dPQ| <= max|dhid| dq_factor, |norm out| <= max|res| + sqrt(rows))."""
import numpy as np
import pytest


def trunc32(x):
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y.astype(np.float64)


def plane_shift(bound):
    """s with bound * 2^s in [2^14, 2^15): csrc/common.cuh plane_shift"""
    if bound == 0 or not np.isfinite(bound):
        return 0
    e = int(np.frexp(np.float32(bound))[1])          # bound = m * 2^e, m in [0.5, 1)
    return max(-110, min(110, 15 - e))


def split_f16(x, s):
    xs = (x.astype(np.float32) * np.float32(2.0 ** s)).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = ((xs - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def gemm_f16x3(a, b, sa, sb, promote=128, kstep=16):
    ah, al = split_f16(a, sa)
    bh, bl = split_f16(b, sb)
    m, k = a.shape
    reg = np.zeros((m, b.shape[0]))
    main = np.zeros_like(reg)
    corr = np.zeros_like(reg)
    for t0 in range(0, k, kstep):
        sl = slice(t0, t0 + kstep)
        main = trunc32(main + ah[:, sl] @ bh[:, sl].T)
        corr = trunc32(trunc32(corr + ah[:, sl] @ bl[:, sl].T) + al[:, sl] @ bh[:, sl].T)
        if (t0 + kstep) % promote == 0 or t0 + kstep >= k:
            part = (main.astype(np.float32) + corr.astype(np.float32) * np.float32(1.0 / 2048.0)).astype(np.float64)
            reg = (reg + part).astype(np.float32).astype(np.float64)
            main[:] = 0
            corr[:] = 0
    return (reg.astype(np.float32) * np.float32(2.0 ** (-(sa + sb)))).astype(np.float64)


def rel_err(c, ref):
    return float(np.abs(c - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("k", [64, 1024, 4096])
@pytest.mark.parametrize("loose_bits", [0, 10, 18])
def test_fp32_class_error_for_any_k_and_loose_bounds(k, loose_bits):
    g = np.random.default_rng(k + loose_bits)
    a = (g.normal(size=(24, k)) + 0.5).astype(np.float32) * np.float32(1e-6)             # gradient-sized, positive mean
    b = (g.normal(size=(16, k)) * np.exp(g.normal(size=(16, k)) * 3)).astype(np.float32)  # heavy tails
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    sa = plane_shift(np.abs(a).max() * 2.0 ** loose_bits)
    sb = plane_shift(np.abs(b).max() * 2.0 ** loose_bits)
    assert rel_err(gemm_f16x3(a, b, sa, sb), ref) < 1.5e-6


def test_exponent_range_is_used_in_full():
    x = np.array([[3.0e-9, 1.7e-3, 0.9]], dtype=np.float32)
    s = plane_shift(np.abs(x).max())
    xs = x * np.float32(2.0 ** s)
    assert 2 ** 14 <= float(np.abs(xs).max()) < 2 ** 15
    hi, lo = split_f16(x, s)
    back = (hi + lo / 2048.0) * 2.0 ** -s
    assert np.all(np.abs(back - x.astype(np.float64)) <= np.abs(x) * 2.0 ** -21 + np.abs(x).max() * 2.0 ** -40)


def test_one_pass_keeps_eleven_bits():
    g = np.random.default_rng(3)
    a = g.normal(size=(8, 512)).astype(np.float32)
    b = g.normal(size=(8, 512)).astype(np.float32)
    ah, _ = split_f16(a, plane_shift(np.abs(a).max()))
    bh, _ = split_f16(b, plane_shift(np.abs(b).max()))
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    one = (ah @ bh.T) * 2.0 ** -(plane_shift(np.abs(a).max()) + plane_shift(np.abs(b).max()))
    assert 1e-5 < rel_err(one, ref) < 2e-3                   # the reduced-precision mode: ~2^-12 per operand
