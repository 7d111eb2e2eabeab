"""A numpy model of the fp32-parity dense layer's arithmetic (csrc/gemm_tc.cu, MODE_TF32X3): operands split into
TF32-rounded hi and truncated lo halves, hi*hi + hi*lo + lo*hi per product, tensor-core accumulation that TRUNCATES when
it adds into its fp32 accumulator, a separate accumulator for the correction terms, and promotion of the partial sums to
round-to-nearest fp32 registers every 128 reduction elements.  It documents why those three design choices are there:
with them the result error is fp32-class (the 1e-5 parity bar) independent of K; without promotion the truncation bias
grows with K (measured on the B200: up to 3e-5 at K = 4096, scripts/exp_promote.sh)."""
import numpy as np
import pytest


def split_tf32(x):
    xf = x.astype(np.float32)
    hi = ((xf.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)      # round to 10 explicit bits
    lo = (xf - hi).astype(np.float32)
    lo = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)                              # the MMA truncates lo
    return hi.astype(np.float64), lo.astype(np.float64)


def trunc32(x):
    """round toward zero to fp32 (the tensor core's accumulate step)"""
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y.astype(np.float64)


def gemm_tf32x3(a, b, promote=128, split_acc=True, kstep=8):
    """C = A @ B.T the way the kernel evaluates it; a [M,K], b [N,K] fp32."""
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    m, k = a.shape
    reg = np.zeros((m, b.shape[0]))
    main = np.zeros_like(reg)
    corr = np.zeros_like(reg)
    for t0 in range(0, k, kstep):
        s = slice(t0, t0 + kstep)
        c = al[:, s] @ bh[:, s].T + ah[:, s] @ bl[:, s].T
        if split_acc:
            corr = trunc32(corr + c)
            main = trunc32(main + ah[:, s] @ bh[:, s].T)
        else:
            main = trunc32(trunc32(main + c) + ah[:, s] @ bh[:, s].T)
        if (t0 + kstep) % promote == 0 or t0 + kstep >= k:
            part = (main + corr).astype(np.float32).astype(np.float64)
            reg = (reg + part).astype(np.float32).astype(np.float64)                                # RN add in registers
            main[:] = 0
            corr[:] = 0
    return reg


def rel_err(c, ref):
    return float(np.abs(c - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("k", [64, 1024, 4096])
def test_fp32_class_error_independent_of_k(k):
    g = np.random.default_rng(k)
    a = g.normal(size=(24, k)).astype(np.float32) + 0.5          # a positive mean makes truncation bias visible
    b = g.normal(size=(16, k)).astype(np.float32) + 0.5
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    e = rel_err(gemm_tf32x3(a, b), ref)
    assert e < 1e-6, e                                           # the parity bar is 1e-5 per tensor


def test_without_promotion_the_truncation_bias_grows_with_k():
    g = np.random.default_rng(7)
    errs = {}
    for k in (256, 4096):
        a = np.abs(g.normal(size=(16, k))).astype(np.float32)
        b = np.abs(g.normal(size=(16, k))).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        errs[k] = (rel_err(gemm_tf32x3(a, b, promote=1 << 30, split_acc=False), ref), rel_err(gemm_tf32x3(a, b), ref))
    assert errs[4096][0] > 4 * errs[256][0]                      # one truncating accumulator: error ~ linear in K
    assert errs[4096][0] > 10 * errs[4096][1]                    # promotion + separate correction accumulator fix it
    assert errs[4096][1] < 1e-6


def test_dropping_the_lo_lo_term_is_harmless():
    g = np.random.default_rng(3)
    a = g.normal(size=(8, 512)).astype(np.float32)
    b = g.normal(size=(8, 512)).astype(np.float32)
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    three = ah @ bh.T + ah @ bl.T + al @ bh.T
    assert rel_err(three, ref) < 2e-7
    assert rel_err(ah @ bh.T, ref) > 1e-5                        # one TF32 pass alone misses the bar
