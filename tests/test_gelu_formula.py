"""The FFN-up epilogue's GELU (csrc/gemm_tc.cu::gelu_erf) is a fitted closed form, not erf: check the formula
itself against HF's erf GELU (the activation of BertIntermediate, hidden_act="gelu") on the CPU."""
import math

import numpy as np

KC, KA, KB = 7.97507884e-01, 3.70056460e-02, -3.51516788e-04   # csrc/gemm_tc.cu


def gelu_kernel_formula(x, tanh=np.tanh):
    x = x.astype(np.float32)
    t = np.minimum(x * x, np.float32(36.0))
    p = (np.float32(KB) * t + np.float32(KA)) * t + np.float32(KC)
    hx = np.float32(0.5) * x
    return hx + hx * tanh((p * x).astype(np.float32)).astype(np.float32)


def gelu_erf(x):
    return np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in x.astype(np.float64)])


def test_formula_matches_erf_gelu():
    x = np.concatenate([np.linspace(-12, 12, 48001), np.array([-100.0, -30.0, 30.0, 100.0])]).astype(np.float32)
    err = np.abs(gelu_kernel_formula(x).astype(np.float64) - gelu_erf(x))
    assert err.max() < 3e-5, err.max()


def test_formula_tolerates_the_hardware_tanh_error():
    # tanh.approx.f32 is specified to 2^-11 relative error: the result moves by at most 2^-11 |x / 2|,
    # i.e. less than half a bf16 ulp (2^-9 relative) of any output with |gelu(x)| >= |x| / 8
    x = np.linspace(-8, 8, 16001).astype(np.float32)
    hi = gelu_kernel_formula(x, tanh=lambda u: np.tanh(u) * (1 + 2.0 ** -11))
    lo = gelu_kernel_formula(x, tanh=lambda u: np.tanh(u) * (1 - 2.0 ** -11))
    ref = gelu_erf(x)
    worst = np.maximum(np.abs(hi - ref), np.abs(lo - ref))
    assert (worst <= 2.0 ** -11 * np.abs(x) / 2 + 3e-5).all()
    assert worst.max() < 2.5e-3
