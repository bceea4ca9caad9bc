"""CPU checks of the closed forms the GEMM epilogues use for the towers' activations (csrc/gemm_tcgen05.cu `activation<ACT>`),
restated in numpy fp32 with the kernel's operation order, against the definitions the reference's towers call
(hf activations.py: GELUActivation = 0.5 x (1 + erf(x / sqrt 2)) for EVA / BLIP-2, QuickGELUActivation = x sigmoid(1.702 x) for
CLIP, gelu_pytorch_tanh for SigLIP). The GPU tests (test_kernels_gpu.py::test_gemm_activation) check the kernels themselves;
these pin the formulas, so a wrong constant cannot hide inside the 16-bit output tolerance."""
import numpy as np
from scipy.special import erf

F = np.float32
X = np.linspace(-12.0, 12.0, 600001).astype(F)


def gelu_erf_kernel_form(x):
    """Abramowitz & Stegun 7.1.26 with the kernel's constants: erf|z| = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p |z|)."""
    az = np.abs(x) * F(0.70710678118654752)
    t = F(1.0) / (az * F(0.3275911) + F(1.0))
    pl = t * F(1.061405429) + F(-1.453152027)
    pl = pl * t + F(1.421413741)
    pl = pl * t + F(-0.284496736)
    pl = pl * t + F(0.254829592)
    pl = pl * t
    e = np.exp2(az * az * F(-1.4426950408889634)).astype(F)
    erf_abs = F(1.0) - pl * e
    hx = F(0.5) * x
    return np.abs(hx) * erf_abs + hx


def test_erf_gelu_closed_form_is_exact_to_fp32_noise():
    ref = 0.5 * X.astype(np.float64) * (1.0 + erf(X.astype(np.float64) / np.sqrt(2.0)))
    got = gelu_erf_kernel_form(X).astype(np.float64)
    err = np.abs(got - ref)
    # a tenth of the smallest normal fp16 step (6.1e-5) the value is rounded to; A&S 7.1.26 is good to 1.5e-7 on erf
    assert err.max() < 6e-7, (err.max(), X[err.argmax()])
    # odd-part symmetry of the two branches (x >= 0: hx (1 + erf|z|); x < 0: hx (1 - erf|z|))
    assert np.all(got[X > 6] == X[X > 6].astype(np.float64))
    assert np.all(np.abs(got[X < -6]) < 1e-6)


def test_quick_gelu_is_one_tanh():
    """x sigmoid(1.702 x) == 0.5 x (1 + tanh(0.851 x)): the identity behind the single-MUFU epilogue of CLIP's fc1."""
    x = X.astype(np.float64)
    ref = x / (1.0 + np.exp(-1.702 * x))
    got = 0.5 * x * (1.0 + np.tanh(0.851 * x))
    assert np.abs(got - ref).max() < 1e-12
    # tanh.approx's worst-case relative error (2^-11, PTX ISA) moves the value by at most 0.5 |x| |tanh| 2^-11: below half an
    # fp16 step for x >= 0 (relative 2^-12 at most); for x < 0 the form 1 + tanh cancels, so the same ABSOLUTE error meets a
    # small value (up to 2e-3 relative at |y| = 0.125) — the error class of SigLIP's gelu_pytorch_tanh epilogue; the live-oracle
    # parity of CLIP did not move with it (logit error 3.09e-3 before, 2.98e-3 after; tests/test_parity_live.py)
    pert = 0.5 * x * (1.0 + np.tanh(0.851 * x) * (1.0 + 2.0 ** -11))
    assert np.all(np.abs(pert - got) <= 0.5 * np.abs(x) * np.abs(np.tanh(0.851 * x)) * 2.0 ** -11 + 1e-12)
    pos = x > 0
    assert (np.abs(pert - got)[pos] / np.abs(got)[pos]).max() <= 2.0 ** -12 + 1e-9


def test_gelu_tanh_constants():
    """gelu_pytorch_tanh: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with the kernel's folded constants."""
    x = X.astype(np.float64)
    c0 = 0.7978845608028654
    c1 = c0 * 0.044715
    u = x * (x * x * c1 + c0)
    got = 0.5 * x + 0.5 * x * np.tanh(u)
    ref = 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))
    assert np.abs(got - ref).max() < 1e-12
