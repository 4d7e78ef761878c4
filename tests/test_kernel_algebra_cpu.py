"""CPU restatements of two algebraic identities the CUDA kernels rely on.  No GPU, no library calls:
these pin the MATH; tests/test_kernels_gpu.py pins the kernels against torch."""
from fractions import Fraction

import numpy as np
import torch
import torch.nn.functional as F


def phase_tap(a, k):
    """conv_tcg.cu::phase_tap: which low-res offset (-1, 0, +1) tap k of output phase a lands on."""
    return (-1 if k == 0 else 0) if a == 0 else (1 if k == 2 else 0)


def phase_weights(w, cg):
    """sgqn_conv_weights_prep_phase (without the TF32 rounding): w (Co,Ci,3,3) -> (4*cg, Ci, 3, 3), phase p = 2a+b owns rows
    [p*cg, p*cg+Co)."""
    co, ci = w.shape[:2]
    out = torch.zeros(4 * cg, ci, 3, 3, dtype=w.dtype)
    for a in range(2):
        for b in range(2):
            for ky in range(3):
                for kx in range(3):
                    out[(2 * a + b) * cg:(2 * a + b) * cg + co, :, phase_tap(a, ky) + 1, phase_tap(b, kx) + 1] += w[:, :, ky, kx]
    return out


def phase_fold(dwp, co, cg):
    """sgqn_conv_phase_fold: chain rule of phase_weights."""
    dw = torch.zeros(co, dwp.shape[1], 3, 3, dtype=dwp.dtype)
    for a in range(2):
        for b in range(2):
            for ky in range(3):
                for kx in range(3):
                    dw[:, :, ky, kx] += dwp[(2 * a + b) * cg:(2 * a + b) * cg + co, :, phase_tap(a, ky) + 1, phase_tap(b, kx) + 1]
    return dw


def depth_to_space(yp, co, cg):
    """(B, 4*cg, H, W) phase layout -> (B, co, 2H, 2W)."""
    B, _, H, W = yp.shape
    y = yp.reshape(B, 2, 2, cg, H, W)[:, :, :, :co]                 # (B, a, b, co, H, W)
    return y.permute(0, 3, 4, 1, 5, 2).reshape(B, co, 2 * H, 2 * W)


def test_conv_after_nearest_upsample_is_a_low_resolution_phase_conv():
    """modules.py:327-337: conv3x3(pad 1) over F.upsample(x, 2) == depth_to_space(conv3x3(pad 1) over x with the phase weights),
    forward and both gradients (float64: the identity is exact up to summation order)."""
    g = torch.Generator().manual_seed(0)
    for (B, ci, co, cg, H) in ((2, 5, 3, 4, 6), (1, 8, 9, 16, 7), (2, 4, 4, 4, 5)):
        x = torch.randn(B, ci, H, H, generator=g, dtype=torch.float64, requires_grad=True)
        w = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
        b = torch.randn(co, generator=g, dtype=torch.float64)
        ref = F.conv2d(F.interpolate(x, scale_factor=2), w, b, padding=1)
        wp = phase_weights(w.detach(), cg).requires_grad_(True)
        bp = torch.zeros(4 * cg, dtype=torch.float64)
        for p in range(4):
            bp[p * cg:p * cg + co] = b
        x2 = x.detach().clone().requires_grad_(True)
        got = depth_to_space(F.conv2d(x2, wp, bp, padding=1), co, cg)
        assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)
        dy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        ref.backward(dy); got.backward(dy)
        assert torch.allclose(x2.grad, x.grad, rtol=1e-12, atol=1e-12)      # the data gradient lands on the low-res tensor
        assert torch.allclose(phase_fold(wp.grad, co, cg), w.grad, rtol=1e-12, atol=1e-12)
        # 4 of the 9 taps of every phase are structurally non-zero
        nz = (phase_weights(torch.ones(1, 1, 3, 3, dtype=torch.float64), 1) != 0).reshape(4, 9).sum(1)
        assert nz.tolist() == [4, 4, 4, 4]


def _f32(fr):
    return np.float32(float(fr))


def test_division_by_255_through_two_fmas_is_correctly_rounded():
    """conv1_tc.cu: q0 = x*rcp; r = fma(-q0, 255, x); q = fma(r, rcp, q0) equals the IEEE quotient x / 255 (what the reference's
    NormalizeImg computes) -- all integer pixel values and a random sample of the [0, 256) range, in exact rational arithmetic."""
    rcp = np.float32(1.0) / np.float32(255.0)
    rs = np.random.RandomState(0)
    xs = np.concatenate([np.arange(256, dtype=np.float32), rs.uniform(0, 256, 4000).astype(np.float32),
                         (rs.uniform(0, 1, 500) ** 8 * 256).astype(np.float32)])
    for x in xs:
        q0 = np.float32(np.float64(x) * np.float64(rcp))                    # fp32 product, correctly rounded (exact in fp64)
        rem = Fraction(float(x)) - Fraction(float(q0)) * 255                 # fma: exact, then one rounding
        rem32 = _f32(rem)
        assert Fraction(float(rem32)) == rem                                 # the residual is exactly representable
        q = _f32(Fraction(float(rem32)) * Fraction(float(rcp)) + Fraction(float(q0)))
        assert q == np.float32(x) / np.float32(255.0), (x, q, np.float32(x) / np.float32(255.0))
