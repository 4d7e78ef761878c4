"""Per-kernel parity (B200): every C-ABI entry point against a plain fp32 torch restatement of the same op,
the oracle's integer/byte semantics bit-exactly, and the golden vectors generated from the reference."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from sgqn_carla_b200._lib import K
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"


def P(t, off=0):
    return t.data_ptr() + off * t.element_size()


def ST():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = a.float().cpu(), b.float().cpu()
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs().max().item()
    assert err <= atol * max(scale, 1.0) + rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def nhwc(x):      # NCHW -> flat NHWC
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x, B, H, W, C):
    return x.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()


def wk(w):        # [Cout][Cin][3][3] -> [Cout][3][3][Cin]
    return w.permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------ Linear
@pytest.mark.parametrize("M,N,K_,relu,split", [(128, 1024, 102, 0, 0), (37, 100, 14112, 0, 1), (256, 1, 1024, 1, 0),
                                               (5, 4, 1024, 1, 0), (130, 14112, 102, 0, 0), (1, 1024, 100, 0, 0), (64, 4, 1024, 1, 2), (128, 1, 1024, 1, 2)])
def test_linear_fwd(M, N, K_, relu, split):
    x, w, b = rnd(M, K_, seed=1), rnd(N, K_, seed=2, scale=0.05), rnd(N, seed=3)
    y = torch.zeros(M, N, device=DEV)
    K.linear_fwd(P(x), K_, 0, P(w), 0, P(b), 0, P(y), N, 0, M, N, K_, relu, 1, split, ST())
    ref = F.linear(F.relu(x) if relu else x, w, b)
    close(y, ref, what="linear_fwd")


def test_linear_fwd_batched_heads():
    M, N, K_ = 64, 1024, 102
    x, w, b = rnd(M, K_, seed=1), rnd(2, N, K_, seed=2, scale=0.1), rnd(2, N, seed=3)
    y = torch.zeros(2, M, N, device=DEV)
    K.linear_fwd(P(x), K_, 0, P(w), N * K_, P(b), N, P(y), N, M * N, M, N, K_, 0, 2, 0, ST())
    for h in range(2):
        close(y[h], F.linear(x, w[h], b[h]), what=f"head{h}")


@pytest.mark.parametrize("M,N,K_,mode", [(128, 1024, 1024, 1), (64, 1, 1024, 2), (200, 100, 14112, 0), (33, 1024, 102, 0)])
def test_linear_dgrad(M, N, K_, mode):
    dy, w, z = rnd(M, N, seed=1), rnd(N, K_, seed=2, scale=0.05), rnd(M, K_, seed=3)
    dx = torch.zeros(M, K_, device=DEV)
    K.linear_dgrad(P(dy), N, 0, P(w), 0, P(z) if mode else 0, K_, 0, P(dx), K_, 0, M, N, K_, mode, 0, 1, ST())
    ref = dy @ w
    if mode == 1:
        ref = ref * (z > 0)
    if mode == 2:
        ref = F.relu(ref) * (z > 0)
    close(dx, ref, what="linear_dgrad")


@pytest.mark.parametrize("M,N,K_,relu", [(256, 1024, 1024, 1), (128, 100, 14112, 0), (64, 1, 1024, 1), (77, 1024, 102, 0), (128, 14112, 102, 0)])
def test_linear_wgrad(M, N, K_, relu):
    x, dy = rnd(M, K_, seed=1), rnd(M, N, seed=2)
    dw = torch.zeros(N, K_, device=DEV); db = torch.zeros(N, device=DEV)
    K.linear_wgrad(P(x), K_, 0, P(dy), N, 0, P(dw), 0, P(db), 0, M, N, K_, relu, 1, ST())
    xa = F.relu(x) if relu else x
    close(dw, dy.t() @ xa, rtol=2e-4, what="linear_wgrad")
    close(db, dy.sum(0), rtol=2e-4, what="linear_bgrad")


# ------------------------------------------------------------------ dense layers on tcgen05 (3xTF32 split precision)
def _err64(a, ref64):
    """max |a - ref| relative to the output scale, both against an fp64 reference."""
    return float((a.double() - ref64).abs().max() / (ref64.abs().max() + 1e-30))


@pytest.mark.parametrize("M,N,K_,relu,split,batch", [(128, 100, 14112, 0, 2, 1), (256, 100, 14112, 0, 2, 1), (128, 1024, 1024, 1, 2, 2),
                                                     (256, 1024, 1024, 1, 2, 2), (128, 1024, 100, 0, 0, 1), (77, 1024, 1024, 1, 0, 1),
                                                     (1, 1024, 1024, 1, 2, 1)])
def test_gemm_tc_fwd(M, N, K_, relu, split, batch):
    """sgqn_linear_fwd_tc vs fp64 F.linear: within 1e-5 of the output scale (plain TF32 is ~1e-3, the fp32 CUDA-core kernel ~1e-6)."""
    x, w, b = rnd(batch, M, K_, seed=1), rnd(batch, N, K_, seed=2, scale=0.05), rnd(batch, N, seed=3)
    y = torch.full((batch, M, N), 7.0, device=DEV) if split != 1 else torch.zeros(batch, M, N, device=DEV)
    y0 = torch.zeros(batch, M, N, device=DEV)
    K.linear_fwd_tc(P(x), K_, M * K_, P(w), N * K_, P(b), N, P(y), N, M * N, M, N, K_, relu, batch, split, ST())
    K.linear_fwd(P(x), K_, M * K_, P(w), N * K_, P(b), N, P(y0), N, M * N, M, N, K_, relu, batch, 2 if split else 0, ST())
    xa = F.relu(x) if relu else x
    ref = torch.einsum("bmk,bnk->bmn", xa.double(), w.double()) + b.double()[:, None, :]
    e_tc, e_simt = _err64(y, ref), _err64(y0, ref)
    # the tensor core adds into its fp32 accumulator with truncation, so a 1024-long chain inside one accumulator is ~1e-5;
    # split-K shortens the chains (the partial sums meet through fp32 red.adds)
    assert e_tc < max(4 * e_simt, 1e-5), (e_tc, e_simt)


@pytest.mark.parametrize("M,N,K_,mode,acc,batch", [(128, 1024, 1024, 1, 2, 2), (256, 1024, 1024, 1, 2, 2), (128, 1024, 1024, 2, 0, 1),
                                                   (200, 100, 14112, 0, 0, 1), (128, 1024, 100, 0, 2, 1), (128, 100, 14112, 0, 1, 1)])
def test_gemm_tc_dgrad(M, N, K_, mode, acc, batch):
    dy, w, z = rnd(batch, M, N, seed=1), rnd(batch, N, K_, seed=2, scale=0.05), rnd(batch, M, K_, seed=3)
    init = rnd(batch, M, K_, seed=4) if acc == 1 else torch.full((batch, M, K_), 3.0, device=DEV)
    dx = init.clone()
    K.linear_dgrad_tc(P(dy), N, M * N, P(w), N * K_, P(z) if mode else 0, K_, M * K_, P(dx), K_, M * K_, M, N, K_, mode, acc, batch, ST())
    ref = torch.einsum("bmn,bnk->bmk", dy.double(), w.double())
    if mode == 1:
        ref = ref * (z > 0)
    if mode == 2:
        ref = F.relu(ref) * (z > 0)
    if acc == 1:
        ref = ref + init.double()
    assert _err64(dx, ref) < 2e-6, _err64(dx, ref)


@pytest.mark.parametrize("M,N,K_,relu,batch", [(256, 1024, 1024, 1, 2), (128, 100, 14112, 0, 1), (256, 100, 14112, 0, 1), (128, 1024, 100, 0, 1),
                                               (77, 1024, 1024, 1, 1)])
def test_gemm_tc_wgrad(M, N, K_, relu, batch):
    x, dy = rnd(batch, M, K_, seed=1), rnd(batch, M, N, seed=2)
    dw0 = rnd(batch, N, K_, seed=5)
    dw = dw0.clone(); db = torch.zeros(batch, N, device=DEV)
    K.linear_wgrad_tc(P(x), K_, M * K_, P(dy), N, M * N, P(dw), N * K_, P(db), N, M, N, K_, relu, batch, ST())
    xa = F.relu(x) if relu else x
    ref = torch.einsum("bmn,bmk->bnk", dy.double(), xa.double()) + dw0.double()
    assert _err64(dw, ref) < 2e-6, _err64(dw, ref)
    close(db, dy.sum(1), rtol=2e-4, what="bias grad")


# ------------------------------------------------------------------ convs (NHWC)
CONVS = [  # B, Hs, Cin, Cout(real), Cout(stored), pad, up
    (3, 41, 32, 32, 32, 0, 1), (2, 23, 32, 32, 32, 0, 1), (2, 21, 32, 128, 128, 1, 1), (2, 21, 128, 64, 64, 1, 2),
    (2, 42, 64, 9, 16, 1, 2)]


@pytest.mark.parametrize("B,Hs,Cin,Co,Cs,pad,up", CONVS)
def test_conv_fwd_dgrad_wgrad(B, Hs, Cin, Co, Cs, pad, up):
    x = rnd(B, Cin, Hs, Hs, seed=1)
    w = rnd(Co, Cin, 3, 3, seed=2, scale=0.1); b = rnd(Co, seed=3)
    ws = torch.zeros(Cs, Cin, 3, 3, device=DEV); ws[:Co] = w
    bs = torch.zeros(Cs, device=DEV); bs[:Co] = b
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    xin = F.relu(xr)
    if up == 2:
        xin = F.interpolate(xin, scale_factor=2)
    xin.retain_grad()
    yr = F.conv2d(xin, wr, br, padding=pad)
    Ho = yr.shape[-1]
    y = torch.zeros(B * Ho * Ho * Cs, device=DEV)
    xh, wsk = nhwc(x), wk(ws)            # keep the NHWC copies alive while kernels read them
    K.conv_fwd(P(xh), P(wsk), P(bs), P(y), B, Hs, Hs, Cin, Cs, pad, up, 1, 0, ST())
    close(nchw(y, B, Ho, Ho, Cs)[:, :Co], yr, what="conv_fwd")
    dy = rnd(B, Co, Ho, Ho, seed=4)
    yr.backward(dy)
    dys = torch.zeros(B, Cs, Ho, Ho, device=DEV); dys[:, :Co] = dy
    dw = torch.zeros(Cs * 9 * Cin, device=DEV); db = torch.zeros(Cs, device=DEV)
    dyh = nhwc(dys)
    K.conv_wgrad(P(xh), P(dyh), P(dw), P(db), B, Hs, Hs, Cin, Cs, pad, up, 1, 0, ST())
    close(dw.reshape(Cs, 3, 3, Cin).permute(0, 3, 1, 2)[:Co], wr.grad, rtol=3e-4, what="conv_wgrad")
    close(db[:Co], br.grad, rtol=3e-4, what="conv_bgrad")
    Hl = Hs * up
    dxl = torch.zeros(B * Hl * Hl * Cin, device=DEV)
    K.conv_dgrad(P(dyh), P(wsk), 0, P(dxl), B, Hl, Hl, Cin, Cs, pad, 0, ST())
    close(nchw(dxl, B, Hl, Hl, Cin), xin.grad, rtol=3e-4, what="conv_dgrad(logical input)")
    if up == 1:
        for mode in (1, 2):
            K.conv_dgrad(P(dyh), P(wsk), P(xh), P(dxl), B, Hl, Hl, Cin, Cs, pad, mode, ST())
            ref = xin.grad * (x > 0) if mode == 1 else F.relu(xin.grad) * (x > 0)
            close(nchw(dxl, B, Hl, Hl, Cin), ref, rtol=3e-4, what=f"conv_dgrad mode {mode}")
    else:
        dx = torch.zeros(B * Hs * Hs * Cin, device=DEV)
        K.upsample2_bwd(P(dxl), P(xh), P(dx), B, Hs, Hs, Cin, ST())
        close(nchw(dx, B, Hs, Hs, Cin), xr.grad, rtol=3e-4, what="upsample2_bwd")


@pytest.mark.parametrize("B,Hin", [(3, 84), (2, 100)])
def test_conv1(B, Hin):
    g = torch.Generator().manual_seed(5)
    obs = torch.randint(0, 256, (B, 9, Hin, Hin), generator=g).float().to(DEV)
    w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
    c = (Hin - 84) // 2
    xr = obs[:, :, c:c + 84, c:c + 84].clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = F.conv2d(xr / 255.0, wr, br, stride=2)
    y = torch.zeros(B * 41 * 41 * 32, device=DEV)
    K.conv1_fwd(P(obs), P(w), P(b), P(y), B, Hin, 9, 32, 0, ST())
    close(nchw(y, B, 41, 41, 32), yr, what="conv1_fwd")
    dy = rnd(B, 32, 41, 41, seed=4)
    yr.backward(dy)
    dw = torch.zeros(32 * 81, device=DEV); db = torch.zeros(32, device=DEV)
    dyh = nhwc(dy)
    K.conv1_wgrad(P(obs), P(dyh), P(dw), P(db), B, Hin, 9, 32, ST())
    close(dw.reshape(32, 9, 3, 3), wr.grad, rtol=3e-4, what="conv1_wgrad")
    close(db, br.grad, rtol=3e-4, what="conv1_bgrad")
    if Hin == 84:
        dobs = torch.zeros(B, 9, 84, 84, device=DEV)
        K.conv1_dgrad(P(dyh), P(w), P(dobs), B, 9, 32, ST())
        close(dobs, xr.grad, rtol=3e-4, what="conv1_dgrad")
        assert float(dobs[:, :, 83].abs().max()) == 0.0 and float(dobs[:, :, :, 83].abs().max()) == 0.0


# ------------------------------------------------------------------ heads / losses
def test_ln_tanh_fwd_bwd():
    M, Pd = 67, 100
    z = rnd(M, Pd, seed=1, scale=2.0); gam = rnd(Pd, seed=2) + 1; bet = rnd(Pd, seed=3)
    zr, gr, br = z.clone().requires_grad_(True), gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    hr = torch.tanh(F.layer_norm(zr, (Pd,), gr, br, 1e-5))
    h = torch.zeros(M, Pd + 2, device=DEV)
    K.ln_tanh_fwd(P(z), P(gam), P(bet), P(h), Pd + 2, M, Pd, ST())
    close(h[:, :Pd], hr, what="ln_tanh_fwd")
    dh = rnd(M, Pd + 2, seed=4)
    hr.backward(dh[:, :Pd])
    dz = torch.zeros(M, Pd, device=DEV); dg = torch.zeros(Pd, device=DEV); dbt = torch.zeros(Pd, device=DEV)
    K.ln_tanh_bwd(P(dh), Pd + 2, P(z), P(h), Pd + 2, P(gam), P(dz), P(dg), P(dbt), M, Pd, ST())
    close(dz, zr.grad, rtol=3e-4, what="ln dz"); close(dg, gr.grad, rtol=3e-4, what="ln dgamma"); close(dbt, br.grad, rtol=3e-4, what="ln dbeta")


def _actor_head_ref(raw, noise, lmin, lmax):
    A = noise.shape[1]
    mu, ls = raw.chunk(2, dim=-1)
    ls = torch.tanh(ls); ls = lmin + 0.5 * (lmax - lmin) * (ls + 1)
    pi = mu + noise * ls.exp()
    logp = (-0.5 * noise.pow(2) - ls).sum(-1, keepdim=True) - 0.5 * np.log(2 * np.pi) * A
    mu_t, pi_t = torch.tanh(mu), torch.tanh(pi)
    logp = logp - torch.log(F.relu(1 - pi_t.pow(2)) + 1e-6).sum(-1, keepdim=True)
    return mu_t, pi_t, logp, ls


@pytest.mark.parametrize("A", [1, 2, 6])
def test_actor_head_fwd_bwd(A):
    M = 50
    raw, noise = rnd(M, 2 * A, seed=1), rnd(M, A, seed=2)
    rr = raw.clone().requires_grad_(True)
    mu_r, pi_r, lp_r, ls_r = _actor_head_ref(rr, noise, -10.0, 2.0)
    mu = torch.zeros(M, A, device=DEV); pi = torch.zeros(M, A + 3, device=DEV); lp = torch.zeros(M, device=DEV); ls = torch.zeros(M, A, device=DEV)
    K.actor_head_fwd(P(raw), P(noise), -10.0, 2.0, P(mu), P(pi), A + 3, P(lp), P(ls), M, A, ST())
    close(mu, mu_r, what="mu"); close(pi[:, :A], pi_r, what="pi"); close(lp, lp_r[:, 0], rtol=2e-4, what="log_pi"); close(ls, ls_r, what="log_std")
    log_alpha = torch.tensor([np.log(0.3)], dtype=torch.float64, device=DEV)
    dpi = rnd(M, A + 3, seed=3)
    loss = (pi_r * dpi[:, :A]).sum() + (0.3 / M) * lp_r.sum()
    loss.backward()
    draw = torch.zeros(M, 2 * A, device=DEV)
    K.actor_head_bwd(P(raw), P(noise), P(dpi), A + 3, P(log_alpha), -10.0, 2.0, P(draw), M, A, M, ST())
    close(draw, rr.grad, rtol=5e-4, what="actor_head_bwd")
    # a data-parallel shard: the entropy term is scaled by 1 / global batch (here 3 M), like the Q term that arrives in dpi
    rr2 = raw.clone().requires_grad_(True)
    _, pi2, lp2, _ = _actor_head_ref(rr2, noise, -10.0, 2.0)
    ((pi2 * dpi[:, :A]).sum() + (0.3 / (3 * M)) * lp2.sum()).backward()
    K.actor_head_bwd(P(raw), P(noise), P(dpi), A + 3, P(log_alpha), -10.0, 2.0, P(draw), M, A, 3 * M, ST())
    close(draw, rr2.grad, rtol=5e-4, what="actor_head_bwd (global batch 3M)")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_critic_loss(mode):
    B = 37
    R = B if mode == 0 else 2 * B
    q = rnd(2, 2 * B, seed=1); tq = rnd(2, B, seed=2); nlp = rnd(B, seed=3); r = rnd(B, seed=4)
    nd = (rnd(B, seed=5) > -1).float(); la = torch.tensor([np.log(0.2)], dtype=torch.float64, device=DEV)
    qr = q.clone().requires_grad_(True)
    t = r + nd * 0.99 * (torch.min(tq[0], tq[1]) - 0.2 * nlp)
    if mode == 0:
        ref = F.mse_loss(qr[0, :B], t) + F.mse_loss(qr[1, :B], t)
    elif mode == 1:
        ref = F.mse_loss(qr[0, :B], t) + F.mse_loss(qr[1, :B], t) + 0.5 * (F.mse_loss(qr[0, :B], qr[0, B:]) + F.mse_loss(qr[1, :B], qr[1, B:]))
    else:
        t2 = torch.cat([t, t])
        ref = (0.5 + 0.5) * (F.mse_loss(qr[0], t2) + F.mse_loss(qr[1], t2))
    ref.backward()
    dq = torch.zeros(2, 2 * B, device=DEV); loss = torch.zeros(1, device=DEV); tqo = torch.zeros(B, device=DEV)
    K.critic_loss(P(q), 2 * B, P(tq), P(tq, B), P(nlp), P(r), P(nd), P(la), 0.99, mode, 0.5, 0.5, P(tqo), P(dq), P(loss), B, B, ST())
    close(loss, ref.detach().reshape(1), what="critic loss"); close(tqo, t, what="target_q")
    close(dq[:, :R], qr.grad[:, :R], rtol=2e-4, what="dq")


def test_actor_loss():
    B = 41
    q = rnd(2, 2 * B, seed=1); lp = rnd(B, seed=2); la = torch.tensor([np.log(0.15)], dtype=torch.float64, device=DEV)
    qr = q.clone().requires_grad_(True); lar = la.clone().requires_grad_(True)
    actor = (0.15 * lp - torch.min(qr[0, :B], qr[1, :B])).mean()
    alpha_loss = (lar.exp() * (-lp - (-2.0))).mean()
    actor.backward(); alpha_loss.backward()
    dq = torch.zeros(2, 2 * B, device=DEV); out = torch.zeros(3, device=DEV); ag = torch.zeros(1, dtype=torch.float64, device=DEV)
    K.actor_loss(P(q), 2 * B, P(lp), P(la), -2.0, P(dq), P(out), P(ag), B, B, ST())
    close(out[0:1], actor.detach().reshape(1), what="actor loss"); close(out[1:2], alpha_loss.detach().float().reshape(1), what="alpha loss")
    close(out[2:3], torch.tensor([0.15]), what="alpha"); close(dq[:, :B], qr.grad[:, :B], what="dq")
    close(ag, lar.grad, what="alpha grad")


def test_bce():
    B, HW, Cs = 3, 84 * 84, 16
    lg = rnd(B, HW, Cs, seed=1, scale=2.0)
    mask = (rnd(B, 3, HW, seed=2) > 0.8).to(torch.uint8)
    x = lg[:, :, :9].permute(0, 2, 1).clone().requires_grad_(True)          # (B,9,HW)
    y = mask.float().repeat_interleave(3, dim=1)
    ref = F.binary_cross_entropy_with_logits(x, y)
    ref.backward()
    loss = torch.zeros(1, device=DEV); d = torch.ones(B, HW, Cs, device=DEV)
    d[:, :, 12:] = 0            # padding channels >= 12 are the caller's to keep zero (the kernel writes channels 0..11 only)
    K.bce(P(lg), P(mask), P(loss), P(d), B, 84, 84, 84, 84, 0, 0, Cs, B, 0, ST())
    close(loss, ref.detach().reshape(1), what="bce loss")
    close(d[:, :, :9].permute(0, 2, 1), x.grad, rtol=2e-4, what="bce grad")
    assert float(d[:, :, 9:].abs().max()) == 0.0


# ------------------------------------------------------------------ saliency / augmentation / replay: bit-exact
def _mask_gpu(grad, q, obs=None, mm=None, u=None, mm_neg=0):
    B = grad.shape[0]
    mask = torch.zeros(B, 3, 84 * 84, dtype=torch.uint8, device=DEV)
    masked = torch.zeros(B, 9, 84, 84, device=DEV) if obs is not None else None
    K.attribution_mask(P(grad), P(obs) if obs is not None else 0, P(mm) if mm is not None else 0, P(u) if u is not None else 0,
                       float(q), P(mask), P(masked) if masked is not None else 0, B, 84 * 84, mm_neg, ST())
    full = mask.reshape(B, 3, 1, 84, 84).expand(B, 3, 3, 84, 84).reshape(B, 9, 84, 84).bool()
    return full, masked


def test_attribution_mask_golden_bit_exact():
    gold = np.load(os.path.join(GOLDEN, "masks.npz"))
    g = torch.Generator().manual_seed(77)
    base = torch.randn(4, 9, 84, 84, generator=g)
    base[1, :3] = 0.0
    base[2, 3:6] = (torch.rand(3, 84, 84, generator=g) < 0.03).float() * base[2, 3:6]
    base[3, 6:9] = torch.round(base[3, 6:9] * 2) / 2
    for q in (0.5, 0.9, 0.95, 0.98, 0.999):
        m, _ = _mask_gpu(base.to(DEV), q)
        assert np.array_equal(np.packbits(m.cpu().numpy().reshape(-1)), gold[f"q{q}"]), q


@pytest.mark.parametrize("q", [0.0, 0.5, 0.95, 0.98, 1.0])
def test_attribution_mask_vs_oracle_and_fill(q):
    from oracle import sgsac_oracle as O
    g = torch.Generator().manual_seed(3)
    grad = torch.randn(6, 9, 84, 84, generator=g) * 1e-3
    grad[0, :3] = 0; grad[1, 3:6, :80] = 0; grad[2] = torch.round(grad[2] * 4000) / 4000
    obs = torch.randint(0, 256, (6, 9, 84, 84), generator=g).float()
    ref = O.compute_attribution_mask(grad, q)
    mm = torch.zeros(4, device=DEV); scratch = torch.zeros(1024, device=DEV)
    obs_d = obs.to(DEV)
    K.minmax(P(obs_d), obs.numel(), P(scratch), P(mm), ST())
    assert mm.cpu().tolist() == [float(obs.min()), float(obs.max()), -float(obs.min()), float(obs.max())]
    u = torch.tensor([0.37], device=DEV)
    m, masked = _mask_gpu(grad.to(DEV), q, obs_d, mm, u)
    m2, masked2 = _mask_gpu(grad.to(DEV), q, obs_d, mm[2:], u, mm_neg=1)        # the all-reduce exchange form {-min, max}
    assert torch.equal(m2, m) and torch.equal(masked2, masked)
    assert torch.equal(m.cpu(), ref)
    fill = obs.min() + (obs.max() - obs.min()) * 0.37
    ref_masked = obs * ref
    ref_masked[ref < 1] = fill
    assert torch.equal(masked.cpu(), ref_masked)


def test_overlay_and_crop_shift_golden():
    gold = np.load(os.path.join(GOLDEN, "aug.npz"))
    rs = np.random.RandomState(11)
    x100 = torch.as_tensor(rs.randint(0, 256, size=(3, 9, 100, 100)).astype(np.float32)).to(DEV)
    w1 = rs.randint(0, 16, size=3); h1 = rs.randint(0, 16, size=3)
    offs = torch.as_tensor(np.stack([w1, h1], 1), dtype=torch.int32).to(DEV).contiguous()
    y = torch.zeros(3, 9, 84, 84, device=DEV)
    K.crop_shift(P(x100), P(offs), P(y), 3, 9, 100, 84, 0, 0, ST())
    assert np.array_equal(y.cpu().numpy().astype(np.uint8), gold["crop"])
    x84 = torch.as_tensor(rs.randint(0, 256, size=(3, 9, 84, 84)).astype(np.float32)).to(DEV)
    dy = rs.randint(0, 9, size=3); dx = rs.randint(0, 9, size=3)
    offs = torch.as_tensor(np.stack([dy, dx], 1), dtype=torch.int32).to(DEV).contiguous()
    K.crop_shift(P(x84), P(offs), P(y), 3, 9, 84, 84, 1, 4, ST())
    assert np.array_equal(y.cpu().numpy().astype(np.uint8), gold["shift"])
    pool = torch.as_tensor(rs.randint(0, 256, size=(8, 3, 84, 84), dtype=np.uint8)).to(DEV)
    ids = torch.as_tensor(rs.randint(0, 8, size=3), dtype=torch.int64).to(DEV)
    K.overlay_u8(P(x84), P(pool), P(ids), float(np.float32(0.8)), float(np.float32(0.2)), P(y), 3, 84 * 84, ST())
    np.testing.assert_allclose(y.cpu().numpy(), gold["overlay"], rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("size,mode", [(84, 0), (84, 1), (100, 0)])
def test_replay_gather_bit_exact(size, mode):
    from oracle import sgsac_oracle as O
    import sgqn_carla_b200 as S
    cap, B = 64, 16
    rep = O.synthetic_replay(cap, 2, size=size, seed=3)
    rb = S.ReplayBuffer((9, size, size), (2,), cap, B)
    rb.load_ring(rep.frames, rep.actions, rep.rewards, rep.not_dones)
    rs = np.random.RandomState(0)
    idxs = rs.randint(0, cap, size=B)
    if mode == 1:
        offs = rs.randint(0, 9, size=(2, B, 2))
        ref = rep.sample_drq(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1]))
        got = rb.sample_drq(idxs=idxs, offs=offs)
    elif size == 100:
        offs = rs.randint(0, 16, size=(2, B, 2))
        ref = rep.sample(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1]))
        got = rb.sample(idxs=idxs, offs=offs)
    else:
        ref = rep.sample(idxs)
        got = rb.sample(idxs=idxs)
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)


def test_replay_add_dedups_lazyframes_and_matches_oracle_stacks():
    import sgqn_carla_b200 as S
    rs = np.random.RandomState(1)
    frames = [rs.randint(0, 256, size=(3, 84, 84), dtype=np.uint8) for _ in range(40)]
    rb = S.ReplayBuffer((9, 84, 84), (2,), 16, 4, frame_capacity=64)
    for i in range(30):          # wraps the 16-transition ring
        rb.add(S.LazyFrames(frames[i:i + 3]), rs.rand(2), float(i), S.LazyFrames(frames[i + 1:i + 4]), False)
    assert rb._next_frame <= 30 + 3 + 1 or rb.F == 64       # ~1 new frame per transition, not 6
    idxs = np.array([0, 5, 13, 15])
    obs, a, r, nxt, nd = rb.sample(idxs=idxs)
    for j, i in enumerate(idxs):
        t = 16 + i if i < 14 else i                           # slot i holds transition 16+i after the wrap (i<14)
        assert np.array_equal(obs[j].cpu().numpy().astype(np.uint8), np.concatenate(frames[t:t + 3]))
        assert np.array_equal(nxt[j].cpu().numpy().astype(np.uint8), np.concatenate(frames[t + 1:t + 4]))
        assert float(r[j]) == float(t)


@pytest.mark.parametrize("storage", ["device", "pinned"])
def test_replay_add_ndarray_stacks_dedup_and_ring_growth(storage):
    """The reference's `add` also takes plain ndarray stacks (utils.py:111-122).  A FrameStack rollout handed over as
    ndarrays still costs ~1 frame per transition (frames found by content in the previous add); unrelated stacks cost 6
    and make the ring GROW instead of overwriting live frames; through the whole capacity every transition reads back."""
    import sgqn_carla_b200 as S
    rs = np.random.RandomState(2)
    cap = 24
    frames = [rs.randint(0, 256, size=(3, 84, 84), dtype=np.uint8) for _ in range(3 * cap + 8)]
    stack = lambda i: np.concatenate(frames[i:i + 3])
    rb = S.ReplayBuffer((9, 84, 84), (2,), cap, 4, storage=storage)
    assert rb.F == 2 * cap + 8
    for i in range(2 * cap + 5):      # rollout: obs_t = next_obs_{t-1}, next_obs_t = obs_t shifted by one frame; wraps twice
        rb.add(stack(i), rs.rand(2), float(i), stack(i + 1), i % 7 == 6)
    assert rb.version == 0 and rb._next_frame <= 2 * cap + 5 + 3       # ~1 new frame per add: the default ring never filled
    idxs = np.arange(cap)
    obs, a, r, nxt, nd = rb.sample(idxs=idxs)
    for j in idxs:
        t = int(r[j])
        assert t >= cap + 5 and t % cap == j
        assert np.array_equal(obs[j].cpu().numpy().astype(np.uint8), stack(t))
        assert np.array_equal(nxt[j].cpu().numpy().astype(np.uint8), stack(t + 1))
        assert float(nd[j]) == float(t % 7 != 6)
    # unrelated stacks: 6 fresh frames per add -> the ring must grow (the old default raised MemoryError at 67 % of capacity)
    rb2 = S.ReplayBuffer((9, 84, 84), (2,), cap, 4, storage=storage)
    stacks = [(rs.randint(0, 256, size=(9, 84, 84), dtype=np.uint8), rs.randint(0, 256, size=(9, 84, 84), dtype=np.uint8)) for _ in range(cap + 6)]
    for i, (o, n) in enumerate(stacks):
        rb2.add(o, rs.rand(2), float(i), n, False)
    assert rb2.version >= 1 and rb2.F >= 6 * cap
    obs, a, r, nxt, nd = rb2.sample(idxs=idxs)
    for j in idxs:
        t = int(r[j])
        assert np.array_equal(obs[j].cpu().numpy().astype(np.uint8), stacks[t][0])
        assert np.array_equal(nxt[j].cpu().numpy().astype(np.uint8), stacks[t][1])


# ------------------------------------------------------------------ optimiser
def test_adam_and_ema_vs_oracle():
    from oracle import sgsac_oracle as O
    n = 4096 + 8
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g); tgt0 = torch.randn(n, generator=g)
    p = {"w": p0.clone()}
    opt = O.Adam(["w"], 1e-3, 0.9)
    pd, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    tgt = tgt0.clone().to(DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV); bc = torch.zeros(2, device=DEV)
    tref = tgt0.clone()
    for it in range(5):
        gr = torch.randn(n, generator=g) * (10.0 ** (it - 3))
        opt.step(p, {"w": gr})
        tref[:1024] = 0.01 * p["w"][:1024] + (1 - 0.01) * tref[:1024]
        tref[1024:] = 0.05 * p["w"][1024:] + (1 - 0.05) * tref[1024:]
        gd = gr.to(DEV)
        K.adam_prep(P(step), P(bc), 0.9, 0.999, ST())
        K.adam(P(pd), P(gd), P(m), P(v), n, P(bc), 1e-3, float(np.float32(0.1)), 0.999, float(np.float32(0.001)), 1e-8,
               P(tgt), 1024, 0.01, 0.05, 0.0, ST())
    assert int(step) == 5
    np.testing.assert_allclose(pd.cpu().numpy(), p["w"].numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(m.cpu().numpy(), opt.m["w"].numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), opt.v["w"].numpy(), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(tgt.cpu().numpy(), tref.numpy(), rtol=0, atol=2e-6)


def test_adam_weight_decay_matches_torch():
    """critic_weight_decay (sac.py:63-65): torch.optim.Adam's L2 form."""
    n = 1024
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01)
    pd, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV); bc = torch.zeros(2, device=DEV)
    for it in range(4):
        gr = torch.randn(n, generator=g)
        ref.grad = gr.clone(); opt.step()
        K.adam_prep(P(step), P(bc), 0.9, 0.999, ST())
        K.adam(P(pd), P(gr.to(DEV)), P(m), P(v), n, P(bc), 1e-3, float(np.float32(0.1)), 0.999, float(np.float32(0.001)), 1e-8,
               0, 0, 0.0, 0.0, 0.01, ST())
    np.testing.assert_allclose(pd.cpu().numpy(), ref.detach().numpy(), rtol=0, atol=2e-6)


def test_alpha_adam_fp64():
    la = torch.tensor([np.log(0.1)], dtype=torch.float64, device=DEV); st = torch.zeros(2, dtype=torch.float64, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    ref = torch.tensor(np.log(0.1), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.5, 0.999))
    for gval in (0.3, -1.2, 0.05):
        ref.grad = torch.tensor(gval, dtype=torch.float64)
        opt.step()
        gd = torch.tensor([gval], dtype=torch.float64, device=DEV)
        K.alpha_adam(P(la), P(gd), P(st), P(step), 1e-4, 0.5, 0.999, 1e-8, ST())
    assert abs(float(la) - float(ref)) < 1e-12


def test_rng_step_statistics():
    B, A = 512, 2
    ctr = torch.zeros(1, dtype=torch.int64, device=DEV); nv = torch.tensor([1000], dtype=torch.int32, device=DEV)
    idxs = torch.zeros(B, dtype=torch.int64, device=DEV); ov = torch.zeros(B, dtype=torch.int64, device=DEV)
    offs = torch.zeros(2, B, 2, dtype=torch.int32, device=DEV)
    n1 = torch.zeros(B, A, device=DEV); n2 = torch.zeros(B, A, device=DEV); u = torch.zeros(1, device=DEV)
    K.rng_step(7, P(ctr), P(nv), P(idxs), P(ov), 256, P(offs), 9, P(n1), P(n2), P(u), B, A, 7, ST())
    first = idxs.clone()
    K.rng_step(7, P(ctr), P(nv), P(idxs), P(ov), 256, P(offs), 9, P(n1), P(n2), P(u), B, A, 7, ST())
    assert int(ctr) == 2 and not torch.equal(first, idxs)
    # two data-parallel ranks: own seed (indices / noise differ), shared seed_u and counter (ONE fill scalar per global batch)
    c0 = torch.zeros(1, dtype=torch.int64, device=DEV); c1 = torch.zeros(1, dtype=torch.int64, device=DEV)
    i0, i1, u0, u1 = idxs.clone(), idxs.clone(), u.clone(), u.clone()
    K.rng_step(100, P(c0), P(nv), P(i0), P(ov), 256, P(offs), 9, P(n1), P(n2), P(u0), B, A, 55, ST())
    K.rng_step(101, P(c1), P(nv), P(i1), P(ov), 256, P(offs), 9, P(n1), P(n2), P(u1), B, A, 55, ST())
    assert float(u0) == float(u1) and not torch.equal(i0, i1)
    assert 0 <= int(idxs.min()) and int(idxs.max()) < 1000 and int(ov.max()) < 256 and int(offs.max()) <= 8 and int(offs.min()) >= 0
    z = torch.cat([n1.flatten(), n2.flatten()])
    assert abs(float(z.mean())) < 0.15 and abs(float(z.std()) - 1.0) < 0.1 and 0.0 <= float(u) < 1.0


# ------------------------------------------------------------------ tcgen05 TF32 conv (32 -> 32)
def tf32_round(x):
    """cvt.rna.tf32.f32 emulation: round-to-nearest (ties away) onto a 10-bit mantissa."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def rows_pad(x_nchw, extra):
    """NCHW -> NHWC with `extra` zero rows appended to every sample ([B][H+extra][W][C], the tcgen05 conv layout)."""
    B, C, H, W = x_nchw.shape
    out = torch.zeros(B, H + extra, W, C, device=DEV)
    out[:, :H] = x_nchw.permute(0, 2, 3, 1)
    return out


def prep_w(w):
    wk_ = wk(w).reshape(1, -1).contiguous()
    wf = torch.zeros(9216, device=DEV); wd = torch.zeros(9216, device=DEV)
    K.conv_weights_prep(P(wk_), 9216, P(wf), P(wd), 1, ST())
    return wf, wd


@pytest.mark.parametrize("B,Hp", [(2, 23), (3, 41), (5, 25), (16, 39)])
def test_conv_tc_forward(B, Hp):
    x = tf32_round(F.relu(rnd(B, 32, Hp, Hp, seed=1)))
    w = rnd(32, 32, 3, 3, seed=2, scale=0.1); b = rnd(32, seed=3)
    wf, wd = prep_w(w)
    assert torch.equal(wf.reshape(32, 3, 3, 32).permute(0, 3, 1, 2), tf32_round(w))
    assert torch.equal(wd.reshape(32, 3, 3, 32).permute(3, 0, 1, 2), tf32_round(w).flip(2, 3))      # wd[ci][t'][co] = w[co][ci][8-t']
    ref = F.conv2d(x.double(), tf32_round(w).double(), b.double())
    xh = rows_pad(x, 2)                                   # [B][Hp+2][Hp][32]
    Ho = Hp - 2
    for flags in (0, 1, 3):
        y = torch.full((B, Ho + 2, Ho, 32), 7.0, device=DEV)          # output with 2 extra rows per sample, never written
        K.conv_tc(P(xh), P(wf), P(b), 0, P(y), 0, B, Hp + 2, Hp, Ho, Ho, 0, Ho + 2, Ho, 0, 0, 0, 0, flags, ST())
        torch.cuda.synchronize()
        r = ref.clone()
        if flags & 1:
            r = F.relu(r)
        got = y[:, :Ho].permute(0, 3, 1, 2)
        assert float((y[:, Ho:] - 7.0).abs().max()) == 0.0
        if flags & 2:
            assert torch.equal(got, tf32_round(got))
            close(got, r.float(), rtol=1e-3, what=f"conv_tc fwd flags={flags}")
        else:
            close(got, r.float(), rtol=2e-5, what=f"conv_tc fwd flags={flags}")


@pytest.mark.parametrize("B,Hl,mode", [(2, 23, 1), (3, 41, 2), (4, 25, 0), (16, 39, 1)])
def test_conv_tc_dgrad_and_wgrad(B, Hl, mode):
    """Data gradient = 3x3 window sums over the zero-bordered dY buffer [B][Ho+4][Ho+2] (2 zero rows above/below, 2 zero
    columns after each row doubling as the next row's left border) with flipped/transposed weights and the ReLU-mask
    epilogue, written into the next layer's buffer of the same kind; weight gradient = tcgen05 reduction over pixels of
    dY x shifted activations, both in the shared [B][Hl+2][Hl] geometry."""
    Ho = Hl - 2
    w = rnd(32, 32, 3, 3, seed=2, scale=0.1)
    dy = tf32_round(rnd(B, 32, Ho, Ho, seed=4))
    act = tf32_round(F.relu(rnd(B, 32, Hl, Hl, seed=5)))
    wf, wd = prep_w(w)
    dyp = torch.zeros(B, Ho + 4, Ho + 2, 32, device=DEV)
    dyh = nhwc(dy)
    K.pad_copy(P(dyh), P(dyp), B, Ho, Ho, 32, Ho + 4, Ho + 2, 2, 0, 1, ST())
    assert torch.equal(dyp[:, 2:2 + Ho, :Ho].permute(0, 3, 1, 2), dy)
    out = torch.zeros(B, Hl + 4, Hl + 2, 32, device=DEV)
    acth = rows_pad(act, 2)                               # [B][Hl+2][Hl][32]
    dbs = torch.zeros(32, device=DEV)
    K.conv_tc(P(dyp), P(wd), 0, P(acth) if mode else 0, P(out), P(dbs), B, Ho + 4, Ho + 2, Hl, Hl, -2, Hl + 4, Hl + 2, 2, 0, Hl + 2, Hl,
              (mode << 2), ST())
    torch.cuda.synchronize()
    close(dbs, out.sum((0, 1, 2)), rtol=1e-4, what="fused bias gradient")
    ref = F.conv_transpose2d(dy.double(), tf32_round(w).double())
    if mode == 1:
        ref = ref * (act > 0)
    if mode == 2:
        ref = F.relu(ref) * (act > 0)
    close(out[:, 2:2 + Hl, :Hl].permute(0, 3, 1, 2), ref.float(), rtol=2e-5, what="conv_tc dgrad")
    border = out.clone(); border[:, 2:2 + Hl, :Hl] = 0
    assert float(border.abs().max()) == 0.0
    # weight gradient
    dw = torch.zeros(9216, device=DEV)
    K.conv_wgrad_tc(P(acth), P(dyp), P(dw), B, Hl + 2, Hl, ST())
    torch.cuda.synchronize()
    wr = w.clone().double().requires_grad_(True)
    F.conv2d(act.double(), wr).backward(dy.double())
    close(dw.reshape(32, 3, 3, 32).permute(0, 3, 1, 2), wr.grad.float(), rtol=2e-5, what="conv_wgrad_tc")
    db = torch.zeros(32, device=DEV)
    K.colsum(P(dyp), 32, B * (Ho + 4) * (Ho + 2), 32, P(db), ST())
    close(db, dy.sum((0, 2, 3)), rtol=1e-5, what="bias grad over the bordered buffer")


def test_conv1_fwd_pitched_output():
    """conv1 writing post-ReLU, TF32-rounded activations with 2 spare rows per sample (what the tcgen05 layers read)."""
    B = 3
    g = torch.Generator().manual_seed(5)
    obs = torch.randint(0, 256, (B, 9, 84, 84), generator=g).float().to(DEV)
    w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
    ref = F.relu(F.conv2d(obs / 255.0, w, b, stride=2))
    y = torch.full((B, 43, 41, 32), -1.0, device=DEV)
    K.conv1_fwd(P(obs), P(w), P(b), P(y), B, 84, 9, 32, 7, ST())
    got = y[:, :41].permute(0, 3, 1, 2)
    close(got, ref, rtol=1e-3, what="conv1 pitched")
    assert torch.equal(got, tf32_round(got)) and float((y[:, 41:] + 1.0).abs().max()) == 0.0


def test_conv1_via_im2col_matches_direct_kernels():
    B = 3
    g = torch.Generator().manual_seed(5)
    obs = torch.randint(0, 256, (B, 9, 100, 100), generator=g).float().to(DEV)
    w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
    col = torch.full((B * 1681, 84), 5.0, device=DEV)
    K.conv1_im2col(P(obs), P(col), B, 100, ST())
    assert float(col[:, 81:].abs().max()) == 0.0
    y0 = torch.zeros(B * 1681 * 32, device=DEV); y1 = torch.zeros(B * 1681 * 32, device=DEV)
    K.conv1_fwd(P(obs), P(w), P(b), P(y0), B, 100, 9, 32, 0, ST())
    K.conv1_fwd_col(P(col), P(w), P(b), P(y1), B, 0, ST())
    close(y1, y0, rtol=1e-6, what="conv1 fwd via col")
    dy = rnd(B * 1681, 32, seed=4)
    dw0 = torch.zeros(32 * 81, device=DEV); db0 = torch.zeros(32, device=DEV); dw1 = torch.zeros(32 * 81, device=DEV); db1 = torch.zeros(32, device=DEV)
    K.conv1_wgrad(P(obs), P(dy), P(dw0), P(db0), B, 100, 9, 32, ST())
    K.conv1_wgrad_col(P(col), P(dy), P(dw1), P(db1), B, ST())
    close(dw1, dw0, rtol=1e-5, what="conv1 wgrad via col"); close(db1, db0, rtol=1e-5, what="conv1 bgrad via col")
    obs84 = obs[:, :, 8:92, 8:92].contiguous()
    d0 = torch.zeros(B, 9, 84, 84, device=DEV); d1 = torch.ones(B, 9, 84, 84, device=DEV); dcol = torch.zeros(B * 1681, 84, device=DEV)
    K.conv1_dgrad(P(dy), P(w), P(d0), B, 9, 32, ST())
    K.conv1_dgrad_col(P(dy), P(w), P(dcol), P(d1), B, ST())
    close(d1, d0, rtol=1e-5, what="conv1 dgrad via col")
    # tcgen05 variant: dcol[pix][96] = tf32(dy)[pix][32] * tf32(W) through the per-position GEMM, then the staged gather
    wp = torch.zeros(32 * 96, device=DEV); wd = torch.zeros(96 * 32, device=DEV)
    K.conv1_weights_prep(P(w), P(wp), P(wd), ST())
    dyr = tf32_round(dy)
    dcol96 = torch.zeros(B * 1681, 96, device=DEV); d2 = torch.ones(B, 9, 84, 84, device=DEV); d3 = torch.zeros(B, 9, 84, 84, device=DEV)
    K.conv_tcg_taps(P(dyr), P(wd), 0, 0, P(dcol96), B, 41, 41, 32, 96, 41, 41, 0, 41, 41, 0, 0, 0, 0, 0, 1, ST())
    K.conv1_col2im(P(dcol96), 96, P(d2), B, ST())
    wr = tf32_round(w.reshape(32, 81)).reshape(w.shape)
    K.conv1_dgrad(P(dyr), P(wr), P(d3), B, 9, 32, ST())
    close(d2, d3, rtol=2e-5, what="conv1 dgrad on tcgen05")
    # the staged gather alone against the element-wise one (same terms, different summation order)
    d4 = torch.ones(B, 9, 84, 84, device=DEV)
    K.conv1_col2im(P(dcol), 84, P(d4), B, ST())
    close(d4, d1, rtol=1e-6, what="staged col2im")


# ------------------------------------------------------------------ generalised tcgen05 convs (decoder shapes)
def bordered(x_nchw, Hq, Wq, oy, ox):
    """NCHW -> zero-initialised NHWC buffer [B][Hq][Wq][C] with the image at rows [oy, oy+H), cols [ox, ox+W)."""
    B, C, H, W = x_nchw.shape
    out = torch.zeros(B, Hq, Wq, C, device=DEV)
    out[:, oy:oy + H, ox:ox + W] = x_nchw.permute(0, 2, 3, 1)
    return out


def prep_wg(w, cs):
    """w (Cout_real, Cin, 3, 3) -> stored [cs][9][Cin] + TF32 operand copies wf [cs][9][Cin], wd [Cin][9][cs]."""
    co, ci = w.shape[:2]
    ws = torch.zeros(cs, ci, 3, 3, device=DEV); ws[:co] = w
    wk_ = wk(ws).contiguous()
    wf = torch.zeros(cs * 9 * ci, device=DEV); wd = torch.zeros(cs * 9 * ci, device=DEV)
    K.conv_weights_prep_g(P(wk_), P(wf), P(wd), cs, ci, co, ST())
    return wf, wd


DEC = [(2, 21, 32, 128, 128), (2, 42, 128, 64, 64), (1, 84, 64, 9, 32), (3, 23, 64, 32, 32)]


@pytest.mark.parametrize("B,H,Cin,Co,Cs", DEC)
def test_conv_tcg_forward_dgrad_wgrad(B, H, Cin, Co, Cs):
    x = tf32_round(F.relu(rnd(B, Cin, H, H, seed=1)))
    w = rnd(Co, Cin, 3, 3, seed=2, scale=0.05); b = rnd(Co, seed=3)
    bs = torch.zeros(Cs, device=DEV); bs[:Co] = b
    wf, wd = prep_wg(w, Cs)
    wt = tf32_round(w)
    xin = bordered(x, H + 2, H + 2, 1, 0)
    # forward, compact output
    y = torch.full((B, H, H, Cs), 3.0, device=DEV)
    K.conv_tcg(P(xin), P(wf), P(bs), 0, P(y), B, H + 2, H + 2, Cin, Cs, H, H, -1, H, H, 0, 0, 0, 0, 0, ST())
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1).float()
    close(y.permute(0, 3, 1, 2)[:, :Co], ref, rtol=3e-5, what="conv_tcg fwd")
    if Cs > Co:
        assert float(y[..., Co:].abs().max()) == 0.0
    # forward with fused ReLU + nearest x2 upsample + TF32 round into the next layer's zero-bordered input
    up = torch.zeros(B, 2 * H + 2, 2 * H + 2, Cs, device=DEV)
    K.conv_tcg(P(xin), P(wf), P(bs), 0, P(up), B, H + 2, H + 2, Cin, Cs, H, H, -1, 2 * H + 2, 2 * H + 2, 1, 0, 0, 0, 1 | 2 | 16, ST())
    torch.cuda.synchronize()
    refu = tf32_round(F.interpolate(F.relu(ref), scale_factor=2))
    close(up[:, 1:2 * H + 1, :2 * H].permute(0, 3, 1, 2)[:, :Co], refu, rtol=1e-3, what="conv_tcg fwd upsample")
    bz = up.clone(); bz[:, 1:2 * H + 1, :2 * H] = 0
    assert float(bz.abs().max()) == 0.0
    # data gradient (no mask / ReLU mask)
    dy = tf32_round(rnd(B, Co, H, H, seed=4))
    dys = torch.zeros(B, Cs, H, H, device=DEV); dys[:, :Co] = dy
    dyb = bordered(dys, H + 2, H + 2, 1, 0)
    refd = F.conv_transpose2d(dy.double(), wt.double(), padding=1).float()
    if Cin in (32, 64, 128):
        dx = torch.zeros(B, H, H, Cin, device=DEV)
        K.conv_tcg(P(dyb), P(wd), 0, 0, P(dx), B, H + 2, H + 2, Cs, Cin, H, H, -1, H, H, 0, 0, 0, 0, 0, ST())
        torch.cuda.synchronize()
        close(dx.permute(0, 3, 1, 2), refd, rtol=3e-5, what="conv_tcg dgrad")
        xm = nhwc(x)
        K.conv_tcg(P(dyb), P(wd), 0, P(xm), P(dx), B, H + 2, H + 2, Cs, Cin, H, H, -1, H, H, 0, 0, H, H, 4, ST())
        torch.cuda.synchronize()
        close(dx.permute(0, 3, 1, 2), refd * (x > 0), rtol=3e-5, what="conv_tcg dgrad masked")
    # weight gradient
    dw = torch.zeros(Cs * 9 * Cin, device=DEV)
    K.conv_wgrad_tcg(P(xin), P(dyb), P(dw), B, H + 2, H + 2, Cin, Cs, -1, -1, ST())
    torch.cuda.synchronize()
    wr = w.clone().double().requires_grad_(True)
    F.conv2d(x.double(), wr, padding=1).backward(dy.double())
    close(dw.reshape(Cs, 3, 3, Cin).permute(0, 3, 1, 2)[:Co], wr.grad.float(), rtol=3e-5, what="conv_wgrad_tcg")


@pytest.mark.parametrize("B,Hl,Cin,Co,Cg", [(2, 42, 64, 9, 16), (3, 10, 64, 9, 16), (1, 21, 32, 5, 8), (2, 21, 128, 64, 64), (1, 7, 128, 64, 64),
                                          (46, 21, 128, 64, 64)])    # 191 position tiles > #SMs: two tiles per weight pass (pair mode), odd tail
def test_phase_conv_equals_conv_after_upsample(B, Hl, Cin, Co, Cg):
    """conv3x3(pad 1) o F.upsample(x, 2) in its sub-pixel form (sgqn_conv_weights_prep_phase + sgqn_conv_tcg at low
    resolution + sgqn_conv_phase_fold) against torch on the materialised upsampled tensor: forward (phase layout, or
    depth-to-space epilogue for the 256-channel conv2 form), masked data gradient w.r.t. the low-res activation (also
    written in space-to-depth form), folded weight / bias gradients."""
    Np, H = 4 * Cg, 2 * Hl
    x = tf32_round(F.relu(rnd(B, Cin, Hl, Hl, seed=1)))
    w = rnd(Co, Cin, 3, 3, seed=2, scale=0.05); b = rnd(Co, seed=3)
    pad_rows = 3 if Co < Cg else 0
    wst = torch.zeros(Co + pad_rows, Cin, 3, 3, device=DEV); wst[:Co] = w         # stored with padding rows
    wk_ = wk(wst).contiguous()
    wf = torch.zeros(Np * 9 * Cin, device=DEV); wd = torch.zeros(Np * 9 * Cin, device=DEV); bp = torch.zeros(Np, device=DEV)
    K.conv_weights_prep_phase(P(wk_), P(b), P(wf), P(wd), P(bp), Cin, Co, Cg, ST())
    xin = bordered(x, Hl + 2, Hl + 2, 1, 0)
    xr = x.double().requires_grad_(True); wr = w.double().requires_grad_(True); br = b.double().requires_grad_(True)
    ref = F.conv2d(F.interpolate(xr, scale_factor=2), wr, br, padding=1)            # (B,Co,H,H)
    tol = dict(rtol=2e-3, atol=2e-3 * float(ref.detach().abs().max()))               # TF32 of summed taps
    if Np == 256:       # depth-to-space epilogue: phase p = 2a+b of low-res (y,x) -> pixel (2y+a, 2x+b) of a Cg-channel buffer
        yd = torch.zeros(B, H + 2, H + 2, Cg, device=DEV)
        K.conv_tcg(P(xin), P(wf), P(bp), 0, P(yd), B, Hl + 2, Hl + 2, Cin, Np, Hl, Hl, -1, H + 2, H + 2, 1, 0, 0, 0, 1 << 5, ST())
        torch.cuda.synchronize()
        close(yd[:, 1:H + 1, :H].permute(0, 3, 1, 2), ref.detach().float(), what="phase conv fwd (depth-to-space)", **tol)
        bz = yd.clone(); bz[:, 1:H + 1, :H] = 0
        assert float(bz.abs().max()) == 0.0
        yr = torch.zeros(B, H + 2, H + 2, Cg, device=DEV)
        K.conv_tcg(P(xin), P(wf), P(bp), 0, P(yr), B, Hl + 2, Hl + 2, Cin, Np, Hl, Hl, -1, H + 2, H + 2, 1, 0, 0, 0, 1 | 2 | (1 << 5), ST())
        torch.cuda.synchronize()
        assert torch.equal(yr, tf32_round(F.relu(yd)))
    else:
        yp = torch.zeros(B, Hl + 2, Hl + 2, Np, device=DEV)
        K.conv_tcg(P(xin), P(wf), P(bp), 0, P(yp), B, Hl + 2, Hl + 2, Cin, Np, Hl, Hl, -1, Hl + 2, Hl + 2, 1, 0, 0, 0, 0, ST())
        torch.cuda.synchronize()
        # depth-to-space of our output: phase p = 2a+b, channels [p*Cg, p*Cg+Co) -> pixel (2y+a, 2x+b)
        got = yp[:, 1:Hl + 1, :Hl].reshape(B, Hl, Hl, 2, 2, Cg)[..., :Co].permute(0, 5, 1, 3, 2, 4).reshape(B, Co, H, H)
        close(got, ref.detach().float(), what="phase conv fwd", **tol)
        if Co < Cg:
            assert float(yp[:, 1:Hl + 1, :Hl].reshape(B, Hl, Hl, 4, Cg)[..., Co:].abs().max()) == 0.0
    # backward
    dy = tf32_round(rnd(B, Co, H, H, seed=4))
    ref.backward(dy.double())
    dyp = torch.zeros(B, Hl + 2, Hl + 2, 4, Cg, device=DEV)
    dyp[:, 1:Hl + 1, :Hl, :, :Co] = dy.reshape(B, Co, Hl, 2, Hl, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Hl, Hl, 4, Co)
    dx = torch.zeros(B, Hl + 2, Hl + 2, Cin, device=DEV)
    K.conv_tcg(P(dyp), P(wd), 0, P(xin, (Hl + 2) * Cin), P(dx), B, Hl + 2, Hl + 2, Np, Cin, Hl, Hl, -1, Hl + 2, Hl + 2, 1, 0,
               Hl + 2, Hl + 2, (1 << 2) | 2, ST())
    torch.cuda.synchronize()
    refd = (xr.grad * (x > 0)).float()
    close(dx[:, 1:Hl + 1, :Hl].permute(0, 3, 1, 2), refd, rtol=2e-3, atol=2e-3 * float(refd.abs().max()), what="phase conv dgrad")
    bz = dx.clone(); bz[:, 1:Hl + 1, :Hl] = 0
    assert float(bz.abs().max()) == 0.0
    if Hl % 2 == 0 and Cin in (32, 64):     # the same gradient written in space-to-depth form (what the next phase conv's backward reads)
        h2 = Hl // 2
        dxs = torch.zeros(B, h2 + 2, h2 + 2, 4, Cin, device=DEV)
        K.conv_tcg(P(dyp), P(wd), 0, P(xin, (Hl + 2) * Cin), P(dxs), B, Hl + 2, Hl + 2, Np, Cin, Hl, Hl, -1, h2 + 2, h2 + 2, 1, 0,
                   Hl + 2, Hl + 2, (1 << 2) | 2 | (2 << 5), ST())
        torch.cuda.synchronize()
        want = dx[:, 1:Hl + 1, :Hl].reshape(B, h2, 2, h2, 2, Cin).permute(0, 1, 3, 2, 4, 5).reshape(B, h2, h2, 4, Cin)
        assert torch.equal(dxs[:, 1:h2 + 1, :h2], want)
        bz = dxs.clone(); bz[:, 1:h2 + 1, :h2] = 0
        assert float(bz.abs().max()) == 0.0
    dwp = torch.zeros(Np * 9 * Cin, device=DEV); dbp = torch.zeros(Np, device=DEV)
    if Np == 256:
        for h in range(2):
            K.conv_wgrad_tcg_ld(P(xin), P(dyp, 128 * h), 256, P(dwp, h * 128 * 9 * Cin), B, Hl + 2, Hl + 2, Cin, 128, -1, -1, ST())
    else:
        K.conv_wgrad_tcg(P(xin), P(dyp), P(dwp), B, Hl + 2, Hl + 2, Cin, Np, -1, -1, ST())
    K.colsum(P(dyp), Np, B * (Hl + 2) * (Hl + 2), Np, P(dbp), ST())
    dw = torch.zeros((Co + pad_rows) * 9 * Cin, device=DEV); db = torch.zeros(Co + pad_rows, device=DEV)
    K.conv_phase_fold(P(dwp), P(dbp), P(dw), P(db), Cin, Co, Cg, ST())
    torch.cuda.synchronize()
    close(dw.reshape(Co + pad_rows, 3, 3, Cin).permute(0, 3, 1, 2)[:Co], wr.grad.float(), rtol=1e-4, atol=1e-4 * float(wr.grad.abs().max()),
          what="phase conv wgrad (folded)")
    close(db[:Co], br.grad.float(), rtol=1e-4, what="phase conv bias grad")
    if pad_rows:
        assert float(dw.reshape(Co + pad_rows, -1)[Co:].abs().max()) == 0.0


def test_bce_phase_equals_bce():
    B, H = 3, 84
    Hl = H // 2
    lg = rnd(B, 9, H, H, seed=1, scale=2.0)
    mask = (rnd(B, 3, H * H, seed=2) > 0.8).to(torch.uint8)
    x = lg.clone().requires_grad_(True)
    y = mask.float().reshape(B, 3, H, H).repeat_interleave(3, dim=1)
    ref = F.binary_cross_entropy_with_logits(x, y)
    ref.backward()
    lgp = torch.zeros(B, Hl + 2, Hl + 2, 4, 16, device=DEV)
    lgp[:, 1:Hl + 1, :Hl, :, :9] = lg.reshape(B, 9, Hl, 2, Hl, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Hl, Hl, 4, 9)
    lgp[..., 12:] = 7.0                                            # padding channels must be ignored
    loss = torch.zeros(1, device=DEV); d = torch.zeros_like(lgp)
    K.bce_phase(P(lgp), P(mask), P(loss), P(d), B, H, H, Hl + 2, Hl + 2, 1, 0, B, 0, ST())
    torch.cuda.synchronize()
    close(loss, ref.detach().reshape(1), what="bce_phase loss")
    got = d[:, 1:Hl + 1, :Hl, :, :9].reshape(B, Hl, Hl, 2, 2, 9).permute(0, 5, 1, 3, 2, 4).reshape(B, 9, H, H)
    close(got, x.grad, rtol=2e-4, what="bce_phase grad")
    assert float(d[..., 9:].abs().max()) == 0.0
    bz = d.clone(); bz[:, 1:Hl + 1, :Hl] = 0
    assert float(bz.abs().max()) == 0.0


def test_pool2_bwd():
    B, H, C = 2, 21, 128
    dup = rnd(B, 2 * H, 2 * H, C, seed=1)
    low = rnd(B, C, H, H, seed=2)
    src = torch.zeros(B, 2 * H + 2, 2 * H + 2, C, device=DEV)
    src[:, 1:2 * H + 1, :2 * H] = F.interpolate(F.relu(low), scale_factor=2).permute(0, 2, 3, 1)
    dst = torch.zeros(B, H + 2, H + 2, C, device=DEV)
    K.pool2_bwd(P(dup), P(src), P(dst), B, H, H, C, ST())
    ref = F.avg_pool2d(dup.permute(0, 3, 1, 2), 2) * 4 * (low > 0)
    close(dst[:, 1:H + 1, :H].permute(0, 3, 1, 2), tf32_round(ref), rtol=1e-3, what="pool2_bwd")
    bz = dst.clone(); bz[:, 1:H + 1, :H] = 0
    assert float(bz.abs().max()) == 0.0


def test_conv1_tcgen05_path_matches_cuda_core_path():
    """First conv as a per-position tcgen05 GEMM over col[.][96] (forward, weight gradient) vs the fp32 im2col kernels."""
    B = 5
    g = torch.Generator().manual_seed(5)
    obs = torch.randint(0, 256, (B, 9, 84, 84), generator=g).float().to(DEV)
    w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
    col = torch.zeros(B * 1681, 84, device=DEV); col96 = torch.full((B * 1681, 96), 9.0, device=DEV)
    K.conv1_im2col(P(obs), P(col), B, 84, ST()); K.conv1_im2col96(P(obs), P(col96), B, 84, ST())
    assert torch.equal(col96[:, :81], tf32_round(col[:, :81])) and float(col96[:, 81:].abs().max()) == 0.0
    wp = torch.zeros(32 * 96, device=DEV)
    K.conv1_weights_prep(P(w), P(wp), 0, ST())
    assert torch.equal(wp.reshape(32, 96)[:, :81], tf32_round(w.reshape(32, 81)))
    y0 = torch.zeros(B, 43, 41, 32, device=DEV); y1 = torch.zeros(B, 43, 41, 32, device=DEV)
    K.conv1_fwd_col(P(col), P(w), P(b), P(y0), B, 7, ST())
    K.conv_tcg_taps(P(col96), P(wp), P(b), 0, P(y1), B, 41, 41, 96, 32, 41, 41, 0, 43, 41, 0, 0, 0, 0, 3, 1, ST())
    torch.cuda.synchronize()
    close(y1, y0, rtol=2e-3, what="conv1 tcgen05 fwd")
    assert float(y1[:, 41:].abs().max()) == 0.0
    dy = tf32_round(rnd(B * 1681, 32, seed=4))
    dw0 = torch.zeros(32 * 81, device=DEV); db0 = torch.zeros(32, device=DEV); dw1 = torch.zeros(32 * 81, device=DEV)
    K.conv1_wgrad_col(P(col), P(dy), P(dw0), P(db0), B, ST())
    K.gemm_wgrad_tcg(P(col96), P(dy), P(dw1), B, 41, 41, 96, 32, 0, 0, 1, 81, ST())
    torch.cuda.synchronize()
    close(dw1, dw0, rtol=2e-3, what="conv1 tcgen05 wgrad")


@pytest.mark.parametrize("B,Hin,col_row0", [(5, 84, 0), (7, 84, 3), (3, 100, 0), (130, 84, 130), (1, 84, 0)])
def test_conv1_fused_matches_im2col_path(B, Hin, col_row0):
    """First conv with the im2col tile built in shared memory (sgqn_conv1_fused_tc) against the materialised-im2col
    tcgen05 path it replaces (same operands, same K order) and against torch; the optional TMA-stored im2col matrix is
    bit-exact for the requested samples."""
    g = torch.Generator().manual_seed(5)
    obs = torch.randint(0, 256, (B, 9, Hin, Hin), generator=g).float().to(DEV)
    obs[0] = obs[0] * 0.37 + 1.3                                   # non-integer pixels too (masked / overlaid observations)
    w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
    wp = torch.zeros(32 * 96, device=DEV)
    K.conv1_weights_prep(P(w), P(wp), 0, ST())
    col96 = torch.zeros(B * 1681, 96, device=DEV)
    K.conv1_im2col96(P(obs), P(col96), B, Hin, ST())
    y1 = torch.zeros(B, 43, 41, 32, device=DEV)
    K.conv_tcg_taps(P(col96), P(wp), P(b), 0, P(y1), B, 41, 41, 96, 32, 41, 41, 0, 43, 41, 0, 0, 0, 0, 3, 1, ST())
    y2 = torch.zeros(B, 43, 41, 32, device=DEV)
    colf = torch.full((B * 1681, 96), -7.0, device=DEV)
    want_col = col_row0 < B
    K.conv1_fused_tc(P(obs), P(wp), P(b), P(y2), P(colf) if want_col else 0, B, Hin, col_row0, ST())
    torch.cuda.synchronize()
    assert torch.equal(y2, y1)
    assert float(y2[:, 41:].abs().max()) == 0.0
    c = (Hin - 84) // 2
    x = obs[:, :, c:c + 84, c:c + 84]
    ref = tf32_round(F.relu(F.conv2d(tf32_round(x / 255.0).double(), tf32_round(w).double(), b.double(), stride=2)).float())
    close(y2[:, :41].permute(0, 3, 1, 2), ref, rtol=2e-3, atol=1e-5, what="conv1 fused vs torch")
    if want_col:
        assert torch.equal(colf[col_row0 * 1681:], col96[col_row0 * 1681:])
        if col_row0:
            assert float((colf[:col_row0 * 1681] + 7.0).abs().max()) == 0.0
    else:
        assert float((colf + 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("B", [1, 3, 130])
def test_conv1_dgrad_fused_matches_materialised_path(B):
    """Observation gradient of the first conv in one kernel (dcol stays in TMEM / shared memory) against the path it replaces
    (per-position GEMM into a dcol matrix + col2im gather: same operands, same summation order -> bit-exact) and torch."""
    d = tf32_round(rnd(B, 41, 41, 32, seed=1) * (rnd(B, 41, 41, 32, seed=2) > 0.3))
    w = rnd(32, 9, 3, 3, seed=3, scale=0.2)
    wp = torch.zeros(32 * 96, device=DEV); wd = torch.zeros(96 * 32, device=DEV)
    K.conv1_weights_prep(P(w), P(wp), P(wd), ST())
    dcol = torch.zeros(B * 1681, 96, device=DEV)
    K.conv_tcg_taps(P(d), P(wd), 0, 0, P(dcol), B, 41, 41, 32, 96, 41, 41, 0, 41, 41, 0, 0, 0, 0, 0, 1, ST())
    d0 = torch.full((B, 9, 84, 84), 5.0, device=DEV)
    K.conv1_col2im(P(dcol), 96, P(d0), B, ST())
    d1 = torch.full((B, 9, 84, 84), -3.0, device=DEV)
    K.conv1_dgrad_fused_tc(P(d), P(wd), P(d1), B, ST())
    torch.cuda.synchronize()
    assert torch.equal(d1, d0)
    ref = F.conv_transpose2d(d.permute(0, 3, 1, 2).double(), tf32_round(w).double(), stride=2, output_padding=1) / 255.0
    close(d1, ref.float(), rtol=2e-3, atol=2e-3 * float(ref.abs().max()), what="conv1 dgrad fused vs torch")
    assert float(d1[:, :, 83].abs().max()) == 0.0 and float(d1[:, :, :, 83].abs().max()) == 0.0


# ------------------------------------------------------------------ the ten 32->32 convs as one persistent launch
@pytest.mark.parametrize("B", [3, 16, 40])
def test_conv_chain_equals_per_layer_launches(B):
    """conv_chain.cu (ticketed tile list over all layers, per-tile producer / consumer flags) against ten sgqn_conv_tc launches:
    forward activations and data gradients (plain and guided) are BIT-identical, bias gradients equal up to the order of the
    per-CTA partial sums; repeated launches on one workspace (epoch / ticket re-arming) stay correct."""
    import sgqn_carla_b200 as S
    from sgqn_carla_b200.engine import _ptr
    from sgqn_carla_b200.layout import ENC_H, FEAT
    args = S.default_args(algorithm="sac", batch_size=B)
    agent = S.make_agent((9, 84, 84), (2,), args)
    eng = agent.engine
    g = torch.Generator().manual_seed(B)
    from oracle import sgsac_oracle as O
    agent.set_parameters(O.init_params((9, 84, 84), 2, O.Args(**vars(args)), g, dense_std=0.05))
    obs = torch.randint(0, 256, (2 * B, 9, 84, 84), generator=g).float().to(DEV)
    dfeat = (torch.randn(2 * B, FEAT, generator=g) * 1e-2).to(DEV)
    outs = {}
    for chain in (False, True, True):                       # the third pass re-uses the workspace of the second
        eng.chain = chain
        for t in eng.actS + [x for x in eng.gpad if x is not None] + [eng.dbuf[0], eng.grads, eng.obs_grad]:
            t.zero_()
        eng.enc_fwd(_ptr(obs), 2 * B, eng.actS, B, col_from=0)
        acts = [a.clone() for a in eng.actS]
        eng.enc_bwd(_ptr(dfeat), 2 * B, eng.actS, B, _ptr(obs), 1, True)
        torch.cuda.synchronize()
        gp = [x.clone() for x in eng.gpad if x is not None] + [eng.dbuf[0].clone()]
        grads = eng.lay.unpack(eng.grads)
        eng.enc_bwd(_ptr(dfeat), B, eng.actS, B, 0, 2, False, dobs=_ptr(eng.obs_grad))
        torch.cuda.synchronize()
        outs[len(outs)] = (acts, gp, grads, eng.obs_grad.clone())
    ref = outs[0]
    assert float(ref[0][10].abs().max()) > 0 and float(ref[3].abs().max()) > 0
    for k in (1, 2):
        acts, gp, grads, og = outs[k]
        for l, (a, b) in enumerate(zip(acts, ref[0])):
            assert torch.equal(a, b), ("act", l, k)
        for l, (a, b) in enumerate(zip(gp, ref[1])):
            assert torch.equal(a, b), ("d act", l, k)
        assert torch.equal(og, ref[3]), ("guided attribution", k)
        for l in range(11):
            for wb in ("weight", "bias"):
                a, b = grads[f"cnn.{l}.{wb}"], ref[2][f"cnn.{l}.{wb}"]
                assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12, (l, wb, k)
