"""TEST INFRASTRUCTURE ONLY -- the parity oracle; never imported by the product path.

A self-contained CPU restatement (plain torch fp32 on CPU, autograd for the
backward) of the reference's SGSAC / SAC / SVEA update path, with every random
draw supplied by the caller.  It travels to the GPU box (the reference itself
cannot: /root/reference does not exist there) and is used only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs, as the checker / the timed CPU baseline.

Pinning: `tests/test_oracle_pin.py` runs this against the UNMODIFIED reference
imported through `oracle/ref_shim.py` (this container only) and against the
golden vectors under `tests/golden/` produced from the reference by
`oracle/make_golden.py` (any box).  Third-party arithmetic restated here:
captum==0.5.0 GuidedBackprop (setup/sgqn-carla.yml:12; call site rl_utils.py:35-39)
and kornia==0.6.6 RandomCrop (setup/sgqn-carla.yml:14; call site
augmentations.py:229-233) -- the reference holds no tests for either, so those
two are pinned only against our own restatement in ref_shim ("parity unpinned"
by reference tests).

TF32 mode (`tf32=True`): the reference runs its convolutions with cuDNN's
`allow_tf32=True` default (SURVEY.md 8c "Oracle numerics"), i.e. every conv operand
(activations, weights, output gradients) is rounded to a 10-bit mantissa before the
multiply and products are accumulated in fp32.  The product path does the same on
tcgen05 (`kind::tf32`), rounding to nearest-away (`cvt.rna.tf32.f32`) at the
producer of every MMA operand.  `tf32=True` restates the convolutions with exactly
those roundings (forward, data gradient and weight gradient; `nn.Linear` stays fp32
as in the reference, matmul TF32 is off), so that this oracle and the product path
differ by fp32 summation order only and can be held to rel 1e-3 end to end.
`tf32=False` is the bit-exact CPU restatement that is pinned against the reference.

All file:line citations are relative to /root/reference/src.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

FEAT = 32 * 21 * 21  # modules.py:318 hard-codes 14112


# --------------------------------------------------------------------------
# default hyper-parameters = arguments.py:6-144 defaults
# --------------------------------------------------------------------------
class Args:
    def __init__(self, **kw):
        self.algorithm = "sgsac"
        self.discount = 0.99
        self.batch_size = 128
        self.hidden_dim = 1024
        self.actor_lr = 1e-3
        self.actor_beta = 0.9
        self.actor_log_std_min = -10.0
        self.actor_log_std_max = 2.0
        self.actor_update_freq = 2
        self.critic_lr = 1e-3
        self.critic_beta = 0.9
        self.critic_tau = 0.01
        self.critic_target_update_freq = 2
        self.critic_weight_decay = 0.0
        self.num_shared_layers = 11
        self.num_head_layers = 0
        self.num_filters = 32
        self.projection_dim = 100
        self.encoder_tau = 0.05
        self.init_temperature = 0.1
        self.alpha_lr = 1e-4
        self.alpha_beta = 0.5
        self.aux_lr = 3e-4
        self.aux_beta = 0.9
        self.aux_update_freq = 2
        self.soda_batch_size = 256
        self.soda_tau = 0.005
        self.svea_alpha = 0.5
        self.svea_beta = 0.5
        self.sgqn_quantile = 0.5
        self.consistency = 1
        self.alpha_blending = 0.2
        self.seed = 10081
        self.log_dir = "logs"
        self.domain_name = "carla"
        self.task_name = "drive"
        self.image_size = 84
        self.image_crop_size = 84
        for k, v in kw.items():
            setattr(self, k, v)


# --------------------------------------------------------------------------
# parameter naming: canonical (oracle / product) name -> reference state_dict keys
# --------------------------------------------------------------------------
def _ref_key_map(num_layers=11):
    """canonical name -> list of (module, reference key) (train.py:207-219 saves these modules)."""
    m = OrderedDict()
    for i in range(num_layers):
        for wb in ("weight", "bias"):
            k = f"encoder.shared_cnn.layers.{2 + 2 * i}.{wb}"
            m[f"cnn.{i}.{wb}"] = [("critic", k), ("actor", k), ("attribution_predictor", k)]
    for j in (0, 1):
        for wb in ("weight", "bias"):
            k = f"encoder.projection.projection.{j}.{wb}"
            m[f"critic_proj.{j}.{wb}"] = [("critic", k), ("attribution_predictor", k)]
            m[f"actor_proj.{j}.{wb}"] = [("actor", k)]
    for q in ("Q1", "Q2"):
        for j in (0, 2, 4):
            for wb in ("weight", "bias"):
                m[f"{q}.{j}.{wb}"] = [("critic", f"{q}.trunk.{j}.{wb}")]
    for j in (0, 2, 4):
        for wb in ("weight", "bias"):
            m[f"actor_mlp.{j}.{wb}"] = [("actor", f"mlp.{j}.{wb}")]
    for l in ("proj", "conv1", "conv2", "conv3"):
        for wb in ("weight", "bias"):
            m[f"dec.{l}.{wb}"] = [("attribution_predictor", f"decoder.{l}.{wb}")]
    for j in (0, 2):
        for wb in ("weight", "bias"):
            m[f"fdec.{j}.{wb}"] = [("attribution_predictor", f"features_decoder.{j}.{wb}")]
    m["curl.W"] = [("curl_head", "W")]                              # modules.py:264-268
    for j in (0, 1):                                                # PAD's own projection + inverse-dynamics MLP (pad.py:18-25)
        for wb in ("weight", "bias"):
            m[f"pad_proj.{j}.{wb}"] = [("pad_head", f"encoder.projection.projection.{j}.{wb}")]
    for j in (0, 2, 4):
        for wb in ("weight", "bias"):
            m[f"pad_mlp.{j}.{wb}"] = [("pad_head", f"mlp.{j}.{wb}")]
    for j in (0, 1, 3):                                             # SODA: encoder projection SODAMLP + predictor SODAMLP (soda.py:19-27)
        for wb in ("weight", "bias"):
            m[f"soda_proj.{j}.{wb}"] = [("predictor", f"encoder.projection.mlp.{j}.{wb}")]
            m[f"soda_pred.{j}.{wb}"] = [("predictor", f"mlp.mlp.{j}.{wb}")]
    return m


CRITIC_GROUP = ("cnn.", "critic_proj.", "Q1.", "Q2.")           # sac.py:63-65 critic.parameters()
ACTOR_GROUP = ("cnn.", "actor_proj.", "actor_mlp.")             # sac.py:60-62 actor.parameters()
AUX_GROUP = ("cnn.", "critic_proj.", "dec.", "fdec.")           # sgsac.py:35-39 attribution_predictor.parameters()
SODA_GROUP = ("cnn.", "soda_proj.", "soda_pred.")               # soda.py:32-34 predictor.parameters()
PAD_GROUP = ("cnn.", "pad_proj.", "pad_mlp.")                    # pad.py:34-37 pad_head.parameters()
CURL_GROUP = ("cnn.", "critic_proj.", "curl.")                  # curl.py:16-20 curl_head.parameters() = critic encoder + W
TARGET_Q = ("Q1.", "Q2.")                                        # sac.py:154-155 (critic_tau)
TARGET_ENC = ("cnn.", "critic_proj.")                            # sac.py:156-158 (encoder_tau)


def _in_group(name, group):
    return any(name.startswith(g) for g in group)


# --------------------------------------------------------------------------
# init (modules.py:53-67 weight_init; torch default init for the decoder)
# --------------------------------------------------------------------------
def init_params(obs_shape, action_dim, args, gen=None, dense_std=None):
    """Returns OrderedDict canonical name -> fp32 tensor.

    dense_std=None : reference init (delta-orthogonal convs, orthogonal linears,
                     zero biases; default torch init for the decoder).
    dense_std=s    : every tensor ~ N(0, s) ("dense" variant of SURVEY.md 8d cfg 2,
                     makes attributions / masks non-degenerate); LayerNorm weight = 1+N(0,s).
    """
    g = gen if gen is not None else torch.Generator().manual_seed(0)
    H, P, nf, A = args.hidden_dim, args.projection_dim, args.num_filters, action_dim
    shapes = OrderedDict()
    for i in range(args.num_shared_layers):
        cin = obs_shape[0] if i == 0 else nf
        shapes[f"cnn.{i}.weight"] = (nf, cin, 3, 3)
        shapes[f"cnn.{i}.bias"] = (nf,)
    for pre in ("critic_proj", "actor_proj"):
        shapes[f"{pre}.0.weight"] = (P, FEAT)
        shapes[f"{pre}.0.bias"] = (P,)
        shapes[f"{pre}.1.weight"] = (P,)
        shapes[f"{pre}.1.bias"] = (P,)
    for q in ("Q1", "Q2"):
        shapes[f"{q}.0.weight"] = (H, P + A); shapes[f"{q}.0.bias"] = (H,)
        shapes[f"{q}.2.weight"] = (H, H); shapes[f"{q}.2.bias"] = (H,)
        shapes[f"{q}.4.weight"] = (1, H); shapes[f"{q}.4.bias"] = (1,)
    shapes["actor_mlp.0.weight"] = (H, P); shapes["actor_mlp.0.bias"] = (H,)
    shapes["actor_mlp.2.weight"] = (H, H); shapes["actor_mlp.2.bias"] = (H,)
    shapes["actor_mlp.4.weight"] = (2 * A, H); shapes["actor_mlp.4.bias"] = (2 * A,)
    shapes["dec.proj.weight"] = (FEAT, P + A); shapes["dec.proj.bias"] = (FEAT,)
    shapes["dec.conv1.weight"] = (128, 32, 3, 3); shapes["dec.conv1.bias"] = (128,)
    shapes["dec.conv2.weight"] = (64, 128, 3, 3); shapes["dec.conv2.bias"] = (64,)
    shapes["dec.conv3.weight"] = (9, 64, 3, 3); shapes["dec.conv3.bias"] = (9,)
    shapes["fdec.0.weight"] = (256, 100); shapes["fdec.0.bias"] = (256,)
    shapes["fdec.2.weight"] = (100, 256); shapes["fdec.2.bias"] = (100,)
    if getattr(args, "algorithm", "") == "pad":
        shapes["pad_proj.0.weight"] = (P, FEAT); shapes["pad_proj.0.bias"] = (P,)
        shapes["pad_proj.1.weight"] = (P,); shapes["pad_proj.1.bias"] = (P,)
        shapes["pad_mlp.0.weight"] = (H, 2 * P); shapes["pad_mlp.0.bias"] = (H,)
        shapes["pad_mlp.2.weight"] = (H, H); shapes["pad_mlp.2.bias"] = (H,)
        shapes["pad_mlp.4.weight"] = (A, H); shapes["pad_mlp.4.bias"] = (A,)
    if getattr(args, "algorithm", "") == "soda":                    # SODAMLP: Linear -> BatchNorm1d -> ReLU -> Linear (modules.py:116-129)
        for pre, fan in (("soda_proj", FEAT), ("soda_pred", P)):
            shapes[f"{pre}.0.weight"] = (P, fan); shapes[f"{pre}.0.bias"] = (P,)
            shapes[f"{pre}.1.weight"] = (P,); shapes[f"{pre}.1.bias"] = (P,)
            shapes[f"{pre}.3.weight"] = (P, P); shapes[f"{pre}.3.bias"] = (P,)
    if getattr(args, "algorithm", "") == "curl":
        shapes["curl.W"] = (P, P)                                   # CURLHead.W = torch.rand(out_dim, out_dim), modules.py:268

    p = OrderedDict()
    for name, shp in shapes.items():
        if dense_std is not None:
            t = torch.randn(*shp, generator=g) * dense_std
            if name.endswith("proj.1.weight"):
                t = t + 1.0
        elif name == "curl.W":
            t = torch.rand(*shp, generator=g)
        elif name.startswith(("dec.", "fdec.")):
            # torch default Conv2d / Linear init: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
            wshape = shapes[name.rsplit(".", 1)[0] + ".weight"]
            fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(*shp, generator=g) * 2 - 1) * bound
        elif name.endswith("proj.1.weight") or name == "soda_pred.1.weight":
            t = torch.ones(*shp)
        elif name.endswith("bias"):
            t = torch.zeros(*shp)
        elif len(shp) == 4:
            t = torch.zeros(*shp)                                   # modules.py:62-67
            w = torch.empty(shp[0], shp[1])
            torch.nn.init.orthogonal_(w, math.sqrt(2.0), generator=g)
            t[:, :, 1, 1] = w
        else:
            t = torch.empty(*shp)
            torch.nn.init.orthogonal_(t, generator=g)               # modules.py:55-58
        p[name] = t.float().contiguous()
    return p


# --------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------
class _GuidedReLU(torch.autograd.Function):
    """captum GuidedBackprop ReLU override: g_in = relu(g_out * 1[x>0])."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x.clamp(min=0)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return F.relu(g * (x > 0).to(g.dtype))


def _relu(x, guided):
    return _GuidedReLU.apply(x) if guided else F.relu(x)


def round_tf32(x):
    """cvt.rna.tf32.f32 on finite fp32 values: round to nearest (ties away from zero) on the 13 dropped mantissa bits."""
    bits = x.detach().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class _TF32Conv(torch.autograd.Function):
    """conv2d whose three contractions (forward, data gradient, weight gradient) take TF32-rounded operands and
    accumulate in fp32 -- cuDNN allow_tf32 / tcgen05 kind::tf32 semantics.  The bias gradient is the sum of the
    ROUNDED output gradient (the product path sums what it stores for the next MMA)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding, acc):
        xr, wr = round_tf32(x).to(acc), round_tf32(w).to(acc)
        ctx.save_for_backward(xr, wr)
        ctx.cfg = (stride, padding, b is not None, acc)
        return F.conv2d(xr, wr, None if b is None else b.to(acc), stride=stride, padding=padding).float()

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        stride, padding, has_b, acc = ctx.cfg
        gr = round_tf32(g).to(acc)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.nn.grad.conv2d_input(xr.shape, wr, gr, stride=stride, padding=padding).float()
        if ctx.needs_input_grad[1]:
            gw = torch.nn.grad.conv2d_weight(xr, wr.shape, gr, stride=stride, padding=padding).float()
        if has_b and ctx.needs_input_grad[2]:
            gb = gr.sum(dim=(0, 2, 3)).float()
        return gx, gw, gb, None, None, None


def conv2d(x, w, b, stride=1, padding=0, tf32=False):
    """tf32: False = plain fp32 (the pinned restatement); True = TF32-rounded operands, fp32 accumulation (what cuDNN /
    tcgen05 compute, up to summation order); "f64" = TF32-rounded operands with the products summed in fp64 and rounded
    to fp32 once -- the summation-order-free centre that both an fp32-accumulating CPU run and the tensor-core run
    scatter around (tests use it to calibrate how far ANY fp32-accumulating TF32 evaluation is from it)."""
    if tf32:
        return _TF32Conv.apply(x, w, b, stride, padding, torch.float64 if tf32 == "f64" else torch.float32)
    return F.conv2d(x, w, b, stride=stride, padding=padding)


def _phase_tap(a, k):
    return (-1 if k == 0 else 0) if a == 0 else (1 if k == 2 else 0)


def phase_weights(w):
    """Sub-pixel form of conv3x3(pad 1) o nearest-x2-upsample (modules.py:327-337): the conv over the upsampled image
    equals a 3x3 pad-1 conv at LOW resolution with 4x the output channels (phase p = 2a+b of low-res pixel (y,x) is
    output pixel (2y+a, 2x+b)); W_phi[(a,b,co)][dy][dx] = sum of the taps (ky,kx) that land on low-res offset (dy,dx).
    w (Cout,Cin,3,3) -> (4*Cout,Cin,3,3), differentiable (the chain rule folds dW_phi back onto the 3x3 taps)."""
    out = []
    for a in (0, 1):
        for b in (0, 1):
            taps = [[None] * 3 for _ in range(3)]
            for ky in range(3):
                for kx in range(3):
                    dy, dx = _phase_tap(a, ky) + 1, _phase_tap(b, kx) + 1
                    taps[dy][dx] = w[:, :, ky, kx] if taps[dy][dx] is None else taps[dy][dx] + w[:, :, ky, kx]
            z = torch.zeros_like(w[:, :, 0, 0])
            out.append(torch.stack([torch.stack([t if t is not None else z for t in row], -1) for row in taps], -2))
    return torch.cat(out, 0)


def conv_after_upsample(x, w, b, tf32=False):
    """conv2d(F.interpolate(x, scale_factor=2), w, b, padding=1) evaluated in sub-pixel form: in TF32 mode the operand that
    is rounded is W_phi (the fp32 sum of up to four taps), as on the product path (conv_tcg.cu)."""
    co = w.size(0)
    y = conv2d(x, phase_weights(w), b.repeat(4), padding=1, tf32=tf32)          # (B, 4*co, H, W), channel = (a, b, co)
    B, _, H, W = y.shape
    return y.view(B, 2, 2, co, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, co, 2 * H, 2 * W)


def center_crop(x):
    """modules.py:70-83"""
    if x.size(2) == 84 and x.size(3) == 84:
        return x
    assert x.size(3) == 100
    return x[:, :, 8:-8, 8:-8]


def cnn_forward(p, x, pre="cnn", guided=False, n_layers=11, tf32=False, keep=None):
    """modules.py:132-152 SharedCNN (+ HeadCNN with 0 layers = Flatten, :155-168).  keep: optional list that receives
    every layer's pre-activation (tests: ReLU sign patterns)."""
    x = center_crop(x) / 255.0
    x = conv2d(x, p[f"{pre}.0.weight"], p[f"{pre}.0.bias"], stride=2, tf32=tf32)
    if keep is not None:
        keep.append(x.detach())
    for i in range(1, n_layers):
        x = _relu(x, guided)
        x = conv2d(x, p[f"{pre}.{i}.weight"], p[f"{pre}.{i}.bias"], stride=1, tf32=tf32)
        if keep is not None:
            keep.append(x.detach())
    return x.reshape(x.size(0), -1)


def projection(p, feat, pre):
    """modules.py:102-113 Linear -> LayerNorm -> Tanh"""
    y = F.linear(feat, p[f"{pre}.0.weight"], p[f"{pre}.0.bias"])
    y = F.layer_norm(y, (y.size(-1),), p[f"{pre}.1.weight"], p[f"{pre}.1.bias"], 1e-5)
    return torch.tanh(y)


def mlp3(p, x, pre, guided=False):
    x = _relu(F.linear(x, p[f"{pre}.0.weight"], p[f"{pre}.0.bias"]), guided)
    x = _relu(F.linear(x, p[f"{pre}.2.weight"], p[f"{pre}.2.bias"]), guided)
    return F.linear(x, p[f"{pre}.4.weight"], p[f"{pre}.4.bias"])


def critic_forward(p, obs, action, detach=False, target=False, guided=False, only_q1=False, tf32=False):
    """modules.py:252-261 Critic.forward (+ :171-184 Encoder.forward)."""
    t = "t_" if target else ""
    feat = cnn_forward(p, obs, pre=t + "cnn", guided=guided, tf32=tf32)
    if detach:
        feat = feat.detach()
    h = projection(p, feat, t + "critic_proj")
    ha = torch.cat([h, action], dim=1)
    q1 = mlp3(p, ha, t + "Q1", guided)
    if only_q1:
        return q1
    return q1, mlp3(p, ha, t + "Q2", guided)


def actor_forward(p, obs, args, noise=None, compute_pi=True, compute_log_pi=True, detach=False, tf32=False):
    """modules.py:201-232 Actor.forward with gaussian_logprob / squash (:20-33)."""
    feat = cnn_forward(p, obs, pre="cnn", tf32=tf32)
    if detach:
        feat = feat.detach()
    h = projection(p, feat, "actor_proj")
    mu, log_std = mlp3(p, h, "actor_mlp").chunk(2, dim=-1)
    log_std = torch.tanh(log_std)
    log_std = args.actor_log_std_min + 0.5 * (args.actor_log_std_max - args.actor_log_std_min) * (log_std + 1)
    if compute_pi:
        std = log_std.exp()
        pi = mu + noise * std
    else:
        pi = None
    if compute_log_pi:
        residual = (-0.5 * noise.pow(2) - log_std).sum(-1, keepdim=True)
        log_pi = residual - 0.5 * np.log(2 * np.pi) * noise.size(-1)
    else:
        log_pi = None
    mu = torch.tanh(mu)
    if pi is not None:
        pi = torch.tanh(pi)
    if log_pi is not None:
        log_pi = log_pi - torch.log(F.relu(1 - pi.pow(2)) + 1e-6).sum(-1, keepdim=True)
    return mu, pi, log_pi, log_std


def decoder_forward(p, h, action, tf32=False):
    """modules.py:315-341 AttributionDecoder (F.upsample default = nearest)."""
    x = torch.cat([h, action], dim=1)
    x = F.linear(x, p["dec.proj.weight"], p["dec.proj.bias"]).view(-1, 32, 21, 21)
    x = F.relu(x)
    if tf32:
        # relu and nearest upsample commute; the convs that follow an upsample run in sub-pixel form on the low-res tensor
        x = F.relu(conv2d(x, p["dec.conv1.weight"], p["dec.conv1.bias"], padding=1, tf32=True))
        x = F.relu(conv_after_upsample(x, p["dec.conv2.weight"], p["dec.conv2.bias"], tf32=True))
        return conv_after_upsample(x, p["dec.conv3.weight"], p["dec.conv3.bias"], tf32=True)
    x = F.conv2d(x, p["dec.conv1.weight"], p["dec.conv1.bias"], padding=1)
    x = F.interpolate(x, scale_factor=2)
    x = F.relu(x)
    x = F.conv2d(x, p["dec.conv2.weight"], p["dec.conv2.bias"], padding=1)
    x = F.interpolate(x, scale_factor=2)
    x = F.relu(x)
    return F.conv2d(x, p["dec.conv3.weight"], p["dec.conv3.bias"], padding=1)


def attribution_predictor_forward(p, obs, action, tf32=False):
    """modules.py:345-354: critic encoder (not detached) -> decoder."""
    feat = cnn_forward(p, obs, pre="cnn", tf32=tf32)
    return decoder_forward(p, projection(p, feat, "critic_proj"), action, tf32=tf32)


# --------------------------------------------------------------------------
# saliency (rl_utils.py:23-39,57-62,76-82)
# --------------------------------------------------------------------------
def compute_attribution(p, obs, action, tf32=False):
    """Guided backprop of sum_b Q1[b] w.r.t. obs; Q1 only (ModelWrapper `[0]`, rl_utils.py:31-32)."""
    x = obs.detach().clone().requires_grad_(True)
    pd = {k: v.detach() for k, v in p.items()}
    with torch.enable_grad():
        q1 = critic_forward(pd, x, action.detach(), guided=True, only_q1=True, tf32=tf32)
        (g,) = torch.autograd.grad(q1.sum(), x)
    return g


def quantile_threshold(a, quantile):
    """torch.quantile(a, q, dim=1) ('linear') restated: a (R, n) fp32 -> (R,) fp32.

    rank = fp32(q)*(n-1) in the input dtype; lo = floor(rank), hi = ceil(rank), w = rank - lo;
    result = lerp(s[lo], s[hi], w) with torch's lerp formula
    (w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w))."""
    s, _ = torch.sort(a, dim=1)
    n = a.size(1)
    rank = torch.tensor(quantile, dtype=a.dtype) * (n - 1)
    lo = torch.floor(rank)
    w = rank - lo
    lo_i = int(lo.item())
    hi_i = int(torch.ceil(rank).item())
    x0, x1 = s[:, lo_i], s[:, hi_i]
    d = x1 - x0
    if float(w) < 0.5:
        return x0 + w * d
    return x1 - d * (1 - w)


def compute_attribution_mask(obs_grad, quantile=0.95, use_torch_quantile=False):
    """rl_utils.py:76-82 -> bool (B,9,H,W)."""
    mask = []
    for i in (0, 3, 6):
        a = obs_grad[:, i:i + 3].abs().max(dim=1)[0]
        if use_torch_quantile:
            q = torch.quantile(a.flatten(1), quantile, 1)
        else:
            q = quantile_threshold(a.flatten(1), quantile)
        mask.append((a >= q[:, None, None]).unsqueeze(1).repeat(1, 3, 1, 1))
    return torch.cat(mask, dim=1)


# --------------------------------------------------------------------------
# augmentations
# --------------------------------------------------------------------------
def random_crop(x, w1, h1, size=84):
    """augmentations.py:236-264: out[b] = x[b,:,w1[b]:w1[b]+size, h1[b]:h1[b]+size] (w1 indexes rows)."""
    if x.shape[-1] - size <= 0:
        return x
    return torch.stack([x[b, :, int(w1[b]):int(w1[b]) + size, int(h1[b]):int(h1[b]) + size] for b in range(x.shape[0])])


def random_shift(x, dy, dx, pad=4):
    """augmentations.py:229-233: replicate-pad then integer crop at (dy,dx) in [0,2*pad]."""
    h, w = x.shape[-2:]
    xp = F.pad(x, (pad, pad, pad, pad), mode="replicate")
    return torch.stack([xp[b, :, int(dy[b]):int(dy[b]) + h, int(dx[b]):int(dx[b]) + w] for b in range(x.shape[0])])


def random_overlay_carla(x, pool, ids, alpha):
    """augmentations.py:65-99 with dataset == 'carla'; pool uint8 (N,3,84,84)."""
    imgs = pool[torch.as_tensor(np.asarray(ids), dtype=torch.long)].repeat(1, 3, 1, 1)
    imgs = imgs / 255.0
    return ((1 - alpha) * (x / 255.0) + alpha * imgs) * 255.0


def random_overlay_places(x, imgs, alpha=0.2):
    """augmentations.py:79-99 default dataset (alpha_blending default 0.2; svea.py:26 passes none): imgs float (B,3,H,W) in [0,1]."""
    imgs = imgs.repeat(1, x.size(1) // 3, 1, 1)
    return ((1 - alpha) * (x / 255.0) + alpha * imgs) * 255.0


# --------------------------------------------------------------------------
# Adam (torch.optim.Adam single-tensor path, amsgrad=False, weight_decay=0)
# --------------------------------------------------------------------------
class Adam:
    def __init__(self, names, lr, beta1, beta2=0.999, eps=1e-8):
        self.names, self.lr, self.b1, self.b2, self.eps = list(names), lr, beta1, beta2, eps
        self.m, self.v, self.t = {}, {}, {}

    def step(self, p, grads):
        for n in self.names:
            g = grads.get(n)
            if g is None:
                continue                                  # params without .grad are skipped
            if n not in self.m:
                self.m[n] = torch.zeros_like(p[n]); self.v[n] = torch.zeros_like(p[n]); self.t[n] = 0
            self.t[n] += 1
            t = self.t[n]
            self.m[n].lerp_(g, 1 - self.b1)
            self.v[n].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            step_size = self.lr / bc1
            denom = (self.v[n].sqrt() / (bc2 ** 0.5)).add_(self.eps)
            p[n].addcdiv_(self.m[n], denom, value=-step_size)


# --------------------------------------------------------------------------
# replay (utils.py:94-198) -- host numpy, idxs supplied
# --------------------------------------------------------------------------
class ReplayOracle:
    """Transitions as a frame ring: obs_i = frames[i:i+3], next_obs_i = frames[i+1:i+4]
    (FrameStack deque semantics, env/wrappers.py:240-304), or explicit stacks via add()."""

    def __init__(self, frames, actions, rewards, not_dones):
        self.frames = frames            # uint8 (N+3, 3, H, W)
        self.actions, self.rewards, self.not_dones = actions, rewards, not_dones

    def stacks(self, idxs):
        f = self.frames
        obs = np.stack([np.concatenate([f[i], f[i + 1], f[i + 2]], 0) for i in idxs])
        nxt = np.stack([np.concatenate([f[i + 1], f[i + 2], f[i + 3]], 0) for i in idxs])
        return obs, nxt

    def sample(self, idxs, crop=None):
        """utils.py:185-198 (crop=(w1,h1,w1n,h1n) for 100->84 buffers, identity at 84)."""
        obs, nxt = self.stacks(idxs)
        obs = torch.as_tensor(obs).float()
        nxt = torch.as_tensor(nxt).float()
        a = torch.as_tensor(self.actions[idxs]); r = torch.as_tensor(self.rewards[idxs])
        nd = torch.as_tensor(self.not_dones[idxs])
        if crop is not None:
            obs = random_crop(obs, crop[0], crop[1]); nxt = random_crop(nxt, crop[2], crop[3])
        return obs, a, r, nxt, nd

    def sample_curl(self, idxs, crop):
        """utils.py:142-156; crop = (w1p, h1p, w1, h1, w1n, h1n): pos, obs, next_obs crops (the reference draws them in this order)."""
        obs, a, r, nxt, nd = self.sample(idxs)
        pos = random_crop(obs, crop[0], crop[1])
        return random_crop(obs, crop[2], crop[3]), a, r, random_crop(nxt, crop[4], crop[5]), nd, pos

    def sample_drq(self, idxs, shift, pad=4):
        """utils.py:158-171; shift=(dy,dx,dyn,dxn)."""
        obs, a, r, nxt, nd = self.sample(idxs)
        return random_shift(obs, shift[0], shift[1], pad), a, r, random_shift(nxt, shift[2], shift[3], pad), nd


def synthetic_replay(capacity, action_dim=2, size=84, seed=0):
    """SURVEY.md 8d cfg 1/2 generator: frames U{0..255}, actions U(-1,1), rewards N(0,1), not_done 1."""
    rs = np.random.RandomState(seed)
    frames = rs.randint(0, 256, size=(capacity + 3, 3, size, size), dtype=np.uint8)
    actions = rs.uniform(-1, 1, size=(capacity, action_dim)).astype(np.float32)
    rewards = rs.randn(capacity, 1).astype(np.float32)
    not_dones = np.ones((capacity, 1), dtype=np.float32)
    return ReplayOracle(frames, actions, rewards, not_dones)


# --------------------------------------------------------------------------
# the agents
# --------------------------------------------------------------------------
class OracleSAC:
    """sac.py:21-169"""

    def __init__(self, obs_shape, action_shape, args, params=None, dense_std=None, seed=0, tf32=False):
        self.args = args
        self.tf32 = tf32                # False | True | "f64": convolutions with TF32-rounded operands (see conv2d)
        self.A = int(action_shape[0])
        self.p = params if params is not None else init_params(
            obs_shape, self.A, args, torch.Generator().manual_seed(seed), dense_std)
        self._make_target()
        self.log_alpha = torch.tensor(np.log(args.init_temperature))       # fp64, sac.py:56
        self.target_entropy = -float(np.prod(action_shape))
        names = list(self.p.keys())
        self.critic_names = [n for n in names if _in_group(n, CRITIC_GROUP)]
        self.actor_names = [n for n in names if _in_group(n, ACTOR_GROUP)]
        self.critic_opt = Adam(self.critic_names, args.critic_lr, args.critic_beta)
        self.actor_opt = Adam(self.actor_names, args.actor_lr, args.actor_beta)
        self.alpha_opt = Adam(["log_alpha"], args.alpha_lr, args.alpha_beta)
        self.training = True
        self.trace = {}

    def _make_target(self):
        for n in list(self.p.keys()):
            if _in_group(n, CRITIC_GROUP):
                self.p["t_" + n] = self.p[n].clone()                       # deepcopy, sac.py:54

    def train(self, training=True):
        self.training = training

    def eval(self):
        self.train(False)

    @property
    def alpha(self):
        return self.log_alpha.exp()

    # ---- acting (sac.py:86-105) ----
    def select_action(self, obs):
        x = torch.as_tensor(np.asarray(obs), dtype=torch.float32).unsqueeze(0)
        with torch.no_grad():
            mu, _, _, _ = actor_forward(self.p, x, self.args, compute_pi=False, compute_log_pi=False, tf32=self.tf32)
        return mu.numpy().flatten()

    def sample_action(self, obs, noise):
        x = torch.as_tensor(np.asarray(obs), dtype=torch.float32).unsqueeze(0)
        with torch.no_grad():
            _, pi, _, _ = actor_forward(self.p, x, self.args, noise=noise, compute_log_pi=False, tf32=self.tf32)
        return pi.numpy().flatten()

    # ---- pieces ----
    def _grad_params(self, names):
        gp = dict(self.p)
        for n in names:
            gp[n] = self.p[n].detach().clone().requires_grad_(True)
        return gp

    def target_q(self, reward, next_obs, not_done, noise):
        """sac.py:108-112"""
        with torch.no_grad():
            _, pa, log_pi, _ = actor_forward(self.p, next_obs, self.args, noise=noise, tf32=self.tf32)
            tq1, tq2 = critic_forward(self.p, next_obs, pa, target=True, tf32=self.tf32)
            tv = torch.min(tq1, tq2) - self.alpha.detach() * log_pi
            tq = reward + (not_done * self.args.discount * tv)
        self.trace.update(next_pi=pa, next_log_pi=log_pi, target_Q=tq, tQ1=tq1, tQ2=tq2)
        return tq

    def critic_loss(self, gp, obs, action, target_q, rnd):
        q1, q2 = critic_forward(gp, obs, action, tf32=self.tf32)
        self.trace.update(Q1=q1.detach(), Q2=q2.detach())
        return F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)

    def update_critic(self, obs, action, reward, next_obs, not_done, rnd, L=None, step=None):
        tq = self.target_q(reward, next_obs, not_done, rnd["noise_next"])
        gp = self._grad_params(self.critic_names)
        loss = self.critic_loss(gp, obs, action, tq, rnd)
        if L is not None:
            L.log("train_critic/loss", loss, step)
        grads = torch.autograd.grad(loss, [gp[n] for n in self.critic_names], allow_unused=True)
        g = {n: gr for n, gr in zip(self.critic_names, grads) if gr is not None}
        self.trace.update(critic_loss=loss.detach(), critic_grads=g)
        with torch.no_grad():
            self.critic_opt.step(self.p, g)

    def update_actor_and_alpha(self, obs, rnd, L=None, step=None):
        """sac.py:125-151"""
        live = [n for n in self.actor_names if not n.startswith("cnn.")]     # detach=True
        gp = self._grad_params(live)
        _, pi, log_pi, log_std = actor_forward(gp, obs, self.args, noise=rnd["noise_pi"], detach=True, tf32=self.tf32)
        aq1, aq2 = critic_forward(gp, obs, pi, detach=True, tf32=self.tf32)
        actor_loss = (self.alpha.detach() * log_pi - torch.min(aq1, aq2)).mean()
        if L is not None:
            L.log("train_actor/loss", actor_loss, step)
        grads = torch.autograd.grad(actor_loss, [gp[n] for n in live])
        g = dict(zip(live, grads))
        with torch.no_grad():
            self.actor_opt.step(self.p, g)
        la = self.log_alpha.detach().clone().requires_grad_(True)
        alpha_loss = (la.exp() * (-log_pi - self.target_entropy).detach()).mean()
        if L is not None:
            L.log("train_alpha/loss", alpha_loss, step)
            L.log("train_alpha/value", la.exp(), step)
        (ga,) = torch.autograd.grad(alpha_loss, la)
        self.trace.update(actor_loss=actor_loss.detach(), actor_grads=g, alpha_loss=alpha_loss.detach(),
                          alpha_grad=ga, pi=pi.detach(), log_pi=log_pi.detach(),
                          actor_Q1=aq1.detach(), actor_Q2=aq2.detach())
        with torch.no_grad():
            box = {"log_alpha": self.log_alpha}
            self.alpha_opt.step(box, {"log_alpha": ga})

    def soft_update_critic_target(self):
        """sac.py:153-158, utils.py:31-33"""
        a = self.args
        with torch.no_grad():
            for n in self.critic_names:
                tau = a.critic_tau if _in_group(n, TARGET_Q) else a.encoder_tau
                t = self.p["t_" + n]
                t.copy_(tau * self.p[n] + (1 - tau) * t)

    def update_from_batch(self, batch, rnd, L, step):
        """sac.py:160-169 after the sample."""
        obs, action, reward, next_obs, not_done = batch
        self.trace = dict(obs=obs, next_obs=next_obs)
        self.update_critic(obs, action, reward, next_obs, not_done, rnd, L, step)
        if step % self.args.actor_update_freq == 0:
            self.update_actor_and_alpha(obs, rnd, L, step)
        if step % self.args.critic_target_update_freq == 0:
            self.soft_update_critic_target()
        return self.trace

    # ---- reference interchange ----
    def load_reference_agent(self, agent):
        """Copy parameters from a live reference agent object (uses its state_dict keys)."""
        sds = {"actor": agent.actor.state_dict(), "critic": agent.critic.state_dict()}
        if hasattr(agent, "attribution_predictor"):
            sds["attribution_predictor"] = agent.attribution_predictor.state_dict()
        for extra in ("curl_head", "pad_head", "predictor"):
            if hasattr(agent, extra):
                sds[extra] = getattr(agent, extra).state_dict()
        for n, refs in _ref_key_map().items():
            mod, key = refs[0]
            if mod in sds:
                self.p[n] = sds[mod][key].detach().clone().float().contiguous()
        for n in list(self.p.keys()):
            if n.startswith("t_"):
                del self.p[n]
        tsd = agent.critic_target.state_dict()
        for n, refs in _ref_key_map().items():
            if _in_group(n, CRITIC_GROUP):
                self.p["t_" + n] = tsd[refs[0][1]].detach().clone().float().contiguous()
        self.log_alpha = agent.log_alpha.detach().clone()
        if hasattr(agent, "predictor_target"):                     # SODA: own EMA copy of the shared CNN + both SODAMLPs
            psd = agent.predictor_target.state_dict()
            for n, refs in _ref_key_map().items():
                if _in_group(n, SODA_GROUP):
                    key = refs[0][1] if refs[0][0] == "predictor" else refs[0][1]
                    self.p["st_" + n] = psd[key].detach().clone().float().contiguous()

    def state_dicts(self):
        out = {"actor": OrderedDict(), "critic": OrderedDict(), "attribution_predictor": OrderedDict()}
        for n, refs in _ref_key_map().items():
            if n in self.p:
                for mod, key in refs:
                    out[mod][key] = self.p[n]
        return out


class OracleSGSAC(OracleSAC):
    """sgsac.py:24-185"""

    def __init__(self, obs_shape, action_shape, args, params=None, dense_std=None, seed=0, overlay_pool=None, tf32=False):
        super().__init__(obs_shape, action_shape, args, params, dense_std, seed, tf32=tf32)
        self.aux_names = [n for n in self.p.keys() if _in_group(n, AUX_GROUP)]
        self.aux_opt = Adam(self.aux_names, args.aux_lr, args.aux_beta)
        self.quantile = args.sgqn_quantile
        self.pool = overlay_pool

    def critic_loss(self, gp, obs, action, target_q, rnd):
        """sgsac.py:59-74"""
        q1, q2 = critic_forward(gp, obs, action, tf32=self.tf32)
        loss = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)
        self.trace.update(Q1=q1.detach(), Q2=q2.detach())
        if self.args.consistency:
            obs_grad = compute_attribution(self.p, obs, action, tf32=self.tf32)
            mask = compute_attribution_mask(obs_grad, self.quantile)
            masked_obs = obs * mask
            lo, hi = obs.view(-1).min(), obs.view(-1).max()
            fill = lo + (hi - lo) * rnd["u"]                                  # random.uniform(lo, hi)
            masked_obs[mask < 1] = fill
            self.trace.update(own_masked_obs=masked_obs)
            if rnd.get("force_masked_obs") is not None:
                # tests only: continue from a given masked observation (the mask is a discontinuous function of the
                # attribution, so two evaluations that agree to rounding can differ in a few threshold pixels)
                masked_obs = rnd["force_masked_obs"]
            mq1, mq2 = critic_forward(gp, masked_obs, action, tf32=self.tf32)
            loss = loss + 0.5 * (F.mse_loss(q1, mq1) + F.mse_loss(q2, mq2))
            self.trace.update(obs_grad1=obs_grad, mask1=mask, masked_obs=masked_obs,
                              mQ1=mq1.detach(), mQ2=mq2.detach(), fill=fill)
        return loss

    def update_aux(self, obs, action, obs_grad, rnd, L=None, step=None):
        """sgsac.py:82-102,163-167 (attribution_augmentation's s_prime feeds logging only)."""
        mask = compute_attribution_mask(obs_grad, self.quantile)
        s_tilde = random_overlay_carla(obs.clone(), self.pool, rnd["overlay_ids"], self.args.alpha_blending)
        live = [n for n in self.aux_names if not n.startswith("fdec.")]
        gp = self._grad_params(live)
        logits = attribution_predictor_forward(gp, s_tilde.detach(), action.detach(), tf32=self.tf32)
        aux_loss = F.binary_cross_entropy_with_logits(logits, mask.float())
        grads = torch.autograd.grad(aux_loss, [gp[n] for n in live])
        g = dict(zip(live, grads))
        self.trace.update(s_tilde=s_tilde, aux_logits=logits.detach(), aux_loss=aux_loss.detach(), aux_grads=g)
        with torch.no_grad():
            self.aux_opt.step(self.p, g)
        if L is not None:
            L.log("train/aux_loss", aux_loss, step)

    def update_from_batch(self, batch, rnd, L, step):
        """sgsac.py:169-185 after the sample."""
        obs, action, reward, next_obs, not_done = batch
        self.trace = dict(obs=obs, next_obs=next_obs)
        self.update_critic(obs, action, reward, next_obs, not_done, rnd, L, step)
        obs_grad = compute_attribution(self.p, obs, action, tf32=self.tf32)
        mask = compute_attribution_mask(obs_grad, self.quantile)
        self.trace.update(obs_grad2=obs_grad, mask2=mask)
        if step % self.args.actor_update_freq == 0:
            self.update_actor_and_alpha(obs, rnd, L, step)
        if step % self.args.critic_target_update_freq == 0:
            self.soft_update_critic_target()
        if step % self.args.aux_update_freq == 0:
            self.update_aux(obs, action, obs_grad, rnd, L, step)
        return self.trace


class OracleSVEA(OracleSAC):
    """svea.py:12-63 (overlay imgs host-supplied as rnd['places'])."""

    def critic_loss(self, gp, obs, action, target_q, rnd):
        a, b = self.args.svea_alpha, self.args.svea_beta
        aug = random_overlay_places(obs.clone(), rnd["places"])
        self.trace.update(obs_aug=aug)
        if a == b:
            o2 = torch.cat([obs, aug], 0); a2 = torch.cat([action, action], 0); t2 = torch.cat([target_q, target_q], 0)
            q1, q2 = critic_forward(gp, o2, a2, tf32=self.tf32)
            self.trace.update(Q1=q1.detach(), Q2=q2.detach())
            return (a + b) * (F.mse_loss(q1, t2) + F.mse_loss(q2, t2))
        q1, q2 = critic_forward(gp, obs, action, tf32=self.tf32)
        loss = a * (F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q))
        q1a, q2a = critic_forward(gp, aug, action, tf32=self.tf32)
        self.trace.update(Q1=q1.detach(), Q2=q2.detach())
        return loss + b * (F.mse_loss(q1a, target_q) + F.mse_loss(q2a, target_q))


class OracleCURL(OracleSAC):
    """curl.py:11-57 (+ CURLHead, modules.py:264-281): SAC on random crops plus the contrastive auxiliary update."""

    def __init__(self, obs_shape, action_shape, args, params=None, dense_std=None, seed=0, tf32=False):
        super().__init__(obs_shape, action_shape, args, params, dense_std, seed, tf32=tf32)
        self.aux_names = [n for n in self.p.keys() if _in_group(n, CURL_GROUP)]
        self.aux_opt = Adam(self.aux_names, args.aux_lr, args.aux_beta)

    def update_curl(self, x, x_pos, L=None, step=None):
        """curl.py:27-43: z_a = critic encoder(x), z_pos = target encoder(x_pos) (no grad); logits = z_a W z_pos^T minus the
        row maximum; cross entropy against the diagonal."""
        gp = self._grad_params(self.aux_names)
        z_a = projection(gp, cnn_forward(gp, x, tf32=self.tf32), "critic_proj")
        with torch.no_grad():
            z_pos = projection(self.p, cnn_forward(self.p, x_pos, pre="t_cnn", tf32=self.tf32), "t_critic_proj")
        Wz = torch.matmul(gp["curl.W"], z_pos.T)
        logits = torch.matmul(z_a, Wz)
        logits = logits - torch.max(logits, 1)[0][:, None]
        labels = torch.arange(logits.shape[0]).long()
        curl_loss = F.cross_entropy(logits, labels)
        grads = torch.autograd.grad(curl_loss, [gp[n] for n in self.aux_names])
        g = dict(zip(self.aux_names, grads))
        self.trace.update(curl_logits=logits.detach(), aux_loss=curl_loss.detach(), aux_grads=g, z_a=z_a.detach(), z_pos=z_pos)
        with torch.no_grad():
            self.aux_opt.step(self.p, g)
        if L is not None:
            L.log("train/aux_loss", curl_loss, step)

    def update_from_batch(self, batch, rnd, L, step):
        """curl.py:45-57 after sample_curl (utils.py:142-156): batch = (obs, action, reward, next_obs, not_done, pos)."""
        obs, action, reward, next_obs, not_done, pos = batch
        self.trace = dict(obs=obs, next_obs=next_obs, pos=pos)
        self.update_critic(obs, action, reward, next_obs, not_done, rnd, L, step)
        if step % self.args.actor_update_freq == 0:
            self.update_actor_and_alpha(obs, rnd, L, step)
        if step % self.args.critic_target_update_freq == 0:
            self.soft_update_critic_target()
        if step % self.args.aux_update_freq == 0:
            self.update_curl(obs, pos, L, step)
        return self.trace


class OraclePAD(OracleSAC):
    """pad.py:11-63 (+ InverseDynamics, modules.py:284-303): SAC on random crops plus the inverse-dynamics auxiliary update
    through the shared CNN, PAD's own projection and a 3-layer MLP."""

    def __init__(self, obs_shape, action_shape, args, params=None, dense_std=None, seed=0, tf32=False):
        super().__init__(obs_shape, action_shape, args, params, dense_std, seed, tf32=tf32)
        self.aux_names = [n for n in self.p.keys() if _in_group(n, PAD_GROUP)]
        self.aux_opt = Adam(self.aux_names, args.aux_lr, args.aux_beta)

    def update_inverse_dynamics(self, obs, obs_next, action, L=None, step=None):
        """pad.py:39-49"""
        gp = self._grad_params(self.aux_names)
        h = projection(gp, cnn_forward(gp, obs, tf32=self.tf32), "pad_proj")
        h_next = projection(gp, cnn_forward(gp, obs_next, tf32=self.tf32), "pad_proj")
        pred_action = mlp3(gp, torch.cat([h, h_next], dim=1), "pad_mlp")
        pad_loss = F.mse_loss(pred_action, action)
        grads = torch.autograd.grad(pad_loss, [gp[n] for n in self.aux_names])
        g = dict(zip(self.aux_names, grads))
        self.trace.update(pad_pred=pred_action.detach(), aux_loss=pad_loss.detach(), aux_grads=g)
        with torch.no_grad():
            self.aux_opt.step(self.p, g)
        if L is not None:
            L.log("train/aux_loss", pad_loss, step)

    def update_from_batch(self, batch, rnd, L, step):
        """pad.py:51-63"""
        obs, action, reward, next_obs, not_done = batch
        super().update_from_batch(batch, rnd, L, step)
        if step % self.args.aux_update_freq == 0:
            self.update_inverse_dynamics(obs, next_obs, action, L, step)
        return self.trace


def soda_mlp(p, x, pre):
    """SODAMLP (modules.py:116-129), BatchNorm1d in training mode: batch mean / biased batch variance, eps 1e-5."""
    x = F.linear(x, p[f"{pre}.0.weight"], p[f"{pre}.0.bias"])
    x = F.batch_norm(x, None, None, p[f"{pre}.1.weight"], p[f"{pre}.1.bias"], training=True, eps=1e-5)
    return F.linear(F.relu(x), p[f"{pre}.3.weight"], p[f"{pre}.3.bias"])


class OracleSODA(OracleSAC):
    """soda.py:12-84: SAC on random crops plus the SODA consistency update -- predictor(overlay-augmented crop) against the EMA
    target encoder of another crop of the same frames, on a separately sampled batch (soda_batch_size)."""

    def __init__(self, obs_shape, action_shape, args, params=None, dense_std=None, seed=0, tf32=False):
        super().__init__(obs_shape, action_shape, args, params, dense_std, seed, tf32=tf32)
        self.aux_names = [n for n in self.p.keys() if _in_group(n, SODA_GROUP)]
        self.aux_opt = Adam(self.aux_names, args.aux_lr, args.aux_beta)
        for n in self.aux_names:                                    # predictor_target = deepcopy(predictor), soda.py:30
            self.p.setdefault("st_" + n, self.p[n].clone())

    def update_soda(self, x, aug_x, L=None, step=None):
        """soda.py:41-69 after the two crops and the overlay: x, aug_x (n,9,84,84)."""
        gp = self._grad_params(self.aux_names)
        h0 = soda_mlp(gp, soda_mlp(gp, cnn_forward(gp, aug_x, tf32=self.tf32), "soda_proj"), "soda_pred")
        with torch.no_grad():
            h1 = soda_mlp(self.p, cnn_forward(self.p, x, pre="st_cnn", tf32=self.tf32), "st_soda_proj")
        soda_loss = F.mse_loss(F.normalize(h0, p=2, dim=1), F.normalize(h1, p=2, dim=1))
        grads = torch.autograd.grad(soda_loss, [gp[n] for n in self.aux_names])
        g = dict(zip(self.aux_names, grads))
        self.trace.update(soda_h0=h0.detach(), soda_h1=h1, aux_loss=soda_loss.detach(), aux_grads=g)
        with torch.no_grad():
            self.aux_opt.step(self.p, g)
            tau = self.args.soda_tau                                # utils.soft_update_params(predictor, predictor_target, soda_tau)
            for n in self.aux_names:
                t = self.p["st_" + n]
                t.copy_(tau * self.p[n] + (1 - tau) * t)
        if L is not None:
            L.log("train/aux_loss", soda_loss, step)

    def update_from_batch(self, batch, rnd, L, step):
        """soda.py:71-84; rnd["soda_x"], rnd["soda_aug_x"]: the two crops of the separately sampled SODA batch (overlay applied)."""
        super().update_from_batch(batch, rnd, L, step)
        if step % self.args.aux_update_freq == 0:
            self.update_soda(rnd["soda_x"], rnd["soda_aug_x"], L, step)
        return self.trace


ALGOS = {"sac": OracleSAC, "rad": OracleSAC, "drq": OracleSAC, "svea": OracleSVEA, "sgsac": OracleSGSAC, "curl": OracleCURL,
         "pad": OraclePAD, "soda": OracleSODA}


def make_oracle(obs_shape, action_shape, args, **kw):
    """factory.py:22-23"""
    return ALGOS[args.algorithm](obs_shape, action_shape, args, **kw)
