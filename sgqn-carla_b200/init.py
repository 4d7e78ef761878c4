"""Parameter initialisation with the reference's rules: `weight_init` (modules.py:53-67: orthogonal Linear,
delta-orthogonal Conv2d with relu gain, zero bias) for the encoder / projections / actor / critic, torch's
default Conv2d / Linear init for the AttributionPredictor (modules.py:315-354 never applies weight_init)."""
import math
from collections import OrderedDict

import torch

from .layout import FEAT


def init_params(action_dim, args, seed=None):
    g = torch.Generator()
    g.manual_seed(int(seed if seed is not None else getattr(args, "seed", 0)))
    H, P, nf, A = int(args.hidden_dim), int(args.projection_dim), int(args.num_filters), int(action_dim)
    p = OrderedDict()

    def ortho(rows, cols, gain=1.0):
        w = torch.empty(rows, cols)
        torch.nn.init.orthogonal_(w, gain, generator=g)
        return w

    def default(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(*shape, generator=g) * 2 - 1) * b

    for i in range(int(args.num_shared_layers)):
        cin = 9 if i == 0 else nf
        w = torch.zeros(nf, cin, 3, 3)
        w[:, :, 1, 1] = ortho(nf, cin, math.sqrt(2.0))
        p[f"cnn.{i}.weight"], p[f"cnn.{i}.bias"] = w, torch.zeros(nf)
    for pre in ("critic_proj", "actor_proj"):
        p[f"{pre}.0.weight"], p[f"{pre}.0.bias"] = ortho(P, FEAT), torch.zeros(P)
        p[f"{pre}.1.weight"], p[f"{pre}.1.bias"] = torch.ones(P), torch.zeros(P)
    for q in ("Q1", "Q2"):
        p[f"{q}.0.weight"], p[f"{q}.0.bias"] = ortho(H, P + A), torch.zeros(H)
        p[f"{q}.2.weight"], p[f"{q}.2.bias"] = ortho(H, H), torch.zeros(H)
        p[f"{q}.4.weight"], p[f"{q}.4.bias"] = ortho(1, H), torch.zeros(1)
    p["actor_mlp.0.weight"], p["actor_mlp.0.bias"] = ortho(H, P), torch.zeros(H)
    p["actor_mlp.2.weight"], p["actor_mlp.2.bias"] = ortho(H, H), torch.zeros(H)
    p["actor_mlp.4.weight"], p["actor_mlp.4.bias"] = ortho(2 * A, H), torch.zeros(2 * A)
    p["dec.proj.weight"], p["dec.proj.bias"] = default((FEAT, P + A), P + A), default((FEAT,), P + A)
    p["dec.conv1.weight"], p["dec.conv1.bias"] = default((128, 32, 3, 3), 288), default((128,), 288)
    p["dec.conv2.weight"], p["dec.conv2.bias"] = default((64, 128, 3, 3), 1152), default((64,), 1152)
    p["dec.conv3.weight"], p["dec.conv3.bias"] = default((9, 64, 3, 3), 576), default((9,), 576)
    p["fdec.0.weight"], p["fdec.0.bias"] = default((256, 100), 100), default((256,), 100)
    p["fdec.2.weight"], p["fdec.2.bias"] = default((100, 256), 256), default((100,), 256)
    if getattr(args, "algorithm", "") == "soda":
        # SODAMLP / SODAPredictor .apply(weight_init) (modules.py:126,309): orthogonal Linear, default BatchNorm (1, 0).  The
        # predictor's apply() also re-initialises the SHARED CNN it wraps -- after critic_target was deep-copied from the critic
        # (sac.py:54) -- so the critic target keeps the first draw and the SODA target (deepcopy of the predictor) gets the second.
        for pre, fan in (("soda_proj", FEAT), ("soda_pred", P)):
            p[f"{pre}.0.weight"], p[f"{pre}.0.bias"] = ortho(P, fan), torch.zeros(P)
            p[f"{pre}.1.weight"], p[f"{pre}.1.bias"] = torch.ones(P), torch.zeros(P)
            p[f"{pre}.3.weight"], p[f"{pre}.3.bias"] = ortho(P, P), torch.zeros(P)
        for n in [k for k in p if k.startswith(("cnn.", "critic_proj.", "Q1.", "Q2."))]:
            p["t_" + n] = p[n].clone()
        for k in range(int(args.num_shared_layers)):
            cin = 9 if k == 0 else nf
            w = torch.zeros(nf, cin, 3, 3)
            w[:, :, 1, 1] = ortho(nf, cin, math.sqrt(2.0))
            p[f"cnn.{k}.weight"] = w
    if getattr(args, "algorithm", "") == "pad":               # InverseDynamics.apply(weight_init), modules.py:296
        p["pad_proj.0.weight"], p["pad_proj.0.bias"] = ortho(P, FEAT), torch.zeros(P)
        p["pad_proj.1.weight"], p["pad_proj.1.bias"] = torch.ones(P), torch.zeros(P)
        p["pad_mlp.0.weight"], p["pad_mlp.0.bias"] = ortho(H, 2 * P), torch.zeros(H)
        p["pad_mlp.2.weight"], p["pad_mlp.2.bias"] = ortho(H, H), torch.zeros(H)
        p["pad_mlp.4.weight"], p["pad_mlp.4.bias"] = ortho(A, H), torch.zeros(A)
    if getattr(args, "algorithm", "") == "curl":
        p["curl.W"] = torch.rand(P, P, generator=g)           # modules.py:268
    return p
