"""Which kernel families sit on the critical path of the update's CUDA graph?  For every family, the family's C-ABI calls are
replaced by no-ops BEFORE the graphs are captured and the step is timed again: `saved` = what the step would gain if the family
cost nothing (results are garbage - timing only).  Run on the B200 box: python tools/ablate_step.py [batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgqn_carla_b200 as S
from sgqn_carla_b200 import _lib
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
FAMILIES = {
    "none": (),
    "linear (all Linear fwd/dgrad/wgrad + colsum)": ("linear_", "colsum"),
    "linear_fwd*": ("linear_fwd",),
    "linear_dgrad*": ("linear_dgrad",),
    "linear_wgrad* + colsum": ("linear_wgrad", "colsum"),
    "ln_tanh + heads + losses": ("ln_tanh", "actor_head", "critic_loss", "actor_loss"),
    "conv_tc (32->32 fwd + dgrad)": ("conv_tc",),
    "conv_wgrad_tc": ("conv_wgrad_tc",),
    "conv1 (fwd + dgrad + wgrad)": ("conv1_",),
    "decoder (conv_tcg, gemm_wgrad_tcg, bce, phase)": ("conv_tcg", "gemm_wgrad_tcg", "bce_", "conv_phase", "conv_wgrad_tcg"),
    "adam + prep": ("adam", "alpha_adam", "conv_weights_prep", "conv1_weights_prep"),
    "saliency (mask, minmax, overlay, guided)": ("attribution_mask", "minmax", "overlay", "guided"),
    "gather + rng": ("replay_gather", "rng_"),
}


class NullLog:
    def log(self, *a, **k):
        pass


def run(prefixes):
    api = _lib.K
    saved = {}
    for full in _lib.SIGNATURES:
        n = full[len("sgqn_"):]
        if any(n.startswith(p) for p in prefixes):
            saved[n] = getattr(api, n)
            setattr(api, n, lambda *a, **k: 0)
    try:
        data = bench.synthetic_sized(20000, 2, 84)
        args = S.default_args(algorithm="sgsac", batch_size=B, sgqn_quantile=0.95, seed=1)
        ag = S.make_agent((9, 84, 84), (2,), args)
        frames, actions, rewards, not_dones, pool = data
        ag.set_overlay_pool(pool)
        rb = S.ReplayBuffer((9, 84, 84), (2,), 20000, B, frame_capacity=20008)
        rb.load_ring(frames, actions, rewards, not_dones)
        L = NullLog()
        step = 1
        for _ in range(6):
            ag.update(rb, L, step); step += 1
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            ag.update(rb, L, step); step += 1
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 40
    finally:
        for n, fn in saved.items():
            setattr(api, n, fn)


base = None
for name, pre in FAMILIES.items():
    ms = run(pre)
    if base is None:
        base = ms
    print(f"{name:55s} {ms:7.3f} ms/step   saved {base - ms:6.3f}", flush=True)
