"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference/src, through oracle/ref_shim.py) on CPU in this container.

    python -m oracle.make_golden            # from the repo root

The fixtures hold only *outputs of the reference* (+ the seeds that regenerate the
inputs deterministically: numpy RandomState / seeded torch.Generator draws are
machine-independent).  Parameters are the 'dense' N(0,0.05) variant because the
reference's orthogonal init goes through LAPACK QR, which is not guaranteed
bit-stable across hosts; the reference-init case is pinned live in
tests/test_oracle_pin.py whenever /root/reference is mounted.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pin, ref_shim as R, sgsac_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def digest(params, n_samples=16, seed=123):
    """per-tensor (sum, abs-sum, sampled elements) in float64."""
    import zlib
    d = {}
    for n in sorted(params):
        rs = np.random.RandomState((seed + zlib.crc32(n.encode())) % (2 ** 31))
        t = params[n].detach().double().reshape(-1).numpy()
        idx = rs.randint(0, t.size, size=min(n_samples, t.size))
        d[n] = np.concatenate([[t.sum(), np.abs(t).sum()], t[idx]])
    return d


def golden_update(algorithm, name, B=2, A=2, steps=(2, 3, 4), quantile=0.95, extra=()):
    agent, rb, orc, rep, args = pin.build_pair(algorithm, B=B, A=A, dense_std=0.05, quantile=quantile, extra_args=extra)
    ns = R.load()
    out = dict(algorithm=algorithm, B=B, A=A, steps=np.array(steps), quantile=quantile)
    # parameters must be regenerable on any host: overwrite with the oracle's seeded dense init
    p0 = O.init_params((9, 84, 84), A, O.Args(**vars(args)), torch.Generator().manual_seed(1234), dense_std=0.05)
    load_canonical(agent, p0)
    rs = np.random.RandomState(5)
    for step in steps:
        idxs = rs.randint(0, 32, size=B)
        rnd = pin.make_rnd(rs, B, A, 16, with_places=(algorithm == "svea"))
        crop = None
        if algorithm in ("svea", "drq"):
            crop = [(rs.randint(0, 9, size=B), rs.randint(0, 9, size=B)) for _ in range(2)]
        if algorithm == "sgsac":
            import algorithms.rl_utils as ru
            obs, action, _, _, _ = rep.sample(idxs)
            g = ru.compute_attribution(agent.critic, obs, action)
            m = ru.compute_attribution_mask(g, quantile)
            out[f"s{step}_attr"] = g.numpy()
            out[f"s{step}_mask"] = np.packbits(m.numpy().reshape(-1))
        logs = pin.ref_step(agent, rb, idxs, rnd, step, algorithm, crop=crop)
        for k, v in logs.items():
            out[f"s{step}_log_{k}"] = v
        for n, v in digest(pin.ref_params(agent)).items():
            out[f"s{step}_p_{n}"] = v
    x = rep.sample(np.array([3]))[0][0].numpy().astype(np.uint8)
    out["select_action"] = agent.select_action(x)
    R.TAPE.noise[:] = [torch.full((1, A), 0.3)]
    out["sample_action"] = agent.sample_action(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: v for k, v in out.items() if "_log_" in k})


def load_canonical(agent, p):
    sds = {"actor": agent.actor, "critic": agent.critic, "attribution_predictor": getattr(agent, "attribution_predictor", None)}
    with torch.no_grad():
        for n, refs in O._ref_key_map().items():
            for mod, key in refs:
                if sds.get(mod) is not None:
                    sds[mod].state_dict()[key].copy_(p[n])
        agent.critic_target.load_state_dict(agent.critic.state_dict())


def golden_masks():
    """compute_attribution_mask (rl_utils.py:76-82) on crafted attributions: ties, zeros, quantiles."""
    R.load()
    import algorithms.rl_utils as ru
    out = {}
    g = torch.Generator().manual_seed(77)
    base = torch.randn(4, 9, 84, 84, generator=g)
    base[1, :3] = 0.0                                  # all-zero frame -> whole frame kept
    base[2, 3:6] = (torch.rand(3, 84, 84, generator=g) < 0.03).float() * base[2, 3:6]   # 97% zeros (ties at 0)
    base[3, 6:9] = torch.round(base[3, 6:9] * 2) / 2   # heavy ties
    for q in (0.5, 0.9, 0.95, 0.98, 0.999):
        m = ru.compute_attribution_mask(base, q)
        out[f"q{q}"] = np.packbits(m.numpy().reshape(-1))
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)
    print("wrote masks")


def golden_aug():
    """random_crop (augmentations.py:236-264), random_shift (:229-233), random_overlay carla (:79-99)."""
    ns = R.load()
    aug = ns["augmentations"]
    rs = np.random.RandomState(11)
    x100 = torch.as_tensor(rs.randint(0, 256, size=(3, 9, 100, 100)).astype(np.float32))
    w1 = torch.as_tensor(rs.randint(0, 16, size=3)); h1 = torch.as_tensor(rs.randint(0, 16, size=3))
    crop = aug.random_crop(x100, 84, w1, h1)
    x84 = torch.as_tensor(rs.randint(0, 256, size=(3, 9, 84, 84)).astype(np.float32))
    dy = rs.randint(0, 9, size=3); dx = rs.randint(0, 9, size=3)
    R.TAPE.crop[:] = [(dy, dx)]
    shift = aug.random_shift(x84, 4)
    pool = torch.as_tensor(rs.randint(0, 256, size=(8, 3, 84, 84), dtype=np.uint8))
    ids = rs.randint(0, 8, size=3)
    R.TAPE.pool = pool; R.TAPE.overlay[:] = [ids]
    ov = aug.random_overlay(x84.clone(), "carla", 0.2)
    np.savez_compressed(os.path.join(OUT, "aug.npz"), w1=w1.numpy(), h1=h1.numpy(), crop=crop.numpy().astype(np.uint8),
                        dy=dy, dx=dx, shift=shift.numpy().astype(np.uint8), ids=ids, overlay=ov.numpy())
    print("wrote aug")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_masks()
    golden_aug()
    golden_update("sgsac", "sgsac_dense")
    golden_update("svea", "svea_dense", steps=(2, 3))
    golden_update("sac", "sac_dense", steps=(2, 3))
    golden_update("drq", "drq_dense", steps=(2, 3))
