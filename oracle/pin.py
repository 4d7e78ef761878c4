"""TEST INFRASTRUCTURE ONLY.  Drives the unmodified reference (through ref_shim) and the
restated oracle on identical inputs / identical host-supplied randomness."""
import tempfile

import numpy as np
import torch

from . import ref_shim as R
from . import sgsac_oracle as O


from .pin_rnd import make_rnd  # noqa: E402,F401


def densify(agent, std, gen):
    """'dense' weight variant: N(0,std) everywhere (LayerNorm weight 1+N) so masks are non-degenerate."""
    seen = set()
    with torch.no_grad():
        for mod in (agent.actor, agent.critic, getattr(agent, "attribution_predictor", None)):
            if mod is None:
                continue
            for n, p in mod.named_parameters():
                if id(p) in seen:
                    continue
                seen.add(id(p))
                t = torch.randn(p.shape, generator=gen) * std
                if n.endswith("projection.1.weight"):
                    t = t + 1.0
                p.copy_(t)
        agent.critic_target.load_state_dict(agent.critic.state_dict())


def build_pair(algorithm="sgsac", B=4, A=2, hidden_dim=1024, quantile=0.95, dense_std=None, capacity=32,
               pool_n=16, seed=0, size=84, extra_args=()):
    """Returns (reference agent, reference replay buffer, oracle agent, oracle replay, args)."""
    ns = R.load()
    args = R.parse_args(["--algorithm", algorithm, "--sgqn_quantile", str(quantile), "--hidden_dim", str(hidden_dim),
                         "--batch_size", str(B), "--log_dir", tempfile.mkdtemp(), *extra_args])
    torch.manual_seed(seed)
    agent = ns["make_agent"]((9, size, size), (A,), args)
    if dense_std is not None:
        densify(agent, dense_std, torch.Generator().manual_seed(seed + 1))
    rep = O.synthetic_replay(capacity, A, size=size, seed=seed)
    utils = ns["utils"]
    rb = utils.ReplayBuffer((9, size, size), (A,), capacity, B, prefill=False)
    for i in range(capacity):
        f = rep.frames
        rb.add(utils.LazyFrames([f[i], f[i + 1], f[i + 2]]), rep.actions[i], rep.rewards[i, 0],
               utils.LazyFrames([f[i + 1], f[i + 2], f[i + 3]]), False)
    pool = torch.as_tensor(np.random.RandomState(seed + 7).randint(0, 256, size=(pool_n, 3, 84, 84), dtype=np.uint8))
    R.TAPE.pool = pool
    oargs = O.Args(**{k: v for k, v in vars(args).items()})
    orc = O.make_oracle((9, size, size), (A,), oargs, seed=seed)
    if isinstance(orc, O.OracleSGSAC):
        orc.pool = pool
    orc.load_reference_agent(agent)
    return agent, rb, orc, rep, args


def ref_step(agent, rb, idxs, rnd, step, algorithm="sgsac", crop=None):
    """One reference update with the tape loaded in the reference's call order."""
    T = R.TAPE
    T.idxs[:] = [idxs]
    T.noise[:] = [rnd["noise_next"], rnd["noise_pi"]]
    T.u[:] = [rnd["u"]]
    T.overlay[:] = [rnd["overlay_ids_unused"], rnd["overlay_ids"]]   # attribution_augmentation first (sgsac.py:84), then random_overlay (:88)
    T.places[:] = [rnd["places"]] if "places" in rnd else []
    T.crop[:] = list(crop) if crop is not None else []
    L = R.NullLogger()
    if algorithm == "sgsac":
        agent.update(rb, L, step, 0)
    else:
        agent.update(rb, L, step)
    return {k: v for k, v, _ in L.rows}


def ref_params(agent):
    """canonical name -> tensor view of the live reference parameters (+ target, log_alpha)."""
    sds = {"actor": agent.actor.state_dict(), "critic": agent.critic.state_dict(),
           "critic_target": agent.critic_target.state_dict()}
    for extra in ("attribution_predictor", "curl_head", "pad_head", "predictor", "predictor_target"):
        if hasattr(agent, extra):
            sds[extra] = getattr(agent, extra).state_dict()
    out = {}
    for n, refs in O._ref_key_map().items():
        mod, key = refs[0]
        if mod in sds:
            out[n] = sds[mod][key]
        if O._in_group(n, O.CRITIC_GROUP):
            out["t_" + n] = sds["critic_target"][key]
        if "predictor_target" in sds and O._in_group(n, O.SODA_GROUP):
            out["st_" + n] = sds["predictor_target"][key if mod == "predictor" else refs[0][1]]
    out["log_alpha"] = agent.log_alpha.detach()
    return out
