"""BASELINE.json configurations at their full sizes (B200): first-update losses against the oracle where the CPU oracle
finishes in seconds, and size-independent properties (mask cardinality, masked-observation structure, gather round trip,
parameter-replica equality) everywhere else."""
import numpy as np
import pytest
import torch

from test_update_parity_gpu import _L, _mk, _rnd, _supply

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [128, 256])
def test_sgsac_full_batch_first_update(B):
    """configs[1] (B=128) and configs[2] (CARLA-shaped 9x84x84, A=2, B=256): one even update, reference init, product (tf32) path."""
    A = 2
    agent, rb, orc, rep, args = _mk(B=B, dense=None, cap=2 * B, precision="tf32")
    eng = agent.engine
    rs = np.random.RandomState(3)
    idxs = rs.randint(0, 2 * B, size=B); rnd = _rnd(rs, B, A, "sgsac")
    L, Lo = _L(), _L()
    orc.update_from_batch(rep.sample(idxs), rnd, Lo, 2)       # the TF32-rounding oracle (_mk): ~1 s (B=128) / ~2 s (B=256) of CPU
    _supply(agent, idxs, rnd)
    agent.update(rb, L, 2)
    torch.cuda.synchronize()
    vals = {k: float(v) for (s, k), v in L.rows.items()}
    assert set(vals) == {"train_critic/loss", "train_actor/loss", "train_alpha/loss", "train_alpha/value", "train/aux_loss"}
    assert all(np.isfinite(v) for v in vals.values()), vals
    for (s, k), v in Lo.rows.items():                         # rel 1e-3 (+ 2e-3 abs for the losses computed after the critic's Adam step)
        np.testing.assert_allclose(vals[k], float(v), rtol=1e-3, atol=1e-5 if k == "train_critic/loss" else 2e-3, err_msg=k)
    # structure of the saliency products (rl_utils.py:76-82, sgsac.py:67-70), any size.  Checked after an ODD step: on
    # even steps the overlay-augmented s_tilde re-uses the masked rows once the critic gradients are out.
    _supply(agent, idxs, rnd)
    agent.update(rb, L, 3)
    torch.cuda.synchronize()
    obs, masked = eng.obs2[:B], eng.obs2[B:]
    torch.testing.assert_close(obs.cpu(), rep.sample(idxs)[0])                      # gather is bit-exact
    m = eng.mask.reshape(B, 3, 84 * 84)
    kept = m.sum(-1)
    assert int(kept.min()) >= 353 and int(kept.max()) <= 7056                      # Q=0.95 keeps >= 353 px / frame (more on ties)
    lo, hi = float(obs.min()), float(obs.max())
    fill = np.float32(lo) + (np.float32(hi) - np.float32(lo)) * np.float32(rnd["u"])
    is_obs = masked == obs
    is_fill = masked == float(fill)
    assert bool((is_obs | is_fill).all())
    per_frame = is_obs.reshape(B, 3, 3, -1).all(2).sum(-1)                        # pixels whose 3 channels kept the observation
    assert int(per_frame.min()) >= 353


def test_svea_config5_first_update():
    B, A = 128, 6
    agent, rb, orc, rep, args = _mk(algorithm="svea", B=B, A=A, dense=None, cap=256, precision="tf32")
    rs = np.random.RandomState(4)
    idxs = rs.randint(0, 256, size=B); rnd = _rnd(rs, B, A, "svea")
    offs = rs.randint(0, 9, size=(2, B, 2))
    L, Lo = _L(), _L()
    orc.update_from_batch(rep.sample_drq(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1])), rnd, Lo, 2)
    _supply(agent, idxs, rnd, offs)
    agent.update(rb, L, 2)
    for (s, k), v in Lo.rows.items():
        np.testing.assert_allclose(float(L.rows[(s, k)]), float(v), rtol=1e-3, atol=1e-5 if k == "train_critic/loss" else 2e-3, err_msg=k)


def test_rad_config5_first_update():
    B, A = 128, 6
    agent, rb, orc, rep, args = _mk(algorithm="rad", B=B, A=A, dense=None, cap=160, size=100, precision="tf32")
    rs = np.random.RandomState(5)
    idxs = rs.randint(0, 160, size=B); rnd = _rnd(rs, B, A, "rad")
    offs = rs.randint(0, 16, size=(2, B, 2))
    L, Lo = _L(), _L()
    orc.update_from_batch(rep.sample(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1])), rnd, Lo, 2)
    _supply(agent, idxs, rnd, offs)
    agent.update(rb, L, 2)
    for (s, k), v in Lo.rows.items():
        np.testing.assert_allclose(float(L.rows[(s, k)]), float(v), rtol=1e-3, atol=1e-5 if k == "train_critic/loss" else 2e-3, err_msg=k)


def test_gather_roundtrip_large():
    """B=1024 (config 4's global batch) random gather out of a 4096-transition ring, with shift, bit-exact vs numpy."""
    import sgqn_carla_b200 as S
    from oracle import sgsac_oracle as O
    cap, B = 4096, 1024
    rep = O.synthetic_replay(cap, 2, seed=9)
    rb = S.ReplayBuffer((9, 84, 84), (2,), cap, B)
    rb.load_ring(rep.frames, rep.actions, rep.rewards, rep.not_dones)
    rs = np.random.RandomState(1)
    idxs = rs.randint(0, cap, size=B); offs = rs.randint(0, 9, size=(2, B, 2))
    got = rb.sample_drq(idxs=idxs, offs=offs)
    ref = rep.sample_drq(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1]))
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)
    # idempotence: a zero shift (offset == pad) is the plain sample
    z = np.full((2, B, 2), 4)
    assert torch.equal(rb.sample_drq(idxs=idxs, offs=z)[0], rb.sample(idxs=idxs)[0])


def test_sharded_batch_equals_full_batch_gradients():
    """Config 4 semantics on one GPU: two 'ranks' of 64 samples with losses scaled by 1/128 sum to the 128-sample gradient
    (the real all-reduce is exercised by bench.py --gpus N and tests/test_dist_cpu.py)."""
    import sgqn_carla_b200 as S
    from oracle import sgsac_oracle as O
    A, Bg = 2, 16
    args = S.default_args(algorithm="sac", batch_size=Bg)
    p0 = O.init_params((9, 84, 84), A, O.Args(**vars(args)), torch.Generator().manual_seed(5), dense_std=0.05)
    rep = O.synthetic_replay(64, A, seed=2)
    rs = np.random.RandomState(0)
    idxs = rs.randint(0, 64, size=Bg); rnd = _rnd(rs, Bg, A, "sac")

    def grads(batch, sl):
        a = S.default_args(algorithm="sac", batch_size=batch)
        ag = S.make_agent((9, 84, 84), (A,), a, precision="fp32", global_batch=Bg)
        ag.set_parameters(p0)
        rb = S.ReplayBuffer((9, 84, 84), (A,), 64, batch)
        rb.load_ring(rep.frames, rep.actions, rep.rewards, rep.not_dones)
        ag.supply(idxs=idxs[sl], noise_next=rnd["noise_next"][sl], noise_pi=rnd["noise_pi"][sl], u=rnd["u"])
        ag._draw(rb); ag._sample_into_engine(rb)
        ag.engine.update_critic(0)
        torch.cuda.synchronize()
        c0, c1 = ag.engine.lay.ranges["critic"]
        return ag.engine.grads[c0:c1].clone(), float(ag.engine.logs[0])

    gf, lf = grads(Bg, slice(0, Bg))
    g0, l0 = grads(Bg // 2, slice(0, Bg // 2))
    g1, l1 = grads(Bg // 2, slice(Bg // 2, Bg))
    assert abs((l0 + l1) - lf) <= 1e-4 * abs(lf)
    rel = float((g0 + g1 - gf).norm() / gf.norm())
    assert rel <= 2e-3, rel


class _FakeSync:
    """In-process stand-in for dist.GradSync when the ranks of a sharded step are looped on one GPU: the min / max exchange
    returns the known global pair, gradient sums are formed by the test."""

    def __init__(self, lo, hi):
        self.lo, self.hi, self.seen = lo, hi, []

    world = 2

    def all_reduce_sum(self, flat, group="main"):
        pass

    def all_reduce_minmax(self, mm, group="minmax"):
        self.seen.append(mm[:2].clone())
        mm[2], mm[3] = -self.lo, self.hi

    def all_reduce_logs(self, logs, group="main"):
        pass

    def log_scale(self, col):
        return 1.0


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_sgsac_sharded_equals_full_batch(precision):
    """BASELINE config 4 semantics for SGSAC: two shards whose observations span DIFFERENT ranges, the fill
    value built from the GLOBAL batch min / max (sgsac.py:68-70) and one shared u; every loss scaled by 1 / global batch.
    Masked observations of the shards are bit-identical to the corresponding rows of the full-batch run; critic, actor
    (incl. the alpha-weighted entropy term) and alpha gradients summed over the shards equal the full-batch gradients."""
    import sgqn_carla_b200 as S
    from oracle import sgsac_oracle as O
    A, Bg = 2, 16
    h = Bg // 2
    args = S.default_args(algorithm="sgsac", batch_size=Bg, sgqn_quantile=0.95)
    p0 = O.init_params((9, 84, 84), A, O.Args(**vars(args)), torch.Generator().manual_seed(5), dense_std=0.05)
    rep = O.synthetic_replay(64, A, seed=2)
    rep.frames[:36] = np.clip(rep.frames[:36], 20, 180)          # transitions 0..31 only see frames in [20, 180]
    rs = np.random.RandomState(0)
    idxs = np.concatenate([rs.randint(0, 32, size=h), rs.randint(36, 64, size=h)])     # shard 0: narrow range, shard 1: full range
    rnd = _rnd(rs, Bg, A, "sgsac")
    full_obs = torch.as_tensor(rep.stacks(idxs)[0]).float()
    lo, hi = float(full_obs.min()), float(full_obs.max())
    assert float(full_obs[:h].min()) > lo and float(full_obs[:h].max()) < hi

    def run(batch, sl, sync):
        a = S.default_args(algorithm="sgsac", batch_size=batch, sgqn_quantile=0.95)
        ag = S.make_agent((9, 84, 84), (A,), a, precision=precision, global_batch=Bg, dist=sync)
        ag.use_cuda_graphs = False
        ag.set_parameters(p0)
        rb = S.ReplayBuffer((9, 84, 84), (A,), 64, batch)
        rb.load_ring(rep.frames, rep.actions, rep.rewards, rep.not_dones)
        ag.supply(idxs=idxs[sl], noise_next=rnd["noise_next"][sl], noise_pi=rnd["noise_pi"][sl], u=rnd["u"])
        ag._draw(rb); ag._sample_into_engine(rb)
        eng = ag.engine
        eng.update_critic(1)
        torch.cuda.synchronize()
        lay = eng.lay
        c0, c1 = lay.ranges["critic"]
        out = dict(masked=eng.obs2[batch:].clone(), mask=eng.mask.clone(), gc=eng.grads[c0:c1].clone(), lc=float(eng.logs[0]))
        eng.shared_obs_fwd()
        eng.update_actor_and_alpha(finish=False)
        torch.cuda.synchronize()
        a0, a1 = lay.ranges["actor"]
        out.update(ga=eng.grads[a0:a1].clone(), la=float(eng.logs[1]), galpha=float(eng.alpha_grad))
        return out

    full = run(Bg, slice(0, Bg), None)
    s0, s1 = _FakeSync(lo, hi), _FakeSync(lo, hi)
    r0 = run(h, slice(0, h), s0)
    r1 = run(h, slice(h, Bg), s1)
    assert len(s0.seen) == 1 and s0.seen[0].tolist() != [lo, hi] and s1.seen[0].tolist() == [lo, hi]    # the exchange mattered
    if precision == "fp32":          # same kernels, same per-sample arithmetic: the shards' rows are the full batch's rows
        assert torch.equal(torch.cat([r0["mask"], r1["mask"]]), full["mask"])
        assert torch.equal(torch.cat([r0["masked"], r1["masked"]]), full["masked"])
    else:                            # tcgen05 tiles straddle samples differently at another batch size: rounding-level attributions
        agree = (torch.cat([r0["mask"], r1["mask"]]) == full["mask"]).float().mean()
        assert float(agree) >= 0.999
        same = (torch.cat([r0["mask"], r1["mask"]]) == full["mask"]).reshape(Bg, 3, 1, 84 * 84).expand(Bg, 3, 3, 84 * 84).reshape(Bg, 9, 84, 84)
        assert torch.equal(torch.cat([r0["masked"], r1["masked"]])[same], full["masked"][same])
    tol = 2e-3 if precision == "fp32" else 5e-2
    assert abs((r0["lc"] + r1["lc"]) - full["lc"]) <= 1e-3 * abs(full["lc"])
    assert abs((r0["la"] + r1["la"]) - full["la"]) <= 1e-3 * abs(full["la"]) + 1e-4
    assert abs((r0["galpha"] + r1["galpha"]) - full["galpha"]) <= 1e-4 * abs(full["galpha"]) + 1e-7
    for k in ("gc", "ga"):
        rel = float((r0[k] + r1[k] - full[k]).norm() / full[k].norm())
        assert rel <= tol, (k, rel)
