"""BASELINE config 4 on real ranks: two processes, one GPU each, NCCL or our own peer-memory collectives.  SGSAC with the batch sharded 8 + 8, shards with
DIFFERENT observation ranges, one shared u: after an even update (critic, actor, alpha and aux buckets, the min / max
exchange, every communicator of dist.GradSync) the replicas are bit-identical to each other and match a single-process
run of the full batch of 16; then graph-captured device-RNG updates keep the replicas bit-identical.  Skipped with < 2 GPUs."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
    import numpy as np, torch, torch.distributed as dist
    import sgqn_carla_b200 as S
    from sgqn_carla_b200.dist import GradSync, P2PGradSync
    from oracle import sgsac_oracle as O
    from oracle.pin_rnd import make_rnd
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p2p = os.environ.get("SGQN_TEST_SYNC") == "p2p"
    sync = P2PGradSync() if p2p else GradSync()
    if p2p:
        # 0. the peer-memory collectives themselves against NCCL: in-place two-shot sum of a sub-range (2 ranks: a + b is
        #    commutative -> bit-exact), untouched neighbours, repeated calls (ticket bookkeeping), the three small reductions
        n = 1 << 20
        g = sync.attach(n, torch.device("cuda", local))
        gen = torch.Generator(device="cuda").manual_seed(10 + rank)
        for it in range(6):
            x = torch.randn(n, device="cuda", generator=gen)
            g.copy_(x)
            ref = x.clone(); dist.all_reduce(ref)
            lo = 4096 * (it + 1)
            hi = lo + (300000 if it % 3 else 40000) + 4 * it        # small ranges on "main" take the one-barrier push form
            torch.cuda.synchronize(); dist.barrier()
            sync.all_reduce_sum(g[lo:hi], "early" if it % 2 else "main")
            torch.cuda.synchronize()
            assert torch.equal(g[lo:hi], ref[lo:hi]), it
            assert torch.equal(g[:lo], x[:lo]) and torch.equal(g[hi:], x[hi:]), it
            mm = torch.tensor([0.0, 0.0, -3.0 - rank - it, 7.0 + 2 * rank + it], device="cuda")
            sync.all_reduce_minmax(mm)
            logs = torch.arange(8, device="cuda", dtype=torch.float32) * (rank + 1) + it
            sync.all_reduce_logs(logs)
            ag64 = torch.tensor([1e-9 * (rank + 1) + it], dtype=torch.float64, device="cuda")
            sync.all_reduce_sum(ag64, "actor")
            torch.cuda.synchronize()
            assert mm.tolist() == [0.0, 0.0, -3.0 - it, 9.0 + it], mm.tolist()
            assert torch.equal(logs.cpu(), torch.arange(8, dtype=torch.float32) * 3 + 2 * it), logs
            assert float(ag64) == 1e-9 * 1 + it + (1e-9 * 2 + it), float(ag64)
        assert sync.timeouts() == 0
    A, Bg, h, cap = 2, 16, 8, 64
    oargs = O.Args(**vars(S.default_args(algorithm="sgsac", batch_size=Bg, sgqn_quantile=0.95)))
    p0 = O.init_params((9, 84, 84), A, oargs, torch.Generator().manual_seed(5), dense_std=None)
    rep = O.synthetic_replay(cap, A, seed=2)
    rep.frames[:36] = np.clip(rep.frames[:36], 20, 180)
    pool = torch.as_tensor(np.random.RandomState(7).randint(0, 256, size=(16, 3, 84, 84), dtype=np.uint8))
    rs = np.random.RandomState(0)
    idxs = np.concatenate([rs.randint(0, 32, size=h), rs.randint(36, 64, size=h)])
    rnd = make_rnd(rs, Bg, A, 16)

    def build(batch, sync_):
        a = S.default_args(algorithm="sgsac", batch_size=batch, sgqn_quantile=0.95)
        ag = S.make_agent((9, 84, 84), (A,), a, global_batch=Bg, dist=sync_)
        ag.set_parameters(p0); ag.set_overlay_pool(pool)
        rb = S.ReplayBuffer((9, 84, 84), (A,), cap, batch)
        rb.load_ring(rep.frames, rep.actions, rep.rewards, rep.not_dones)
        return ag, rb

    class Log:
        def __init__(self): self.rows = {}
        def log(self, k, v, step, n=1): self.rows[k] = v

    sl = slice(rank * h, (rank + 1) * h)
    ag, rb = build(h, sync)
    ag.engine.seed, ag.engine.seed_shared = 100 + rank, 100
    L = Log()
    ag.supply(idxs=idxs[sl], noise_next=rnd["noise_next"][sl], noise_pi=rnd["noise_pi"][sl], u=rnd["u"], overlay_ids=rnd["overlay_ids"][sl])
    ag.update(rb, L, 2)
    torch.cuda.synchronize()
    full, rbf = build(Bg, None)
    Lf = Log()
    full.supply(idxs=idxs, noise_next=rnd["noise_next"], noise_pi=rnd["noise_pi"], u=rnd["u"], overlay_ids=rnd["overlay_ids"])
    full.update(rbf, Lf, 2)
    torch.cuda.synchronize()
    # 2. logged losses are the global batch's
    for k, v in Lf.rows.items():
        np.testing.assert_allclose(float(L.rows[k]), float(v), rtol=2e-2, atol=2e-3, err_msg=k)
    # 3. parameters after the update: replicas bit-identical across ranks, and within the Adam bound of the full-batch run
    mine, ref = ag.get_parameters(), full.get_parameters()
    flat = torch.cat([t.reshape(-1).float() for t in mine.values()])
    both = [torch.empty_like(flat) for _ in range(2)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1]), "replicas diverged"
    for n in ref:
        d = (mine[n].double() - ref[n].double()).abs()
        assert float(d.max()) <= 2.1 * 1.3e-3 + 1e-6, (n, float(d.max()))
        assert float(d.mean()) <= 0.1 * 1.3e-3 + 4.2e-3 / d.numel(), (n, float(d.mean()))
    # 4. graph-captured updates with device RNG (own indices / noise per rank, shared u): replicas stay bit-identical
    for step in range(3, 9):
        ag.update(rb, L, step)
    torch.cuda.synchronize()
    assert len(ag._graphs) == 2, ag._graphs.keys()
    flat = torch.cat([t.reshape(-1).float() for t in ag.get_parameters().values()])
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1]) and bool(torch.isfinite(flat).all())
    us = [torch.empty(1, device="cuda") for _ in range(2)]
    dist.all_gather(us, ag.engine.u.clone())
    assert float(us[0]) == float(us[1])
    ix = [torch.empty(h, dtype=torch.int64, device="cuda") for _ in range(2)]
    dist.all_gather(ix, ag.engine.idxs.clone())
    assert not torch.equal(ix[0], ix[1])
    if p2p:
        assert sync.timeouts() == 0 and sync._offset(ag.engine.grads) is not None
    print("rank", rank, "ok", flush=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    os._exit(0)
""")


@pytest.mark.parametrize("sync", ["nccl", "p2p"])
def test_two_rank_sgsac_matches_full_batch(tmp_path, sync):
    """sync = nccl: dist.GradSync (NCCL communicators); p2p: dist.P2PGradSync (csrc/p2p.cu over NVLink peer memory)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", SGQN_TEST_SYNC=sync)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29741" if sync == "nccl" else "29743", str(script), ROOT], capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
