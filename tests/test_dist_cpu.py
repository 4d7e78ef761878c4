"""Data-parallel glue on CPU: world_size-2 gloo run of sgqn-carla_b200/dist.py (the engine itself needs a GPU)."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, importlib.util
    import torch, torch.distributed as dist
    spec = importlib.util.spec_from_file_location("gs", os.path.join(sys.argv[1], "sgqn-carla_b200", "dist.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    dist.init_process_group("gloo")
    r, w = dist.get_rank(), dist.get_world_size()
    s = m.GradSync()
    # gradient buckets: shard-local sums scaled by 1/B_global add up to the full-batch gradient
    g = torch.Generator().manual_seed(0)
    per_sample = torch.randn(8, 1000, generator=g)                 # same on both ranks
    local = per_sample[r * 4:(r + 1) * 4].sum(0) / 8.0
    s.all_reduce_sum(local)
    assert torch.allclose(local, per_sample.mean(0), atol=1e-6)
    lo, hi = 3.0 + r, 200.0 - 10 * r
    mm = torch.tensor([lo, hi, -lo, hi])                          # what sgqn_minmax writes
    s.all_reduce_minmax(mm)
    assert mm.tolist() == [lo, hi, -3.0, 200.0], mm               # the second pair is global {-min, max}
    logs = torch.tensor([1.0 + r, 2.0, 3.0, 0.1, 4.0, 0, 0, 0])
    s.all_reduce_logs(logs)
    assert abs(float(logs[0]) * s.log_scale(0) - 3.0) < 1e-6 and abs(float(logs[3]) * s.log_scale(3) - 0.1) < 1e-6
    # one communicator per issuing stream; each one reduces independently
    for name in s.GROUPS:
        t = torch.tensor([1.0 + r])
        s.all_reduce_sum(t, name)
        assert float(t) == 3.0, (name, t)
    assert len({id(g) for g in s.groups.values()}) == len(s.GROUPS)
    # P2PGradSync host logic (the kernels need GPUs: tests/test_dist_gpu.py): which slot / form a call takes, and what falls back
    # to the communicators.  A CPU tensor stands in for the symmetric arena, a recorder for the C ABI.
    calls = []
    class Rec:
        def p2p_allreduce_sum(self, *a): calls.append(("sum",) + a)
        def p2p_small(self, *a): calls.append(("small",) + a)
    p = m.P2PGradSync()
    p._k = lambda: Rec()
    p._stream = lambda: 0
    head = 8192
    p.arena = torch.zeros(head + (1 << 20)); p.bases = None
    p.flags_off, p.ctl_off, p.small_off, p.stage_off, p.data_off = 0, 64, 128, 4096, 4 * head
    g = p.arena[head:]
    S = p.SLOTS
    p.all_reduce_sum(g[1000:1000 + 96000], "main")                 # 0.38 MB on main: push form, staging area given
    p.all_reduce_sum(g[0:400000], "main")                          # 1.6 MB on main: two-shot on its own slot, no staging
    p.all_reduce_sum(g[4000:4000 + 96000], "early")                # other slots: always two-shot
    assert [c[6] for c in calls] == [S["main"], S["main_big"], S["early"]], calls
    assert [c[10] for c in calls] == [p.stage_off, -1, -1] and calls[0][7] == 4 * (head + 1000) and calls[0][8] == 96000
    t = g[2:2 + 96000].clone(); t.fill_(1.0 + r)                   # not inside the arena -> the communicator (gloo here)
    p.all_reduce_sum(t, "main")
    assert float(t[0]) == 3.0 and len(calls) == 3
    u = g[2:2 + 96000]; u.fill_(1.0 + r)                           # inside, but not 16-byte aligned -> the communicator
    p.all_reduce_sum(u, "main")
    assert float(u[0]) == 3.0 and len(calls) == 3
    p.all_reduce_sum(torch.zeros(1, dtype=torch.float64), "actor")
    p.all_reduce_minmax(torch.zeros(4)); p.all_reduce_logs(torch.zeros(8))
    assert [(c[0], c[7], c[10], c[11]) for c in calls[3:]] == [("small", S["alpha"], 1, 2), ("small", S["minmax"], 2, 1), ("small", S["logs"], 8, 0)]
    dist.destroy_process_group()
    print("rank", r, "ok")
""")


def test_gradsync_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script), ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2
