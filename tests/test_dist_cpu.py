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
