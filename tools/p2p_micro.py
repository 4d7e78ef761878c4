"""All-reduce of n floats: own peer-memory kernel (csrc/p2p.cu) vs NCCL, device-timed, idle GPU (under torchrun, 2+ ranks)."""
import datetime, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgqn_carla_b200 as S  # noqa: F401
from sgqn_carla_b200.dist import P2PGradSync

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
sync = P2PGradSync()
g = sync.attach(4 << 20, torch.device("cuda", local))
x = torch.randn(4 << 20, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for n in (1 << 10, 96 << 10, 1400 << 10, 2300 << 10):
    for ctas in (16, 32, 74, 148):
        sync.ctas = ctas
        t = timed(lambda: sync.all_reduce_sum(g[:n], "main"))
        if rank == 0:
            print(f"n={n * 4 / 1e6:7.3f} MB  p2p ctas={ctas:3d}: {t:7.1f} us", flush=True)
    t = timed(lambda: dist.all_reduce(x[:n]))
    if rank == 0:
        print(f"n={n * 4 / 1e6:7.3f} MB  nccl        : {t:7.1f} us", flush=True)
mm = torch.zeros(4, device="cuda")
t = timed(lambda: sync.all_reduce_minmax(mm))
t2 = timed(lambda: dist.all_reduce(mm[2:4], op=dist.ReduceOp.MAX))
if rank == 0:
    print(f"min/max pair: p2p {t:.1f} us, nccl {t2:.1f} us; time-outs {sync.timeouts()}", flush=True)
dist.destroy_process_group()
