"""Which collectives of the batch-sharded update cost step time?  Under torchrun: times the update with the collectives of one
communicator at a time turned into no-ops (timing only - the replicas drift apart).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ablate_coll.py"""
import datetime, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgqn_carla_b200 as S
from sgqn_carla_b200.dist import GradSync, P2PGradSync
import bench

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
B = 128


Base = P2PGradSync if os.environ.get("SGQN_P2P", "1") == "1" else GradSync


class SkipSync(Base):
    skip = frozenset()

    def all_reduce_sum(self, flat, group="main"):
        if flat.dtype == torch.float64 and "alpha" in self.skip:
            return
        if group + "_tiny" in self.skip:            # keep the launch and the barriers, move (almost) no data
            return super().all_reduce_sum(flat[:4], group)
        if group not in self.skip:
            super().all_reduce_sum(flat, group)

    def all_reduce_minmax(self, mm, group="minmax"):
        if group not in self.skip:
            super().all_reduce_minmax(mm, group)

    def all_reduce_logs(self, logs, group="main"):
        if "logs" not in self.skip:
            super().all_reduce_logs(logs, group)


class NullLog:
    def log(self, *a, **k):
        pass


sync = SkipSync()
data = bench.synthetic_sized(20000, 2, 84)
if len(sys.argv) > 1:
    sync.ctas = int(sys.argv[1])
for skip in ((), ("early_tiny",), ("early",), ("actor",), ("alpha",), ("main",), ("minmax",), ("logs",), ("early", "actor", "alpha", "main", "minmax", "logs")):
    SkipSync.skip = frozenset(skip)
    args = S.default_args(algorithm="sgsac", batch_size=B, sgqn_quantile=0.95, seed=1 + rank)
    ag = S.make_agent((9, 84, 84), (2,), args, dist=sync, global_batch=B * world)
    ag.engine.seed, ag.engine.seed_shared = 1234 + rank, 1234
    ag.sync_from_rank0()
    frames, actions, rewards, not_dones, pool = data
    ag.set_overlay_pool(pool)
    rb = S.ReplayBuffer((9, 84, 84), (2,), 20000, B, frame_capacity=20008)
    rb.load_ring(frames, actions, rewards, not_dones)
    L, step = NullLog(), 1
    for _ in range(6):
        ag.update(rb, L, step); step += 1
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        ag.update(rb, L, step); step += 1
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 40], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} {Base.__name__} without {','.join(skip) or '-':32s} {float(t):.3f} ms/step", flush=True)
    del ag, rb
dist.destroy_process_group()
