import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import sgqn_carla_b200 as S
import bench
from dbg_pdl import timeit
B = 128
args = S.default_args(algorithm="sgsac", batch_size=B, sgqn_quantile=0.95, seed=1)
frames, actions, rewards, not_dones, pool = bench.synthetic(20000, 2, seed=0)
ag = S.make_agent((9, 84, 84), (2,), args)
ag.set_overlay_pool(pool)
rb = S.ReplayBuffer((9, 84, 84), (2,), 20000, B, storage="pinned", frame_capacity=20008)
rb.load_ring(frames, actions, rewards, not_dones)
L = bench.NullLog()
for s in range(1, 5):
    ag.update(rb, L, s)
torch.cuda.synchronize()
pf = ag._pf
print("prefetch issue alone us:", round(timeit(lambda: ag._prefetch_issue(rb, pf), n=5), 1))
eng = ag.engine
print("direct pinned gather us:", round(timeit(lambda: ag._sample_into_engine(rb), n=5), 1))
def run(agent, rb, n=100, s0=5):
    for s in range(s0, s0 + 10): agent.update(rb, L, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(s0 + 10, s0 + 10 + n): agent.update(rb, L, s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("pinned + prefetch  ms/update", round(run(ag, rb), 4))
ag.prefetch = False; ag._graphs.clear(); ag._eager_runs.clear()
print("pinned no prefetch ms/update", round(run(ag, rb), 4))
rbd = S.ReplayBuffer((9, 84, 84), (2,), 20000, B, frame_capacity=20008)
rbd.load_ring(frames, actions, rewards, not_dones)
print("device ring        ms/update", round(run(ag, rbd), 4))
