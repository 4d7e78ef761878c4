"""Per-ABI-call timing of one phase: record the calls, then time each one alone as a CUDA graph of 5 repeats."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import sgqn_carla_b200 as S
from sgqn_carla_b200 import _lib
import bench
B = 128
args = S.default_args(algorithm="sgsac", batch_size=B, sgqn_quantile=0.95, seed=1)
frames, actions, rewards, not_dones, pool = bench.synthetic(4000, 2, seed=0)
ag = S.make_agent((9, 84, 84), (2,), args)
ag.set_overlay_pool(pool)
rb = S.ReplayBuffer((9, 84, 84), (2,), 4000, B, frame_capacity=4008)
rb.load_ring(frames, actions, rewards, not_dones)
L = bench.NullLog()
for s in range(1, 7):
    ag.update(rb, L, s)
torch.cuda.synchronize()
eng = ag.engine
eng.overlap = False
which = sys.argv[1] if len(sys.argv) > 1 else "aux"
phase = {"aux": eng.update_aux, "critic": lambda: eng.update_critic(1), "actor": eng.update_actor_and_alpha,
         "shared": lambda: eng.shared_obs_fwd(with_aux=True), "attr2": lambda: eng.attribution2(True)}[which]
api = _lib.K
recs = []
saved = {}
for full in _lib.SIGNATURES:
    n = full[5:]
    fn = getattr(api, n); saved[n] = fn
    def wrap(*a, _fn=fn, _n=n):
        recs.append((_n, a, _fn)); _fn(*a)
    setattr(api, n, wrap)
phase()
torch.cuda.synchronize()
for n, fn in saved.items():
    setattr(api, n, fn)
st = torch.cuda.Stream()
tot = 0.0
with torch.cuda.stream(st):
    for n, a, fn in recs:
        a = list(a); a[-1] = st.cuda_stream
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(5):
                fn(*a)
        ts = []
        for i in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 5)
        t = sorted(ts)[2]; tot += t
        ints = [x for x in a[:-1] if isinstance(x, int) and abs(x) < (1 << 31)]
        print(f"{n:22s} {t:8.1f} us  {ints[:16]}", flush=True)
print("sum", round(tot, 1), "us over", len(recs), "calls")
