import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *
B, Hp = 128, 41
x = tf32_round(F.relu(rnd(B, 32, Hp, Hp, seed=1)))
w = rnd(32, 32, 3, 3, seed=2, scale=0.1); b = rnd(32, seed=3)
wf, wd = prep_w(w)
xh = rows_pad(x, 2); Ho = Hp - 2
y = torch.zeros(B, Ho + 2, Ho, 32, device=DEV)
# dgrad operands (layer geometry Hl = 41 -> dY 39x39 padded)
Hl = 41; Hod = Hl - 2
dyp = torch.zeros(B, Hod + 4, Hod + 2, 32, device=DEV); dyp[:, 2:2 + Hod, :Hod] = tf32_round(rnd(B, Hod, Hod, 32, seed=4))
acth = rows_pad(x, 2); out = torch.zeros(B, Hl + 4, Hl + 2, 32, device=DEV)

def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return sorted(ts)[2]

f = lambda: K.conv_tc(P(xh), P(wf), P(b), 0, P(y), 0, B, Hp + 2, Hp, Ho, Ho, 0, Ho + 2, Ho, 0, 0, 0, 0, 3, ST())
d = lambda: K.conv_tc(P(dyp), P(wd), 0, P(acth), P(out), 0, B, Hod + 4, Hod + 2, Hl, Hl, -2, Hl + 4, Hl + 2, 2, 0, Hl + 2, Hl, 2 | (1 << 2), ST())
print("debug", os.environ.get("SGQN_TC_DEBUG"), "fwd us/launch", round(timeit(f), 2), "dgrad us/launch", round(timeit(d), 2))
