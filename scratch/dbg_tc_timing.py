import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *
B, Hp = 256, 41
x = tf32_round(F.relu(rnd(B, 32, Hp, Hp, seed=1)))
w = rnd(32, 32, 3, 3, seed=2, scale=0.1); b = rnd(32, seed=3)
wf, wd = prep_w(w)
xh = rows_pad(x, 2); Ho = Hp - 2
y = torch.zeros(B, Ho + 2, Ho, 32, device=DEV)
big = torch.zeros(64 << 20, device=DEV)
ts = []
for i in range(8):
    big.add_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.conv_tc(P(xh), P(wf), P(b), 0, P(y), 0, B, Hp + 2, Hp, Ho, Ho, 0, Ho + 2, Ho, 0, 0, 0, 0, 3, ST())
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("debug", os.environ.get("SGQN_TC_DEBUG"), "us:", [round(t, 1) for t in ts])

