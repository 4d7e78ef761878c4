import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *
B, Hl = 2, 23
Ho = Hl - 2
dy = tf32_round(rnd(B, 32, Ho, Ho, seed=4))
act = tf32_round(F.relu(rnd(B, 32, Hl, Hl, seed=5)))
dyp = torch.zeros(B, Ho + 4, Ho + 2, 32, device=DEV)
dyh = nhwc(dy)
K.pad_copy(P(dyh), P(dyp), B, Ho, Ho, 32, Ho + 4, Ho + 2, 2, 0, 1, ST())
acth = rows_pad(act, 2)
dw = torch.zeros(9216, device=DEV)
K.conv_wgrad_tc(P(acth), P(dyp), P(dw), B, Hl + 2, Hl, ST())
torch.cuda.synchronize()
wr = torch.zeros(32, 32, 3, 3, device=DEV, dtype=torch.double, requires_grad=True)
F.conv2d(act.double(), wr).backward(dy.double())
ref = wr.grad.float().permute(0, 2, 3, 1).reshape(32, 9, 32)   # [co][tap][ci]
got = dw.reshape(32, 9, 32)
print("nonzero", int((got != 0).sum()), "nan", int(torch.isnan(got).sum()), "absmax", float(got.abs().max()), "ref absmax", float(ref.abs().max()))
print("got[0,:,0:4]\n", got[0, :, :4]); print("ref[0,:,0:4]\n", ref[0, :, :4])
for name, cand in [("identity", ref), ("swap co/ci", ref.permute(2, 1, 0)), ("flip taps", ref.flip(1)), ("swap+flip", ref.permute(2, 1, 0).flip(1))]:
    print(name, float((got - cand).norm() / cand.norm()))
for t in range(9):
    print("tap", t, "rel", float((got[:, t] - ref[:, t]).norm() / ref[:, t].norm()), "relT", float((got[:, t] - ref[:, t].t()).norm() / ref[:, t].norm()))
