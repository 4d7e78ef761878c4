import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *
torch.backends.cudnn.allow_tf32 = False
for (B, Hs) in [(2, 23), (8, 23), (16, 23), (16, 25), (16, 41), (32, 25)]:
    Cin = Cs = Co = 32; pad = 0
    x = rnd(B, Cin, Hs, Hs, seed=1); w = rnd(Co, Cin, 3, 3, seed=2, scale=0.1)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    xin = F.relu(xr); xin.retain_grad()
    yr = F.conv2d(xin, wr)
    Ho = yr.shape[-1]
    dy = rnd(B, Co, Ho, Ho, seed=4)
    yr.backward(dy)
    xh, wsk, dyh = nhwc(x), wk(w), nhwc(dy)
    dw = torch.zeros(Cs * 9 * Cin, device=DEV); db = torch.zeros(Cs, device=DEV)
    K.conv_wgrad(P(xh), P(dyh), P(dw), P(db), B, Hs, Hs, Cin, Cs, pad, 1, 1, ST())
    e1 = (dw.reshape(Cs, 3, 3, Cin).permute(0, 3, 1, 2) - wr.grad).norm() / wr.grad.norm()
    dxl = torch.zeros(B * Hs * Hs * Cin, device=DEV)
    K.conv_dgrad(P(dyh), P(wsk), P(xh), P(dxl), B, Hs, Hs, Cin, Cs, pad, 1, ST())
    ref = xin.grad * (x > 0)
    e2 = (nchw(dxl, B, Hs, Hs, Cin) - ref).norm() / ref.norm()
    e3 = (db - dy.sum((0, 2, 3))).norm() / dy.sum((0, 2, 3)).norm()
    print(B, Hs, "wgrad rel", float(e1), "dgrad rel", float(e2), "bgrad rel", float(e3))
