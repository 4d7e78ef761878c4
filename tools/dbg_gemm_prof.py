import sys, os, ctypes
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_kernels_gpu import *
lib = ctypes.CDLL(os.path.join(R, "sgqn-carla_b200", "libsgqn_b200.so"))
buf = (ctypes.c_ulonglong * 8)()
for (M, N, K_, batch, relu, split) in [(128, 1024, 100, 1, 0, 0), (128, 1024, 1024, 2, 1, 2)]:
    x, w, b = rnd(batch, M, K_, seed=1), rnd(batch, N, K_, seed=2, scale=0.05), rnd(batch, N, seed=3)
    y = torch.zeros(batch, M, N, device=DEV)
    for i in range(3):
        K.linear_fwd_tc(P(x), K_, M * K_, P(w), N * K_, P(b), N, P(y), N, M * N, M, N, K_, relu, batch, split, ST())
        lib.sgqn_debug_gt_prof(buf)
        t = list(buf)
        print((M, N, K_), "ns since entry: setup", t[1]-t[0], "first full", t[2]-t[0], "first conv", t[3]-t[0], "done", t[4]-t[0], "epi end", t[5]-t[0], "exit", t[6]-t[0])
