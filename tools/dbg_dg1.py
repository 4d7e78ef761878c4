import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "scratch"))
import torch
from test_kernels_gpu import *
from dbg_pdl import timeit
for B in (128, 256):
    d = tf32_round(rnd(B, 41, 41, 32, seed=1))
    w = rnd(32, 9, 3, 3, seed=3, scale=0.2)
    wp = torch.zeros(32 * 96, device=DEV); wd = torch.zeros(96 * 32, device=DEV)
    K.conv1_weights_prep(P(w), P(wp), P(wd), ST())
    d1 = torch.zeros(B, 9, 84, 84, device=DEV)
    print("dbg", os.environ.get("SGQN_DG_DEBUG"), "B", B, "conv1_dgrad_fused us", round(timeit(lambda: K.conv1_dgrad_fused_tc(P(d), P(wd), P(d1), B, ST())), 1), flush=True)
