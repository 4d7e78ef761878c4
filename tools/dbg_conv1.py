import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_kernels_gpu import *
from dbg_pdl import timeit
for B in (128, 256):
    obs = torch.randint(0, 256, (B, 9, 84, 84)).float().to(DEV)
    w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
    wp = torch.zeros(32 * 96, device=DEV); K.conv1_weights_prep(P(w), P(wp), 0, ST())
    y = torch.zeros(B, 43, 41, 32, device=DEV); col = torch.zeros(B * 1681, 96, device=DEV)
    f0 = lambda: K.conv1_fused_tc(P(obs), P(wp), P(b), P(y), 0, B, 84, B, ST())
    f1 = lambda: K.conv1_fused_tc(P(obs), P(wp), P(b), P(y), P(col), B, 84, 0, ST())
    def old():
        K.conv1_im2col96(P(obs), P(col), B, 84, ST())
        K.conv_tcg_taps(P(col), P(wp), P(b), 0, P(y), B, 41, 41, 96, 32, 41, 41, 0, 43, 41, 0, 0, 0, 0, 3, 1, ST())
    print("B", B, "fused no col", round(timeit(f0), 1), "fused + col", round(timeit(f1), 1), "im2col + gemm", round(timeit(old), 1), flush=True)
