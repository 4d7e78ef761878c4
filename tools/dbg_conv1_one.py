import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_kernels_gpu import *
B = 256
obs = torch.randint(0, 256, (B, 9, 84, 84)).float().to(DEV)
w = rnd(32, 9, 3, 3, seed=2, scale=0.2); b = rnd(32, seed=3)
wp = torch.zeros(32 * 96, device=DEV); K.conv1_weights_prep(P(w), P(wp), 0, ST())
y = torch.zeros(B, 43, 41, 32, device=DEV)
for _ in range(4):
    K.conv1_fused_tc(P(obs), P(wp), P(b), P(y), 0, B, 84, B, ST())
torch.cuda.synchronize()
print("ok")
