import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *
B, Hl = 256, 41
x = tf32_round(F.relu(rnd(B, 32, Hl, Hl, seed=1)))
w = rnd(32, 32, 3, 3, seed=2, scale=0.1)
wf, wd = prep_w(w)
Hod = Hl - 2
dyp = torch.zeros(B, Hod + 4, Hod + 2, 32, device=DEV); dyp[:, 2:2 + Hod, :Hod] = tf32_round(rnd(B, Hod, Hod, 32, seed=4))
acth = rows_pad(x, 2); out = torch.zeros(B, Hl + 4, Hl + 2, 32, device=DEV)
for flags in (2 | (1 << 2), 2, 2 | (1 << 2), 2):
    K.conv_tc(P(dyp), P(wd), 0, P(acth), P(out), 0, B, Hod + 4, Hod + 2, Hl, Hl, -2, Hl + 4, Hl + 2, 2, 0, Hl + 2, Hl, flags, ST())
torch.cuda.synchronize()
print("ok")
