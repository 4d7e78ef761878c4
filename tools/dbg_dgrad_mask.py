import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "scratch"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *
from dbg_pdl import timeit
for B, Hl in ((128, 41), (256, 41), (128, 25), (256, 25)):
    x = tf32_round(F.relu(rnd(B, 32, Hl, Hl, seed=1)))
    w = rnd(32, 32, 3, 3, seed=2, scale=0.1)
    wf, wd = prep_w(w)
    Hod = Hl - 2
    dyp = torch.zeros(B, Hod + 4, Hod + 2, 32, device=DEV); dyp[:, 2:2 + Hod, :Hod] = tf32_round(rnd(B, Hod, Hod, 32, seed=4))
    acth = rows_pad(x, 2); out = torch.zeros(B, Hl + 4, Hl + 2, 32, device=DEV); db = torch.zeros(32, device=DEV)
    def call(flags, dbias):
        return lambda: K.conv_tc(P(dyp), P(wd), 0, P(acth), P(out), P(db) if dbias else 0, B, Hod + 4, Hod + 2, Hl, Hl, -2, Hl + 4, Hl + 2, 2, 0, Hl + 2, Hl, flags, ST())
    print("B", B, "H", Hl, "dgrad masked", round(timeit(call(2 | (1 << 2), False)), 2), "masked+dbias", round(timeit(call(2 | (1 << 2), True)), 2),
          "no mask (16 warps)", round(timeit(call(2, False)), 2), "no mask + dbias", round(timeit(call(2, True)), 2), flush=True)
