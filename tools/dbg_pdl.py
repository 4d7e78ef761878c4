import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *

def timeit(fn, n=20, graph=True):
    fn(); torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(n):
                    fn()
        run = g.replay
    else:
        def run():
            for _ in range(n):
                fn()
    ts = []
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return sorted(ts)[3]

if __name__ == "__main__":
  for B, Hp in ((128, 41), (256, 41), (128, 23), (256, 23)):
      x = tf32_round(F.relu(rnd(B, 32, Hp, Hp, seed=1)))
      w = rnd(32, 32, 3, 3, seed=2, scale=0.1); b = rnd(32, seed=3)
      wf, wd = prep_w(w)
      xh = rows_pad(x, 2); Ho = Hp - 2
      y = torch.zeros(B, Ho + 2, Ho, 32, device=DEV)
      f = lambda: K.conv_tc(P(xh), P(wf), P(b), 0, P(y), 0, B, Hp + 2, Hp, Ho, Ho, 0, Ho + 2, Ho, 0, 0, 0, 0, 3, ST())
      print("PDL", os.environ.get("SGQN_PDL"), "B", B, "H", Hp, "fwd us/launch graph", round(timeit(f), 2), "eager", round(timeit(f, graph=False), 2), flush=True)
