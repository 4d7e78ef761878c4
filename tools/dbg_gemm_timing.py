import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.nn.functional as F
from test_kernels_gpu import *

def timeit(fn, n=20):
    """GPU time per call from a CUDA-graph replay of n back-to-back calls (no CPU launch overhead in the number)."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return sorted(ts)[len(ts) // 2]

for (M, N, K_, batch, relu) in [(128, 100, 14112, 1, 0), (256, 100, 14112, 1, 0), (128, 1024, 1024, 2, 1), (256, 1024, 1024, 2, 1), (128, 1024, 100, 1, 0)]:
    x, w, b = rnd(batch, M, K_, seed=1), rnd(batch, N, K_, seed=2, scale=0.05), rnd(batch, N, seed=3)
    y = torch.zeros(batch, M, N, device=DEV); dy = rnd(batch, M, N, seed=4); dx = torch.zeros(batch, M, K_, device=DEV)
    dw = torch.zeros(batch, N, K_, device=DEV); db = torch.zeros(batch, N, device=DEV)
    for name, tc in (("simt", ""), ("tc", "_tc")):
        f = getattr(K, "linear_fwd" + tc); d = getattr(K, "linear_dgrad" + tc); g = getattr(K, "linear_wgrad" + tc)
        t1 = timeit(lambda: f(P(x), K_, M * K_, P(w), N * K_, P(b), N, P(y), N, M * N, M, N, K_, relu, batch, 2, ST()))
        t2 = timeit(lambda: d(P(dy), N, M * N, P(w), N * K_, P(x), K_, M * K_, P(dx), K_, M * K_, M, N, K_, 1, 2, batch, ST()))
        t3 = timeit(lambda: g(P(x), K_, M * K_, P(dy), N, M * N, P(dw), N * K_, P(db), N, M, N, K_, relu, batch, ST()))
        print(f"M={M} N={N} K={K_} batch={batch} {name}: fwd {t1:.1f} us  dgrad {t2:.1f} us  wgrad {t3:.1f} us")

print("--- splitk=0 (plain store, no memset)")
for (M, N, K_, batch, relu) in [(128, 1024, 100, 1, 0), (128, 1024, 1024, 2, 1), (128, 100, 14112, 1, 0)]:
    x, w, b = rnd(batch, M, K_, seed=1), rnd(batch, N, K_, seed=2, scale=0.05), rnd(batch, N, seed=3)
    y = torch.zeros(batch, M, N, device=DEV)
    for name, tc in (("simt", ""), ("tc", "_tc")):
        f = getattr(K, "linear_fwd" + tc)
        t1 = timeit(lambda: f(P(x), K_, M * K_, P(w), N * K_, P(b), N, P(y), N, M * N, M, N, K_, relu, batch, 0, ST()))
        print(f"M={M} N={N} K={K_} batch={batch} {name}: fwd {t1:.1f} us")
def nop():
    K.zero(P(y), 4, ST())
print("zero kernel", timeit(nop))
