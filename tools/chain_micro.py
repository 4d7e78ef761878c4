"""One 32->32 layer (41x41 -> 39x39, n samples): sgqn_conv_tc vs sgqn_conv_chain with a 1-layer list (run on the B200 box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgqn_carla_b200 as S
from sgqn_carla_b200._lib import K, conv_layers
from sgqn_carla_b200.engine import _ptr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = torch.randn(n * 43 * 41 * 32, device="cuda")
outs = [torch.zeros(n * 43 * 41 * 32, device="cuda") for _ in range(nl)]
w = torch.randn(10 * 9216, device="cuda") * 0.05
b = torch.zeros(32, device="cuda")
ws = torch.zeros(4 + n * 16 * nl + 64, dtype=torch.int32, device="cuda")
H = [41, 39, 37, 35, 33, 31, 29, 27, 25, 23, 21]
rows = []
src = x
for l in range(nl):
    hi, ho = H[l], H[l + 1]
    rows.append((_ptr(src), _ptr(w, l * 9216), _ptr(b), 0, _ptr(outs[l]), 0, n, hi + 2, hi, ho, ho, 0, ho + 2, ho, 0, 0, 0, 0, 3))
    src = outs[l]


def per_layer():
    st = torch.cuda.current_stream().cuda_stream
    for r in rows:
        K.conv_tc(*r, st)


def chain():
    arr, addr, k = conv_layers(rows)
    K.conv_chain(addr, k, _ptr(ws), ws.numel(), torch.cuda.current_stream().cuda_stream)


for name, fn in (("conv_tc", per_layer), ("chain", chain)):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} layers={nl} {name:8s}: {e0.elapsed_time(e1) * 10:.1f} us per pass", flush=True)
