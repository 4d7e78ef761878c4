"""The AttributionDecoder's tcgen05 launches at the benchmark size (B samples), each timed as CUDA-graph replays of 10 back-to-back
launches (python tools/tcg_micro.py [B] [only]); under ncu: `ncu --set full -k regex:tcg -c 8 python tools/tcg_micro.py 128 conv3_fwd`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgqn_carla_b200 as S  # noqa: F401
from sgqn_carla_b200._lib import K
from sgqn_carla_b200.engine import _ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
only = sys.argv[2] if len(sys.argv) > 2 else None
dev = "cuda"


def buf(*shape):
    return torch.randn(*shape, device=dev) * 0.1


xin1, xin2, xin3 = buf(B, 23, 23, 32), buf(B, 23, 23, 128), buf(B, 44, 44, 64)
lgp, dlgp, dd2s, dd1g, ddl, dl = buf(B, 44, 44, 64), buf(B, 44, 44, 64), buf(B, 23, 23, 256), buf(B, 23, 23, 128), buf(B, 21, 21, 32), buf(B, 21, 21, 32)
w1, w2, w3 = buf(128 * 9 * 32), buf(256 * 9 * 128), buf(64 * 9 * 64)
b1, b2, b3 = buf(128), buf(256), buf(64)
dw3p, dw2p, dw1 = buf(64 * 9 * 64), buf(256 * 9 * 128), buf(128 * 9 * 32)
R = 3


def st():
    return torch.cuda.current_stream().cuda_stream


CASES = {
    "conv1_fwd  32->128 @21": lambda: K.conv_tcg(_ptr(xin1), _ptr(w1), _ptr(b1), 0, _ptr(xin2), B, 23, 23, 32, 128, 21, 21, -1, 23, 23, 1, 0, 0, 0, R, st()),
    "conv2_fwd 128->256 @21": lambda: K.conv_tcg(_ptr(xin2), _ptr(w2), _ptr(b2), 0, _ptr(xin3), B, 23, 23, 128, 256, 21, 21, -1, 44, 44, 1, 0, 0, 0, R | (1 << 5), st()),
    "conv3_fwd  64->64  @42": lambda: K.conv_tcg(_ptr(xin3), _ptr(w3), _ptr(b3), 0, _ptr(lgp), B, 44, 44, 64, 64, 42, 42, -1, 44, 44, 1, 0, 0, 0, 0, st()),
    "conv3_dgrad 64->64 @42": lambda: K.conv_tcg(_ptr(dlgp), _ptr(w3), 0, _ptr(xin3, 44 * 64), _ptr(dd2s), B, 44, 44, 64, 64, 42, 42, -1, 23, 23, 1, 0, 44, 44, (1 << 2) | 2 | (2 << 5), st()),
    "conv2_dgrad 256->128 @21": lambda: K.conv_tcg(_ptr(dd2s), _ptr(w2), 0, _ptr(xin2, 23 * 128), _ptr(dd1g), B, 23, 23, 256, 128, 21, 21, -1, 23, 23, 1, 0, 23, 23, (1 << 2) | 2, st()),
    "conv1_dgrad 128->32 @21": lambda: K.conv_tcg(_ptr(dd1g), _ptr(w1), 0, _ptr(dl), _ptr(ddl), B, 23, 23, 128, 32, 21, 21, -1, 21, 21, 0, 0, 21, 21, 1 << 2, st()),
    "conv3_wgrad 64x64 @42": lambda: K.conv_wgrad_tcg(_ptr(xin3), _ptr(dlgp), _ptr(dw3p), B, 44, 44, 64, 64, -1, -1, st()),
    "conv2_wgrad 128x128 @21 (x2)": lambda: [K.conv_wgrad_tcg_ld(_ptr(xin2), _ptr(dd2s, 128 * h), 256, _ptr(dw2p, h * 128 * 9 * 128), B, 23, 23, 128, 128, -1, -1, st()) for h in range(2)],
    "conv1_wgrad 32x128 @21": lambda: K.conv_wgrad_tcg(_ptr(xin1), _ptr(dd1g), _ptr(dw1), B, 23, 23, 32, 128, -1, -1, st()),
}

for name, fn in CASES.items():
    if only and not name.startswith(only):
        continue
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} {name:30s}: {e0.elapsed_time(e1) * 20:.1f} us per launch", flush=True)
