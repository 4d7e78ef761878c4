"""Encoder pass of the ten 32->32 layers: one persistent chain launch vs ten per-layer launches, CUDA-graph replays
(run on the B200 box).  Usage: python tools/chain_timing.py [n_samples ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import sgqn_carla_b200 as S
    from sgqn_carla_b200.engine import _ptr
    from sgqn_carla_b200.layout import ENC_H, FEAT
    sizes = [int(x) for x in sys.argv[1:]] or [128, 256]
    for n in sizes:
        B = n // 2
        args = S.default_args(algorithm="sac", batch_size=B)
        agent = S.make_agent((9, 84, 84), (2,), args)
        eng = agent.engine
        obs = torch.randint(0, 256, (n, 9, 84, 84)).float().cuda()
        dfeat = (torch.randn(n, FEAT) * 1e-2).cuda()
        alg_f = sum(128.0 * n * (ENC_H[l - 1] ** 2 + ENC_H[l] ** 2) for l in range(1, 11))
        alg_d = sum(128.0 * n * ((ENC_H[l] + 2) ** 2 + 2 * ENC_H[l - 1] ** 2) for l in range(1, 11))
        for what in ("fwd", "dgrad"):
            for chain in (False, True):
                eng.chain = chain

                def body():
                    if what == "fwd":
                        eng._convs(rows_f, eng.wsS, eng.st)
                    else:
                        eng.overlap = False
                        eng.enc_bwd(_ptr(dfeat), n, eng.actS, B, 0, 1, False)
                rows_f = []
                for l in range(1, 11):
                    hi, ho = ENC_H[l - 1], ENC_H[l]
                    last = l == 10
                    rows_f.append((_ptr(eng.actS[l - 1], B * (hi + 2) * hi * 32), _ptr(eng.wf, (l - 1) * 9216), eng.P(f"cnn.{l}.bias"), 0,
                                   _ptr(eng.actS[l], B * (ho * ho if last else (ho + 2) * ho) * 32), 0, n, hi + 2, hi, ho, ho, 0,
                                   ho if last else ho + 2, ho, 0, 0, 0, 0, 0 if last else 3))
                eng.enc_fwd(_ptr(obs), n, eng.actS, B)
                body(); torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(5):
                        body()
                for _ in range(3):
                    g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    g.replay()
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / 50
                alg = alg_f if what == "fwd" else alg_d
                print(f"n={n} {what:5s} {'chain    ' if chain else 'per-layer'}: {us:8.1f} us  ({alg / us / 1e3:7.1f} GB/s algorithmic, "
                      f"{alg / us / 1e3 / 6541.5:.2f} of measured HBM)", flush=True)


if __name__ == "__main__":
    main()
