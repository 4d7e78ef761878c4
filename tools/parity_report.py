"""Measured spread of the CUDA path against the oracle (run on the B200 box: `python tools/parity_report.py`).

For every (algorithm, initialisation, precision) it runs three updates on host-supplied randomness and prints, per
update: the relative error of every logged loss, and the distance of the updated parameters in units of one Adam step
-- once against the oracle that rounds conv operands to TF32 where the tcgen05 kernels do (`tf32=True`: summation order
is the only difference) and once against the plain fp32 oracle (what the reference's CPU run computes).  The numbers
quoted in DESIGN.md ("measured spread") come from this script; tests/test_update_parity_gpu.py asserts the bars.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import sgqn_carla_b200 as S
    from oracle import sgsac_oracle as O
    from oracle.pin_rnd import make_rnd

    class Log:
        def __init__(self):
            self.rows = {}

        def log(self, k, v, step, n=1):
            self.rows[(step, k)] = float(v)

    B, A, cap = 8, 2, 48
    for algorithm in ("sgsac", "svea", "drq", "sac"):
        for dense in (0.05, None):
            for precision in ("tf32", "fp32"):
                args = S.default_args(algorithm=algorithm, batch_size=B, sgqn_quantile=0.95)
                oargs = O.Args(**vars(args))
                p0 = O.init_params((9, 84, 84), A, oargs, torch.Generator().manual_seed(11), dense_std=dense)
                pool = torch.as_tensor(np.random.RandomState(7).randint(0, 256, size=(16, 3, 84, 84), dtype=np.uint8))
                oracles = {}
                for name, tf in (("emu", precision == "tf32"), ("fp32", False)):
                    if name == "fp32" and precision == "fp32":
                        continue
                    o = O.make_oracle((9, 84, 84), (A,), oargs, params={k: v.clone() for k, v in p0.items()}, tf32=tf)
                    o.pool = pool
                    oracles[name] = (o, Log())
                agent = S.make_agent((9, 84, 84), (A,), args, precision=precision)
                agent.set_parameters(p0)
                if algorithm == "sgsac":
                    agent.set_overlay_pool(pool)
                rep = O.synthetic_replay(cap, A, seed=0)
                rb = S.ReplayBuffer((9, 84, 84), (A,), cap, B)
                rb.load_ring(rep.frames, rep.actions, rep.rewards, rep.not_dones)
                rs = np.random.RandomState(9)
                L = Log()
                print(f"== {algorithm} init={'dense' if dense else 'reference'} precision={precision}")
                for step in (2, 3, 4):
                    idxs = rs.randint(0, cap, size=B)
                    rnd = make_rnd(rs, B, A, 16, with_places=(algorithm == "svea"))
                    offs = None
                    if algorithm in ("svea", "drq"):
                        offs = rs.randint(0, 9, size=(2, B, 2))
                        batch = rep.sample_drq(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1]))
                    else:
                        batch = rep.sample(idxs)
                    for o, lo in oracles.values():
                        o.update_from_batch(batch, rnd, lo, step)
                    agent.supply(idxs=idxs, noise_next=rnd["noise_next"], noise_pi=rnd["noise_pi"], u=rnd["u"],
                                 overlay_ids=rnd["overlay_ids"], offs=offs, places=rnd.get("places"))
                    agent.update(rb, L, step)
                    torch.cuda.synchronize()
                    mine = agent.get_parameters()
                    for name, (o, lo) in oracles.items():
                        errs = {k.split("/")[-2][6:] + "/" + k.split("/")[-1]: abs(L.rows[(s, k)] - v) / (abs(v) + 1e-12)
                                for (s, k), v in lo.rows.items() if s == step}
                        dmax, dmean, worst = 0.0, 0.0, ""
                        for n, ref in o.p.items():
                            if n in mine:
                                d = (mine[n].cpu().double() - ref.double()).abs()
                                if float(d.max()) > dmax:
                                    dmax, worst = float(d.max()), n
                                dmean = max(dmean, float(d.mean()))
                        print(f"  step {step} vs {name:4s}: loss relerr " + " ".join(f"{k}={e:.1e}" for k, e in errs.items())
                              + f" | param max {dmax / 1e-3:.3f} lr ({worst}) worst-mean {dmean / 1e-3:.4f} lr")


if __name__ == "__main__":
    main()
