"""Run-to-run spread of the free-running RAD parity case (tests/test_update_parity_gpu.py::test_rad_crop_and_actions_at_100):
prints the step-3 losses of N fresh runs next to the oracle's (python tools/flaky_rad.py [N]; SGQN_PDL=0 for plain stream order)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_update_parity_gpu as T

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for it in range(N):
    B, A = 4, 6
    agent, rb, orc, rep, args = T._mk(algorithm="rad", B=B, A=A, size=100, dense=None)
    rs = np.random.RandomState(4)
    L, Lo = T._L(), T._L()
    for step in (2, 3):
        idxs = rs.randint(0, 48, size=B); rnd = T._rnd(rs, B, A, "rad")
        offs = rs.randint(0, 16, size=(2, B, 2))
        batch = rep.sample(idxs, (offs[0, :, 0], offs[0, :, 1], offs[1, :, 0], offs[1, :, 1]))
        orc.update_from_batch(batch, rnd, Lo, step)
        T._supply(agent, idxs, rnd, offs)
        agent.update(rb, L, step)
    print(it, " ".join(f"{k.split('/')[-1]}:{float(L.rows[(s, k)]):.6f}/{float(v):.6f}" for (s, k), v in Lo.rows.items()), flush=True)
