"""SVEA critic-stage gradient diagnostic (tf32 product path vs the TF32-emulating oracle); run on the B200 box."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_update_parity_gpu import _mk, _rnd, _supply, _relerr, _batch_for  # noqa: E402

for precision in ("tf32", "fp32"):
    B, A = 8, 2
    agent, rb, orc, rep, args = _mk(algorithm="svea", B=B, dense=0.05, precision=precision)
    eng = agent.engine
    rs = np.random.RandomState(5)
    idxs, offs, batch = _batch_for("svea", rep, rs, B)
    rnd = _rnd(rs, B, A, "svea")
    orc.trace = {}
    tq = orc.target_q(batch[2], batch[3], batch[4], rnd["noise_next"])
    gp = orc._grad_params(orc.critic_names)
    loss = orc.critic_loss(gp, batch[0], batch[1], tq, rnd)
    grads = dict(zip(orc.critic_names, torch.autograd.grad(loss, [gp[n] for n in orc.critic_names])))
    _supply(agent, idxs, rnd, offs)
    agent._draw(rb); agent._sample_into_engine(rb)
    eng.update_critic(2)
    torch.cuda.synchronize()
    got = eng.lay.unpack(eng.grads)
    print("==", precision, "loss", float(eng.logs[0]), float(loss), "aug err", _relerr(eng.obs2[B:], orc.trace["obs_aug"]),
          "obs eq", torch.equal(eng.obs2[:B].cpu(), batch[0]))
    print("  Q1", _relerr(eng.q[0, :2 * B], orc.trace["Q1"][:, 0]), "tq", _relerr(eng.target_q, orc.trace["target_Q"][:, 0]))
    for n, gr in grads.items():
        print(f"   {n:24s} relerr {_relerr(got[n], gr):.2e} |g| {float(gr.norm()):.2e}")
