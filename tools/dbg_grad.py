import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_update_parity_gpu import _mk, _rnd, _supply, _relerr
from oracle import sgsac_oracle as O
B, A = 8, 2
for feed, prec in ((False, 'fp32'), (False, 'tf32'), (True, 'tf32')):
    agent, rb, orc, rep, args = _mk(B=B, dense=0.05, quantile=0.95, precision=prec)
    eng = agent.engine
    rs = np.random.RandomState(2)
    idxs = rs.randint(0, 48, size=B); rnd = _rnd(rs, B, A, "sgsac")
    batch = rep.sample(idxs)
    orc.trace = {}
    tq = orc.target_q(batch[2], batch[3], batch[4], rnd["noise_next"])
    gp = orc._grad_params(orc.critic_names)
    loss = orc.critic_loss(gp, batch[0], batch[1], tq, rnd)
    grads = torch.autograd.grad(loss, [gp[n] for n in orc.critic_names])
    tr = orc.trace
    _supply(agent, idxs, rnd)
    agent._draw(rb); agent._sample_into_engine(rb)
    if feed:
        eng.debug_masked_obs = tr["masked_obs"].cuda()
    eng.update_critic(1)
    torch.cuda.synchronize()
    g_ref = tr["obs_grad1"]
    print("feed", feed, prec, "attr maxerr", float((eng.obs_grad.cpu() - g_ref).abs().max()), "max", float(g_ref.abs().max()), "relnorm", _relerr(eng.obs_grad, g_ref))
    mask = eng.mask.reshape(B, 3, 1, 84, 84).expand(B, 3, 3, 84, 84).reshape(B, 9, 84, 84).bool().cpu()
    print(" mask mismatches", int((mask != tr["mask1"]).sum()) // 3, "of", B * 3 * 7056, "kept", int(mask.sum()) // 3, int(tr["mask1"].sum()) // 3)
    print(" loss", float(eng.logs[0]), float(loss))
    print(" Q1", _relerr(eng.q[0, :B], tr["Q1"][:, 0]), "mQ1", _relerr(eng.q[0, B:], tr["mQ1"][:, 0]))
    got = eng.lay.unpack(eng.grads)
    for n, gr in zip(orc.critic_names, grads):
        print("  %-24s rel %.3e  norm %.3e" % (n, _relerr(got[n], gr), float(gr.norm())))
