"""Per-tensor gradient / activation error of update_critic on the tf32 product path against the TF32-emulating oracle
and the fp32 oracle (diagnostic; run on the B200 box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_update_parity_gpu import _mk, _rnd, _supply, _relerr  # noqa: E402


def main():
    from oracle import sgsac_oracle as O
    from sgqn_carla_b200.layout import ENC_H
    B, A = 8, 2
    for dense in (0.05, None):
        agent, rb, orc, rep, args = _mk(B=B, dense=dense, quantile=0.95, precision="tf32")
        eng = agent.engine
        rs = np.random.RandomState(2)
        idxs = rs.randint(0, 48, size=B); rnd = _rnd(rs, B, A, "sgsac")
        batch = rep.sample(idxs)
        res = {}
        for name, tf in (("emu", True), ("emu64", "f64"), ("fp32", False)):
            orc.tf32 = tf
            orc.trace = {}
            tq = orc.target_q(batch[2], batch[3], batch[4], rnd["noise_next"])
            gp = orc._grad_params(orc.critic_names)
            loss = orc.critic_loss(gp, batch[0], batch[1], tq, rnd)
            grads = torch.autograd.grad(loss, [gp[n] for n in orc.critic_names])
            res[name] = (dict(orc.trace), loss.detach(), dict(zip(orc.critic_names, grads)))
        _supply(agent, idxs, rnd)
        agent._draw(rb); agent._sample_into_engine(rb)
        eng.debug_masked_obs = res["emu"][0]["masked_obs"].to("cuda")
        eng.update_critic(1)
        torch.cuda.synchronize()
        got = eng.lay.unpack(eng.grads)
        print(f"== dense={dense}")
        # activations of the obs rows, per layer, against the emulating oracle
        keep = []
        O.cnn_forward(orc.p, batch[0], tf32=True, keep=keep)
        for l in range(11):
            h = ENC_H[l]
            if l < 10:
                mine = eng.actS[l].reshape(3 * B, h + 2, h, 32)[B:2 * B, :h].permute(0, 3, 1, 2).cpu()
                ref = O.round_tf32(torch.relu(keep[l]))
            else:
                mine = eng.actS[l][:3 * B * h * h * 32].reshape(3 * B, h, h, 32)[B:2 * B].permute(0, 3, 1, 2).cpu()
                ref = keep[l]
            ne = int((mine != ref).sum())
            print(f"  act {l}: relerr {_relerr(mine, ref):.2e}  differing {ne}/{ref.numel()}  signflips {int(((mine > 0) != (ref > 0)).sum())}")
        tr64, loss64, grads64 = res["emu64"]
        tr32, loss32, grads32 = res["emu"]
        print(f"  emu vs emu64: loss {abs(float(loss32) - float(loss64)) / abs(float(loss64)):.2e} attribution {_relerr(tr32['obs_grad1'], tr64['obs_grad1']):.2e}"
              f" maxabs/max {float((tr32['obs_grad1'] - tr64['obs_grad1']).abs().max() / tr64['obs_grad1'].abs().max()):.2e}")
        for n in grads64:
            print(f"     {n:24s} relerr {_relerr(grads32[n], grads64[n]):.2e}")
        for name in ("emu", "emu64", "fp32"):
            tr, loss, grads = res[name]
            print("  attribution maxabs/max", float((eng.obs_grad.cpu() - tr['obs_grad1']).abs().max() / tr['obs_grad1'].abs().max()))
            print(f"  vs {name}: loss {abs(float(eng.logs[0]) - float(loss)) / abs(float(loss)):.2e}  attribution {_relerr(eng.obs_grad, tr['obs_grad1']):.2e}"
                  f"  Q1 {_relerr(eng.q[0, :B], tr['Q1'][:, 0]):.2e}")
            for n, gr in grads.items():
                a, b = got[n].double().cpu().reshape(-1), gr.double().reshape(-1)
                sf = float(((a > 0) != (b > 0)).double().mean())
                print(f"     {n:24s} relerr {_relerr(got[n], gr):.2e}  signflip-frac {sf:.4f}  |g| {float(b.norm()):.2e}")


if __name__ == "__main__":
    main()
