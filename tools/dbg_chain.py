import sys, os
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R_); sys.path.insert(0, os.path.join(R_, "tests"))
import numpy as np, torch, torch.nn.functional as F
from test_update_parity_gpu import _mk, _rnd, _supply, _relerr
from oracle import sgsac_oracle as O
from sgqn_carla_b200 import _lib
from sgqn_carla_b200.layout import ENC_H
B, A = 8, 2
agent, rb, orc, rep, args = _mk(B=B, dense=0.05, quantile=0.95)
eng = agent.engine
rs = np.random.RandomState(2)
idxs = rs.randint(0, 48, size=B); rnd = _rnd(rs, B, A, "sgsac")
batch = rep.sample(idxs)
orc.trace = {}
tq = orc.target_q(batch[2], batch[3], batch[4], rnd["noise_next"])
# oracle forward with per-layer activations retained
p = {k: v.detach() for k, v in orc.p.items()}
obs_grad = O.compute_attribution(orc.p, batch[0], batch[1]); mask = O.compute_attribution_mask(obs_grad, 0.95)
masked = batch[0] * mask; lo, hi = batch[0].min(), batch[0].max(); masked[mask < 1] = lo + (hi - lo) * rnd["u"]
x2 = torch.cat([batch[0], masked], 0)
acts = []
x = F.conv2d(x2 / 255.0, p["cnn.0.weight"], p["cnn.0.bias"], stride=2); x.requires_grad_(True); acts.append(x)
for i in range(1, 11):
    x = F.conv2d(F.relu(x), p[f"cnn.{i}.weight"], p[f"cnn.{i}.bias"]); x.retain_grad(); acts.append(x)
h = O.projection(p, x.reshape(2 * B, -1), "critic_proj")
a2 = torch.cat([batch[1], batch[1]], 0)
ha = torch.cat([h, a2], 1)
q1 = O.mlp3(p, ha, "Q1"); q2 = O.mlp3(p, ha, "Q2")
loss = F.mse_loss(q1[:B], tq) + F.mse_loss(q2[:B], tq) + 0.5 * (F.mse_loss(q1[:B], q1[B:]) + F.mse_loss(q2[:B], q2[B:]))
loss.backward()
# engine with capture
caps = []
orig = _lib.K.conv_dgrad
import ctypes
def view(ptr, numel):
    buf = (ctypes.c_float * numel).from_address(0)  # placeholder
    return None
def cap(*a):
    n, hl, cin = a[4], a[5], a[7]
    if n == 2 * B and hl == 23:
        # find the tensors the pointers belong to
        dsrc = eng.dbuf[1] if a[0] == eng.dbuf[1].data_ptr() else eng.dbuf[0]
        dy = dsrc[:n * 21 * 21 * 32].clone().reshape(n, 21, 21, 32).permute(0, 3, 1, 2).contiguous()
        assert a[2] == eng.actS[9].data_ptr()
        act9 = eng.actS[9][:n * 23 * 23 * 32].clone().reshape(n, 23, 23, 32).permute(0, 3, 1, 2).contiguous()
        w = eng.lay.unpack(eng.params)["cnn.10.weight"]
        ref = F.conv_transpose2d(dy, w) * (act9 > 0)
        globals()["ref10"] = ref.cpu(); globals()["dy10"] = dy.cpu(); globals()["act9e"] = act9.cpu()
    orig(*a); torch.cuda.synchronize()
    buf = eng.dbuf[0] if a[3] == eng.dbuf[0].data_ptr() else eng.dbuf[1]
    caps.append((n, hl, buf[:n * hl * hl * cin].clone()))
_lib.K.conv_dgrad = cap
_supply(agent, idxs, rnd)
agent._draw(rb); agent._sample_into_engine(rb)
eng.update_critic(1)
torch.cuda.synchronize()
print("loss", float(eng.logs[0]), float(loss))
crit = [c for c in caps if c[0] == 2 * B]
for k, (n, hl, d) in enumerate(crit):
    l = 9 - k
    ref = acts[l].grad                     # (2B,32,hl,hl)
    mine = d.reshape(n, hl, hl, 32).permute(0, 3, 1, 2).cpu()
    print("layer act", l, "hl", hl, ref.shape[-1], "rel all %.3e clean %.3e masked %.3e" % (_relerr(mine, ref), _relerr(mine[:B], ref[:B]), _relerr(mine[B:], ref[B:])))

mine = crit[0][2].reshape(2*B, 23, 23, 32).permute(0, 3, 1, 2).cpu()
print("kernel vs torch on same inputs: clean %.3e masked %.3e" % (_relerr(mine[:B], ref10[:B]), _relerr(mine[B:], ref10[B:])))
print("engine dfeat vs oracle: clean %.3e masked %.3e" % (_relerr(dy10[:B], acts[10].grad[:B]), _relerr(dy10[B:], acts[10].grad[B:])))
print("engine act9 vs oracle: clean %.3e masked %.3e" % (_relerr(act9e[:B], acts[9].detach()[:B]), _relerr(act9e[B:], acts[9].detach()[B:])))
print("sign mismatches act9 clean", int(((act9e[:B] > 0) != (acts[9].detach()[:B] > 0)).sum()), "masked", int(((act9e[B:] > 0) != (acts[9].detach()[B:] > 0)).sum()))
