"""Time the phases of one SGSAC update as separately captured CUDA graphs (B=128)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import sgqn_carla_b200 as S
import bench
B = 128
args = S.default_args(algorithm="sgsac", batch_size=B, sgqn_quantile=0.95, seed=1)
frames, actions, rewards, not_dones, pool = bench.synthetic(4000, 2, seed=0)
ag = S.make_agent((9, 84, 84), (2,), args)
ag.set_overlay_pool(pool)
rb = S.ReplayBuffer((9, 84, 84), (2,), 4000, B, frame_capacity=4008)
rb.load_ring(frames, actions, rewards, not_dones)
L = bench.NullLog()
for s in range(1, 7):
    ag.update(rb, L, s)
torch.cuda.synchronize()
eng = ag.engine

def timeg(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]

def sample():
    ag._draw(rb); ag._sample_into_engine(rb)

def aux_and_actor():
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(main); eng.side2.wait_event(ev)
    with torch.cuda.stream(eng.side2):
        eng.update_actor_and_alpha(finish=False)
        ev2 = torch.cuda.Event(); ev2.record(eng.side2)
    eng.update_aux()
    main.wait_event(ev2)
    eng.actor_finish()

def crit_fwd_part():
    eng.target_q_pass(); eng.critic_fwd_rows(0, B, encode=False)

phases = [("sample", sample), ("target_q_pass+critic_fwd(obs)", crit_fwd_part),
          ("update_critic (whole)", lambda: eng.update_critic(1)),
          ("critic_step(ema)", lambda: eng.critic_step(True)), ("critic_step(no ema)", lambda: eng.critic_step(False)),
          ("shared_obs_fwd(aux)", lambda: eng.shared_obs_fwd(with_aux=True)),
          ("attribution2+mask", lambda: eng.attribution2(True)),
          ("actor_and_alpha alone", lambda: eng.update_actor_and_alpha()),
          ("update_aux alone", lambda: eng.update_aux()),
          ("aux || actor", aux_and_actor),
          ("enc_fwd 2B", lambda: eng.enc_fwd(eng.obs3.data_ptr(), 2 * B, eng.actS, 0)),
          ("enc_fwd B target", lambda: eng.enc_fwd(eng.next_obs.data_ptr(), B, eng.actT, target=True)),
          ("attribution (B)", lambda: eng.attribution(B, eng.haS.data_ptr(), eng.zS.data_ptr(), eng.obs_grad.data_ptr())),
          ("enc_bwd 2B wgrad", lambda: eng.enc_bwd(eng.dbuf[1].data_ptr(), 2 * B, eng.actS, B, eng.obs2.data_ptr(), 1, True)),
          ("enc_bwd B wgrad", lambda: eng.enc_bwd(eng.dbuf[1].data_ptr(), B, eng.actS, 2 * B, eng.s_tilde.data_ptr(), 1, True)),
          ("whole even step", lambda: (sample(), eng.update_sgsac(2))), ("whole odd step", lambda: (sample(), eng.update_sgsac(3)))]
for name, fn in phases:
    print(f"{name:36s} {timeg(fn):9.1f} us", flush=True)
