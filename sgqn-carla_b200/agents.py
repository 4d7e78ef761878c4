"""Drop-in agents with the reference's API (src/algorithms/{factory,sac,sgsac,svea,rad,drq}.py):

    agent = make_agent(obs_shape, action_shape, args)
    agent.update(replay_buffer, L, step[, count]); agent.select_action(obs); agent.sample_action(obs)
    agent.train(bool) / agent.eval() / agent.training / agent.alpha
    agent.actor / agent.critic / agent.attribution_predictor  -> .state_dict() with the reference's keys

Everything underneath runs on the hand-written sm_100a kernels of libsgqn_b200.so (engine.py); there is no
PyTorch-op fallback for the update path.
"""
import contextlib
import gc
import os
from collections import OrderedDict

import numpy as np
import torch

from . import init as _init
from . import _lib
from ._lib import K
from .engine import UpdateEngine, _ptr
from .layout import reference_key_map
from .lazylog import LazyScalar, _Ring
from .replay import ReplayBuffer


@contextlib.contextmanager
def _capture(graph):
    """torch.cuda.graph(graph) with the cyclic garbage collector parked for the duration of the capture.  A collection that
    happens to run in the middle of a capture can finalise objects of EARLIER agents / buffers (reference cycles keep them
    alive until a collection): releasing a pinned host tensor records a CUDA event and destroying a CUDAGraph resets it --
    both are illegal while a stream is capturing, invalidate the capture and abort the process from a destructor."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was_enabled:
            gc.enable()


class _ModuleView(object):
    """Stands in for `agent.actor` / `.critic` / `.critic_target` / `.attribution_predictor`: state_dict
    interchange with reference checkpoints (train.py:206-219), `parameters()`, `train()`."""

    def __init__(self, agent, module, target=False):
        self._agent, self._module, self._target = agent, module, target
        self.training = True

    def _names(self):
        m = "critic" if self._target is True else self._module
        return [(n, key) for n, refs in reference_key_map().items() for mod, key in refs if mod == m]

    def _slice(self, n):
        """Stored (flat) segment of parameter n: the online arena, the critic target, or SODA's predictor target."""
        eng = self._agent.engine
        if self._target == "soda":
            return eng.soda_target_slice(n)
        o, st, _ = eng.lay.entries[n]
        if self._target:
            c0 = eng.lay.ranges["critic"][0]
            return eng.target[o - c0:o - c0 + st]
        return eng.params[o:o + st]

    def state_dict(self):
        lay = self._agent.engine.lay
        out = OrderedDict()
        for n, key in self._names():
            if n in lay.entries:
                out[key] = lay.from_stored(n, self._slice(n))
        return out

    def load_state_dict(self, sd, strict=True):
        eng = self._agent.engine
        lay = eng.lay
        for n, key in self._names():
            if n not in lay.entries:
                continue
            if key not in sd:
                if strict:
                    raise KeyError(key)
                continue
            seg = self._slice(n)
            seg.copy_(lay.to_stored(n, sd[key]).to(seg.device))
        if self._target == "soda":
            eng.prep_soda_target_weights()
        else:
            eng.prep_conv_weights(target=bool(self._target))
        eng.prep_dec_weights()

    def parameters(self):
        return list(self.state_dict().values())

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)


class SAC(object):
    """sac.py:21-169"""
    algorithm = "sac"
    critic_mode = 0
    sample_mode = "crop"            # replay_buffer.sample() (sac.py:161)

    def __init__(self, obs_shape, action_shape, args, device="cuda", dist=None, global_batch=None, precision="tf32"):
        self.args = args
        self.obs_shape, self.action_shape = tuple(obs_shape), tuple(action_shape)
        self.discount = args.discount
        self.critic_tau, self.encoder_tau = args.critic_tau, args.encoder_tau
        self.actor_update_freq = args.actor_update_freq
        self.critic_target_update_freq = args.critic_target_update_freq
        self.batch_size = int(args.batch_size)
        self.engine = UpdateEngine(action_shape[0], args, self.batch_size, device=device, algorithm=self.algorithm,
                                   dist=dist, global_batch=global_batch, precision=precision)
        self.set_parameters(_init.init_params(action_shape[0], args))
        self.actor = _ModuleView(self, "actor")
        self.critic = _ModuleView(self, "critic")
        self.critic_target = _ModuleView(self, "critic", target=True)
        self.target_entropy = self.engine.target_entropy
        self._ring = _Ring()
        self._supplied = None
        self.defer_logs = True
        # CUDA graphs: after one eager update of each kind (odd / even step pattern) the whole update -- device RNG,
        # replay gather, ~200 kernels, optimiser steps -- is captured once and replayed with a single launch.
        # (the NCCL all-reduces of the sharded configuration are captured too; SGQN_DIST_GRAPHS=0 keeps that case eager)
        self.use_cuda_graphs = dist is None or os.environ.get("SGQN_DIST_GRAPHS", "1") == "1"
        self._graphs, self._eager_runs, self._graph_rb, self._graph_nodes = {}, {}, None, {}
        self._act = {}                  # (H, sample) -> staging buffers + captured batch-1 actor graph
        self._pf = None                 # prefetch state of a host-resident (pinned) replay buffer
        self._prefetch = os.environ.get("SGQN_PREFETCH", "1") == "1"
        self.train()

    # ---- parameters
    def set_parameters(self, canonical, sync_target=True):
        """canonical: name -> tensor in the reference's shapes (layout.reference_key_map names)."""
        eng = self.engine
        eng.lay.pack(canonical, eng.params)
        if sync_target:                                            # deepcopy(critic), sac.py:54
            c0, c1 = eng.lay.ranges["critic"]
            eng.target.copy_(eng.params[c0:c1])
        if "log_alpha" in canonical:
            eng.log_alpha.copy_(torch.as_tensor(canonical["log_alpha"], dtype=torch.float64).reshape(1))
        eng.prep_conv_weights()
        eng.prep_conv_weights(target=True)
        eng.prep_dec_weights()

    def set_training_state(self, canonical, optim=None, alpha_optim=None):
        """Overwrite the whole training state from reference-shaped tensors: parameters (`name`), target-network
        parameters (`t_name`), `log_alpha`, and per optimiser ("critic" / "actor" / "aux") the Adam moments as
        {name: tensor} dicts plus the step count -- what a `torch.optim.Adam.state_dict()` of the reference holds.
        Tensors an optimiser has no state for (never received a gradient) get zero moments, as in torch."""
        eng = self.engine
        lay = eng.lay
        self.set_parameters({k: v for k, v in canonical.items() if not k.startswith("t_")}, sync_target=False)
        c0, c1 = lay.ranges["critic"]
        crit = [n for n in lay.entries if c0 <= lay.off(n) < c1]
        lay.pack({n: canonical["t_" + n] for n in crit if "t_" + n in canonical}, eng.target, names=crit, base=c0)
        eng.prep_conv_weights(target=True)
        if eng.algorithm == "soda":                                # SODA's predictor target: st_<name>
            for n in lay.entries:
                if "st_" + n in canonical:
                    eng.soda_target_slice(n).copy_(lay.to_stored(n, canonical["st_" + n]).to(eng.dev))
            eng.prep_soda_target_weights()
        for which, st in (optim or {}).items():
            opt = self._optims()[which]
            r0, r1 = lay.ranges[which]
            names = [n for n in lay.entries if r0 <= lay.off(n) < r1]
            opt.m.zero_(); opt.v.zero_()
            lay.pack(st["m"], opt.m, names=names, base=r0)
            lay.pack(st["v"], opt.v, names=names, base=r0)
            opt.step.fill_(int(st["step"]))
        if alpha_optim is not None:
            eng.alpha_st.copy_(torch.tensor([float(alpha_optim["m"]), float(alpha_optim["v"])], dtype=torch.float64))
            eng.alpha_step.fill_(int(alpha_optim["step"]))

    def sync_from_rank0(self):
        """Data-parallel start-up (SURVEY.md 8e 'identical replicated parameters'): broadcast rank 0's parameters, targets,
        log_alpha and optimiser states, then rebuild every derived tensor-core operand copy from them."""
        import torch.distributed as dist
        eng = self.engine
        ts = [eng.params, eng.target, eng.log_alpha, eng.alpha_st, eng.alpha_step]
        for o in self._optims().values():
            ts += [o.m, o.v, o.step]
        for t in ts:
            dist.broadcast(t, 0)
        eng.prep_conv_weights(); eng.prep_conv_weights(target=True); eng.prep_dec_weights()
        self._graphs.clear(); self._eager_runs.clear()

    def get_parameters(self):
        eng = self.engine
        out = eng.lay.unpack(eng.params)
        c0, c1 = eng.lay.ranges["critic"]
        crit = [n for n in eng.lay.entries if c0 <= eng.lay.off(n) < c1]
        for n, t in eng.lay.unpack(eng.target, crit, base=c0).items():
            out["t_" + n] = t
        if eng.algorithm == "soda":
            c0s, c1s = eng.lay.ranges["cnn"]; s0, s1 = eng.lay.ranges["soda"]
            for n in eng.lay.entries:
                if c0s <= eng.lay.off(n) < c1s or s0 <= eng.lay.off(n) < s1:
                    out["st_" + n] = eng.lay.from_stored(n, eng.soda_target_slice(n))
        out["log_alpha"] = eng.log_alpha.clone().reshape(())
        return out

    # ---- checkpoints: reference-keyed module state_dicts (what train.py:206-219 saves one file each)
    #      + what the reference never saves: the optimisers' moments / step counts, log_alpha, the device RNG counter
    def _modules(self):
        mods = {"actor": self.actor, "critic": self.critic, "critic_target": self.critic_target}
        if hasattr(self, "attribution_predictor"):
            mods["attribution_predictor"] = self.attribution_predictor
        return mods

    def _optims(self):
        eng = self.engine
        return {"critic": eng.opt_critic, "actor": eng.opt_actor, "aux": eng.opt_aux}

    def checkpoint(self):
        eng = self.engine
        ck = {"format": "sgqn_b200/1", "algorithm": self.algorithm, "layout": [eng.lay.total, eng.A, eng.H],
              "modules": {k: OrderedDict((n, t.detach().cpu().clone()) for n, t in m.state_dict().items())
                          for k, m in self._modules().items()},
              "log_alpha": eng.log_alpha.cpu().clone(),
              "optim": {k: {"m": o.m.cpu().clone(), "v": o.v.cpu().clone(), "step": o.step.cpu().clone()}
                        for k, o in self._optims().items()},
              "alpha_optim": {"state": eng.alpha_st.cpu().clone(), "step": eng.alpha_step.cpu().clone()},
              "rng_counter": eng.rng_counter.cpu().clone(), "seed": int(eng.seed)}
        return ck

    def load_checkpoint_dict(self, ck, strict=True):
        eng = self.engine
        if ck.get("format") != "sgqn_b200/1":
            raise ValueError("not a sgqn_b200 checkpoint (for reference checkpoints use agent.actor.load_state_dict(...) etc.)")
        if list(ck["layout"]) != [eng.lay.total, eng.A, eng.H]:
            raise ValueError(f"checkpoint layout {ck['layout']} != agent layout {[eng.lay.total, eng.A, eng.H]}")
        for k, m in self._modules().items():
            if k in ck["modules"]:
                m.load_state_dict(ck["modules"][k], strict=strict)
            elif strict:
                raise KeyError(k)
        eng.log_alpha.copy_(ck["log_alpha"])
        for k, o in self._optims().items():
            st = ck["optim"][k]
            o.m.copy_(st["m"]); o.v.copy_(st["v"]); o.step.copy_(st["step"])
        eng.alpha_st.copy_(ck["alpha_optim"]["state"]); eng.alpha_step.copy_(ck["alpha_optim"]["step"])
        eng.rng_counter.copy_(ck["rng_counter"]); eng.seed = int(ck["seed"])
        self._graphs.clear(); self._eager_runs.clear(); self._act.clear()      # the seed is baked into the captured graphs

    def save_checkpoint(self, path):
        torch.save(self.checkpoint(), path)

    def load_checkpoint(self, path, strict=True):
        self.load_checkpoint_dict(torch.load(path, map_location="cpu", weights_only=False), strict=strict)

    def train(self, training=True):
        self.training = training
        self.actor.train(training)
        self.critic.train(training)

    def eval(self):
        self.train(False)

    @property
    def alpha(self):
        return self.engine.log_alpha.exp().reshape(())

    @property
    def log_alpha(self):
        return self.engine.log_alpha.reshape(())

    # ---- acting (sac.py:86-105)
    def _obs_to_input(self, obs):
        _obs = np.asarray(obs)
        return torch.as_tensor(_obs, dtype=torch.float32).to(self.engine.dev).unsqueeze(0)

    def _act_fast(self, obs, sample):
        """Batch-1 actor as ONE CUDA graph launch: the uint8 stack goes host -> pinned staging -> device
        (63.5 KB instead of the reference's 254 KB fp32 pageable copy, sac.py:86-93), uint8 -> fp32 conversion, encoder,
        projection, MLP, head and the device -> pinned-host copy of the action are nodes of the same graph; the host does
        one memcpy into the staging buffer, one cudaGraphLaunch and one stream synchronise per environment step."""
        arr = np.asarray(obs)
        if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[0] != 9 or arr.shape[1] != arr.shape[2] or arr.shape[1] % 4:
            return None
        H = int(arr.shape[1])
        eng, A = self.engine, self.action_shape[0]
        slot = self._act.get((H, sample))
        if slot is None:
            dev = eng.dev
            slot = dict(stage=torch.zeros(9, H, H, dtype=torch.uint8).pin_memory(),
                        frames=torch.zeros(9, H, H, dtype=torch.uint8, device=dev),
                        fidx=torch.tensor([0, 1, 2, 0, 1, 2], dtype=torch.int32, device=dev),
                        idx=torch.zeros(1, dtype=torch.int64, device=dev),
                        obs=torch.zeros(2, 9, H, H, device=dev),
                        out=torch.zeros(A).pin_memory(), graph=None, runs=0)
            slot["stage_np"] = slot["stage"].numpy()
            slot["out_np"] = slot["out"].numpy()
            self._act[(H, sample)] = slot

        def body():
            slot["frames"].copy_(slot["stage"], non_blocking=True)
            K.replay_gather(_ptr(slot["frames"]), _ptr(slot["fidx"]), _ptr(slot["idx"]), 0, _ptr(slot["obs"]), _ptr(slot["obs"][1]),
                            1, H, H, 0, 4, eng.st)
            res = eng.act(slot["obs"][:1], H, sample=sample)        # (sampling noise: the engine's Philox kernel)
            slot["out"].copy_(res, non_blocking=True)

        np.copyto(slot["stage_np"], arr)
        if slot["graph"] is None and slot["runs"] >= 1 and self.use_cuda_graphs:
            g = torch.cuda.CUDAGraph()
            c0 = _lib.launch_count
            with _capture(g):
                body()
            slot["graph"], slot["nodes"] = g, _lib.launch_count - c0
            _lib.launch_count = c0
        if slot["graph"] is not None:
            slot["graph"].replay()
            _lib.launch_count += slot["nodes"]
        else:
            slot["runs"] += 1
            body()
        torch.cuda.current_stream().synchronize()
        return slot["out_np"].copy()

    def select_action(self, obs):
        a = self._act_fast(obs, sample=False)
        if a is not None:
            return a
        x = self._obs_to_input(obs)
        return self.engine.act(x, x.shape[-1], sample=False).cpu().numpy().flatten()

    def sample_action(self, obs, noise=None):
        if noise is None:
            a = self._act_fast(obs, sample=True)
            if a is not None:
                return a
        x = self._obs_to_input(obs)
        if noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self.engine.dev).reshape(1, -1)
        return self.engine.act(x, x.shape[-1], sample=True, noise=noise).cpu().numpy().flatten()

    # ---- one step's randomness
    def supply(self, idxs=None, noise_next=None, noise_pi=None, u=None, overlay_ids=None, offs=None, places=None):
        """Host-supplied randomness for the NEXT update (parity runs; SURVEY.md 5 'RNG').  Anything left None is
        drawn on the device."""
        self._supplied = dict(idxs=idxs, noise_next=noise_next, noise_pi=noise_pi, u=u, overlay_ids=overlay_ids,
                              offs=offs, places=places)

    def _pool_n(self):
        """Size of the overlay image pool the step's `overlay_ids` index (SGSAC: carla frames; SVEA: places images)."""
        pool = self.engine.overlay_pool
        return int(pool.shape[0]) if pool is not None else 1

    def _draw(self, replay_buffer, skip_idxs=False):
        eng, B = self.engine, self.batch_size
        pool_n = self._pool_n()
        off_n = 9 if self.sample_mode == "shift" else max(1, getattr(replay_buffer, "Hs", 84) - 84)
        n_valid = replay_buffer.n_valid if isinstance(replay_buffer, ReplayBuffer) else eng.rng_counter.to(torch.int32)
        K.rng_step(eng.seed, _ptr(eng.rng_counter), _ptr(n_valid), 0 if skip_idxs else _ptr(eng.idxs), _ptr(eng.overlay_ids), pool_n,
                   _ptr(eng.offs), off_n, _ptr(eng.noise_next), _ptr(eng.noise_pi), _ptr(eng.u), B, eng.A,
                   eng.seed_shared if eng.seed_shared is not None else eng.seed, eng.st)
        s, self._supplied = self._supplied, None
        if s:
            dev = eng.dev
            if s["idxs"] is not None:
                eng.idxs.copy_(torch.as_tensor(np.asarray(s["idxs"]), dtype=torch.int64))
            if s["overlay_ids"] is not None:
                eng.overlay_ids.copy_(torch.as_tensor(np.asarray(s["overlay_ids"]), dtype=torch.int64))
            if s["offs"] is not None:
                eng.offs.copy_(torch.as_tensor(np.asarray(s["offs"]), dtype=torch.int32).reshape(2, B, 2))
            if s["noise_next"] is not None:
                eng.noise_next.copy_(torch.as_tensor(s["noise_next"], dtype=torch.float32))
            if s["noise_pi"] is not None:
                eng.noise_pi.copy_(torch.as_tensor(s["noise_pi"], dtype=torch.float32))
            if s["u"] is not None:
                eng.u.fill_(float(np.float32(s["u"])))
            if s["places"] is not None and hasattr(eng, "places"):
                eng.places.copy_(torch.as_tensor(s["places"], dtype=torch.float32).reshape(B, 3, -1))

    def _sample_into_engine(self, replay_buffer):
        eng, B = self.engine, self.batch_size
        if isinstance(replay_buffer, ReplayBuffer):
            hs = replay_buffer.Hs
            if self.sample_mode == "shift":
                mode, offs = 1, eng.offs
            else:
                mode, offs = 0, (eng.offs if hs > 84 else None)
            replay_buffer.gather_into(eng.idxs, eng.obs2[:B], eng.next_obs, eng.action, eng.reward, eng.not_done,
                                      offs, mode, 4, 84)
        else:       # a foreign buffer with the reference's surface: use its own sample*() (CUDA fp32 tensors)
            fn = replay_buffer.sample_drq if self.sample_mode == "shift" else replay_buffer.sample
            obs, a, r, nxt, nd = fn()
            eng.obs2[:B].copy_(obs); eng.next_obs.copy_(nxt); eng.action.copy_(a); eng.reward.copy_(r); eng.not_done.copy_(nd)

    def _emit_logs(self, L, step, cols):
        if L is None:
            return
        slot, serial = self._ring.push(self.engine.logs)
        sync = self.engine.dist
        for key, col in cols:
            v = LazyScalar([(self._ring, slot, serial, col, sync.log_scale(col) if sync is not None else 1.0)])
            L.log(key, v if self.defer_logs else float(v), step)

    def _log_cols(self, step):
        cols = [("train_critic/loss", 0)]
        if step % self.actor_update_freq == 0:
            cols += [("train_actor/loss", 1), ("train_alpha/loss", 2), ("train_alpha/value", 3)]
        return cols

    def _step_kind(self, step):
        return (step % self.actor_update_freq == 0, step % self.critic_target_update_freq == 0)

    @property
    def prefetch(self):
        return self._prefetch

    @prefetch.setter
    def prefetch(self, on):
        if bool(on) != self._prefetch:                    # the choice is baked into the captured update graphs
            self._prefetch = bool(on)
            self._graphs.clear(); self._eager_runs.clear()
            if self._pf is not None:
                self._pf["primed"] = False

    # ---- host-resident replay (storage="pinned"): the NEXT update's batch crosses PCIe under the CURRENT update
    def _prefetch_issue(self, rb, pf):
        """Draw the next batch's indices and pull its frames (raw uint8) + action / reward / not_done rows into the
        device staging buffers, on the current stream."""
        eng, B = self.engine, self.batch_size
        st = eng.st
        K.rng_step(eng.seed ^ 0x5DEECE66D, _ptr(pf["counter"]), _ptr(rb.n_valid), _ptr(pf["idxs"]), 0, 1, 0, 1, 0, 0, 0, B, eng.A, 0, st)
        K.frames_copy(_ptr(rb.frames), _ptr(rb.fidx), _ptr(pf["idxs"]), _ptr(pf["frames"]), B, 3 * rb.Hs * rb.Hs, st)
        K.take_rows(_ptr(rb.actions), _ptr(pf["idxs"]), _ptr(pf["action"]), B, eng.A, st)
        K.take_rows(_ptr(rb.rewards), _ptr(pf["idxs"]), _ptr(pf["reward"]), B, 1, st)
        K.take_rows(_ptr(rb.not_dones), _ptr(pf["idxs"]), _ptr(pf["not_done"]), B, 1, st)

    def _run_update_prefetched(self, rb, step):
        eng, B = self.engine, self.batch_size
        pf = self._pf
        if pf is None or pf["rb"] is not rb or pf["version"] != rb.version:
            dev = eng.dev
            pf = self._pf = dict(rb=rb, version=rb.version, primed=False, frames=torch.zeros(B * 6, 3 * rb.Hs * rb.Hs, dtype=torch.uint8, device=dev),
                                 fidx=torch.arange(6 * B, dtype=torch.int32, device=dev).reshape(B, 6),
                                 arange=torch.arange(B, dtype=torch.int64, device=dev),
                                 idxs=torch.zeros(B, dtype=torch.int64, device=dev),
                                 counter=torch.zeros(1, dtype=torch.int64, device=dev),
                                 action=torch.zeros(B, eng.A, device=dev), reward=torch.zeros(B, 1, device=dev),
                                 not_done=torch.zeros(B, 1, device=dev), stream=torch.cuda.Stream(device=dev))
        if not pf["primed"]:                              # first update with this buffer: fetch synchronously
            self._prefetch_issue(rb, pf)
            pf["primed"] = True
        # consume the staged batch: uint8 -> fp32 (+ crop / shift) on the device
        self._draw(rb, skip_idxs=True)
        if self.sample_mode == "shift":
            mode, offs = 1, eng.offs
        else:
            mode, offs = 0, (eng.offs if rb.Hs > 84 else None)
        K.replay_gather(_ptr(pf["frames"]), _ptr(pf["fidx"]), _ptr(pf["arange"]), _ptr(offs) if offs is not None else 0,
                        _ptr(eng.obs2[:B]), _ptr(eng.next_obs), B, rb.Hs, 84, mode, 4, eng.st)
        eng.idxs.copy_(pf["idxs"]); eng.action.copy_(pf["action"]); eng.reward.copy_(pf["reward"]); eng.not_done.copy_(pf["not_done"])
        # the staging buffers are free again: fetch the next batch beside this update
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main); pf["stream"].wait_event(ev)
        with torch.cuda.stream(pf["stream"]):
            self._prefetch_issue(rb, pf)
            ev2 = torch.cuda.Event(); ev2.record(pf["stream"])
        self._engine_update(step)
        main.wait_event(ev2)

    def _run_update(self, replay_buffer, step):
        if (self.prefetch and self._supplied is None and isinstance(replay_buffer, ReplayBuffer)
                and replay_buffer.storage == "pinned"):
            self._run_update_prefetched(replay_buffer, step)
            return
        self._draw(replay_buffer)
        self._sample_into_engine(replay_buffer)
        self._engine_update(step)

    def _engine_update(self, step):
        self.engine.update_sac(step, self.critic_mode)

    def _update_maybe_graphed(self, replay_buffer, step):
        """Eager the first time a step kind is seen (and whenever randomness is host-supplied or the buffer is foreign);
        afterwards one cudaGraphLaunch per update."""
        kind = self._step_kind(step)
        graphable = (self.use_cuda_graphs and self._supplied is None and isinstance(replay_buffer, ReplayBuffer)
                     and self._graphable())
        if not graphable:
            self._run_update(replay_buffer, step)
            return
        key = (id(replay_buffer), replay_buffer.version)   # pointers are baked into the graphs (the frame ring can grow)
        if self._graph_rb != key:
            self._graphs.clear(); self._eager_runs.clear(); self._graph_rb = key
        g = self._graphs.get(kind)
        if g is None:
            if self._eager_runs.get(kind, 0) < 1:
                self._eager_runs[kind] = self._eager_runs.get(kind, 0) + 1
                self._run_update(replay_buffer, step)
                return
            g = torch.cuda.CUDAGraph()
            c0 = _lib.launch_count
            with _capture(g):
                self._run_update(replay_buffer, step)
            self._graphs[kind] = g
            self._graph_nodes[kind] = _lib.launch_count - c0
            _lib.launch_count = c0                          # capturing launched nothing
        g.replay()
        _lib.launch_count += self._graph_nodes[kind]        # kernels-launching ABI calls replayed by this graph

    def _graphable(self):
        return True

    def update(self, replay_buffer, L, step, count=0):
        self.count = count
        self._update_maybe_graphed(replay_buffer, step)
        self._emit_logs(L, step, self._log_cols(step))


class RAD(SAC):
    """rad.py:11-13: SAC on 100x100 frames with sample()'s random crop to 84."""
    algorithm = "rad"


class DrQ(SAC):
    """drq.py:11-24: SAC with sample_drq() (random shift, K=1, M=1)."""
    algorithm = "drq"
    sample_mode = "shift"


class SODA(SAC):
    """soda.py:12-84: SAC on random crops (100 -> 84) plus the SODA consistency update on a separately sampled batch of
    `soda_batch_size` observations (two crops, one of them overlaid with a places image; SODAMLPs with BatchNorm1d; EMA target)."""
    algorithm = "soda"

    def __init__(self, obs_shape, action_shape, args, **kw):
        super().__init__(obs_shape, action_shape, args, **kw)
        self.aux_update_freq = args.aux_update_freq
        self.soda_batch_size, self.soda_tau = args.soda_batch_size, args.soda_tau
        self.predictor = _ModuleView(self, "predictor")
        self.predictor_target = _ModuleView(self, "predictor", target="soda")
        self.places_pool = None
        p = _init.init_params(action_shape[0], args)
        p.update({"st_" + n: p[n].clone() for n in p if n.startswith(("cnn.", "soda_"))})     # deepcopy(predictor), soda.py:30
        self.set_training_state(p)

    def _modules(self):
        return dict(super()._modules(), predictor=self.predictor, predictor_target=self.predictor_target)

    def set_parameters(self, canonical, sync_target=True):
        super().set_parameters(canonical, sync_target)
        if sync_target and hasattr(self.engine, "soda_target"):   # a fresh parameter set: the predictor target is its copy
            eng = self.engine
            for n in eng.lay.entries:
                if n in canonical and n.startswith(("cnn.", "soda_")):
                    eng.soda_target_slice(n).copy_(eng.lay.to_stored(n, canonical[n]).to(eng.dev))
            eng.prep_soda_target_weights()

    def set_places_pool(self, imgs):
        t = torch.as_tensor(imgs, dtype=torch.float32)
        assert t.dim() == 4 and tuple(t.shape[1:]) == (3, 84, 84), "places pool: float images (N,3,84,84) in [0,1]"
        self.places_pool = t.to(self.engine.dev).contiguous()
        self.engine.places_pool = self.places_pool.reshape(t.shape[0], 3, -1)
        self._graphs.clear(); self._eager_runs.clear()

    def load_places_dir(self, data_dirs, n=4096, use_val=False, seed=0):
        from .datasets import load_places_pool
        self.set_places_pool(load_places_pool(data_dirs, n, 84, use_val, seed))

    def _log_cols(self, step):
        cols = super()._log_cols(step)
        if step % self.aux_update_freq == 0:
            cols.append(("train/aux_loss", 4))
        return cols

    def _step_kind(self, step):
        return super()._step_kind(step) + (step % self.aux_update_freq == 0,)

    def supply(self, soda_idxs=None, soda_offs=None, soda_places=None, **kw):
        """+ the SODA batch of the next update: indices (n,), crop offsets (2,n,2) for x / aug_x, overlay images (n,3,84,84)."""
        super().supply(**kw)
        self._supplied.update(soda_idxs=soda_idxs, soda_offs=soda_offs, soda_places=soda_places)

    def update(self, replay_buffer, L, step, count=0):
        eng = self.engine
        if not isinstance(replay_buffer, ReplayBuffer):
            raise TypeError("SODA samples its own second batch from the buffer: use sgqn_carla_b200.ReplayBuffer")
        s = self._supplied
        if step % self.aux_update_freq == 0:
            if s is not None and s.get("soda_idxs") is not None:
                n = eng.ns
                eng.soda_idxs.copy_(torch.as_tensor(np.asarray(s["soda_idxs"]), dtype=torch.int64))
                eng.soda_offs[:2].copy_(torch.as_tensor(np.asarray(s["soda_offs"]), dtype=torch.int32).reshape(2, n, 2))
                eng.soda_places.copy_(torch.as_tensor(s["soda_places"], dtype=torch.float32).reshape(n, 3, -1))
                eng.soda_supplied = dict(places=True)
            elif self.places_pool is None:
                raise RuntimeError("SODA needs an overlay image pool: agent.set_places_pool(float images (N,3,84,84) in [0,1])")
        eng.soda_src = dict(frames=_ptr(replay_buffer.frames), fidx=_ptr(replay_buffer.fidx), n_valid=_ptr(replay_buffer.n_valid),
                            Hs=replay_buffer.Hs)
        super().update(replay_buffer, L, step, count)

    def _run_update(self, replay_buffer, step):           # (the SODA batch is gathered straight from the ring: no staged prefetch)
        self._draw(replay_buffer)
        self._sample_into_engine(replay_buffer)
        self._engine_update(step)


class PAD(SAC):
    """pad.py:11-63: SAC on random crops (100 -> 84) plus the inverse-dynamics auxiliary update (obs, next_obs -> action)."""
    algorithm = "pad"

    def __init__(self, obs_shape, action_shape, args, **kw):
        super().__init__(obs_shape, action_shape, args, **kw)
        self.aux_update_freq = args.aux_update_freq
        self.pad_head = _ModuleView(self, "pad_head")

    def _modules(self):
        return dict(super()._modules(), pad_head=self.pad_head)

    def _log_cols(self, step):
        cols = super()._log_cols(step)
        if step % self.aux_update_freq == 0:
            cols.append(("train/aux_loss", 4))
        return cols

    def _step_kind(self, step):
        return super()._step_kind(step) + (step % self.aux_update_freq == 0,)


class CURL(SAC):
    """curl.py:11-57: SAC on random crops (100 -> 84) plus the contrastive auxiliary update on a second crop of obs
    (`replay_buffer.sample_curl()`, utils.py:142-156)."""
    algorithm = "curl"

    def __init__(self, obs_shape, action_shape, args, **kw):
        super().__init__(obs_shape, action_shape, args, **kw)
        self.aux_update_freq = args.aux_update_freq
        self.curl_head = _ModuleView(self, "curl_head")

    def _modules(self):
        return dict(super()._modules(), curl_head=self.curl_head)

    def _log_cols(self, step):
        cols = super()._log_cols(step)
        if step % self.aux_update_freq == 0:
            cols.append(("train/aux_loss", 4))
        return cols

    def _step_kind(self, step):
        return super()._step_kind(step) + (step % self.aux_update_freq == 0,)

    def supply(self, offs_pos=None, **kw):
        """+ offs_pos: (B,2) crop offsets of the second crop of obs."""
        super().supply(**kw)
        self._supplied["offs_pos"] = offs_pos

    def _run_update(self, replay_buffer, step):
        eng, B = self.engine, self.batch_size
        if not isinstance(replay_buffer, ReplayBuffer):
            raise TypeError("CURL samples obs / next_obs / pos from the same indices: use sgqn_carla_b200.ReplayBuffer")
        s = self._supplied
        off_n = max(1, replay_buffer.Hs - 84)
        self._draw(replay_buffer)                         # idxs, obs / next_obs crop offsets, noise (consumes self._supplied)
        # the third crop (`pos`): own Philox stream; only the offsets are used
        K.rng_step(eng.seed ^ 0x2545F491, _ptr(eng.rng_counter), 0, 0, 0, 1, _ptr(eng.offs_pos), off_n, 0, 0, 0, B, eng.A, 0, eng.st)
        if s and s.get("offs_pos") is not None:
            eng.offs_pos[0].copy_(torch.as_tensor(np.asarray(s["offs_pos"]), dtype=torch.int32).reshape(B, 2))
        self._sample_into_engine(replay_buffer)
        hs = replay_buffer.Hs
        K.replay_gather(_ptr(replay_buffer.frames), _ptr(replay_buffer.fidx), _ptr(eng.idxs), _ptr(eng.offs_pos) if hs > 84 else 0,
                        _ptr(eng.pos), _ptr(eng.pos_scratch), B, hs, 84, 0, 4, eng.st)
        self._engine_update(step)


class SVEA(SAC):
    """svea.py:12-63: critic on cat(obs, random_overlay(obs)); places365 images are host-supplied / pooled."""
    algorithm = "svea"
    critic_mode = 2
    sample_mode = "shift"

    def __init__(self, obs_shape, action_shape, args, **kw):
        super().__init__(obs_shape, action_shape, args, **kw)
        self.svea_alpha, self.svea_beta = args.svea_alpha, args.svea_beta
        self.places_pool = None         # float (N,3,84,84) in [0,1] on the device (what _get_places_batch yields)

    def set_places_pool(self, imgs):
        t = torch.as_tensor(imgs, dtype=torch.float32)
        assert t.dim() == 4 and tuple(t.shape[1:]) == (3, 84, 84), "places pool: float images (N,3,84,84) in [0,1]"
        self.places_pool = t.to(self.engine.dev).contiguous()
        self.engine.places_pool = self.places_pool.reshape(t.shape[0], 3, -1)
        self._graphs.clear(); self._eager_runs.clear()          # the pool pointer / size are baked into the captured graphs

    def _pool_n(self):
        return int(self.places_pool.shape[0]) if self.places_pool is not None else 1

    def load_places_dir(self, data_dirs, n=4096, use_val=False, seed=0):
        """Places365 as the reference reads it (augmentations.py:17-62), n transformed images drawn once into the device pool."""
        from .datasets import load_places_pool
        self.set_places_pool(load_places_pool(data_dirs, n, 84, use_val, seed))

    def update(self, replay_buffer, L, step, count=0):
        """svea.py:54-63.  The overlay images of a step are rows `overlay_ids` (drawn on the device with everything else)
        of the device-resident places pool, read by the overlay kernel itself -- the whole update is one CUDA graph, like
        SAC's.  Parity runs hand the images over with supply(places=...)."""
        supplied_places = self._supplied is not None and self._supplied.get("places") is not None
        if not supplied_places and self.places_pool is None:
            raise RuntimeError("SVEA needs an overlay image pool: agent.set_places_pool(float images (N,3,84,84) in [0,1])")
        self.engine.places_from_pool = not supplied_places
        super().update(replay_buffer, L, step, count)


class SGSAC(SAC):
    """sgsac.py:24-185"""
    algorithm = "sgsac"
    critic_mode = 1

    def __init__(self, obs_shape, action_shape, args, **kw):
        super().__init__(obs_shape, action_shape, args, **kw)
        self.attribution_predictor = _ModuleView(self, "attribution_predictor")
        self.quantile = args.sgqn_quantile
        self.aux_update_freq = args.aux_update_freq
        self.consistency = args.consistency
        self.alpha_blending = args.alpha_blending
        self.count = 0
        self._writer = None             # SummaryWriter, created on first use (sgsac.py:41-48)

    def set_overlay_pool(self, frames_u8):
        """uint8 (N,3,84,84) frames: what `datasets/carla/*.npy` hold (utils.py:325-327), loaded once to the device
        instead of B `np.load`s per aux update (augmentations.py:65-76)."""
        t = torch.as_tensor(frames_u8)
        assert t.dtype == torch.uint8 and t.shape[1:] == (3, 84, 84)
        self.engine.overlay_pool = t.to(self.engine.dev).reshape(t.shape[0], 3, -1).contiguous()

    def load_overlay_dir(self, path, limit=None):
        """`datasets/carla` of the reference (one uint8 (3,84,84) `.npy` per frame, utils.py:325-327), read once."""
        from .datasets import load_carla_frames
        self.set_overlay_pool(load_carla_frames(path, limit))

    def _log_cols(self, step):
        cols = super()._log_cols(step)
        if step % self.aux_update_freq == 0:
            cols.append(("train/aux_loss", 4))
        return cols

    def update(self, replay_buffer, L, step, count=0):
        self.count = count
        if step % self.aux_update_freq == 0 and self.engine.overlay_pool is None:
            raise RuntimeError("SGSAC.update_aux needs the overlay pool: agent.set_overlay_pool(uint8 frames (N,3,84,84))")
        self._update_maybe_graphed(replay_buffer, step)
        self._emit_logs(L, step, self._log_cols(step))

    def _step_kind(self, step):
        return super()._step_kind(step) + (step % self.aux_update_freq == 0,)

    def _engine_update(self, step):
        self.engine.update_sgsac(step)

    # ---- the eval hook of the reference's train loops (train.py:36-50, train_carla.py:45-52)
    @property
    def writer(self):
        if self._writer is None:
            from .viz import make_writer
            a = self.args
            self._writer = make_writer(os.path.join(str(getattr(a, "log_dir", "logs")),
                                                    f"{getattr(a, 'domain_name', 'carla')}_{getattr(a, 'task_name', 'drive')}",
                                                    str(getattr(a, "algorithm", "sgsac")), str(getattr(a, "seed", 0)), "tensorboard"))
        return self._writer

    @writer.setter
    def writer(self, w):
        self._writer = w

    def log_tensorboard(self, obs, action, step, prefix="original"):
        """sgsac.py:104-135: observation grid, guided-backprop attribution grid, the observation masked by the attribution
        predictor's output, the predicted attribution, and the observation masked at five quantiles -- images go to
        `self.writer` and (best effort, like the reference's try/except) to output/<prefix>/...png."""
        from .viz import make_obs_grad_grid, make_obs_grid
        obs = torch.as_tensor(obs, dtype=torch.float32, device=self.engine.dev)
        action = torch.as_tensor(action, dtype=torch.float32, device=self.engine.dev).reshape(obs.shape[0], -1)
        obs = obs[..., 8:-8, 8:-8] if obs.shape[-1] == 100 else obs            # CenterCrop (modules.py:70-83)
        n = min(4, obs.shape[0])
        obs_grad = self.compute_attribution(obs, action)
        attrib = self.predict_attribution(obs, action)
        images = [("observation", "grid", make_obs_grid(obs, n)),
                  ("attributions", "grad_grid", make_obs_grad_grid(obs_grad.abs(), n)),
                  ("masked_obs", "masked_obs", make_obs_grid(obs * (torch.sigmoid(attrib) > 0.5).float(), n)),
                  ("predicted_attrib", "attrib_grid", make_obs_grad_grid(torch.sigmoid(attrib), n))]
        for q in (0.95, 0.975, 0.9, 0.995, 0.999):
            images.append((f"attrib_q{q}", "masked_obs", make_obs_grid(obs * self.compute_attribution_mask(obs_grad, q).float(), n)))
        for tag, name, img in images:
            self.writer.add_image(f"{prefix}/{tag}", img, global_step=step)
            self.save_image(f"{prefix}/{tag}", name, img, step)

    def save_image(self, folder, name, obj, step, plot=False):
        """sgsac.py:137-161 writes output/<folder>/<name>_<step>_<count>.png through matplotlib inside a try/except; here through
        PIL when it is importable (as an HWC image -- the reference's `.view(H, W, 3)` of a CHW tensor scrambles it)."""
        try:
            from PIL import Image
            path = os.path.join("output", folder)
            os.makedirs(path, exist_ok=True)
            img = (obj.detach().clamp(0, 1) * 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
            Image.fromarray(img).save(os.path.join(path, f"{name}_{step}_{self.count}.png"))
        except Exception:
            pass

    def _rows_into_engine(self, obs, action):
        """Put n <= batch_size observation rows (+ actions) into the engine's obs slot (eval-time entry points run the
        batch-sized kernels; the unused rows keep whatever the last update left there)."""
        eng, B = self.engine, self.batch_size
        n = int(obs.shape[0])
        if n > B:
            raise ValueError(f"{n} observations > batch_size {B}: call in chunks")
        eng.obs2[:n].copy_(obs); eng.action[:n].copy_(action)
        return n

    # stage-wise entry points mirroring rl_utils (used by the parity tests and by eval-time visualisation)
    def compute_attribution(self, obs, action):
        """rl_utils.compute_attribution(self.critic, obs, action): (n,9,84,84) guided-backprop attribution, n <= batch_size."""
        eng = self.engine
        n = self._rows_into_engine(obs, action)
        eng.shared_obs_fwd()
        eng.attribution2(want_mask=False)
        return eng.obs_grad[:n].clone()

    def predict_attribution(self, obs, action):
        """self.attribution_predictor(obs, action) (modules.py:345-354): logits (n,9,84,84), n <= batch_size."""
        eng = self.engine
        n = self._rows_into_engine(obs, action)
        eng.shared_obs_fwd()
        return eng.predict_attribution()[:n].clone()

    def compute_attribution_mask(self, obs_grad, quantile=None):
        """rl_utils.compute_attribution_mask: bool (B,9,84,84)."""
        eng, B = self.engine, obs_grad.shape[0]
        g = obs_grad.contiguous().float()
        mask = torch.empty(B, 3, 84 * 84, dtype=torch.uint8, device=eng.dev)
        K.attribution_mask(_ptr(g), 0, 0, 0, float(self.quantile if quantile is None else quantile), _ptr(mask), 0, B, 84 * 84, 0, eng.st)
        return mask.reshape(B, 3, 1, 84, 84).expand(B, 3, 3, 84, 84).reshape(B, 9, 84, 84).bool()


algorithm = {"sac": SAC, "rad": RAD, "drq": DrQ, "svea": SVEA, "sgsac": SGSAC, "curl": CURL, "pad": PAD, "soda": SODA}


def make_agent(obs_shape, action_shape, args, **kw):
    """factory.py:22-23"""
    if args.algorithm not in algorithm:
        raise KeyError(f'algorithm "{args.algorithm}" is outside the B200 hot-path scope (SURVEY.md 8): {sorted(algorithm)}')
    return algorithm[args.algorithm](obs_shape, action_shape, args, **kw)
