"""Device-side engine of the SAC / SGSAC / SVEA update: owns the flat parameter / gradient / Adam arenas and the
activation workspaces, and issues the kernel schedule of one update through the C ABI (include/sgqn_b200.h).

Reference call stack being replaced: sgsac.py:169-185 `SGSAC.update` -> update_critic (:52-80),
compute_attribution x2 (rl_utils.py:35-39,57-62), update_actor_and_alpha (sac.py:125-151),
soft_update_critic_target (sac.py:153-158), update_aux (sgsac.py:82-102).

Forward passes that the reference repeats on identical inputs and weights are shared:
`critic(obs)` and attribution #1 share one encoder forward; attribution #2, `actor(obs, detach)` and
`critic(obs, pi, detach)` share one.  The soft target update rides on the critic Adam launch of even steps
(nothing reads the target or writes the critic in between).
"""
import math
import os

import numpy as np
import torch

from ._lib import K, conv_layers
from .layout import DEC_C3, ENC_H, FEAT, ParamLayout


def _ptr(t, off=0):
    return t.data_ptr() + off * t.element_size()


class _Opt:
    def __init__(self, dev, n, lr, b1, b2=0.999, eps=1e-8, wd=0.0):
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bc = torch.zeros(2, device=dev)
        self.lr, self.b1, self.b2, self.eps, self.wd = float(lr), float(b1), float(b2), float(eps), float(wd)


class UpdateEngine:
    LOG_KEYS = ("train_critic/loss", "train_actor/loss", "train_alpha/loss", "train_alpha/value", "train/aux_loss")

    def __init__(self, action_dim, args, batch_size, device="cuda", algorithm="sgsac", dist=None, global_batch=None,
                 precision="tf32"):
        # precision: "tf32" = tcgen05 TF32 tensor-core convs for SharedCNN layers 2..11 (the product path; the reference
        # runs cuDNN with allow_tf32=True); "fp32" = the same schedule on the fp32 CUDA-core conv kernels (used by the
        # parity tests to check the schedule to 1e-3 without TF32 rounding / ReLU sign flips in the way).
        assert precision in ("tf32", "fp32")
        self.precision = precision
        # dense layers whose operands meet TMA's alignment rules (leading dimensions multiples of 4 floats: the 1024x1024
        # trunks and the 14112 <-> 100 projections) run on tcgen05 with split-precision ("3xTF32") operands in the product
        # mode; the (P + A = 102)-wide layers, the 1-/2A-wide output layers and precision="fp32" use the CUDA-core GEMM
        self.tc_dense = precision == "tf32"
        # (self.overlap / self.side) Independent kernels run on a side stream so that one kernel's prologue (TMEM allocation,
        # weight / first-tile loads) overlaps the other's last tiles: the weight gradient of layer l beside the data-gradient
        # chain, the target encoder beside the online one.  Fork / join with events; in a CUDA graph: parallel branches.
        if not torch.cuda.is_available():
            raise RuntimeError("sgqn-carla_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.dev = torch.device(device)
        self.overlap = precision == "tf32"
        self.side = torch.cuda.Stream(device=self.dev) if self.overlap else None
        self.side2 = torch.cuda.Stream(device=self.dev) if self.overlap else None
        self.side3 = torch.cuda.Stream(device=self.dev) if self.overlap else None     # obs min / max (+ its exchange)
        self.cstream = torch.cuda.Stream(device=self.dev) if (self.overlap and dist is not None) else None   # early gradient pieces
        self._mm_ev, self._early_ev = None, None
        self.args, self.A, self.B = args, int(action_dim), int(batch_size)
        self.Bg = int(global_batch) if global_batch else self.B
        self.dist = dist
        self.algorithm = algorithm
        self.H = int(args.hidden_dim)
        self.lay = ParamLayout(self.A, self.H, int(args.projection_dim), int(args.num_shared_layers), int(args.num_filters),
                               algorithm=algorithm)
        L, dev, B, A, H = self.lay, self.dev, self.B, self.A, self.H
        self.params = torch.zeros(L.total, device=dev)
        # gradient arena: inside the ranks' symmetric (peer-mapped) allocation when the exchange runs over NVLink peer memory
        self.grads = dist.attach(L.total, dev) if hasattr(dist, "attach") else torch.zeros(L.total, device=dev)
        c0, c1 = L.ranges["critic"]
        self.target = torch.zeros(c1 - c0, device=dev)
        self.log_alpha = torch.tensor([math.log(args.init_temperature)], dtype=torch.float64, device=dev)
        self.alpha_st = torch.zeros(2, dtype=torch.float64, device=dev)
        self.alpha_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.alpha_grad = torch.zeros(1, dtype=torch.float64, device=dev)
        self.target_entropy = -float(np.prod((A,)))
        self.opt_critic = _Opt(dev, c1 - c0, args.critic_lr, args.critic_beta, wd=getattr(args, "critic_weight_decay", 0.0))
        a0, a1 = L.ranges["actor"]
        self.opt_actor = _Opt(dev, a1 - a0, args.actor_lr, args.actor_beta)
        x0, x1 = L.ranges["aux"]
        self.opt_aux = _Opt(dev, x1 - x0, getattr(args, "aux_lr", 3e-4), getattr(args, "aux_beta", 0.9))
        self.quantile = float(getattr(args, "sgqn_quantile", 0.5))

        R = 2 * B                                       # rows of the critic heads: [clean | masked] or [obs | aug]
        E = 3 * B                                       # encoder rows of the critic slot: [next_obs | obs | masked / aug]
        self.ns = int(getattr(args, "soda_batch_size", 256)) if algorithm == "soda" else 0     # SODA's own batch (soda.py:16)
        Ec, Rc, Bt = max(E, self.ns), max(R, self.ns), max(B, self.ns)                       # conv workspaces must hold it too
        f32 = lambda *s: torch.zeros(*s, device=dev)
        # One observation buffer so that encoder passes which share weights run as ONE batch: [next_obs ; obs] before
        # the critic update (actor(next_obs) and critic(obs), sac.py:109,114) and [obs ; s_tilde] after it (attribution
        # #2 / the actor update and the aux forward, sgsac.py:175-184) -- s_tilde overwrites the dead masked rows.
        self.obs3 = f32(E, 9, 84, 84)
        self.next_obs = self.obs3[:B]
        self.obs2 = self.obs3[B:]                       # [obs ; masked_obs / overlay-augmented obs]
        self.action = f32(B, A); self.reward = f32(B, 1); self.not_done = f32(B, 1)
        # activations: NHWC, post-ReLU; layers 0..9 carry 2 spare zero rows per sample ([n][h+2][h][32], tcgen05 path)
        self.actS = [f32(Ec * (h + 2) * h * 32) for h in ENC_H]  # critic slot (encoder rows = obs3 rows)
        self.actT = [f32(Bt * (h + 2) * h * 32) for h in ENC_H]  # target-network / acting slot
        self.dbuf = [f32(Rc * 41 * 41 * 32), f32(Rc * 41 * 41 * 32)]
        # im2col matrices of the first conv (col[n*1681][84]) per slot: built once per observation batch by enc_fwd and
        # re-used by that slot's weight gradient; dcol is the attribution's data-gradient workspace
        self.colS, self.colT = f32(Ec * 1681 * 96), f32(B * 1681 * 96)
        self.dcol = f32(B * 1681 * 96) if precision != "tf32" else None
        self.w1p, self.w1p_t = f32(32 * 96), f32(32 * 96)          # TF32 operand copies of cnn.0 ([32][96]) / target
        self.w1d = f32(96 * 32)                                    # ... transposed ([96][32]): data-gradient operand
        # tcgen05 conv path (conv_tc.cu): TF32-rounded operand copies of the 32->32 conv weights (forward; flipped +
        # transposed for the data gradient; forward copy of the target net) and one zero-bordered (pad 2) gradient
        # buffer per layer -- borders are written once here (zeros) and never again.
        self.wf, self.wd, self.wf_t, self.wd_t = f32(10 * 9216), f32(10 * 9216), f32(10 * 9216), f32(10 * 9216)
        self.gpad = [None] + [f32(Rc * (h + 4) * (h + 2) * 32) for h in ENC_H[1:]]    # d(act_l): [n][h+4][h+2][32]
        P1 = L.P + A
        self.zS, self.haS, self.dzS, self.dhaS = f32(R, L.P), f32(R, P1), f32(R, L.P), f32(R, P1)
        self.zT, self.haT, self.dzT, self.dhaT = f32(B, L.P), f32(B, P1), f32(B, L.P), f32(B, P1)
        self.z_a, self.h_a, self.dz_a, self.dh_a = f32(B, L.P), f32(B, L.P), f32(B, L.P), f32(B, L.P)   # actor projection
        self.q = f32(2, R); self.dq = f32(2, R)
        self.z1 = f32(2, R, H); self.z2 = f32(2, R, H); self.dz1 = f32(2, R, H); self.dz2 = f32(2, R, H)
        self.z1t = f32(2, B, H); self.z2t = f32(2, B, H)          # target trunks: own scratch (they run beside the online ones)
        self.az1 = f32(B, H); self.az2 = f32(B, H); self.daz1 = f32(B, H); self.daz2 = f32(B, H)
        self.raw = f32(B, 2 * A); self.draw = f32(B, 2 * A)
        self.mu = f32(B, A); self.pi = f32(B, A); self.log_pi = f32(B); self.log_std = f32(B, A)
        self.next_log_pi = f32(B); self.tq = f32(2, B); self.target_q = f32(B)
        self.ones = torch.ones(B, device=dev)
        self.obs_grad = f32(B, 9, 84, 84)
        self.mask = torch.zeros(B, 3, 84 * 84, dtype=torch.uint8, device=dev)
        self.mm = f32(4); self.mm_scratch = f32(1024)      # {min, max, -min, max} of the obs batch (sgqn_minmax)
        self.logs = f32(8)          # critic_loss, actor_loss, alpha_loss, alpha, aux_loss
        # host-suppliable randomness of one step (SURVEY.md 5 'RNG')
        self.idxs = torch.zeros(B, dtype=torch.int64, device=dev)
        self.overlay_ids = torch.zeros(B, dtype=torch.int64, device=dev)
        self.offs = torch.zeros(2, B, 2, dtype=torch.int32, device=dev)
        self.noise_next = f32(B, A); self.noise_pi = f32(B, A); self.u = f32(1)
        self.rng_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.noise_act = f32(1, A); self.act_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seed = int(getattr(args, "seed", 0))
        self.seed_shared = None         # data-parallel ranks: a seed common to all ranks for the per-batch fill scalar u
        if algorithm == "sgsac":
            self.s_tilde = self.obs2[B:]                 # overlay-augmented obs (written after the critic update)
            self.dl = f32(B, FEAT); self.ddl = f32(B, FEAT)
            if precision != "tf32":
                self.d1 = f32(B * 21 * 21 * 128); self.dd1 = f32(B * 21 * 21 * 128)
                self.d2 = f32(B * 42 * 42 * 64); self.dd2 = f32(B * 42 * 42 * 64)
                self.dup2 = f32(B * 42 * 42 * 128)
                self.lg = f32(B * 86 * 86 * DEC_C3); self.dlg = f32(B * 86 * 86 * DEC_C3)
                self.dup3 = f32(B * 84 * 84 * 64)
            if precision == "tf32":
                # tcgen05 decoder: zero-bordered pitch-linear buffers [B][H+2][W+2][C], image at rows [1,H+1), cols [0,W)
                # (borders are zero from this allocation on; kernels only ever write the interior).  conv2 and conv3 run in
                # their sub-pixel ("phase") form on the PRE-upsample activations (conv_tcg.cu, sgqn_conv_weights_prep_phase):
                # the nearest-x2 upsampled tensors (127 MB + 242 MB at B = 128) and their gradients are never materialised.
                self.xin1 = f32(B * 23 * 23 * 32)        # relu(proj output)
                self.xin2 = f32(B * 23 * 23 * 128)       # relu(conv1)
                self.xin3 = f32(B * 44 * 44 * 64)        # relu(conv2)
                self.lgp = f32(B * 44 * 44 * 64); self.dlgp = f32(B * 44 * 44 * 64)   # logits / d logits, [4 phases][16]
                self.dd2s = f32(B * 23 * 23 * 256)       # d conv2 output in space-to-depth form [4 phases][64]
                self.dd1g = f32(B * 23 * 23 * 128)       # d conv1 output
                self.dwf, self.dwd = f32(128 * 9 * 32), f32(128 * 9 * 32)             # TF32 operand copies of conv1 (fwd / dgrad)
                self.w2f, self.w2d, self.b2p = f32(256 * 9 * 128), f32(256 * 9 * 128), f32(256)
                self.w3f, self.w3d, self.b3p = f32(64 * 9 * 64), f32(64 * 9 * 64), f32(64)
                self.dw2p, self.db2p = f32(256 * 9 * 128), f32(256)
                self.dw3p, self.db3p = f32(64 * 9 * 64), f32(64)
        if algorithm == "soda":
            ns, P = self.ns, L.P
            c0s, c1s = L.ranges["cnn"]; s0, s1 = L.ranges["soda"]
            self.soda_target = f32((c1s - c0s) + (s1 - s0))            # EMA copy of [SharedCNN | soda_proj | soda_pred] (soda.py:30)
            self.wf_st, self.wd_st, self.w1p_st = f32(10 * 9216), f32(10 * 9216), f32(32 * 96)
            self.soda_x, self.soda_aug, self.soda_scratch = f32(ns, 9, 84, 84), f32(ns, 9, 84, 84), f32(ns, 9, 84, 84)
            self.soda_idxs = torch.zeros(ns, dtype=torch.int64, device=dev)
            self.soda_ovl = torch.zeros(ns, dtype=torch.int64, device=dev)
            self.soda_offs = torch.zeros(3, ns, 2, dtype=torch.int32, device=dev)      # [x crop | aug_x crop | zeros]
            self.soda_places = f32(ns, 3, 84 * 84)                                       # host-supplied overlay images (parity runs)
            self.soda_counter = torch.zeros(1, dtype=torch.int64, device=dev)
            self.sy1, self.sa1, self.ss, self.sy2, self.sa2, self.sh0 = (f32(ns, P) for _ in range(6))
            self.sy1t, self.sa1t, self.sh1 = (f32(ns, P) for _ in range(3))
            self.dsh0, self.dsa2, self.dsy2, self.dss, self.dsa1, self.dsy1 = (f32(ns, P) for _ in range(6))
            self.sstat1, self.sstat2, self.sstat_t = f32(2 * P), f32(2 * P), f32(2 * P)
            self.places_pool, self.soda_src, self.soda_supplied = None, None, None
        if algorithm == "pad":
            self.z_p = f32(2 * B, L.P); self.dz_p = f32(2 * B, L.P)     # PAD projection pre-activations of [next_obs ; obs]
            self.joint = f32(B, 2 * L.P); self.djoint = f32(B, 2 * L.P)  # cat[h(obs), h(next_obs)]
            self.pad_pred = f32(B, A); self.dpad_pred = f32(B, A)
        if algorithm == "curl":
            self.pos = f32(B, 9, 84, 84)                # the second random crop of obs (utils.py:152)
            self.pos_scratch = f32(B, 9, 84, 84)
            self.offs_pos = torch.zeros(2, B, 2, dtype=torch.int32, device=dev)
            self.cu = f32(B, L.P); self.dcu = f32(B, L.P)              # u = z_pos W^T and its gradient
            self.clog = f32(B, B); self.dclog = f32(B, B)              # logits z_a u^T and their gradient
            self.dh_c = f32(B, L.P)
        if algorithm == "svea":
            self.places = f32(B, 3, 84 * 84)            # host-supplied overlay images of one step
            self.places_pool, self.places_from_pool = None, False      # float (N,3,84*84) in [0,1] on the device
        # conv_chain.cu: the ten 32->32 layers of an encoder pass (or their data gradients) as ONE persistent launch.  One int
        # workspace (ticket counter, epoch, per-tile flags) per stream that issues chains: main / critic slot, side / target slot.
        # Measured: bit-identical, but at 128-256 samples still 5-10 % slower per pass than ten PDL-chained
        # launches inside a CUDA graph, and the weight gradients lose their overlap with the data-gradient launches -> opt-in.
        self.chain = precision == "tf32" and os.environ.get("SGQN_CHAIN", "0") == "1"
        ws_ints = 4 + max(2 * B, self.ns) * 104 + 64
        self.wsS = torch.zeros(ws_ints, dtype=torch.int32, device=dev)
        self.wsT = torch.zeros(ws_ints, dtype=torch.int32, device=dev)
        self.debug_masked_obs = None
        self.overlay_pool = None       # uint8 (N,3,84*84) device pool for the 'carla' overlay
        self._p = self.params.data_ptr(); self._g = self.grads.data_ptr(); self._t = self.target.data_ptr()
        self._c0 = c0

    # ------------------------------------------------------------------ pointers
    def P(self, name):
        return self._p + 4 * self.lay.off(name)

    def G(self, name):
        return self._g + 4 * self.lay.off(name)

    def T(self, name):                                  # target copy of a critic-range parameter
        return self._t + 4 * (self.lay.off(name) - self._c0)

    def ST(self, name):                                 # SODA target copy: [cnn range | soda range] of the parameter arena
        c0s, c1s = self.lay.ranges["cnn"]
        o = self.lay.off(name)
        if c0s <= o < c1s:
            return self.soda_target.data_ptr() + 4 * (o - c0s)
        return self.soda_target.data_ptr() + 4 * ((c1s - c0s) + o - self.lay.ranges["soda"][0])

    def soda_target_slice(self, name):
        c0s, c1s = self.lay.ranges["cnn"]
        o, stn, _ = self.lay.entries[name]
        base = (o - c0s) if c0s <= o < c1s else (c1s - c0s) + o - self.lay.ranges["soda"][0]
        return self.soda_target[base:base + stn]

    def prep_soda_target_weights(self):
        ls = self.lay.off("cnn.2.weight") - self.lay.off("cnn.1.weight")
        K.conv_weights_prep(self.ST("cnn.1.weight"), ls, _ptr(self.wf_st), _ptr(self.wd_st), 10, self.st)
        K.conv1_weights_prep(self.ST("cnn.0.weight"), _ptr(self.w1p_st), 0, self.st)

    @property
    def st(self):
        return torch.cuda.current_stream().cuda_stream

    def lin_fwd(self, *a):                              # late-bound so that bench.py's per-call profiler sees them
        (K.linear_fwd_tc if self.tc_dense else K.linear_fwd)(*a)

    def lin_dgrad(self, *a):
        (K.linear_dgrad_tc if self.tc_dense else K.linear_dgrad)(*a)

    def lin_wgrad(self, *a):
        (K.linear_wgrad_tc if self.tc_dense else K.linear_wgrad)(*a)

    def _fork(self):
        """Stream for work that only depends on what the current stream has enqueued so far (the side stream once it has
        waited for that point), or the current stream itself when overlap is off."""
        if not self.overlap:
            return self.st
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream()); self.side.wait_event(ev)
        return self.side.cuda_stream

    def _join(self):
        if self.overlap:
            ev = torch.cuda.Event(); ev.record(self.side); torch.cuda.current_stream().wait_event(ev)

    # ------------------------------------------------------------------ building blocks
    def enc_fwd(self, x_ptr, n, acts, row0=0, target=False, hin=84, col_from=None):
        """SharedCNN forward (modules.py:132-152): x (n,9,hin,hin) fp32 NCHW -> acts[0..10] rows [row0, row0+n).
        col_from (tf32 path): the first conv also leaves the im2col matrix of the samples >= col_from in the slot's col
        buffer for the weight gradient of a backward pass over these rows (None: no backward follows)."""
        # target: False = online weights, True = the critic target, "soda" = SODA's own EMA copy of the SharedCNN
        W = self.ST if target == "soda" else (self.T if target else self.P)
        wf = self.wf_st if target == "soda" else (self.wf_t if target else self.wf)
        w1p = self.w1p_st if target == "soda" else (self.w1p_t if target else self.w1p)
        st = self.st
        # activations are stored AFTER the ReLU that follows each conv (the last conv has none) and rounded to TF32,
        # the operand format of the next layer's tcgen05 MMA; 1[x>0] for the backward is 1[relu(x)>0].
        tc = self.precision == "tf32"
        col = _ptr(self.colS if acts is self.actS else self.colT, row0 * 1681 * (96 if tc else 84))
        if tc:
            # pitch-linear layout [n][h+2][h][32]: the 2 spare rows per sample stay zero (never written) so that the
            # weight-gradient kernel can pair activations and the zero-bordered output gradient row by row
            K.conv1_fused_tc(x_ptr, _ptr(w1p), W("cnn.0.bias"), _ptr(acts[0], row0 * 43 * 41 * 32),
                             col if col_from is not None else 0, n, hin, col_from if col_from is not None else n, st)
            rows = []
            for l in range(1, 11):
                hi, ho = ENC_H[l - 1], ENC_H[l]
                last = l == 10                                  # the feature map that feeds the projection is compact
                rows.append((_ptr(acts[l - 1], row0 * (hi + 2) * hi * 32), _ptr(wf, (l - 1) * 9216), W(f"cnn.{l}.bias"), 0,
                             _ptr(acts[l], row0 * (ho * ho if last else (ho + 2) * ho) * 32), 0, n, hi + 2, hi, ho, ho, 0,
                             ho if last else ho + 2, ho, 0, 0, 0, 0, (0 if last else 3) | 16))     # bit 4: weights are old (prep_conv_weights)
            self._convs(rows, self.wsS if acts is self.actS else self.wsT, st)
            return
        K.conv1_im2col(x_ptr, col, n, hin, st)
        K.conv1_fwd_col(col, W("cnn.0.weight"), W("cnn.0.bias"), _ptr(acts[0], row0 * 41 * 41 * 32), n, 1, st)
        for l in range(1, 11):                                  # fp32 CUDA-core path, compact [n][h][h][32] layout
            hi, ho = ENC_H[l - 1], ENC_H[l]
            K.conv_fwd(_ptr(acts[l - 1], row0 * hi * hi * 32), W(f"cnn.{l}.weight"), W(f"cnn.{l}.bias"),
                       _ptr(acts[l], row0 * ho * ho * 32), n, hi, hi, 32, 32, 0, 1, 0, 1 if l < 10 else 0, st)

    def _convs(self, rows, ws, st):
        """A chain of 32->32 tcgen05 convs (each row = the arguments of one sgqn_conv_tc call, each reading what the one before
        wrote): one persistent launch (conv_chain.cu), or one launch per layer with SGQN_CHAIN=0."""
        if self.chain and rows[0][6] >= 4:                   # (batch-1 acting: ten tiny launches are cheaper than the ticket traffic)
            arr, addr, n = conv_layers(rows)
            K.conv_chain(addr, n, _ptr(ws), ws.numel(), st)
        else:
            for r in rows:
                K.conv_tc(*r, st)

    def prep_dec_weights(self):
        if self.algorithm != "sgsac" or self.precision != "tf32":
            return
        K.conv_weights_prep_g(self.P("dec.conv1.weight"), _ptr(self.dwf), _ptr(self.dwd), 128, 32, 128, self.st)
        K.conv_weights_prep_phase(self.P("dec.conv2.weight"), self.P("dec.conv2.bias"), _ptr(self.w2f), _ptr(self.w2d),
                                  _ptr(self.b2p), 128, 64, 64, self.st)
        K.conv_weights_prep_phase(self.P("dec.conv3.weight"), self.P("dec.conv3.bias"), _ptr(self.w3f), _ptr(self.w3d),
                                  _ptr(self.b3p), 64, 9, 16, self.st)

    def prep_conv_weights(self, target=False):
        """Refresh the TF32 operand copies after the 32->32 conv weights changed (optimiser step / EMA / load)."""
        L = self.lay
        ls = L.off("cnn.2.weight") - L.off("cnn.1.weight")
        if target:
            K.conv_weights_prep(self.T("cnn.1.weight"), ls, _ptr(self.wf_t), _ptr(self.wd_t), 10, self.st)
            K.conv1_weights_prep(self.T("cnn.0.weight"), _ptr(self.w1p_t), 0, self.st)
        else:
            K.conv_weights_prep(self.P("cnn.1.weight"), ls, _ptr(self.wf), _ptr(self.wd), 10, self.st)
            K.conv1_weights_prep(self.P("cnn.0.weight"), _ptr(self.w1p), _ptr(self.w1d), self.st)

    def proj_fwd(self, feat_ptr, n, pre, z, h, ldh, target=False):
        """RLProjection (modules.py:102-113): Linear(14112->100) (split-K) -> LayerNorm -> tanh, h row stride ldh."""
        W = self.T if target else self.P
        st = self.st
        (self.lin_fwd if n >= 32 else K.linear_fwd)(feat_ptr, FEAT, 0, W(f"{pre}.0.weight"), 0, W(f"{pre}.0.bias"), 0, z, self.lay.P, 0,
                                                    n, self.lay.P, FEAT, 0, 1, 2, st)
        K.ln_tanh_fwd(z, W(f"{pre}.1.weight"), W(f"{pre}.1.bias"), h, ldh, n, self.lay.P, st)

    def q_fwd(self, ha, n, row0, nheads=2, target=False):
        """Q1/Q2 trunks (modules.py:235-261) on ha (n, P+A); both heads in one launch per layer."""
        W = self.T if target else self.P
        L, H, R, st = self.lay, self.H, 2 * self.B, self.st
        P1, qs = L.P + self.A, L.q_stride
        z1, z2 = _ptr(self.z1, row0 * H), _ptr(self.z2, row0 * H)
        out = _ptr(self.tq) if target else _ptr(self.q, row0)
        obs_ = self.B if target else R
        zs = R * H                                      # head stride of the hidden activations
        if target:
            z1, z2, zs = _ptr(self.z1t), _ptr(self.z2t), self.B * H
        K.linear_fwd(ha, P1, 0, W("Q1.0.weight"), qs, W("Q1.0.bias"), qs, z1, H, zs, n, H, P1, 0, nheads, 0, st)
        self.lin_fwd(z1, H, zs, W("Q1.2.weight"), qs, W("Q1.2.bias"), qs, z2, H, zs, n, H, H, 1, nheads, 2, st)
        K.linear_fwd(z2, H, zs, W("Q1.4.weight"), qs, W("Q1.4.bias"), qs, out, 1, obs_, n, 1, H, 1, nheads, 2, st)

    def q_dgrad(self, dq, dq_bs, n, row0, nheads, mode, dha):
        """Backward of the Q trunks to their input (n, P+A), heads summed.  mode 1 plain, 2 guided."""
        L, H, R, st = self.lay, self.H, 2 * self.B, self.st
        P1, qs = L.P + self.A, L.q_stride
        z1, z2 = _ptr(self.z1, row0 * H), _ptr(self.z2, row0 * H)
        dz1, dz2 = _ptr(self.dz1, row0 * H), _ptr(self.dz2, row0 * H)
        K.linear_dgrad(dq, 1, dq_bs, self.P("Q1.4.weight"), qs, z2, H, R * H, dz2, H, R * H, n, 1, H, mode, 0, nheads, st)
        self.lin_dgrad(dz2, H, R * H, self.P("Q1.2.weight"), qs, z1, H, R * H, dz1, H, R * H, n, H, H, mode,
                       2 if mode == 1 else 0, nheads, st)
        K.zero(dha, 4 * n * P1, st)
        K.linear_dgrad(dz1, H, R * H, self.P("Q1.0.weight"), qs, 0, 0, 0, dha, P1, 0, n, H, P1, 0, 1, nheads, st)

    def q_wgrad(self, ha, dq, n, row0, st=None):
        L, H, R = self.lay, self.H, 2 * self.B
        st = st if st is not None else self.st
        P1, qs = L.P + self.A, L.q_stride
        z1, z2 = _ptr(self.z1, row0 * H), _ptr(self.z2, row0 * H)
        dz1, dz2 = _ptr(self.dz1, row0 * H), _ptr(self.dz2, row0 * H)
        K.linear_wgrad(z2, H, R * H, dq, 1, R, self.G("Q1.4.weight"), qs, self.G("Q1.4.bias"), qs, n, 1, H, 1, 2, st)
        self.lin_wgrad(z1, H, R * H, dz2, H, R * H, self.G("Q1.2.weight"), qs, self.G("Q1.2.bias"), qs, n, H, H, 1, 2, st)
        K.linear_wgrad(ha, P1, 0, dz1, H, R * H, self.G("Q1.0.weight"), qs, self.G("Q1.0.bias"), qs, n, H, P1, 0, 2, st)

    def proj_bwd(self, dh, lddh, n, z, h, ldh, pre, dz, feat_ptr=0, dfeat=0, wgrad=True):
        st, P = self.st, self.lay.P
        K.ln_tanh_bwd(dh, lddh, z, h, ldh, self.P(f"{pre}.1.weight"), dz,
                      self.G(f"{pre}.1.weight") if wgrad else 0, self.G(f"{pre}.1.bias") if wgrad else 0, n, P, st)
        if wgrad:                                         # beside the data gradient when one follows (joined by enc_bwd)
            self.lin_wgrad(feat_ptr, FEAT, 0, dz, P, 0, self.G(f"{pre}.0.weight"), 0, self.G(f"{pre}.0.bias"), 0,
                           n, P, FEAT, 0, 1, self._fork() if dfeat else st)
        if dfeat:
            self.lin_dgrad(dz, P, 0, self.P(f"{pre}.0.weight"), 0, 0, 0, 0, dfeat, FEAT, 0, n, P, FEAT, 0, 0, 1, st)

    def enc_bwd(self, dfeat, n, acts, row0, x_ptr, mode, wgrad, dobs=0):
        """Backward through SharedCNN.  dfeat: (n,21,21,32).  mode 1: plain ReLU backward (+ wgrad into the grad
        arena); mode 2: guided backprop to the observation (rl_utils.py:35-39)."""
        st = self.st
        if self.precision != "tf32":
            d = dfeat
            for l in range(10, 0, -1):
                hi = ENC_H[l - 1]
                a_in = _ptr(acts[l - 1], row0 * hi * hi * 32)
                if wgrad:
                    K.conv_wgrad(a_in, d, self.G(f"cnn.{l}.weight"), self.G(f"cnn.{l}.bias"), n, hi, hi, 32, 32, 0, 1, 0, 0, st)
                dx = _ptr(self.dbuf[l & 1])
                K.conv_dgrad(d, self.P(f"cnn.{l}.weight"), a_in, dx, n, hi, hi, 32, 32, 0, mode, st)
                d = dx
            self._conv1_bwd(d, n, acts, row0, wgrad, dobs)
            return
        K.pad_copy(dfeat, _ptr(self.gpad[10]), n, 21, 21, 32, 25, 23, 2, 0, 1, st)
        side = self.side if (wgrad and self.overlap) else None
        main = torch.cuda.current_stream()
        if self.chain:
            # the ten data gradients as one launch; the weight gradients (they need every d(act_l)) follow on the side stream
            # beside the first-conv backward
            rows = []
            for l in range(10, 0, -1):
                hi, ho = ENC_H[l - 1], ENC_H[l]
                a_in = _ptr(acts[l - 1], row0 * (hi + 2) * hi * 32)
                db = self.G(f"cnn.{l - 1}.bias") if wgrad else 0
                if l > 1:
                    rows.append((_ptr(self.gpad[l]), _ptr(self.wd, (l - 1) * 9216), 0, a_in, _ptr(self.gpad[l - 1]), db, n, ho + 4, ho + 2,
                                 hi, hi, -2, hi + 4, hi + 2, 2, 0, hi + 2, hi, 2 | (mode << 2) | 16))
                else:
                    rows.append((_ptr(self.gpad[1]), _ptr(self.wd), 0, a_in, _ptr(self.dbuf[0]), db, n, ho + 4, ho + 2, hi, hi, -2,
                                 hi, hi, 0, 0, hi + 2, hi, 2 | (mode << 2) | 16))
            self._convs(rows, self.wsS, st)                  # (backward passes run on the update's main stream only)
            if wgrad:
                ws = st
                if side is not None:
                    ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
                    ws = side.cuda_stream
                for l in range(10, 0, -1):
                    hi = ENC_H[l - 1]
                    K.conv_wgrad_tc(_ptr(acts[l - 1], row0 * (hi + 2) * hi * 32), _ptr(self.gpad[l]), self.G(f"cnn.{l}.weight"), n, hi + 2, hi, ws)
                K.colsum(_ptr(self.gpad[10]), 32, n * 25 * 23, 32, self.G("cnn.10.bias"), ws)
            self._conv1_bwd(_ptr(self.dbuf[0]), n, acts, row0, wgrad, dobs)
            if side is not None:
                ev = torch.cuda.Event(); ev.record(side); main.wait_event(ev)
            return
        for l in range(10, 0, -1):
            hi, ho = ENC_H[l - 1], ENC_H[l]
            a_in = _ptr(acts[l - 1], row0 * (hi + 2) * hi * 32)     # [n][hi+2][hi][32], post-ReLU
            d = _ptr(self.gpad[l])                                  # d(act_l): [n][ho+4][ho+2][32] == [n][hi+2][hi][32]
            if wgrad:
                ws = st
                if side is not None:                    # d(act_l) is complete on the main stream: fork
                    ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
                    ws = side.cuda_stream
                K.conv_wgrad_tc(a_in, d, self.G(f"cnn.{l}.weight"), n, hi + 2, hi, ws)
                if l == 10:                             # the other layers' bias gradients ride on the data-gradient epilogues
                    K.colsum(d, 32, n * (hi + 2) * hi, 32, self.G("cnn.10.bias"), ws)
            db = self.G(f"cnn.{l - 1}.bias") if wgrad else 0    # sum of d(act_{l-1}) = bias gradient of layer l-1
            if l > 1:
                K.conv_tc(d, _ptr(self.wd, (l - 1) * 9216), 0, a_in, _ptr(self.gpad[l - 1]), db, n, ho + 4, ho + 2, hi, hi, -2,
                          hi + 4, hi + 2, 2, 0, hi + 2, hi, 2 | (mode << 2) | 16, st)
            else:                                       # d(act_0) compact: consumed by the first-conv kernels
                K.conv_tc(d, _ptr(self.wd), 0, a_in, _ptr(self.dbuf[0]), db, n, ho + 4, ho + 2, hi, hi, -2,
                          hi, hi, 0, 0, hi + 2, hi, 2 | (mode << 2) | 16, st)
        self._conv1_bwd(_ptr(self.dbuf[0]), n, acts, row0, wgrad, dobs)
        if side is not None:                            # join: the gradient buffers are reused by the next backward
            ev = torch.cuda.Event(); ev.record(side); main.wait_event(ev)

    def _conv1_bwd(self, d, n, acts, row0, wgrad, dobs):
        """Backward of the first conv from d = d(act_0) (compact [n][41][41][32]) through the slot's im2col matrix."""
        st = self.st
        tc = self.precision == "tf32"
        col = _ptr(self.colS if acts is self.actS else self.colT, row0 * 1681 * (96 if tc else 84))
        if wgrad and tc:
            K.gemm_wgrad_tcg(col, d, self.G("cnn.0.weight"), n, 41, 41, 96, 32, 0, 0, 1, 81, st)   # (bias gradient: see enc_bwd)
        elif wgrad:
            K.conv1_wgrad_col(col, d, self.G("cnn.0.weight"), self.G("cnn.0.bias"), n, st)
        if dobs and tc:                                   # dcol = d(act_0) W on tcgen05 + the gather to NCHW, one kernel
            K.conv1_dgrad_fused_tc(d, _ptr(self.w1d), dobs, n, st)
        elif dobs:
            K.conv1_dgrad_col(d, self.P("cnn.0.weight"), _ptr(self.dcol), dobs, n, st)

    def attribution(self, erow, ha, z, obs_grad):
        """compute_attribution (rl_utils.py:57-62): guided backprop of sum_b Q1[b] to the observation, re-using the
        forward activations of critic(obs): critic-slot encoder rows [erow, erow+B), head rows [0, B)."""
        B, L, st = self.B, self.lay, self.st
        acts = self.actS
        self.q_dgrad(_ptr(self.ones), 0, B, 0, 1, 2, _ptr(self.dhaT))
        K.ln_tanh_bwd(_ptr(self.dhaT), L.P + self.A, z, ha, L.P + self.A, self.P("critic_proj.1.weight"), _ptr(self.dzT),
                      0, 0, B, L.P, st)
        dfeat = _ptr(self.dbuf[1])
        self.lin_dgrad(_ptr(self.dzT), L.P, 0, self.P("critic_proj.0.weight"), 0, 0, 0, 0, dfeat, FEAT, 0, B, L.P, FEAT,
                       0, 0, 1, st)
        self.enc_bwd(dfeat, B, acts, erow, 0, 2, False, dobs=obs_grad)

    def adam(self, opt, rng, target=None, n_tau0=0, tau0=0.0, tau1=0.0):
        st = self.st
        o0, o1 = rng
        K.adam_prep(_ptr(opt.step), _ptr(opt.bc), opt.b1, opt.b2, st)
        K.adam(self._p + 4 * o0, self._g + 4 * o0, _ptr(opt.m), _ptr(opt.v), o1 - o0, _ptr(opt.bc), opt.lr,
               float(np.float32(1 - opt.b1)), opt.b2, float(np.float32(1 - opt.b2)), opt.eps,
               target if target else 0, n_tau0, tau0, tau1, opt.wd, st)

    def allreduce_grads(self, rng, group="main"):
        if self.dist is not None:
            self.dist.all_reduce_sum(self.grads[rng[0]:rng[1]], group)

    def _early_reduce(self, rng, after=None):
        """Sharded configuration: sum-all-reduce grads[rng] as soon as the stream `after` (default: the current one) has
        produced it, on the communication stream / the "early" communicator, beside the rest of the backward pass.  The
        pieces of a bucket that are ready before its encoder backward starts (Q heads 9.2 MB, projection 5.6 MB, decoder
        6.9 MB) overlap the data-gradient chain; only the conv range (0.38 MB) is left for the main stream."""
        if self.dist is None:
            return
        if self.cstream is None:
            self.dist.all_reduce_sum(self.grads[rng[0]:rng[1]], "early")
            return
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream()); self.cstream.wait_event(ev)
        if after is not None:                           # pieces of the range were produced on a forked stream
            ev = torch.cuda.Event(); ev.record(after); self.cstream.wait_event(ev)
        with torch.cuda.stream(self.cstream):
            self.dist.all_reduce_sum(self.grads[rng[0]:rng[1]], "early")
            self._early_ev = torch.cuda.Event(); self._early_ev.record(self.cstream)

    def _early_join(self):
        if self._early_ev is not None:
            torch.cuda.current_stream().wait_event(self._early_ev)
            self._early_ev = None

    # ------------------------------------------------------------------ the update
    def target_q_pass(self, critic_rows=False):
        """sac.py:108-112 / sgsac.py:53-57 (no grad): actor(next_obs), critic_target(next_obs, a'); with critic_rows also
        critic(obs, action) (sac.py:114 / sgsac.py:59) on the obs half of the shared encoder pass.
        After the online encoder three chains of small launches are independent -- actor(next_obs) (5), the target
        projection (2) and the online critic heads on obs (6) -- and each fills a fraction of the SMs: with overlap they run on
        three streams and meet at the target Q trunks / at the end."""
        B, A, L, st = self.B, self.A, self.lay, self.st
        a = self.args
        nx = _ptr(self.next_obs)
        ev_t = ev_c = None
        if self.overlap:                                # target encoder (B rows) beside the online one (2B rows)
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event(); ev.record(main); self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                self.enc_fwd(nx, B, self.actT, target=True)
                self.proj_fwd(_ptr(self.actT[10]), B, "critic_proj", _ptr(self.zT), _ptr(self.haT), L.P + A, target=True)
                ev_t = torch.cuda.Event(); ev_t.record(self.side)
        self.enc_fwd(_ptr(self.obs3), 2 * B, self.actS, 0, col_from=B)   # online encoder over [next_obs ; obs] in one batch (the obs half is
                                                                         # differentiated by the critic backward)
        if critic_rows and self.overlap:
            ev = torch.cuda.Event(); ev.record(main); self.side2.wait_event(ev)
            with torch.cuda.stream(self.side2):
                self.critic_fwd_rows(0, B, encode=False)
                ev_c = torch.cuda.Event(); ev_c.record(self.side2)
        self.proj_fwd(_ptr(self.actS[10]), B, "actor_proj", _ptr(self.z_a), _ptr(self.h_a), L.P)
        self.actor_mlp_fwd(B)
        K.actor_head_fwd(_ptr(self.raw), _ptr(self.noise_next), float(a.actor_log_std_min), float(a.actor_log_std_max),
                         0, _ptr(self.haT, L.P), L.P + A, _ptr(self.next_log_pi), 0, B, A, st)
        if ev_t is not None:
            torch.cuda.current_stream().wait_event(ev_t)
        else:
            self.enc_fwd(nx, B, self.actT, target=True)
            self.proj_fwd(_ptr(self.actT[10]), B, "critic_proj", _ptr(self.zT), _ptr(self.haT), L.P + A, target=True)
        self.q_fwd(_ptr(self.haT), B, 0, 2, target=True)
        if ev_c is not None:
            torch.cuda.current_stream().wait_event(ev_c)
        elif critic_rows:
            self.critic_fwd_rows(0, B, encode=False)

    def actor_mlp_fwd(self, n):
        L, H, A, st = self.lay, self.H, self.A, self.st
        K.linear_fwd(_ptr(self.h_a), L.P, 0, self.P("actor_mlp.0.weight"), 0, self.P("actor_mlp.0.bias"), 0,
                     _ptr(self.az1), H, 0, n, H, L.P, 0, 1, 0, st)
        (self.lin_fwd if n >= 32 else K.linear_fwd)(_ptr(self.az1), H, 0, self.P("actor_mlp.2.weight"), 0, self.P("actor_mlp.2.bias"), 0,
                                                    _ptr(self.az2), H, 0, n, H, H, 1, 1, 2, st)
        K.linear_fwd(_ptr(self.az2), H, 0, self.P("actor_mlp.4.weight"), 0, self.P("actor_mlp.4.bias"), 0,
                     _ptr(self.raw), 2 * A, 0, n, 2 * A, H, 1, 1, 2, st)

    def critic_fwd_rows(self, row0, n, encode=True):
        """critic(obs2[row0:row0+n], action) with activations kept in the critic slot (encoder rows B + row0 ...)."""
        L, A, st, B = self.lay, self.A, self.st, self.B
        P1 = L.P + A
        if encode:
            self.enc_fwd(_ptr(self.obs2, row0 * 9 * 84 * 84), n, self.actS, B + row0, col_from=0)
        K.set_cols(_ptr(self.haS, row0 * P1), P1, L.P, _ptr(self.action), A, n, A, st)
        self.proj_fwd(_ptr(self.actS[10], (B + row0) * FEAT), n, "critic_proj", _ptr(self.zS, row0 * L.P),
                      _ptr(self.haS, row0 * P1), P1)
        self.q_fwd(_ptr(self.haS, row0 * P1), n, row0)

    def _obs_minmax_fork(self):
        """Global min / max of the sampled obs batch (sgsac.py:68-69).  It depends on the batch only, so it -- and, in the
        sharded configuration, its 2-float all-reduce -- is issued beside the target / critic forward passes (own stream,
        own communicator) and joined just before the mask kernel needs it."""
        B = self.B
        if not self.overlap:
            K.minmax(_ptr(self.obs2), B * 9 * 84 * 84, _ptr(self.mm_scratch), _ptr(self.mm), self.st)
            if self.dist is not None:
                self.dist.all_reduce_minmax(self.mm)
            self._mm_ev = None
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main); self.side3.wait_event(ev)
        with torch.cuda.stream(self.side3):
            K.minmax(_ptr(self.obs2), B * 9 * 84 * 84, _ptr(self.mm_scratch), _ptr(self.mm), self.st)
            if self.dist is not None:
                self.dist.all_reduce_minmax(self.mm)
            self._mm_ev = torch.cuda.Event(); self._mm_ev.record(self.side3)

    def _obs_minmax_join(self):
        if self._mm_ev is not None:
            torch.cuda.current_stream().wait_event(self._mm_ev)
            self._mm_ev = None

    def update_critic(self, mode):
        """mode 0: SAC (sac.py:107-123); 1: SGSAC with consistency (sgsac.py:52-80); 2: SVEA (svea.py:19-52)."""
        B, A, L, st, a = self.B, self.A, self.lay, self.st, self.args
        P1 = L.P + A
        if mode == 1:
            self._obs_minmax_fork()
        self.target_q_pass(critic_rows=True)            # + critic(obs, action): obs went through the encoder with next_obs
        R = B
        if mode == 1:
            self.attribution(B, _ptr(self.haS), _ptr(self.zS), _ptr(self.obs_grad))
            self._obs_minmax_join()
            sharded = self.dist is not None             # then mm[2:4] = {-min, max} of the GLOBAL batch (sgsac.py:68-70)
            K.attribution_mask(_ptr(self.obs_grad), _ptr(self.obs2), _ptr(self.mm, 2 if sharded else 0), _ptr(self.u), self.quantile,
                               _ptr(self.mask), _ptr(self.obs2, B * 9 * 84 * 84), B, 84 * 84, 1 if sharded else 0, st)
            if self.debug_masked_obs is not None:           # parity tests: feed the oracle's masked obs forward
                self.debug_own_masked_obs = self.obs2[B:].clone()
                self.obs2[B:].copy_(self.debug_masked_obs)
                self.debug_masked_obs = None
            self.critic_fwd_rows(B, B)
            R = 2 * B
        elif mode == 2:
            al = 0.2                                             # augmentations.py:79 default (svea.py:26 passes none)
            if self.places_from_pool:                            # rows overlay_ids of the device-resident places pool
                K.overlay_f32(_ptr(self.obs2), _ptr(self.places_pool), _ptr(self.overlay_ids), float(np.float32(1 - al)),
                              float(np.float32(al)), _ptr(self.obs2, B * 9 * 84 * 84), B, 84 * 84, st)
            else:                                                # host-supplied batch of images (parity runs)
                K.overlay_f32(_ptr(self.obs2), _ptr(self.places), 0, float(np.float32(1 - al)), float(np.float32(al)),
                              _ptr(self.obs2, B * 9 * 84 * 84), B, 84 * 84, st)
            self.critic_fwd_rows(B, B)
            R = 2 * B
        wa, wb = float(getattr(a, "svea_alpha", 0.5)), float(getattr(a, "svea_beta", 0.5))
        K.critic_loss(_ptr(self.q), 2 * B, _ptr(self.tq), _ptr(self.tq, B), _ptr(self.next_log_pi), _ptr(self.reward),
                      _ptr(self.not_done), _ptr(self.log_alpha), float(a.discount), mode, wa, wb, _ptr(self.target_q),
                      _ptr(self.dq), _ptr(self.logs, 0), B, self.Bg, st)
        # backward over all R rows at once
        c0, c1 = L.ranges["critic"]
        K.zero(self._g + 4 * c0, 4 * (c1 - c0), st)
        self.q_dgrad(_ptr(self.dq), 2 * B, R, 0, 2, 1, _ptr(self.dhaS))
        wside = self.side if self.overlap else None
        self.q_wgrad(_ptr(self.haS), _ptr(self.dq), R, 0, st=self._fork())     # joined at the end of enc_bwd
        self._early_reduce(L.ranges["critic_q"], after=wside)
        dfeat = _ptr(self.dbuf[1])
        self.proj_bwd(_ptr(self.dhaS), P1, R, _ptr(self.zS), _ptr(self.haS), P1, "critic_proj", _ptr(self.dzS),
                      feat_ptr=_ptr(self.actS[10], B * FEAT), dfeat=dfeat)
        self._early_reduce(L.ranges["critic_proj"], after=wside)
        self.enc_bwd(dfeat, R, self.actS, B, _ptr(self.obs2), 1, True)
        self.allreduce_grads(L.ranges["cnn"])
        self._early_join()

    def critic_step(self, with_ema):
        L, a = self.lay, self.args
        c0, c1 = L.ranges["critic"]
        q0, q1 = L.ranges["critic_q"]
        self.adam(self.opt_critic, (c0, c1), target=self._t if with_ema else None, n_tau0=q1 - q0,
                  tau0=float(a.critic_tau), tau1=float(a.encoder_tau))
        self.prep_conv_weights()
        if with_ema:
            self.prep_conv_weights(target=True)

    def shared_obs_fwd(self, with_aux=False):
        """One encoder + projection forward with the updated critic weights of obs -- shared by attribution #2
        (sgsac.py:175), actor(obs, detach) and critic(obs, pi, detach) (sac.py:126-127) -- and, batched with it when
        the aux update follows, of the overlay-augmented s_tilde (sgsac.py:84-89; the actor update in between does not
        touch these weights).  Re-uses the critic slot, whose contents are dead once the critic gradients are out:
        encoder rows [B, 2B) / head rows [0, B) = obs, encoder rows [2B, 3B) / head rows [B, 2B) = s_tilde."""
        B, A, L, st, a = self.B, self.A, self.lay, self.st, self.args
        P1 = L.P + A
        n = B
        if with_aux:
            al = float(getattr(a, "alpha_blending", 0.2))
            K.overlay_u8(_ptr(self.obs2), _ptr(self.overlay_pool), _ptr(self.overlay_ids), float(np.float32(1 - al)),
                         float(np.float32(al)), _ptr(self.s_tilde), B, 84 * 84, st)
            n = 2 * B
        # only the s_tilde half gets a backward pass (the aux update); obs feeds attribution #2 / the detached actor update
        self.enc_fwd(_ptr(self.obs2), n, self.actS, B, col_from=B if with_aux else None)
        K.set_cols(_ptr(self.haS), P1, L.P, _ptr(self.action), A, B, A, st)
        if with_aux:
            K.set_cols(_ptr(self.haS, B * P1), P1, L.P, _ptr(self.action), A, B, A, st)
        self.proj_fwd(_ptr(self.actS[10], B * FEAT), n, "critic_proj", _ptr(self.zS), _ptr(self.haS), P1)

    def attribution2(self, want_mask):
        B, st = self.B, self.st
        self.q_fwd(_ptr(self.haS), B, 0, 1)
        self.attribution(B, _ptr(self.haS), _ptr(self.zS), _ptr(self.obs_grad))
        if want_mask:
            K.attribution_mask(_ptr(self.obs_grad), 0, 0, 0, self.quantile, _ptr(self.mask), 0, B, 84 * 84, 0, st)

    def update_actor_and_alpha(self, finish=True, fork_wgrad=True):
        """sac.py:125-151; expects shared_obs_fwd() state (critic slot, head rows [0,B)).  fork_wgrad: the MLP's weight
        gradients on the side stream beside its data gradients (off when the whole update already runs beside the aux update)."""
        B, A, L, H, st, a = self.B, self.A, self.lay, self.H, self.st, self.args
        P1 = L.P + A
        lmin, lmax = float(a.actor_log_std_min), float(a.actor_log_std_max)
        self.proj_fwd(_ptr(self.actS[10], B * FEAT), B, "actor_proj", _ptr(self.z_a), _ptr(self.h_a), L.P)
        self.actor_mlp_fwd(B)
        K.actor_head_fwd(_ptr(self.raw), _ptr(self.noise_pi), lmin, lmax, 0, _ptr(self.haS, L.P), P1, _ptr(self.log_pi),
                         0, B, A, st)
        self.q_fwd(_ptr(self.haS), B, 0, 2)
        K.actor_loss(_ptr(self.q), 2 * B, _ptr(self.log_pi), _ptr(self.log_alpha), self.target_entropy, _ptr(self.dq),
                     _ptr(self.logs, 1), _ptr(self.alpha_grad), B, self.Bg, st)
        self.q_dgrad(_ptr(self.dq), 2 * B, B, 0, 2, 1, _ptr(self.dhaS))
        K.actor_head_bwd(_ptr(self.raw), _ptr(self.noise_pi), _ptr(self.dhaS, L.P), P1, _ptr(self.log_alpha), lmin, lmax,
                         _ptr(self.draw), B, A, self.Bg, st)
        a0, a1 = L.ranges["actor"]
        K.zero(self._g + 4 * a0, 4 * (a1 - a0), st)
        # actor MLP backward
        fork = fork_wgrad and self.overlap
        K.linear_dgrad(_ptr(self.draw), 2 * A, 0, self.P("actor_mlp.4.weight"), 0, _ptr(self.az2), H, 0, _ptr(self.daz2), H, 0,
                       B, 2 * A, H, 1, 0, 1, st)
        ws = self._fork() if fork else st
        K.linear_wgrad(_ptr(self.az2), H, 0, _ptr(self.draw), 2 * A, 0, self.G("actor_mlp.4.weight"), 0,
                       self.G("actor_mlp.4.bias"), 0, B, 2 * A, H, 1, 1, ws)
        self.lin_wgrad(_ptr(self.az1), H, 0, _ptr(self.daz2), H, 0, self.G("actor_mlp.2.weight"), 0,
                       self.G("actor_mlp.2.bias"), 0, B, H, H, 1, 1, ws)
        self.lin_dgrad(_ptr(self.daz2), H, 0, self.P("actor_mlp.2.weight"), 0, _ptr(self.az1), H, 0, _ptr(self.daz1), H, 0,
                       B, H, H, 1, 2, 1, st)
        ws = self._fork() if fork else st
        K.linear_wgrad(_ptr(self.h_a), L.P, 0, _ptr(self.daz1), H, 0, self.G("actor_mlp.0.weight"), 0,
                       self.G("actor_mlp.0.bias"), 0, B, H, L.P, 0, 1, ws)
        K.linear_dgrad(_ptr(self.daz1), H, 0, self.P("actor_mlp.0.weight"), 0, 0, 0, 0, _ptr(self.dh_a), L.P, 0,
                       B, H, L.P, 0, 2, 1, st)
        self.proj_bwd(_ptr(self.dh_a), L.P, B, _ptr(self.z_a), _ptr(self.h_a), L.P, "actor_proj", _ptr(self.dz_a),
                      feat_ptr=_ptr(self.actS[10], B * FEAT), dfeat=0)
        if fork:
            self._join()
        if finish:
            self.actor_finish()

    def actor_finish(self):
        """Gradient exchange + optimiser steps of the actor / alpha update (separate so that, when the backward ran on the
        second stream, every collective is still issued from the main stream in one fixed order on all ranks)."""
        a, st = self.args, self.st
        a0, a1 = self.lay.ranges["actor"]
        self.allreduce_grads((a0, a1), "actor")         # (own communicator: issued from the stream the actor update runs on)
        if self.dist is not None:
            self.dist.all_reduce_sum(self.alpha_grad, "actor")
        self.adam(self.opt_actor, (a0, a1))
        K.alpha_adam(_ptr(self.log_alpha), _ptr(self.alpha_grad), _ptr(self.alpha_st), _ptr(self.alpha_step),
                     float(a.alpha_lr), float(a.alpha_beta), 0.999, 1e-8, st)

    def update_aux(self, phase="all"):
        """sgsac.py:82-102,163-167: overlay -> attribution predictor -> BCE vs mask of attribution #2.  The overlay, the
        encoder and the projection of s_tilde ran in shared_obs_fwd(with_aux=True): head rows [B, 2B).
        phase "fwd": the predictor's forward up to the logits (needs neither the mask nor anything attribution #2 touches, so
        update_sgsac runs it beside attribution #2's chain of small head kernels); "bwd": BCE against the mask onwards."""
        B, A, L, st, a = self.B, self.A, self.lay, self.st, self.args
        P1 = L.P + A
        ha, z = _ptr(self.haS, B * P1), _ptr(self.zS, B * L.P)
        feat = _ptr(self.actS[10], 2 * B * FEAT)
        Wp, G = self.P, self.G
        x0, x1 = L.ranges["aux"]
        if phase in ("all", "fwd"):
            K.linear_fwd(ha, P1, 0, Wp("dec.proj.weight"), 0, Wp("dec.proj.bias"), 0, _ptr(self.dl), FEAT, 0,
                         B, FEAT, P1, 0, 1, 0, st)
            if self.precision == "tf32":
                self._decoder_tc_fwd(B, st, Wp)
            if phase == "fwd":
                return
        if self.precision == "tf32":
            self._decoder_tc_bwd(B, st, Wp, G, x0, x1)
        else:
            self._decoder_simt(B, st, Wp, G, x0, x1)
        wside = self.side if self.overlap else None     # (the decoder's weight gradients were issued there, _decoder_tc_bwd)
        K.linear_wgrad(ha, P1, 0, _ptr(self.ddl), FEAT, 0, G("dec.proj.weight"), 0, G("dec.proj.bias"), 0,
                       B, FEAT, P1, 0, 1, self._fork())
        self._early_reduce(L.ranges["dec"], after=wside)     # decoder gradients are complete: exchange them under the encoder backward
        K.linear_dgrad(_ptr(self.ddl), FEAT, 0, Wp("dec.proj.weight"), 0, 0, 0, 0, _ptr(self.dhaT), P1, 0, B, FEAT, P1, 0, 2, 1, st)
        dfeat = _ptr(self.dbuf[1])
        self.proj_bwd(_ptr(self.dhaT), P1, B, z, ha, P1, "critic_proj", _ptr(self.dzT), feat_ptr=feat, dfeat=dfeat)
        self._early_reduce(L.ranges["critic_proj"], after=wside)
        self.enc_bwd(dfeat, B, self.actS, 2 * B, _ptr(self.s_tilde), 1, True)
        self.allreduce_grads(L.ranges["cnn"])           # (fdec.* never receives a gradient: not exchanged)
        self._early_join()
        self.adam(self.opt_aux, (x0, x1))
        self.prep_conv_weights()
        self.prep_dec_weights()

    def predict_attribution(self):
        """attribution_predictor(obs, action) forward only (modules.py:345-354; sgsac.py:107 in log_tensorboard): expects
        shared_obs_fwd() state (critic-slot head rows [0,B) = obs); returns the logits as (B,9,84,84)."""
        B, A, L, st = self.B, self.A, self.lay, self.st
        P1 = L.P + A
        Wp = self.P
        K.linear_fwd(_ptr(self.haS), P1, 0, Wp("dec.proj.weight"), 0, Wp("dec.proj.bias"), 0, _ptr(self.dl), FEAT, 0,
                     B, FEAT, P1, 0, 1, 0, st)
        if self.precision == "tf32":
            R = 1 | 2
            K.pad_copy(_ptr(self.dl), _ptr(self.xin1), B, 21, 21, 32, 23, 23, 1, 0, 3, st)
            K.conv_tcg(_ptr(self.xin1), _ptr(self.dwf), Wp("dec.conv1.bias"), 0, _ptr(self.xin2), B, 23, 23, 32, 128, 21, 21, -1,
                       23, 23, 1, 0, 0, 0, R, st)
            K.conv_tcg(_ptr(self.xin2), _ptr(self.w2f), _ptr(self.b2p), 0, _ptr(self.xin3), B, 23, 23, 128, 256, 21, 21, -1,
                       44, 44, 1, 0, 0, 0, R | (1 << 5), st)
            K.conv_tcg(_ptr(self.xin3), _ptr(self.w3f), _ptr(self.b3p), 0, _ptr(self.lgp), B, 44, 44, 64, 64, 42, 42, -1,
                       44, 44, 1, 0, 0, 0, 0, st)
            # phase layout [B][44][44][2a+b][16] (image at rows [1,43), cols [0,42)) -> NCHW pixels (2y+a, 2x+b)
            lg = self.lgp.reshape(B, 44, 44, 2, 2, 16)[:, 1:43, :42, :, :, :9]
            return lg.permute(0, 5, 1, 3, 2, 4).reshape(B, 9, 84, 84).contiguous()
        K.conv_fwd(_ptr(self.dl), Wp("dec.conv1.weight"), Wp("dec.conv1.bias"), _ptr(self.d1), B, 21, 21, 32, 128, 1, 1, 1, 0, st)
        K.conv_fwd(_ptr(self.d1), Wp("dec.conv2.weight"), Wp("dec.conv2.bias"), _ptr(self.d2), B, 21, 21, 128, 64, 1, 2, 1, 0, st)
        K.conv_fwd(_ptr(self.d2), Wp("dec.conv3.weight"), Wp("dec.conv3.bias"), _ptr(self.lg), B, 42, 42, 64, DEC_C3, 1, 2, 1, 0, st)
        return self.lg[:B * 84 * 84 * DEC_C3].reshape(B, 84, 84, DEC_C3)[..., :9].permute(0, 3, 1, 2).contiguous()

    def _soda_mlp_fwd(self, x, ldx, n, pre, y, a, stats, out, W, big):
        """SODAMLP (modules.py:116-129): Linear -> BatchNorm1d (training statistics) -> ReLU -> Linear."""
        P, st = self.lay.P, self.st
        Kdim = FEAT if big else P
        (self.lin_fwd if big else K.linear_fwd)(x, ldx, 0, W(f"{pre}.0.weight"), 0, W(f"{pre}.0.bias"), 0, y, P, 0, n, P, Kdim, 0, 1,
                                                2 if big else 0, st)
        K.bn_relu_fwd(y, W(f"{pre}.1.weight"), W(f"{pre}.1.bias"), a, stats, n, P, st)
        K.linear_fwd(a, P, 0, W(f"{pre}.3.weight"), 0, W(f"{pre}.3.bias"), 0, out, P, 0, n, P, P, 0, 1, 0, st)

    def _soda_mlp_bwd(self, dout, x, ldx, n, pre, y, a, stats, da, dy, dx, big):
        """Backward of _soda_mlp_fwd: parameter gradients into the arena; dx (n, K) is written (zero-filled here for the split-K path)."""
        P, st, G, Wp = self.lay.P, self.st, self.G, self.P
        Kdim = FEAT if big else P
        K.linear_wgrad(a, P, 0, dout, P, 0, G(f"{pre}.3.weight"), 0, G(f"{pre}.3.bias"), 0, n, P, P, 0, 1, st)
        K.linear_dgrad(dout, P, 0, Wp(f"{pre}.3.weight"), 0, 0, 0, 0, da, P, 0, n, P, P, 0, 0, 1, st)
        K.bn_relu_bwd(da, y, a, Wp(f"{pre}.1.weight"), stats, dy, G(f"{pre}.1.weight"), G(f"{pre}.1.bias"), n, P, st)
        (self.lin_wgrad if big else K.linear_wgrad)(x, ldx, 0, dy, P, 0, G(f"{pre}.0.weight"), 0, G(f"{pre}.0.bias"), 0, n, P, Kdim, 0, 1, st)
        (self.lin_dgrad if big else K.linear_dgrad)(dy, P, 0, Wp(f"{pre}.0.weight"), 0, 0, 0, 0, dx, Kdim, 0, n, P, Kdim, 0, 0, 1, st)

    def update_soda(self):
        """soda.py:41-69: a separately sampled batch of soda_batch_size observations, two random crops, places overlay on one of
        them; predictor(aug_x) against the EMA target encoder of x (BatchNorm1d batch statistics on both), normalised MSE; Adam
        over (SharedCNN, both SODAMLPs); EMA of the target copy.  BatchNorm couples the batch: single-GPU only."""
        L, st, a, ns, P = self.lay, self.st, self.args, self.ns, self.lay.P
        if self.dist is not None:
            raise RuntimeError("SODA's BatchNorm1d statistics are not shardable over the batch")
        src, sup = self.soda_src, self.soda_supplied
        self.soda_supplied = None
        hs = src["Hs"]
        if sup is None:
            K.rng_step(self.seed ^ 0x534F4441, _ptr(self.soda_counter), src["n_valid"], _ptr(self.soda_idxs), _ptr(self.soda_ovl),
                       int(self.places_pool.shape[0]), _ptr(self.soda_offs), max(1, hs - 84), 0, 0, 0, ns, self.A, 0, st)
        offs = lambda k: (_ptr(self.soda_offs, k * ns * 2) if hs > 84 else 0)
        K.replay_gather(src["frames"], src["fidx"], _ptr(self.soda_idxs), offs(0), _ptr(self.soda_x), _ptr(self.soda_scratch), ns, hs, 84, 0, 4, st)
        K.replay_gather(src["frames"], src["fidx"], _ptr(self.soda_idxs), offs(1), _ptr(self.soda_aug), _ptr(self.soda_scratch), ns, hs, 84, 0, 4, st)
        al = 0.2                                                   # random_overlay(aug_x): default alpha_blending (augmentations.py:79)
        if sup is not None and sup.get("places") is not None:
            K.overlay_f32(_ptr(self.soda_aug), _ptr(self.soda_places), 0, float(np.float32(1 - al)), float(np.float32(al)), _ptr(self.soda_aug), ns, 84 * 84, st)
        else:
            K.overlay_f32(_ptr(self.soda_aug), _ptr(self.places_pool), _ptr(self.soda_ovl), float(np.float32(1 - al)), float(np.float32(al)),
                          _ptr(self.soda_aug), ns, 84 * 84, st)
        # target branch (no gradient) beside the online one
        ev_t = None
        def target_branch():
            self.enc_fwd(_ptr(self.soda_x), ns, self.actT, target="soda")
            self._soda_mlp_fwd(_ptr(self.actT[10]), FEAT, ns, "soda_proj", _ptr(self.sy1t), _ptr(self.sa1t), _ptr(self.sstat_t), _ptr(self.sh1), self.ST, True)
        if self.overlap:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event(); ev.record(main); self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                target_branch()
                ev_t = torch.cuda.Event(); ev_t.record(self.side)
        self.enc_fwd(_ptr(self.soda_aug), ns, self.actS, 0, col_from=0)
        feat = _ptr(self.actS[10])
        self._soda_mlp_fwd(feat, FEAT, ns, "soda_proj", _ptr(self.sy1), _ptr(self.sa1), _ptr(self.sstat1), _ptr(self.ss), self.P, True)
        self._soda_mlp_fwd(_ptr(self.ss), P, ns, "soda_pred", _ptr(self.sy2), _ptr(self.sa2), _ptr(self.sstat2), _ptr(self.sh0), self.P, False)
        if ev_t is not None:
            torch.cuda.current_stream().wait_event(ev_t)
        else:
            target_branch()
        K.zero(_ptr(self.logs, 4), 4, st)
        K.soda_loss(_ptr(self.sh0), _ptr(self.sh1), _ptr(self.logs, 4), _ptr(self.dsh0), ns, P, ns, st)
        x0, x1 = L.ranges["aux"]
        K.zero(self._g + 4 * x0, 4 * (x1 - x0), st)
        self._soda_mlp_bwd(_ptr(self.dsh0), _ptr(self.ss), P, ns, "soda_pred", _ptr(self.sy2), _ptr(self.sa2), _ptr(self.sstat2),
                           _ptr(self.dsa2), _ptr(self.dsy2), _ptr(self.dss), False)
        dfeat = _ptr(self.dbuf[1])
        self._soda_mlp_bwd(_ptr(self.dss), feat, FEAT, ns, "soda_proj", _ptr(self.sy1), _ptr(self.sa1), _ptr(self.sstat1),
                           _ptr(self.dsa1), _ptr(self.dsy1), dfeat, True)
        self.enc_bwd(dfeat, ns, self.actS, 0, _ptr(self.soda_aug), 1, True)
        self.adam(self.opt_aux, (x0, x1))
        tau = float(a.soda_tau)
        c0s, c1s = L.ranges["cnn"]; s0, s1 = L.ranges["soda"]
        K.ema(self._p + 4 * c0s, _ptr(self.soda_target), c1s - c0s, 0, tau, tau, st)       # soft_update_params(predictor, predictor_target, soda_tau)
        K.ema(self._p + 4 * s0, _ptr(self.soda_target, c1s - c0s), s1 - s0, 0, tau, tau, st)
        self.prep_conv_weights()
        self.prep_soda_target_weights()

    def update_pad(self):
        """pad.py:39-49 (InverseDynamics.forward, modules.py:298-303): h = encoder(obs), h' = encoder(next_obs) through the shared
        CNN and PAD's own projection, MLP(cat[h, h']) -> predicted action, MSE against the taken action; Adam over (SharedCNN,
        PAD projection, MLP).  The shared CNN runs once over [next_obs ; obs] (one 2B batch)."""
        B, A, L, H, st, a = self.B, self.A, self.lay, self.H, self.st, self.args
        P = L.P
        Wp, G = self.P, self.G
        feat = self.actS[10]
        self.enc_fwd(_ptr(self.obs3), 2 * B, self.actS, 0, col_from=0)
        # encoder rows [0,B) = next_obs -> joint[:, P:2P], rows [B,2B) = obs -> joint[:, :P]
        self.proj_fwd(_ptr(feat), B, "pad_proj", _ptr(self.z_p), _ptr(self.joint, P), 2 * P)
        self.proj_fwd(_ptr(feat, B * FEAT), B, "pad_proj", _ptr(self.z_p, B * P), _ptr(self.joint), 2 * P)
        K.linear_fwd(_ptr(self.joint), 2 * P, 0, Wp("pad_mlp.0.weight"), 0, Wp("pad_mlp.0.bias"), 0, _ptr(self.az1), H, 0, B, H, 2 * P, 0, 1, 0, st)
        (self.lin_fwd if B >= 32 else K.linear_fwd)(_ptr(self.az1), H, 0, Wp("pad_mlp.2.weight"), 0, Wp("pad_mlp.2.bias"), 0,
                                                    _ptr(self.az2), H, 0, B, H, H, 1, 1, 2, st)
        K.linear_fwd(_ptr(self.az2), H, 0, Wp("pad_mlp.4.weight"), 0, Wp("pad_mlp.4.bias"), 0, _ptr(self.pad_pred), A, 0, B, A, H, 1, 1, 2, st)
        K.zero(_ptr(self.logs, 4), 4, st)
        K.mse_loss(_ptr(self.pad_pred), _ptr(self.action), _ptr(self.logs, 4), _ptr(self.dpad_pred), B, A, self.Bg, st)
        x0, x1 = L.ranges["aux"]
        K.zero(self._g + 4 * x0, 4 * (x1 - x0), st)
        K.linear_dgrad(_ptr(self.dpad_pred), A, 0, Wp("pad_mlp.4.weight"), 0, _ptr(self.az2), H, 0, _ptr(self.daz2), H, 0, B, A, H, 1, 0, 1, st)
        self.lin_dgrad(_ptr(self.daz2), H, 0, Wp("pad_mlp.2.weight"), 0, _ptr(self.az1), H, 0, _ptr(self.daz1), H, 0, B, H, H, 1, 2, 1, st)
        K.linear_dgrad(_ptr(self.daz1), H, 0, Wp("pad_mlp.0.weight"), 0, 0, 0, 0, _ptr(self.djoint), 2 * P, 0, B, H, 2 * P, 0, 2, 1, st)
        K.linear_wgrad(_ptr(self.az2), H, 0, _ptr(self.dpad_pred), A, 0, G("pad_mlp.4.weight"), 0, G("pad_mlp.4.bias"), 0, B, A, H, 1, 1, st)
        self.lin_wgrad(_ptr(self.az1), H, 0, _ptr(self.daz2), H, 0, G("pad_mlp.2.weight"), 0, G("pad_mlp.2.bias"), 0, B, H, H, 1, 1, st)
        K.linear_wgrad(_ptr(self.joint), 2 * P, 0, _ptr(self.daz1), H, 0, G("pad_mlp.0.weight"), 0, G("pad_mlp.0.bias"), 0, B, H, 2 * P, 0, 1, st)
        dfeat = self.dbuf[1]
        self.proj_bwd(_ptr(self.djoint, P), 2 * P, B, _ptr(self.z_p), _ptr(self.joint, P), 2 * P, "pad_proj", _ptr(self.dz_p),
                      feat_ptr=_ptr(feat), dfeat=_ptr(dfeat))
        self.proj_bwd(_ptr(self.djoint), 2 * P, B, _ptr(self.z_p, B * P), _ptr(self.joint), 2 * P, "pad_proj", _ptr(self.dz_p, B * P),
                      feat_ptr=_ptr(feat, B * FEAT), dfeat=_ptr(dfeat, B * FEAT))
        self.enc_bwd(_ptr(dfeat), 2 * B, self.actS, 0, _ptr(self.obs3), 1, True)
        if self.dist is not None:
            self.allreduce_grads((x0, x1))
        self.adam(self.opt_aux, (x0, x1))
        self.prep_conv_weights()

    def update_curl(self):
        """curl.py:27-43 (CURLHead.compute_logits, modules.py:270-281): z_a = critic encoder(obs) with gradient, z_pos = target
        encoder(pos) without; logits = z_a W z_pos^T, cross entropy against the diagonal; Adam over (SharedCNN, critic
        projection, W).  The contrastive loss couples every sample of the batch with every other: single-GPU only."""
        B, A, L, st, a = self.B, self.A, self.lay, self.st, self.args
        P, P1 = L.P, L.P + A
        if self.dist is not None:
            raise RuntimeError("CURL's (B,B) contrastive logits are not shardable over the batch")
        # z_pos: target encoder + target projection of the second crop (no gradient)
        ev_t = None
        if self.overlap:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event(); ev.record(main); self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                self.enc_fwd(_ptr(self.pos), B, self.actT, target=True)
                self.proj_fwd(_ptr(self.actT[10]), B, "critic_proj", _ptr(self.zT), _ptr(self.haT), P1, target=True)
                ev_t = torch.cuda.Event(); ev_t.record(self.side)
        # z_a: the critic's encoder (updated weights) on obs, activations kept for the backward pass
        self.enc_fwd(_ptr(self.obs2), B, self.actS, B, col_from=0)
        self.proj_fwd(_ptr(self.actS[10], B * FEAT), B, "critic_proj", _ptr(self.zS), _ptr(self.haS), P1)
        if ev_t is not None:
            torch.cuda.current_stream().wait_event(ev_t)
        else:
            self.enc_fwd(_ptr(self.pos), B, self.actT, target=True)
            self.proj_fwd(_ptr(self.actT[10]), B, "critic_proj", _ptr(self.zT), _ptr(self.haT), P1, target=True)
        W, dW = self.P("curl.W"), self.G("curl.W")
        K.linear_fwd(_ptr(self.haT), P1, 0, W, 0, 0, 0, _ptr(self.cu), P, 0, B, P, P, 0, 1, 0, st)             # u = z_pos W^T  (B,P)
        K.linear_fwd(_ptr(self.haS), P1, 0, _ptr(self.cu), 0, 0, 0, _ptr(self.clog), B, 0, B, B, P, 0, 1, 0, st)  # logits = z_a u^T  (B,B)
        K.zero(_ptr(self.logs, 4), 4, st)
        K.ce_diag(_ptr(self.clog), B, _ptr(self.logs, 4), _ptr(self.dclog), B, B, self.Bg, st)
        x0, x1 = L.ranges["aux"]
        K.zero(self._g + 4 * x0, 4 * (x1 - x0), st)
        K.zero(_ptr(self.dcu), 4 * B * P, st)
        K.linear_dgrad(_ptr(self.dclog), B, 0, _ptr(self.cu), 0, 0, 0, 0, _ptr(self.dh_c), P, 0, B, B, P, 0, 0, 1, st)    # d z_a = dlogits u
        K.linear_wgrad(_ptr(self.haS), P1, 0, _ptr(self.dclog), B, 0, _ptr(self.dcu), 0, 0, 0, B, B, P, 0, 1, st)         # d u = dlogits^T z_a
        K.linear_wgrad(_ptr(self.haT), P1, 0, _ptr(self.dcu), P, 0, dW, 0, 0, 0, B, P, P, 0, 1, st)                       # d W = d u^T z_pos
        dfeat = _ptr(self.dbuf[1])
        self.proj_bwd(_ptr(self.dh_c), P, B, _ptr(self.zS), _ptr(self.haS), P1, "critic_proj", _ptr(self.dzS),
                      feat_ptr=_ptr(self.actS[10], B * FEAT), dfeat=dfeat)
        self.enc_bwd(dfeat, B, self.actS, B, _ptr(self.obs2), 1, True)
        self.adam(self.opt_aux, (x0, x1))
        self.prep_conv_weights()

    def _decoder_simt(self, B, st, Wp, G, x0, x1):
        """AttributionDecoder convs + BCE + their backward on the fp32 CUDA-core kernels (compact NHWC buffers)."""
        K.conv_fwd(_ptr(self.dl), Wp("dec.conv1.weight"), Wp("dec.conv1.bias"), _ptr(self.d1), B, 21, 21, 32, 128, 1, 1, 1, 0, st)
        K.conv_fwd(_ptr(self.d1), Wp("dec.conv2.weight"), Wp("dec.conv2.bias"), _ptr(self.d2), B, 21, 21, 128, 64, 1, 2, 1, 0, st)
        K.conv_fwd(_ptr(self.d2), Wp("dec.conv3.weight"), Wp("dec.conv3.bias"), _ptr(self.lg), B, 42, 42, 64, DEC_C3, 1, 2, 1, 0, st)
        K.zero(_ptr(self.logs, 4), 4, st)
        K.bce(_ptr(self.lg), _ptr(self.mask), _ptr(self.logs, 4), _ptr(self.dlg), B, 84, 84, 84, 84, 0, 0, DEC_C3, self.Bg, 0, st)
        K.zero(self._g + 4 * x0, 4 * (x1 - x0), st)
        K.conv_wgrad(_ptr(self.d2), _ptr(self.dlg), G("dec.conv3.weight"), G("dec.conv3.bias"), B, 42, 42, 64, DEC_C3, 1, 2, 1, 0, st)
        K.conv_dgrad(_ptr(self.dlg), Wp("dec.conv3.weight"), 0, _ptr(self.dup3), B, 84, 84, 64, DEC_C3, 1, 0, st)
        K.upsample2_bwd(_ptr(self.dup3), _ptr(self.d2), _ptr(self.dd2), B, 42, 42, 64, st)
        K.conv_wgrad(_ptr(self.d1), _ptr(self.dd2), G("dec.conv2.weight"), G("dec.conv2.bias"), B, 21, 21, 128, 64, 1, 2, 1, 0, st)
        K.conv_dgrad(_ptr(self.dd2), Wp("dec.conv2.weight"), 0, _ptr(self.dup2), B, 42, 42, 128, 64, 1, 0, st)
        K.upsample2_bwd(_ptr(self.dup2), _ptr(self.d1), _ptr(self.dd1), B, 21, 21, 128, st)
        K.conv_wgrad(_ptr(self.dl), _ptr(self.dd1), G("dec.conv1.weight"), G("dec.conv1.bias"), B, 21, 21, 32, 128, 1, 1, 1, 0, st)
        K.conv_dgrad(_ptr(self.dd1), Wp("dec.conv1.weight"), _ptr(self.dl), _ptr(self.ddl), B, 21, 21, 32, 128, 1, 1, st)

    def _decoder_tc_fwd(self, B, st, Wp):
        """The same on the generalised tcgen05 kernels (conv_tcg.cu).  Every conv reads a zero-bordered pitch-linear
        buffer [B][H+2][W+2][C] (image at rows [1,H+1), cols [0,W)).  conv2 / conv3 -- the convs that follow F.upsample
        (modules.py:327-337) -- run in sub-pixel form at the resolution of their PRE-upsample input: 4x the output
        channels (one block per output phase), ReLU + TF32 rounding fused into the producer, depth-to-space in conv2's
        epilogue, and the BCE / data gradients work on the phase layout."""
        R = 1 | 2                                             # ReLU, TF32 round
        K.pad_copy(_ptr(self.dl), _ptr(self.xin1), B, 21, 21, 32, 23, 23, 1, 0, 3, st)
        K.conv_tcg(_ptr(self.xin1), _ptr(self.dwf), Wp("dec.conv1.bias"), 0, _ptr(self.xin2), B, 23, 23, 32, 128, 21, 21, -1,
                   23, 23, 1, 0, 0, 0, R, st)                 # relu(conv1) at 21x21
        K.conv_tcg(_ptr(self.xin2), _ptr(self.w2f), _ptr(self.b2p), 0, _ptr(self.xin3), B, 23, 23, 128, 256, 21, 21, -1,
                   44, 44, 1, 0, 0, 0, R | (1 << 5), st)      # relu(conv2(up2(.))) at 42x42 (depth-to-space epilogue)
        K.conv_tcg(_ptr(self.xin3), _ptr(self.w3f), _ptr(self.b3p), 0, _ptr(self.lgp), B, 44, 44, 64, 64, 42, 42, -1,
                   44, 44, 1, 0, 0, 0, 0, st)                 # conv3(up2(.)) logits, kept in phase layout

    def _decoder_tc_bwd(self, B, st, Wp, G, x0, x1):
        """BCE of the phase-layout logits against the attribution mask and the decoder's backward (weight gradients forked)."""
        K.zero(_ptr(self.logs, 4), 4, st)
        K.bce_phase(_ptr(self.lgp), _ptr(self.mask), _ptr(self.logs, 4), _ptr(self.dlgp), B, 84, 84, 44, 44, 1, 0, self.Bg, 1, st)
        K.zero(self._g + 4 * x0, 4 * (x1 - x0), st)
        # conv3 backward at 42x42: phase weight gradient folded back onto the 3x3 taps; the data gradient lands on
        # relu(conv2) directly (sum over phases = the 2x2 sum-pool of the upsample backward) with its ReLU mask fused, and
        # is written in space-to-depth form = the phase layout of conv2's output gradient
        ws = self._fork()                                     # weight / bias gradients beside the data-gradient chain
        K.zero(_ptr(self.dw3p), 4 * self.dw3p.numel(), ws)
        K.zero(_ptr(self.db3p), 4 * 64, ws)
        K.conv_wgrad_tcg(_ptr(self.xin3), _ptr(self.dlgp), _ptr(self.dw3p), B, 44, 44, 64, 64, -1, -1, ws)
        K.colsum(_ptr(self.dlgp), 64, B * 44 * 44, 64, _ptr(self.db3p), ws)
        K.conv_phase_fold(_ptr(self.dw3p), _ptr(self.db3p), G("dec.conv3.weight"), G("dec.conv3.bias"), 64, 9, 16, ws)
        K.conv_tcg(_ptr(self.dlgp), _ptr(self.w3d), 0, _ptr(self.xin3, 44 * 64), _ptr(self.dd2s), B, 44, 44, 64, 64, 42, 42, -1,
                   23, 23, 1, 0, 44, 44, (1 << 2) | 2 | (2 << 5), st)
        # conv2 backward at 21x21 (256 phase channels: the weight gradient runs as two blocks of 128)
        ws = self._fork()
        K.zero(_ptr(self.dw2p), 4 * self.dw2p.numel(), ws)
        K.zero(_ptr(self.db2p), 4 * 256, ws)
        for h in range(2):
            K.conv_wgrad_tcg_ld(_ptr(self.xin2), _ptr(self.dd2s, 128 * h), 256, _ptr(self.dw2p, h * 128 * 9 * 128), B, 23, 23, 128, 128,
                                -1, -1, ws)
        K.colsum(_ptr(self.dd2s), 256, B * 23 * 23, 256, _ptr(self.db2p), ws)
        K.conv_phase_fold(_ptr(self.dw2p), _ptr(self.db2p), G("dec.conv2.weight"), G("dec.conv2.bias"), 128, 64, 64, ws)
        K.conv_tcg(_ptr(self.dd2s), _ptr(self.w2d), 0, _ptr(self.xin2, 23 * 128), _ptr(self.dd1g), B, 23, 23, 256, 128, 21, 21, -1,
                   23, 23, 1, 0, 23, 23, (1 << 2) | 2, st)
        # conv1 backward (ReLU mask of the projection output)
        ws = self._fork()
        K.conv_wgrad_tcg(_ptr(self.xin1), _ptr(self.dd1g), G("dec.conv1.weight"), B, 23, 23, 32, 128, -1, -1, ws)
        K.colsum(_ptr(self.dd1g), 128, B * 23 * 23, 128, G("dec.conv1.bias"), ws)
        K.conv_tcg(_ptr(self.dd1g), _ptr(self.dwd), 0, _ptr(self.dl), _ptr(self.ddl), B, 23, 23, 128, 32, 21, 21, -1, 21, 21, 0, 0,
                   21, 21, 1 << 2, st)
        self._join()

    def update_sgsac(self, step):
        """sgsac.py:169-185 after the sample (obs2[:B], next_obs, action, reward, not_done and the step's randomness
        are already in place)."""
        a = self.args
        do_actor = step % a.actor_update_freq == 0
        do_target = step % a.critic_target_update_freq == 0
        do_aux = step % a.aux_update_freq == 0
        self.update_critic(1 if a.consistency else 0)
        self.critic_step(with_ema=do_target)
        if do_actor or do_aux:
            self.shared_obs_fwd(with_aux=do_aux)
        ev_f = None
        if do_aux:
            if self.overlap and self.precision == "tf32":
                # the predictor's forward (big kernels) beside attribution #2's head chain (a dozen small dependent launches)
                main = torch.cuda.current_stream()
                ev = torch.cuda.Event(); ev.record(main); self.side3.wait_event(ev)
                with torch.cuda.stream(self.side3):
                    self.update_aux("fwd")
                    ev_f = torch.cuda.Event(); ev_f.record(self.side3)
            # attribution #2 with the updated critic feeds only update_aux's mask (sgsac.py:175-176,83); on steps
            # without an aux update the reference computes it and discards it (no side effects).
            self.attribution2(want_mask=True)
        if do_actor and do_aux and self.overlap:
            # the actor / alpha update (~40 small launches on the heads) is independent of the aux update (encoder + decoder
            # heavy): disjoint parameter / gradient ranges and head rows -> run it beside the aux update
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event(); ev.record(main); self.side2.wait_event(ev)
            with torch.cuda.stream(self.side2):         # incl. its gradient exchange ("actor" communicator) and optimiser steps
                self.update_actor_and_alpha(finish=True, fork_wgrad=False)
                ev2 = torch.cuda.Event(); ev2.record(self.side2)
            if ev_f is not None:
                main.wait_event(ev_f)
            self.update_aux("bwd" if ev_f is not None else "all")
            main.wait_event(ev2)
        else:
            if do_actor:
                self.update_actor_and_alpha()
            if do_aux:
                if ev_f is not None:
                    torch.cuda.current_stream().wait_event(ev_f)
                self.update_aux("bwd" if ev_f is not None else "all")
        self._finish_logs()

    def update_sac(self, step, mode=0):
        """sac.py:160-169 (SAC / RAD / DrQ) and svea.py:54-63 (mode 2)."""
        a = self.args
        do_actor = step % a.actor_update_freq == 0
        do_target = step % a.critic_target_update_freq == 0
        self.update_critic(mode)
        self.critic_step(with_ema=do_target)
        if do_actor:
            self.shared_obs_fwd()
            self.update_actor_and_alpha()
        if self.algorithm == "curl" and step % a.aux_update_freq == 0:
            self.update_curl()
        if self.algorithm == "pad" and step % a.aux_update_freq == 0:
            self.update_pad()
        if self.algorithm == "soda" and step % a.aux_update_freq == 0:
            self.update_soda()
        self._finish_logs()

    def _finish_logs(self):
        if self.dist is not None:
            self.dist.all_reduce_logs(self.logs)

    # ------------------------------------------------------------------ acting (sac.py:86-105)
    def act(self, obs_dev, hin, sample, noise=None):
        """obs_dev: (1,9,hin,hin) fp32 device tensor -> (A,) device tensor tanh(mu) or tanh(pi)."""
        L, A, st, a = self.lay, self.A, self.st, self.args
        self.enc_fwd(_ptr(obs_dev), 1, self.actT, hin=hin)
        self.proj_fwd(_ptr(self.actT[10]), 1, "actor_proj", _ptr(self.z_a), _ptr(self.h_a), L.P)
        self.actor_mlp_fwd(1)
        if sample:
            if noise is None:                           # torch.randn_like (modules.py:219) -> the device Philox stream
                noise = self.noise_act
                K.rng_step(self.seed ^ 0x41C64E6D, _ptr(self.act_counter), 0, 0, 0, 1, 0, 1, _ptr(noise), 0, 0, 1, A, 0, st)
            K.actor_head_fwd(_ptr(self.raw), _ptr(noise), float(a.actor_log_std_min), float(a.actor_log_std_max),
                             _ptr(self.mu), _ptr(self.pi), A, 0, 0, 1, A, st)
            return self.pi[0]
        K.actor_head_fwd(_ptr(self.raw), 0, float(a.actor_log_std_min), float(a.actor_log_std_max),
                         _ptr(self.mu), 0, A, 0, 0, 1, A, st)
        return self.mu[0]
