"""Data-parallel glue for the batch-sharded configuration (new functionality, the reference is
single-GPU): every rank holds identical replicated parameters / Adam state, processes B/R samples with losses
scaled by 1/B_global, and the flat gradient ranges (critic, actor, aux) are sum-all-reduced over NCCL/NVLink.
`sgsac.py:68-70` uses the min/max of the *whole* batch for the fill scalar, hence the 2-float min/max exchange.
One process per GPU; torch.distributed is the plumbing.

Communicators.  Collectives that the engine issues from DIFFERENT streams (so that they overlap each other and the
backward passes) must not share a communicator -- NCCL orders the operations of one communicator, and two streams give
no order.  There is one process group per issuing stream:
    main   : what the optimiser steps wait for on the update's main stream (the last, small piece of each bucket; the logs)
    early  : the pieces of a gradient bucket that are complete before its encoder backward starts (Q heads, projection,
             decoder), issued from the communication stream while the data-gradient chain runs
    actor  : the actor / alpha gradients, issued from the stream the actor update runs on beside the aux update
    minmax : the obs-batch min / max, issued right after the sample beside the target / critic forward passes
Every rank issues the collectives of one group in the same program order."""
import torch
import torch.distributed as dist


class GradSync(object):
    GROUPS = ("main", "early", "actor", "minmax")

    def __init__(self, group=None, extra_groups=True):
        assert dist.is_initialized()
        self.world = dist.get_world_size(group)
        self.groups = {"main": group}
        ranks = dist.get_process_group_ranks(group) if group is not None else None
        for name in self.GROUPS[1:]:
            self.groups[name] = dist.new_group(ranks=ranks) if extra_groups else group
        self.group = group

    def all_reduce_sum(self, flat, group="main"):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.groups[group])

    def all_reduce_minmax(self, mm, group="minmax"):
        """mm = {min, max, -min, max} as written by sgqn_minmax: ONE max-all-reduce over mm[2:4] makes the second pair
        the global batch's {-min, max} (no extra kernels; the mask kernel reads that form directly)."""
        dist.all_reduce(mm[2:4], op=dist.ReduceOp.MAX, group=self.groups[group])

    def all_reduce_logs(self, logs, group="main"):
        """Loss scalars are local sums / B_global: one sum.  Replicated entries (alpha, col 3) come out multiplied by the
        world size; readers divide (`log_scale`)."""
        dist.all_reduce(logs, op=dist.ReduceOp.SUM, group=self.groups[group])

    def log_scale(self, col):
        return 1.0 / self.world if col == 3 else 1.0


class P2PGradSync(GradSync):
    """The same exchanges through our own kernels over NVLink peer memory (csrc/p2p.cu) instead of NCCL kernels.

    The update's kernels are persistent, one CTA per SM with ~all of its shared memory; an NCCL CTA cannot share an SM with
    them, so a collective that overlaps the backward pass costs whole SMs.  `sgqn_p2p_allreduce_sum` (two-shot, in place,
    rank r reduces slice r in rank order -> every replica holds bit-identical sums) and `sgqn_p2p_small` (min / max, loss
    vector, alpha gradient) use no shared memory and few registers, so their CTAs run BESIDE the resident compute CTAs.

    The gradient arena is carved out of one symmetric allocation (torch.distributed._symmetric_memory: same layout on every
    rank, every peer's copy mapped into this process); `attach()` is collective.  One slot (flags + call counter) per
    issuing stream, like the NCCL communicators above.  Tensors outside the arena fall back to NCCL."""
    SLOTS = {"main": 0, "early": 1, "actor": 2, "minmax": 3, "logs": 4, "alpha": 5, "main_big": 6}
    STAGE_STRIDE = 512 << 10          # bytes per rank of the one-barrier ("push") form: the "main" slot's small ranges

    def __init__(self, group=None, extra_groups=True, ctas=148):
        super().__init__(group, extra_groups)
        self.rank = dist.get_rank(group)
        self.ctas = ctas
        self.arena = None

    def attach(self, n_floats, device):
        """Collective.  Returns a zero-filled flat fp32 tensor of n_floats inside the symmetric arena."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        lay = (C.c_longlong * 3)()
        self._k().p2p_layout(lay)
        stage_off = (sum(lay) + 255) // 256 * 256
        header = stage_off + 2 * 8 * self.STAGE_STRIDE
        need = header // 4 + (n_floats + 3) // 4 * 4
        if self.arena is None or self.arena.numel() < need:
            ok, self.fallback_reason = 1, None
            try:
                arena = symm.empty(need, dtype=torch.float32, device=device)
                hdl = symm.rendezvous(arena, self.group if self.group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                assert len(ptrs) == self.world and ptrs[self.rank] == arena.data_ptr()
            except Exception as e:                     # no peer mapping on this box: every rank falls back to the NCCL communicators
                ok, self.fallback_reason = 0, repr(e)
            flag = torch.tensor([ok], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag) == 0:
                self.arena = None
                return torch.zeros(n_floats, device=device)
            self.arena, self.hdl = arena, hdl
            self.arena.zero_()
            torch.cuda.synchronize(device)
            dist.barrier(self.group)                   # nobody signals into a header that is still being cleared
            self.bases = (C.c_void_p * 8)(*(ptrs + [0] * (8 - self.world)))
            self.flags_off, self.ctl_off, self.small_off, self.data_off = 0, lay[0], lay[0] + lay[1], header
            self.stage_off = stage_off
        g = self.arena[self.data_off // 4: self.data_off // 4 + n_floats]
        g.zero_()
        return g

    @staticmethod
    def _k():
        from ._lib import K                            # (late: dist.py itself stays importable without the extension)
        return K

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _offset(self, t):
        if self.arena is None or t.dtype != torch.float32 or not t.is_contiguous():
            return None
        off = t.data_ptr() - self.arena.data_ptr()
        if off < self.data_off or off + 4 * t.numel() > 4 * self.arena.numel() or (off & 15) or (t.numel() & 3):
            return None
        return off

    def _small(self, t, slot, op):
        self._k().p2p_small(self.bases, self.rank, self.world, self.flags_off, self.ctl_off, self.small_off, self.SLOTS[slot],
                    t.data_ptr(), t.data_ptr(), t.numel(), op, self._stream())

    def all_reduce_sum(self, flat, group="main"):
        off = self._offset(flat)
        if off is not None:
            # the one-barrier push form and the two-shot form count their barrier tickets differently (e vs 2e-1, 2e), so they never
            # share a slot: "main" ranges that fit the staging area push on slot "main", larger ones (PAD's whole aux range) go
            # two-shot on their own slot -- same stream, same order on every rank
            push = group == "main" and 4 * flat.numel() <= self.STAGE_STRIDE
            slot = self.SLOTS["main_big" if (group == "main" and not push) else group]
            self._k().p2p_allreduce_sum(self.bases, self.rank, self.world, self.flags_off, self.ctl_off, slot, off,
                                flat.numel(), self.ctas, self.stage_off if push else -1, self.STAGE_STRIDE,
                                self._stream())
        elif self.arena is not None and flat.dtype == torch.float64 and flat.numel() <= 16 and flat.is_contiguous():
            self._small(flat, "alpha", 2)              # the fp64 alpha gradient (issued from the actor update's stream)
        else:
            super().all_reduce_sum(flat, group)

    def all_reduce_minmax(self, mm, group="minmax"):
        if self.arena is None:
            return super().all_reduce_minmax(mm, group)
        self._small(mm[2:4], "minmax", 1)

    def all_reduce_logs(self, logs, group="main"):
        if self.arena is None or logs.numel() > 32:
            return super().all_reduce_logs(logs, group)
        self._small(logs, "logs", 0)

    def timeouts(self):
        """Barrier time-outs recorded by the kernels (a peer that never arrived); 0 in a healthy run."""
        lay = (self.small_off - self.ctl_off) // 4
        ctl = self.arena[self.ctl_off // 4: self.ctl_off // 4 + lay].view(torch.int32).reshape(-1, 4)
        return int(ctl[:, 2].sum())
