"""Data-parallel glue for the batch-sharded configuration (new functionality, the reference is
single-GPU): every rank holds identical replicated parameters / Adam state, processes B/R samples with losses
scaled by 1/B_global, and the three flat gradient buckets (critic, actor, aux) are sum-all-reduced over
NCCL/NVLink.  `sgsac.py:68-70` uses the min/max of the *whole* batch for the fill scalar, hence the 2-float
min/max exchange.  One process per GPU; torch.distributed is the plumbing."""
import torch
import torch.distributed as dist


class GradSync(object):
    def __init__(self, group=None):
        assert dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group)
        self._mm = None

    def all_reduce_sum(self, flat):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_minmax(self, mm):
        """mm = [min, max] -> global [min, max] with one collective: max over [-min, max]."""
        if self._mm is None:
            self._mm = torch.empty_like(mm)
        self._mm[0] = -mm[0]
        self._mm[1] = mm[1]
        dist.all_reduce(self._mm, op=dist.ReduceOp.MAX, group=self.group)
        mm[0] = -self._mm[0]
        mm[1] = self._mm[1]

    def all_reduce_logs(self, logs):
        """Loss scalars are local sums / B_global (alpha, col 3, is replicated: average it)."""
        dist.all_reduce(logs, op=dist.ReduceOp.SUM, group=self.group)
        logs[3] /= self.world
