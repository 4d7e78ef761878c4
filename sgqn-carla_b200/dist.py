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
