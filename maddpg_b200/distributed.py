"""Multi-GPU data parallelism for the MADDPG hot path: one process per GPU (torchrun), env instances
and replay shards are rank-local, parameters / Adam state / targets are replicated, and the ONLY
collective is an all-reduce of the gradient bucket of the network being stepped.

The reference has no multi-process code at all (single-threaded session, maddpg/common/
tf_util.py:202-204); this module is the new-work half of BASELINE.json's north_star
("env instances and replay shards partition across the 8 GPUs ... the only collective is an NCCL
allreduce of the critic/actor gradients").

Sequential-agent semantics (maddpg/trainer/maddpg.py:181-194) are kept: the actor gradient of agent
j flows through the critic AFTER its Adam step, so one agent update needs two all-reduces
(critic bucket, then actor bucket); per-variable clip_by_norm acts on the REDUCED gradient
(grad_scale = 1/world_size inside the fused clip+Adam+polyak kernel).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of ``total`` env instances / replay rows owned by ``rank``."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed, rank):
    return int(seed) + int(rank)


class DataParallelUpdater(object):
    """Drives MADDPGCore's split update entry points with an all-reduce between gradient and step."""

    def __init__(self, core=None, group=None, grads=None, segment_fn=None):
        self.core = core
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.allreduce_bytes = 0

    def broadcast_params(self, params, adam_m=None, adam_v=None):
        """Replicas start from rank 0's weights (the reference initialises once, train.py:89)."""
        if self.world > 1:
            for t in (params, adam_m, adam_v):
                if t is not None:
                    dist.broadcast(t, src=0, group=self.group)

    def allreduce_sum(self, segment):
        if self.world > 1:
            dist.all_reduce(segment, op=dist.ReduceOp.SUM, group=self.group)
            self.allreduce_bytes += segment.numel() * segment.element_size()
        return segment

    def update_agent(self, j, batch, u_target=None, u_actor=None, idx=None):
        c = self.core
        scale = 1.0 / self.world
        y = c.td_target(j, batch, u_target, idx=idx)
        c.critic_grads(j, batch, y, idx=idx)
        self.allreduce_sum(c.train_segment(c.grads, j, 1))
        c.clip_adam_polyak(j, 1, grad_scale=scale)
        c.actor_grads(j, batch, u_actor, idx=idx)
        self.allreduce_sum(c.train_segment(c.grads, j, 0))
        c.clip_adam_polyak(j, 0, grad_scale=scale)

    def reduce_stats(self, stats):
        """Sum the per-agent float64 accumulators over ranks (lazy: only when statistics are read)."""
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return stats
