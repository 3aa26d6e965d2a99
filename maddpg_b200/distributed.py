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
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of ``total`` env instances / replay rows owned by ``rank``."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed, rank):
    return int(seed) + int(rank)


class PeerGradExchange(object):
    """Fused gradient exchange over NVLink peer memory (include/maddpg_b200.h: mdp_core_bind_peers).

    The core's gradient bucket is re-homed in a symmetric-memory allocation that every rank of the node maps
    (torch.distributed._symmetric_memory); the clip+Adam+polyak kernel then sums the bucket over ranks itself with
    peer loads and flag barriers, so an update round has NO separate collective launch and no host synchronisation --
    it can be captured in a CUDA graph.  Replicas stay bit-identical (fixed summation order)."""

    SLOT_WORDS = 8

    LL_MAX_FLOATS = 1 << 21  # low-latency push mode up to 2 M gradient floats (receive buffers: 16 B x world per float)

    def __init__(self, core, group=None, low_latency=None):
        import torch.distributed._symmetric_memory as symm
        self.core = core
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = core.device
        n = int(core.layout.total_train)
        self.grads = symm.empty(n, dtype=torch.float32, device=dev)
        self.flags = symm.empty(12 * core.n * self.SLOT_WORDS, dtype=torch.int32, device=dev)  # slot = 12 * agent + 6 * net + variable
        self.grads.zero_()
        self.flags.zero_()
        self.gh = symm.rendezvous(self.grads, group)
        self.fh = symm.rendezvous(self.flags, group)
        self.epoch = torch.zeros(12 * core.n, dtype=torch.int32, device=dev)
        self.low_latency = (n <= self.LL_MAX_FLOATS) if low_latency is None else bool(low_latency)
        self.recv = self.rh = None
        if self.low_latency:
            self.recv = symm.empty(2 * self.world * n * 2, dtype=torch.int32, device=dev)
            self.recv.zero_()
            self.rh = symm.rendezvous(self.recv, group)
        # re-home the gradient bucket: same layout, symmetric allocation
        core.grads = self.grads
        _lib.check(_lib.lib.mdp_core_bind(core._h, _lib.ptr(core.params), _lib.ptr(core.grads), _lib.ptr(core.adam_m),
                                          _lib.ptr(core.adam_v), _lib.ptr(core.adam_t), _lib.ptr(core.stats)), "mdp_core_bind")
        gp = (C.c_void_p * self.world)(*[int(p) for p in self.gh.buffer_ptrs])
        fp = (C.c_void_p * self.world)(*[int(p) for p in self.fh.buffer_ptrs])
        assert int(self.gh.buffer_ptrs[self.rank]) == self.grads.data_ptr()
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's buckets and flags are zero before anyone signals
        rp = (C.c_void_p * self.world)(*[int(p) for p in self.rh.buffer_ptrs]) if self.low_latency else None
        self._tables = (gp, fp, rp)
        self.open()

    def open(self):
        """(Re)binds the peer tables: from here on the optimizer kernels sum the gradient bucket over the ranks.  CUDA graphs
        captured before a close()/open() pair hold stale table pointers and must be re-captured."""
        gp, fp, rp = self._tables
        _lib.check(_lib.lib.mdp_core_bind_peers(self.core._h, self.world, self.rank, gp, fp, _lib.ptr(self.epoch), rp),
                   "mdp_core_bind_peers")
        self.core.peer_world = self.world

    def close(self):
        """Unbinds the peers: every rank steps on its own gradients (lock-step across ranks is the caller's business: all
        ranks must close and re-open at the same point of their update sequence)."""
        self.core.peer_world = 1
        _lib.check(_lib.lib.mdp_core_bind_peers(self.core._h, 1, 0, None, None, None, None), "mdp_core_bind_peers")


class DataParallelUpdater(object):
    """Drives MADDPGCore's split update entry points with an all-reduce between gradient and step.
    Lock-step contract of the fused peer exchange: every rank issues the SAME sequence of optimizer launches (same agents,
    same nets, same order); a rank that skips one makes its peers wait for the matching epoch and trap after ~4 s.
    ``peer=True`` (NCCL process group on one NVLink node): the all-reduce is fused into the clip+Adam+polyak kernel
    (PeerGradExchange) and ``update_agent`` issues kernels only."""

    def __init__(self, core=None, group=None, grads=None, segment_fn=None, peer=False, low_latency=None):
        self.core = core
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.allreduce_bytes = 0
        self.peer = PeerGradExchange(core, group, low_latency) if (peer and self.world > 1) else None

    def set_exchange(self, on):
        """Fused peer exchange on / off (bench.py measures the same update rounds without it)."""
        if self.peer is not None:
            torch.cuda.synchronize(self.core.device)
            dist.barrier(self.group)
            (self.peer.open if on else self.peer.close)()

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
        if self.peer is not None:  # gradient sum over ranks happens inside the clip+Adam+polyak kernels
            y = c.td_target(j, batch, u_target, idx=idx)
            c.critic_grads(j, batch, y, idx=idx)
            c.clip_adam_polyak(j, 1, grad_scale=scale)
            c.actor_grads(j, batch, u_actor, idx=idx)
            c.clip_adam_polyak(j, 0, grad_scale=scale)
            seg = c.train_segment(c.grads, j, 1).numel() + c.train_segment(c.grads, j, 0).numel()
            self.allreduce_bytes += 4 * seg
            return
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
