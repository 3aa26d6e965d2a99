"""Device-resident replay: the reference-facing mirror of ``maddpg.trainer.replay_buffer.ReplayBuffer``.

Reference: maddpg/trainer/replay_buffer.py -- ``add`` :25-32, ``_encode_sample`` :34-44,
``make_index`` :46-47, ``make_latest_index`` :49-53, ``sample_index`` :55-56, ``sample`` :58-82,
``collect`` :84-85, ``__len__`` :18-19, ``clear`` :21-23.

The reference keeps one python list per agent; all agents insert every step and every update
gathers ALL agents' buffers at one index set (maddpg/trainer/maddpg.py:173-178).  Here one
``JointReplayRing`` holds a joint row per transition on the GPU (layout in include/maddpg_b200.h)
and each agent's ``DeviceReplayBuffer`` is a column view of it with its own cursor, so the
reference's per-agent surface (and its index stream, drawn with python's ``random`` exactly like
the reference) is preserved while a sampled index is one contiguous row copy.
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib


def make_ring_layout(obs_dims, act_dims):
    n = len(obs_dims)
    lay = _lib.RingLayout()
    od = (C.c_int32 * n)(*[int(x) for x in obs_dims])
    ad = (C.c_int32 * n)(*[int(x) for x in act_dims])
    _lib.check(_lib.lib.mdp_ring_make_layout(n, od, ad, C.byref(lay)), "mdp_ring_make_layout")
    return lay


class JointReplayRing(object):
    def __init__(self, obs_dims, act_dims, capacity=int(1e6), device="cuda", gather_mode=0):
        self.n = len(obs_dims)
        self.obs_dims, self.act_dims = [int(x) for x in obs_dims], [int(x) for x in act_dims]
        self.layout = make_ring_layout(obs_dims, act_dims)
        self.row_stride = int(self.layout.row_stride)
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.gather_mode = gather_mode
        # torch.empty: pages are touched as the ring fills (1e6 rows x row_stride floats)
        self.ring = torch.empty((self.capacity, self.row_stride), dtype=torch.float32, device=self.device)
        self.next_idx = [0] * self.n   # per-agent cursor (replay_buffer.py:16,32)
        self.length = [0] * self.n     # per-agent len(_storage)
        self._stage = None

    # column views of a (rows, row_stride) array ------------------------------------------------
    def cols(self, agent):
        L = self.layout
        o, a = int(L.obs_off[agent]), int(L.act_off[agent])
        D, K = self.obs_dims[agent], self.act_dims[agent]
        return dict(obs=(o, o + D), act=(int(L.obs_sum) + a, int(L.obs_sum) + a + K),
                    next_obs=(int(L.nx_off) + o, int(L.nx_off) + o + D),
                    rew=int(L.rw_off) + agent, done=int(L.dn_off) + agent)

    def aligned(self):
        return len(set(self.next_idx)) == 1 and len(set(self.length)) == 1

    def clear(self, agent=None):
        for i in (range(self.n) if agent is None else [agent]):
            self.next_idx[i] = 0
            self.length[i] = 0

    def _advance(self, agent, E):
        cur = self.next_idx[agent]
        self.length[agent] = min(self.capacity, max(self.length[agent], cur + E) if cur + E <= self.capacity else self.capacity)
        self.next_idx[agent] = (cur + E) % self.capacity
        return cur

    def reserve_joint(self, E):
        """Advance every agent's cursor by E rows (they must be aligned); returns the start row."""
        if not self.aligned():
            raise RuntimeError("joint insert needs index-aligned agents (every agent inserts every step)")
        cur = 0
        for i in range(self.n):
            cur = self._advance(i, E)
        return cur

    def advance_all(self, rows):
        """Host mirror of rows inserted on device by a replayed CUDA graph."""
        for i in range(self.n):
            self.length[i] = min(self.capacity, self.length[i] + rows)
            self.next_idx[i] = (self.next_idx[i] + rows) % self.capacity

    def insert_joint(self, obs, act, rew, next_obs, done):
        """E lockstep transitions of all agents from joint device arrays (obs (E,obs_stride), ...)."""
        E = obs.shape[0]
        cur = self.reserve_joint(E)
        _lib.check(_lib.lib.mdp_replay_insert(C.byref(self.layout), _lib.ptr(self.ring), self.capacity, cur, E, -1,
                                              _lib.ptr(obs), obs.stride(0), _lib.ptr(act), act.stride(0),
                                              _lib.ptr(rew), rew.stride(0), _lib.ptr(next_obs), next_obs.stride(0),
                                              _lib.ptr(done), done.stride(0), _lib.current_stream()),
                   "mdp_replay_insert")

    def insert_agent(self, agent, obs, act, rew, next_obs, done):
        """Agent-local device arrays obs (E,D_i), act (E,K_i), rew (E,), next_obs (E,D_i), done (E,) uint8."""
        E = obs.shape[0]
        cur = self._advance(agent, E)
        _lib.check(_lib.lib.mdp_replay_insert(C.byref(self.layout), _lib.ptr(self.ring), self.capacity, cur, E, agent,
                                              _lib.ptr(obs), obs.stride(0), _lib.ptr(act), act.stride(0),
                                              _lib.ptr(rew), rew.stride(0), _lib.ptr(next_obs), next_obs.stride(0),
                                              _lib.ptr(done), done.stride(0), _lib.current_stream()),
                   "mdp_replay_insert")

    def gather(self, idx, out=None, mode=None):
        """out[b, :] = ring[idx[b], :]; idx: int64 CUDA tensor (B,)."""
        B = idx.shape[0]
        if out is None:
            out = torch.empty((B, self.row_stride), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.mdp_replay_gather(_lib.ptr(self.ring), self.capacity, self.row_stride, _lib.ptr(idx), B,
                                              _lib.ptr(out), self.gather_mode if mode is None else mode,
                                              _lib.current_stream()), "mdp_replay_gather")
        return out

    def index_tensor(self, idxes):
        if isinstance(idxes, torch.Tensor):
            return idxes.to(device=self.device, dtype=torch.int64)
        return torch.as_tensor(np.asarray(list(idxes), dtype=np.int64)).to(self.device, non_blocking=True)


class DeviceReplayBuffer(object):
    """One agent's view of the joint ring with the reference ``ReplayBuffer`` surface."""

    def __init__(self, ring, agent, numpy_io=True):
        self.ring = ring
        self.agent = agent
        self.numpy_io = numpy_io
        self._maxsize = ring.capacity
        D, K = ring.obs_dims[agent], ring.act_dims[agent]
        self._pack = 2 * D + K + 2
        self._h = torch.zeros(self._pack, dtype=torch.float32)
        if ring.device.type == "cuda":
            self._h = self._h.pin_memory()
        self._d = torch.zeros(self._pack, dtype=torch.float32, device=ring.device)
        self._d_done = torch.zeros(1, dtype=torch.uint8, device=ring.device)
        self._staged = None  # event marking the end of the last async copy out of the pinned buffer

    def __len__(self):
        return self.ring.length[self.agent]

    @property
    def _next_idx(self):
        return self.ring.next_idx[self.agent]

    def clear(self):
        self.ring.clear(self.agent)

    def add(self, obs_t, action, reward, obs_tp1, done):
        """replay_buffer.py:25-32.  Host scalars/arrays (one transition) or device tensors with a
        leading env axis (E lockstep transitions)."""
        r, i = self.ring, self.agent
        D, K = r.obs_dims[i], r.act_dims[i]
        if isinstance(obs_t, torch.Tensor) and obs_t.is_cuda:
            E = obs_t.shape[0]
            done_t = done if isinstance(done, torch.Tensor) else torch.full((E,), int(bool(done)), dtype=torch.uint8, device=r.device)
            if done_t.dtype != torch.uint8:
                done_t = (done_t != 0).to(torch.uint8)
            rew_t = reward if isinstance(reward, torch.Tensor) else torch.full((E,), float(reward), dtype=torch.float32, device=r.device)
            r.insert_agent(i, obs_t, action, rew_t.contiguous(), obs_tp1, done_t.contiguous())
            return
        # host path: E transitions (E = 1 for the reference's per-step call) packed into ONE pinned
        # staging buffer -> one H2D copy -> one insert kernel
        obs_h = np.asarray(obs_t, dtype=np.float32)
        E = 1 if obs_h.ndim == 1 else obs_h.shape[0]
        pack = self._pack
        if self._h.shape[0] < E * pack:
            self._h = torch.zeros(E * pack, dtype=torch.float32)
            if r.device.type == "cuda":
                self._h = self._h.pin_memory()
            self._d = torch.zeros(E * pack, dtype=torch.float32, device=r.device)
            self._d_done = torch.zeros(E, dtype=torch.uint8, device=r.device)
            self._staged = None
        if self._staged is not None:
            self._staged.synchronize()  # the previous H2D copy must have left the pinned buffer
        h = self._h[:E * pack].view(E, pack).numpy()
        h[:, :D] = obs_h.reshape(E, D)
        h[:, D:D + K] = np.asarray(action, dtype=np.float32).reshape(E, K)
        h[:, D + K] = np.asarray(reward, dtype=np.float32).reshape(E)
        h[:, D + K + 1:2 * D + K + 1] = np.asarray(obs_tp1, dtype=np.float32).reshape(E, D)
        h[:, 2 * D + K + 1] = np.asarray(done, dtype=np.float32).reshape(-1)
        d = self._d[:E * pack].view(E, pack)
        d.copy_(self._h[:E * pack].view(E, pack), non_blocking=True)
        if r.device.type == "cuda":
            self._staged = torch.cuda.Event()
            self._staged.record()
        dn = self._d_done[:E]
        torch.ne(d[:, 2 * D + K + 1], 0, out=self._ne_buf(E))
        dn.copy_(self._ne_buf(E))
        r.insert_agent(i, d[:, :D], d[:, D:D + K], d[:, D + K], d[:, D + K + 1:2 * D + K + 1], dn)

    def _ne_buf(self, E):
        if getattr(self, "_ne", None) is None or self._ne.shape[0] < E:
            self._ne = torch.zeros(E, dtype=torch.bool, device=self.ring.device)
        return self._ne[:E]

    def make_index(self, batch_size):
        # replay_buffer.py:46-47 -- the same python MT19937 stream as the reference
        return [random.randint(0, len(self) - 1) for _ in range(batch_size)]

    def make_latest_index(self, batch_size):
        # replay_buffer.py:49-53
        idx = [(self._next_idx - 1 - i) % self._maxsize for i in range(batch_size)]
        np.random.shuffle(idx)
        return idx

    def sample_index(self, idxes):
        """replay_buffer.py:55-56 -> (obs (B,D), act (B,K), rew (B,), next_obs (B,D), done (B,))."""
        rows = self.ring.gather(self.ring.index_tensor(idxes))
        c = self.ring.cols(self.agent)
        out = (rows[:, c["obs"][0]:c["obs"][1]], rows[:, c["act"][0]:c["act"][1]], rows[:, c["rew"]],
               rows[:, c["next_obs"][0]:c["next_obs"][1]], rows[:, c["done"]])
        if self.numpy_io:
            host = rows.cpu().numpy()
            return (host[:, c["obs"][0]:c["obs"][1]].copy(), host[:, c["act"][0]:c["act"][1]].copy(),
                    host[:, c["rew"]].copy(), host[:, c["next_obs"][0]:c["next_obs"][1]].copy(),
                    host[:, c["done"]].copy())
        return out

    def sample(self, batch_size):
        # replay_buffer.py:58-82
        if batch_size > 0:
            idxes = self.make_index(batch_size)
        else:
            idxes = range(0, len(self))
        return self.sample_index(idxes)

    def collect(self):
        return self.sample(-1)
