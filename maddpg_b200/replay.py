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
        self._stage_E = 0           # staging buffers of the host path are sized at the first add()
        self._staged = None         # event marking the end of the last async copy out of the pinned buffer

    def _staging(self, E):
        """Page-locked staging buffer + device mirror for E rows: `pack - 1` floats per row (obs | act | rew | next_obs) followed by
        the E done flags as bytes, so that ONE host-to-device copy carries everything the insert kernel reads."""
        if self._stage_E != E:
            dev = self.ring.device
            nf = E * (self._pack - 1)
            nbytes = (4 * nf + E + 15) // 16 * 16
            hb = torch.zeros(nbytes, dtype=torch.uint8)
            if dev.type == "cuda":
                hb = hb.pin_memory()
            db = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self._hb, self._db = hb, db
            self._hf = hb[:4 * nf].view(torch.float32).view(E, self._pack - 1).numpy()
            self._hd = hb[4 * nf:4 * nf + E].numpy()
            self._df = db[:4 * nf].view(torch.float32).view(E, self._pack - 1)
            self._dd = db[4 * nf:4 * nf + E]
            self._stage_E = E
            self._staged = torch.cuda.Event() if dev.type == "cuda" else None
            self._staged_pending = False
        return self._hf, self._hd, self._df, self._dd

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
            r.insert_agent(i, obs_t, action, rew_t, obs_tp1, done_t)  # column views are fine: the kernel takes their strides
            return
        # host path: E transitions (E = 1 for the reference's per-step call) packed into ONE pinned
        # staging buffer -> one H2D copy -> one insert kernel
        obs_h = np.asarray(obs_t, dtype=np.float32)
        E = 1 if obs_h.ndim == 1 else obs_h.shape[0]
        hf, hd, df, dd = self._staging(E)
        if self._staged is not None and self._staged_pending:
            self._staged.synchronize()  # the previous H2D copy must have left the pinned buffer
        hf[:, :D] = obs_h.reshape(E, D)
        hf[:, D:D + K] = np.asarray(action, dtype=np.float32).reshape(E, K)
        hf[:, D + K] = np.asarray(reward, dtype=np.float32).reshape(E)
        hf[:, D + K + 1:2 * D + K + 1] = np.asarray(obs_tp1, dtype=np.float32).reshape(E, D)
        hd[:] = np.asarray(done).reshape(-1) != 0
        self._db.copy_(self._hb, non_blocking=True)
        if self._staged is not None:
            self._staged.record()
            self._staged_pending = True
        r.insert_agent(i, df[:, :D], df[:, D:D + K], df[:, D + K], df[:, D + K + 1:2 * D + K + 1], dd)

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


class DevicePrioritizedReplayMemory(object):
    """Device-resident mirror of ``PrioritizedReplayMemory`` (maddpg/trainer/prioritized_replay_buffer.py:149-201; SumTree :19-146).

    Same surface: ``add(obs_t, action, reward, obs_tp1, done)``, ``sample(n) -> (b_idx, b_memory, ISWeights)``,
    ``batch_update(tree_idx, abs_errors)`` and the class constants.  The rows live in a one-agent ``JointReplayRing`` (the same
    insert / gather kernels as the uniform buffer), the sum tree is the reference's float64 array on the GPU
    (``mdp_sumtree_*``, csrc/mdp_prio.cu).  Given the same uniforms the returned tree indices are bit-exact, quirks included
    (oracle/prioritized.py lists them; tests/golden/prioritized_ref.npz pins them to the real class):

    * ``add`` defers the tree update to the next ``sample`` exactly like ``SumTree.add(update=False)``;
    * ``sample`` raises ``IndexError`` where the reference does (a descent through data slot 0's node) when ``strict``;
    * ``sample`` draws its uniforms from ``np.random`` like the reference (one ``random_sample`` per stratum: the same global
      stream as the reference's ``np.random.uniform(a, b)`` calls) unless ``uniforms`` is passed;
    * ``batch_update`` adds ``epsilon`` IN PLACE to a numpy ``abs_errors`` argument, as the reference does.

    Differences kept visible: ``b_idx`` / ``ISWeights`` are arrays instead of python lists, ``b_memory`` is
    ``[obs (n, D), act (n, K), rew (n,), obs_tp1 (n, D), done (n,)]`` (float32 rows) instead of tuples of the stored python
    objects, and ``add`` also accepts device tensors with a leading env axis (E lockstep transitions = E consecutive adds).
    Priorities: host inputs are converted with numpy's ``power`` (bit-identical to the reference); device inputs use CUDA's
    ``pow`` (<= 2 ulp, so tree values may differ from the reference's in the last bit)."""
    epsilon = 0.01
    alpha = 0.6
    beta = 0.4
    beta_increment_per_sampling = 0.001
    abs_err_upper = 1.0
    max_p = 1e6  # :165

    def __init__(self, capacity, device="cuda", numpy_io=True, strict=True):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.numpy_io = numpy_io
        self.strict = strict
        size, k, scratch = C.c_int64(), C.c_int32(), C.c_int64()
        _lib.check(_lib.lib.mdp_sumtree_layout(self.capacity, C.byref(size), C.byref(k), C.byref(scratch)), "mdp_sumtree_layout")
        self.k, self.tree_size = int(k.value), int(size.value)
        self.parent_nodes = 2 ** self.k - 1
        self.tree = torch.zeros(self.tree_size, dtype=torch.float64, device=self.device)
        self._scratch = torch.zeros(int(scratch.value), dtype=torch.float64, device=self.device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ring = None            # allocated at the first add (row shapes are the caller's)
        self._buf = None
        self.dirty_start, self.dirty_count = 0, 0

    @property
    def data_pointer(self):
        return 0 if self.ring is None else self.ring.next_idx[0]

    @property
    def total_p(self):
        return float(self.tree[0].item())

    def _ensure_ring(self, D, K):
        if self.ring is None:
            self.ring = JointReplayRing([D], [K], capacity=self.capacity, device=self.device)
            self._buf = DeviceReplayBuffer(self.ring, 0, numpy_io=self.numpy_io)

    def add(self, obs_t, action, reward, obs_tp1, done):
        if isinstance(obs_t, torch.Tensor) and obs_t.is_cuda:
            E, D, K = obs_t.shape[0], obs_t.shape[1], action.shape[1]
        else:
            o = np.asarray(obs_t)
            E = 1 if o.ndim <= 1 else o.shape[0]
            D, K = o.size // E, np.asarray(action).size // E
            if o.ndim == 0:
                obs_t, obs_tp1 = o.reshape(1), np.asarray(obs_tp1).reshape(1)
            if np.asarray(action).ndim == 0:
                action = np.asarray(action).reshape(1)
        if E > self.capacity:  # more rows than slots: consecutive adds, the later rows overwrite the earlier ones
            sl = lambda x, a, b: x[a:b] if (isinstance(x, (torch.Tensor, np.ndarray)) and x.ndim > 0) else x
            for lo in range(0, E, self.capacity):
                hi = min(E, lo + self.capacity)
                self.add(sl(obs_t, lo, hi), sl(action, lo, hi), sl(reward, lo, hi), sl(obs_tp1, lo, hi), sl(done, lo, hi))
            return
        self._ensure_ring(D, K)
        if self.dirty_count == 0:
            self.dirty_start = self.ring.next_idx[0]
        self._buf.add(obs_t, action, reward, obs_tp1, done)
        self.dirty_count = min(self.capacity, self.dirty_count + E)
        if self.dirty_count == self.capacity:
            self.dirty_start = 0

    def flush(self):
        """SumTree.update_all for the pending adds (sample() does this itself, like the reference's get_leaf)."""
        if self.dirty_count:
            _lib.check(_lib.lib.mdp_sumtree_flush(_lib.ptr(self.tree), self.capacity, self.dirty_start, self.dirty_count,
                                                  self.max_p, _lib.ptr(self._scratch), _lib.current_stream()), "mdp_sumtree_flush")
            self.dirty_count = 0

    def sample(self, n, uniforms=None):
        n = int(n)
        if uniforms is None:
            uniforms = np.random.random_sample(n)
        if isinstance(uniforms, torch.Tensor):
            u = uniforms.to(device=self.device, dtype=torch.float64)
        else:
            u = torch.as_tensor(np.ascontiguousarray(uniforms, dtype=np.float64)).to(self.device)
        self.beta = float(np.min([1.0, self.beta + self.beta_increment_per_sampling]))
        tidx = torch.empty(n, dtype=torch.int64, device=self.device)
        didx = torch.empty(n, dtype=torch.int64, device=self.device)
        isw = torch.empty(n, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib.mdp_sumtree_sample(_lib.ptr(self.tree), self.capacity, self.dirty_start, self.dirty_count, self.max_p,
                                               n, _lib.ptr(u), self.beta, _lib.ptr(tidx), _lib.ptr(didx), _lib.ptr(isw),
                                               _lib.ptr(self.flag), _lib.ptr(self._scratch), _lib.current_stream()),
                   "mdp_sumtree_sample")
        self.dirty_count = 0
        if self.strict:
            if int(self.flag.item()) & 1:
                raise IndexError("list index out of range")  # prioritized_replay_buffer.py:142 (self.data[data_idx])
        else:
            didx = didx.clamp(max=self.capacity - 1)
        self.last_data_idx = didx
        if self.ring is None:
            raise IndexError("sample from an empty memory")
        b_memory = list(self._buf.sample_index(didx))
        if self.numpy_io:
            return tidx.cpu().numpy(), b_memory, isw.cpu().numpy()
        return tidx, b_memory, isw

    def priorities(self, abs_errors):
        """:197-199 on the host (numpy's power: bit-identical to the reference)."""
        e = np.asarray(abs_errors, dtype=np.float64) + self.epsilon
        return np.power(np.minimum(e, self.abs_err_upper), self.alpha)

    def batch_update(self, tree_idx, abs_errors):
        if isinstance(tree_idx, torch.Tensor):
            ti = tree_idx.to(device=self.device, dtype=torch.int64)
        else:
            ti_h = np.asarray(tree_idx, dtype=np.int64)
            if ti_h.size and (ti_h.max() >= self.tree_size or ti_h.min() < -self.tree_size):
                raise IndexError("list index out of range")
            ti = torch.as_tensor(ti_h).to(self.device)
        B = int(ti.shape[0])
        if isinstance(abs_errors, torch.Tensor) and abs_errors.is_cuda:
            err, prio = abs_errors.to(torch.float64).contiguous(), None
        else:
            ps = self.priorities(abs_errors)
            if isinstance(abs_errors, np.ndarray) and abs_errors.dtype == np.float64:
                abs_errors += self.epsilon  # the reference's in-place side effect (:197)
            err, prio = None, torch.as_tensor(ps).to(self.device)
        _lib.check(_lib.lib.mdp_sumtree_update(_lib.ptr(self.tree), self.capacity, _lib.ptr(ti), B, _lib.ptr(err), _lib.ptr(prio),
                                               self.epsilon, self.abs_err_upper, self.alpha, _lib.ptr(self.flag),
                                               _lib.ptr(self._scratch), _lib.current_stream()), "mdp_sumtree_update")
