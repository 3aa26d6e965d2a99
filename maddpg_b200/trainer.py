"""MADDPG trainers on the B200 kernels: the reference-facing mirror of ``MADDPGAgentTrainer``.

Reference: maddpg/trainer/maddpg.py:112-196 (``__init__``, ``action`` :151, ``experience`` :154,
``preupdate`` :158, ``update`` :161) and the abstract surface maddpg/__init__.py:1-15.  The n
per-agent trainer objects the reference builds in ``get_trainers`` (experiments/train.py:63-75) are
thin views over one shared ``MADDPGCore``: a flat device parameter buffer (four MLPs per agent:
p, target_p, q, target_q), Adam state, a flat gradient bucket and one joint replay ring.  The core
is created lazily when the first trainer method is used, after all n trainers of the group have
been constructed (``local_q_func`` differs per agent, train.py:70,74).

All arithmetic runs in libmaddpg_b200.so (csrc/mdp_train.cu, csrc/mdp_optim.cu); there is no
PyTorch or CPU fallback for any of it.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .env import EnvBatchFlag, EnvBatchFloat
from .replay import DeviceReplayBuffer, JointReplayRing
from .spaces import act_heads

UPDATE_PERIOD = 100        # maddpg.py:164
REPLAY_CAPACITY = int(1e6)  # maddpg.py:147


class AgentTrainer(object):
    """maddpg/__init__.py:1-15."""

    def __init__(self, name, model, obs_shape, act_space, args):
        raise NotImplementedError()

    def action(self, obs):
        raise NotImplementedError()

    def process_experience(self, obs, act, rew, new_obs, done, terminal):
        raise NotImplementedError()

    def preupdate(self):
        raise NotImplementedError()

    def update(self, agents):
        raise NotImplementedError()


class MADDPGCore(object):
    """Device state shared by the n trainers of one experiment."""

    def __init__(self, obs_dims, act_spaces, local_q, num_units=64, lr=1e-2, gamma=0.95, device="cuda", seed=0,
                 replay_capacity=REPLAY_CAPACITY, polyak=1.0 - 1e-2, grad_clip=0.5, actor_reg=1e-3,
                 beta1=0.9, beta2=0.999, adam_eps=1e-8, gather_mode=0):
        self.n = len(obs_dims)
        self.device = torch.device(device)
        self.obs_dims = [int(d) for d in obs_dims]
        self.heads = [act_heads(s) for s in act_spaces]
        self.act_dims = [sum(h) for h in self.heads]
        self.local_q = [bool(x) for x in local_q]
        self.num_units = int(num_units)
        self.seed = int(seed)
        cfg = _lib.CoreCfg()
        cfg.n_agents, cfg.num_units = self.n, self.num_units
        for i in range(self.n):
            cfg.obs_dim[i], cfg.act_dim[i] = self.obs_dims[i], self.act_dims[i]
            cfg.n_heads[i] = len(self.heads[i])
            for h, hd in enumerate(self.heads[i]):
                cfg.head_dim[i][h] = hd
            cfg.local_q[i] = int(self.local_q[i])
        cfg.lr, cfg.gamma, cfg.polyak, cfg.grad_clip = lr, gamma, polyak, grad_clip
        cfg.actor_reg, cfg.beta1, cfg.beta2, cfg.adam_eps = actor_reg, beta1, beta2, adam_eps
        self.cfg = cfg
        self._h = C.c_void_p()
        _lib.check(_lib.lib.mdp_core_create(C.byref(cfg), C.byref(self._h)), "mdp_core_create")
        self.layout = _lib.CoreLayout()
        _lib.check(_lib.lib.mdp_core_get_layout(self._h, C.byref(self.layout)), "mdp_core_get_layout")
        dev = self.device
        self.params = torch.zeros(int(self.layout.total_params), dtype=torch.float32, device=dev)
        self.grads = torch.zeros(int(self.layout.total_train), dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros_like(self.grads)
        self.adam_v = torch.zeros_like(self.grads)
        self.adam_t = torch.zeros(2 * self.n, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(8 * self.n, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.mdp_core_bind(self._h, _lib.ptr(self.params), _lib.ptr(self.grads), _lib.ptr(self.adam_m),
                                          _lib.ptr(self.adam_v), _lib.ptr(self.adam_t), _lib.ptr(self.stats)),
                   "mdp_core_bind")
        self.obs_off = np.concatenate([[0], np.cumsum(self.obs_dims)[:-1]]).astype(int).tolist()
        self.act_off = np.concatenate([[0], np.cumsum(self.act_dims)[:-1]]).astype(int).tolist()
        self.obs_sum, self.act_sum = sum(self.obs_dims), sum(self.act_dims)
        self.obs_stride = (self.obs_sum + 3) // 4 * 4
        self.act_stride = (self.act_sum + 3) // 4 * 4
        self.ring = JointReplayRing(self.obs_dims, self.act_dims, replay_capacity, dev, gather_mode)
        self.counter = 0  # Philox stream position (advanced once per noise-consuming launch)
        self._y = {}
        self._batch = {}
        self.init_weights(np.random.RandomState(self.seed))

    def __del__(self):
        try:
            _lib.lib.mdp_core_destroy(self._h)
        except Exception:
            pass

    # -- parameters ----------------------------------------------------------------------------
    def net_shapes(self, agent, net):
        i, o, U = int(self.layout.net_in[agent][net]), int(self.layout.net_out[agent][net]), self.num_units
        return [(i, U), (U,), (U, U), (U,), (U, o), (o,)]

    def net_view(self, agent, net):
        """List of six tensors [W1,b1,W2,b2,W3,b3] viewing the flat parameter buffer."""
        off = int(self.layout.net_off[agent][net])
        out = []
        for shp in self.net_shapes(agent, net):
            n = int(np.prod(shp))
            out.append(self.params[off:off + n].view(*shp))
            off += n
        return out

    def train_view(self, buf, agent, which):
        off = int(self.layout.train_off[agent][which])
        out = []
        for shp in self.net_shapes(agent, _lib.NET_P if which == 0 else _lib.NET_Q):
            n = int(np.prod(shp))
            out.append(buf[off:off + n].view(*shp))
            off += n
        return out

    def train_segment(self, buf, agent, which):
        """Contiguous slice of a grads/adam buffer for one net (the allreduce bucket of that net)."""
        off = int(self.layout.train_off[agent][which])
        size = int(self.layout.net_size[agent][_lib.NET_P if which == 0 else _lib.NET_Q])
        return buf[off:off + size]

    def set_weights(self, agent, net, arrays):
        for dst, src in zip(self.net_view(agent, net), arrays):
            dst.copy_(torch.as_tensor(np.asarray(src, dtype=np.float32)).view_as(dst))

    def get_weights(self, agent, net):
        return [t.detach().cpu().numpy().copy() for t in self.net_view(agent, net)]

    def init_weights(self, rng):
        """xavier_initializer (uniform +-sqrt(6/(fan_in+fan_out))) and zero biases for all four nets of
        every agent, each drawn independently (train.py:89: targets are NOT copies of the running nets)."""
        for a in range(self.n):
            for net in (_lib.NET_Q, _lib.NET_TARGET_Q, _lib.NET_P, _lib.NET_TARGET_P):
                arrs = []
                for shp in self.net_shapes(a, net):
                    if len(shp) == 2:
                        lim = math.sqrt(6.0 / (shp[0] + shp[1]))
                        arrs.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
                    else:
                        arrs.append(np.zeros(shp, np.float32))
                self.set_weights(a, net, arrs)

    # -- kernels -------------------------------------------------------------------------------
    def next_counter(self):
        self.counter += 1
        return self.counter

    def set_ctl(self, ctl):
        """Attach / detach (None) a device control block (include/maddpg_b200.h: mdp_core_set_ctl)."""
        _lib.check(_lib.lib.mdp_core_set_ctl(self._h, _lib.ptr(ctl)), "mdp_core_set_ctl")

    def set_tensor_cores(self, mode):
        """0 = automatic, 1 = tcgen05 kernels wherever supported, -1 = fp32 SIMT kernels only."""
        _lib.check(_lib.lib.mdp_core_set_tensor_cores(self._h, int(mode)), "mdp_core_set_tensor_cores")

    def set_fused_update(self, on):
        """TD target + critic step in one launch (default) or two (include/maddpg_b200.h: mdp_core_set_fused_update)."""
        _lib.check(_lib.lib.mdp_core_set_fused_update(self._h, int(bool(on))), "mdp_core_set_fused_update")

    def act(self, obs_joint, act_joint, agent_begin=0, agent_count=None, use_target=False, u=None, logits_out=None,
            counter=None):
        """Grouped actor inference + Gumbel-softmax on joint device arrays (see header).  ``counter``
        overrides the host Philox counter (a graph-relative offset while a control block is attached)."""
        agent_count = self.n - agent_begin if agent_count is None else agent_count
        _lib.check(_lib.lib.mdp_actor_act(self._h, agent_begin, agent_count, int(use_target), obs_joint.shape[0],
                                          _lib.ptr(obs_joint), obs_joint.stride(0), _lib.ptr(act_joint), act_joint.stride(0),
                                          _lib.ptr(u), self.seed, self.next_counter() if counter is None else counter,
                                          _lib.ptr(logits_out), _lib.current_stream()), "mdp_actor_act")

    def act_agent(self, agent, obs, use_target=False, u=None, want_logits=False):
        """One agent on its own (E, D_i) device array -> (E, K_i).  The kernel addresses agent i's
        columns as base + off_i, so standalone arrays are passed with the base shifted back."""
        E = obs.shape[0]
        if obs.stride(1) != 1:  # rows may be strided (a column view of a joint observation array), elements must be adjacent
            obs = obs.contiguous()
        K = self.act_dims[agent]
        sh_o, sh_a = 4 * self.obs_off[agent], 4 * self.act_off[agent]
        if use_target or want_logits or u is not None:
            act = torch.empty((E, K), dtype=torch.float32, device=self.device)
            act_base, act_ld = C.c_void_p(act.data_ptr() - sh_a), K
        else:
            # the rollout call (MADDPGAgentTrainer.action): the result is agent i's column block of a joint (E, act_stride)
            # array, so that env.step can take the n results of a step without copying them together.  Two arrays alternate:
            # a result stays valid until the agent's action() call after the next one.
            act = self._joint_action_view(agent, E)
            act_base, act_ld = C.c_void_p(act.data_ptr() - sh_a), self.act_stride
        logits = torch.empty((E, K), dtype=torch.float32, device=self.device) if want_logits else None
        up = None
        if u is not None:
            u = u if u.is_contiguous() else u.contiguous()
            assert u.shape == (E, K)
            up = C.c_void_p(u.data_ptr() - sh_a)
        _lib.check(_lib.lib.mdp_actor_act(self._h, agent, 1, int(use_target), E, C.c_void_p(obs.data_ptr() - sh_o),
                                          obs.stride(0), act_base, act_ld, up, self.seed,
                                          self.next_counter(),
                                          None if logits is None else C.c_void_p(logits.data_ptr() - sh_a),
                                          _lib.current_stream()), "mdp_actor_act")
        return (act, logits) if want_logits else act

    def _joint_action_view(self, agent, E):
        st = getattr(self, "_jact", None)
        if st is None or st["buf"].shape[1] != E:
            st = self._jact = {"buf": torch.zeros((2, E, self.act_stride), dtype=torch.float32, device=self.device), "slot": 0,
                               "last": self.n}
        if agent <= st["last"]:  # a new round of action() calls (agents are asked in order, train.py:112)
            st["slot"] ^= 1
        st["last"] = agent
        o = self.act_off[agent]
        return st["buf"][st["slot"], :, o:o + self.act_dims[agent]]

    def critic_q(self, agent, x, use_target=False):
        B = x.shape[0]
        q = torch.empty(B, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.mdp_critic_q(self._h, agent, int(use_target), B, _lib.ptr(x), x.stride(0), _lib.ptr(q),
                                         _lib.current_stream()), "mdp_critic_q")
        return q

    def _scratch(self, B):
        if B not in self._y:
            self._y[B] = torch.empty(B, dtype=torch.float32, device=self.device)
            self._batch[B] = torch.empty((B, self.ring.row_stride), dtype=torch.float32, device=self.device)
        return self._y[B], self._batch[B]

    def td_target(self, agent, batch, u_target=None, want_target_act=False, idx=None):
        """``idx`` (int64 CUDA tensor): rows are ``batch[idx[b]]`` (pass the ring: the gather is fused in)."""
        B = batch.shape[0] if idx is None else idx.shape[0]
        y, _ = self._scratch(B)
        ta = torch.zeros((B, self.act_stride), dtype=torch.float32, device=self.device) if want_target_act else None
        _lib.check(_lib.lib.mdp_td_target(self._h, agent, C.byref(self.ring.layout), B, _lib.ptr(batch), _lib.ptr(idx),
                                          _lib.ptr(u_target), self.act_stride, self.seed, self.next_counter(), _lib.ptr(y), _lib.ptr(ta),
                                          _lib.current_stream()), "mdp_td_target")
        return (y, ta) if want_target_act else y

    def td_target_all(self, batch, idx=None, counter=None):
        """TD targets of every agent in one grouped launch -> (n_agents, B) (include/maddpg_b200.h: mdp_td_target_all)."""
        if idx is None:
            B, stride = batch.shape[0], 0
        else:
            B, stride = idx.shape[-1], (idx.stride(0) if idx.dim() == 2 else 0)
        key = ("all", B)
        if key not in self._y:
            self._y[key] = torch.empty((self.n, B), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.mdp_td_target_all(self._h, C.byref(self.ring.layout), B, _lib.ptr(batch), _lib.ptr(idx), stride,
                                              self.seed, self.next_counter() if counter is None else counter,
                                              _lib.ptr(self._y[key]), _lib.current_stream()), "mdp_td_target_all")
        return self._y[key]

    def critic_grads(self, agent, batch, y, want_q=False, idx=None):
        B = batch.shape[0] if idx is None else idx.shape[0]
        q = torch.empty(B, dtype=torch.float32, device=self.device) if want_q else None
        _lib.check(_lib.lib.mdp_critic_grads(self._h, agent, C.byref(self.ring.layout), B, _lib.ptr(batch), _lib.ptr(idx),
                                             _lib.ptr(y), _lib.ptr(q), _lib.current_stream()), "mdp_critic_grads")
        return q

    def actor_grads(self, agent, batch, u_actor=None, idx=None):
        B = batch.shape[0] if idx is None else idx.shape[0]
        _lib.check(_lib.lib.mdp_actor_grads(self._h, agent, C.byref(self.ring.layout), B, _lib.ptr(batch), _lib.ptr(idx),
                                            _lib.ptr(u_actor),
                                            self.act_stride, self.seed, self.next_counter(), _lib.current_stream()),
                   "mdp_actor_grads")

    def clip_adam_polyak(self, agent, which, grad_scale=1.0, do_polyak=True):
        _lib.check(_lib.lib.mdp_clip_adam_polyak(self._h, agent, which, float(grad_scale), int(do_polyak),
                                                 _lib.current_stream()), "mdp_clip_adam_polyak")

    def update_agent(self, agent, batch, u_target=None, u_actor=None, counter=None, idx=None):
        """maddpg.py:181-194 for one agent, five kernels on the current stream.  With ``idx`` the rows are
        ``batch[idx[b]]``: pass ``self.ring.ring`` and the sampled indices and no gather kernel is needed."""
        B = batch.shape[0] if idx is None else idx.shape[0]
        y, _ = self._scratch(B)
        _lib.check(_lib.lib.mdp_update_agent(self._h, agent, C.byref(self.ring.layout), B, _lib.ptr(batch), _lib.ptr(idx),
                                             _lib.ptr(u_target), _lib.ptr(u_actor), self.act_stride, self.seed,
                                             self.next_counter() if counter is None else counter, _lib.ptr(y),
                                             _lib.current_stream()), "mdp_update_agent")

    def update_all(self, batch, idx=None, counter=None, grad_scale=None):
        """Grouped ("Jacobi") round for all agents in five launches (include/maddpg_b200.h: mdp_update_all).
        idx: None, (B,) shared index set or (n_agents, B) per-agent sets (int64 CUDA)."""
        if idx is None:
            B, stride = batch.shape[0], 0
        else:
            B, stride = idx.shape[-1], (idx.stride(0) if idx.dim() == 2 else 0)
        if grad_scale is None:  # fused peer all-reduce bound (PeerGradExchange): average over ranks
            grad_scale = 1.0 / getattr(self, "peer_world", 1)
        key = ("all", B)
        if key not in self._y:
            self._y[key] = torch.empty((self.n, B), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.mdp_update_all(self._h, C.byref(self.ring.layout), B, _lib.ptr(batch), _lib.ptr(idx), stride,
                                           self.seed, self.next_counter() if counter is None else counter,
                                           _lib.ptr(self._y[key]), float(grad_scale), _lib.current_stream()),
                   "mdp_update_all")

    def make_index(self, idx_out, length=None, counter=None, ctl=None, for_update=None):
        """Device-side ``ReplayBuffer.make_index`` (replay_buffer.py:46-47): B uniform draws in [0, len).
        ``for_update=(agent, count)``: the draw of the update of these agents that is launched NEXT -- one kernel also resets
        their statistics accumulators and the update's first kernel starts as its programmatic dependent (mdp_update_prepare);
        ``ctl`` must then be the control block attached to the core (or None)."""
        if for_update is not None:
            agent, count = for_update
            _lib.check(_lib.lib.mdp_update_prepare(self._h, agent, count, _lib.ptr(idx_out), idx_out.shape[0],
                                                   int(self.ring.length[0] if length is None else length), self.seed,
                                                   self.next_counter() if counter is None else counter,
                                                   _lib.current_stream()), "mdp_update_prepare")
            return
        _lib.check(_lib.lib.mdp_replay_make_index(_lib.ptr(idx_out), idx_out.shape[0],
                                                  int(self.ring.length[0] if length is None else length), self.seed,
                                                  self.next_counter() if counter is None else counter, _lib.ptr(ctl),
                                                  _lib.current_stream()), "mdp_replay_make_index")

    def snapshot_stats(self, agent, B=None):
        """The agent's statistics accumulators copied on the DEVICE (no synchronisation): a LazyStats that turns into the six
        numbers of maddpg.py:196 when it is first read."""
        return LazyStats(self, agent, self.stats[8 * agent:8 * agent + 8].clone(), B)

    def read_stats(self, agent, B=None, raw=None):
        """[q_loss, p_loss, mean(target_q), mean(rew), mean(target_q_next), std(target_q)] (maddpg.py:196)."""
        s = (self.stats[8 * agent:8 * agent + 8] if raw is None else raw).cpu().numpy()
        n = s[7] if s[7] > 0 else float(B or 1)
        K = self.act_dims[agent]
        mean_y = s[3] / n
        var_y = max(s[4] / n - mean_y * mean_y, 0.0)
        q_loss = s[0] / n
        p_loss = s[1] / n + float(self.cfg.actor_reg) * s[2] / (n * K)
        return [np.float32(q_loss), np.float32(p_loss), mean_y, s[5] / n, s[6] / n, math.sqrt(var_y)]


class LazyStats(object):
    """What ``MADDPGAgentTrainer.update`` returns: the reference's list of six statistics (maddpg.py:196), materialised
    (one device-to-host copy + synchronisation) only when it is indexed, iterated or converted -- experiments/train.py:161
    discards the value, so the training loop itself never synchronises on it."""

    def __init__(self, core, agent, raw, B):
        self._core, self._agent, self._raw, self._B, self._v = core, agent, raw, B, None

    def _get(self):
        if self._v is None:
            self._v = self._core.read_stats(self._agent, self._B, raw=self._raw)
            self._raw = None
        return self._v

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return 6

    def __getitem__(self, k):
        return self._get()[k]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._get(), dtype=dtype)

    def __repr__(self):
        return repr(self._get())


class _Group(object):
    """All trainers constructed against the same (obs_shape_n, act_space_n)."""
    registry = {}

    def __init__(self, obs_shape_n, act_space_n, args):
        self.obs_dims = [int(s[0]) for s in obs_shape_n]
        self.act_space_n = list(act_space_n)
        self.args = args
        self.members = {}
        self.core = None

    @classmethod
    def get(cls, obs_shape_n, act_space_n, args, agent_index):
        """The group a new trainer joins: same spaces (by structure, not object identity -- callers often rebuild the
        list), same observation dims, same ``args`` object, still collecting members.  A group whose core already exists,
        or that already has this agent index, belongs to an earlier experiment: start a new one."""
        key = (tuple((type(s).__name__,) + tuple(act_heads(s)) for s in act_space_n), tuple(int(s[0]) for s in obs_shape_n))
        g = cls.registry.get(key)
        if g is None or g.args is not args or g.core is not None or agent_index in g.members:
            g = cls(obs_shape_n, act_space_n, args)
            cls.registry[key] = g
        return g

    def finalize(self):
        if self.core is not None:
            return self.core
        n = len(self.obs_dims)
        missing = [i for i in range(n) if i not in self.members]
        if missing:
            raise RuntimeError("MADDPGAgentTrainer: trainers for agents %s have not been constructed yet; the shared "
                               "device core is built once all %d agents exist (train.py:63-75)" % (missing, n))
        a = self.args
        self.core = MADDPGCore(self.obs_dims, self.act_space_n, [self.members[i].local_q_func for i in range(n)],
                               num_units=a.num_units, lr=a.lr, gamma=a.gamma,
                               device=getattr(a, "device", "cuda"), seed=getattr(a, "seed", 0),
                               replay_capacity=int(getattr(a, "replay_capacity", REPLAY_CAPACITY)))
        # a finalized group only lives through its trainers: the registry must not keep the core (and its replay ring,
        # 0.5 GB for simple_spread N=3, 28 GB for N=24 at 1e6 rows) alive after the experiment's trainers are gone
        for k in [k for k, g in _Group.registry.items() if g is self]:
            del _Group.registry[k]
        return self.core


class MADDPGAgentTrainer(AgentTrainer):
    def __init__(self, name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=False):
        self.name = name
        self.n = len(obs_shape_n)
        self.agent_index = agent_index
        self.args = args
        self.model = model  # accepted for signature parity; the MLP (train.py:39-46) lives in the kernels
        self.local_q_func = bool(local_q_func)
        for s in act_space_n:
            act_heads(s)  # NotImplementedError for unsupported spaces, like make_pdtype (distributions.py:422)
        self._group = _Group.get(obs_shape_n, act_space_n, args, agent_index)
        self._group.members[agent_index] = self
        self.max_replay_buffer_len = args.batch_size * args.max_episode_len
        self.replay_sample_index = None
        self._replay = None
        self._idx_dev = None
        self._noise = {}
        self.p_debug = {"p_values": self._p_values, "target_act": self._target_act}
        self.q_debug = {"q_values": self._q_values, "target_q_values": self._target_q_values}

    # -- shared state ----------------------------------------------------------------------------
    @property
    def core(self):
        return self._group.finalize()

    @property
    def replay_buffer(self):
        if self._replay is None:
            self._replay = DeviceReplayBuffer(self.core.ring, self.agent_index, numpy_io=True)
        return self._replay

    def inject_noise(self, u_target=None, u_actor=None):
        """Parity hook: U[0,1) draws replacing the in-kernel Philox streams for the next
        update: u_target (B, sum K) joint layout for all agents' target actors, u_actor (B, K_j)."""
        self._noise = {"u_target": u_target, "u_actor": u_actor}

    # -- helpers -----------------------------------------------------------------------------------
    def _dev(self, x, dtype=torch.float32):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.core.device, dtype=dtype)
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(self.core.device)

    def _out(self, t, like):
        return t if isinstance(like, torch.Tensor) else t.cpu().numpy()

    def _joint_x(self, args):
        core = self.core
        cols = [self._dev(a) for a in args]
        B = cols[0].shape[0]
        x = torch.zeros((B, core.obs_sum + core.act_sum), dtype=torch.float32, device=core.device)
        off = 0
        for c in cols:
            x[:, off:off + c.shape[1]] = c
            off += c.shape[1]
        return x

    def _p_values(self, obs):
        return self._out(self.core.act_agent(self.agent_index, self._dev(obs), want_logits=True)[1], obs)

    def _target_act(self, obs, u=None):
        return self._out(self.core.act_agent(self.agent_index, self._dev(obs), use_target=True,
                                             u=None if u is None else self._dev(u)), obs)

    def _q_values(self, *args):
        return self._out(self.core.critic_q(self.agent_index, self._joint_x(args)), args[0])

    def _target_q_values(self, *args):
        return self._out(self.core.critic_q(self.agent_index, self._joint_x(args), use_target=True), args[0])

    def act(self, obs, u=None):
        """``U.function([obs_ph], act_sample)`` (maddpg.py:62): batched (E, D_i) -> (E, K_i)."""
        return self._out(self.core.act_agent(self.agent_index, self._dev(obs), u=None if u is None else self._dev(u)), obs)

    # -- reference surface ---------------------------------------------------------------------------
    def action(self, obs):
        """maddpg.py:151-152: ``self.act(obs[None])[0]``; a CUDA (E, D_i) tensor is treated as a batch."""
        if isinstance(obs, torch.Tensor) and obs.dim() == 2:
            return self.core.act_agent(self.agent_index, obs)
        obs = np.asarray(obs, dtype=np.float32)
        if obs.ndim == 2:  # host batch (E, D_i) -> (E, K_i)
            return self.act(obs)
        return self.act(obs[None])[0]

    def experience(self, obs, act, rew, new_obs, done, terminal):
        """maddpg.py:154-156 (``terminal`` is ignored by the reference too)."""
        if isinstance(rew, EnvBatchFloat):  # the batched loop's scalar wrappers carry the per-instance vectors
            rew = rew.values
        if isinstance(done, EnvBatchFlag):
            done = done.values
        self.replay_buffer.add(obs, act, rew, new_obs, done if isinstance(done, torch.Tensor) or np.ndim(done) else float(done))

    def process_experience(self, obs, act, rew, new_obs, done, terminal):
        return self.experience(obs, act, rew, new_obs, done, terminal)

    def preupdate(self):
        self.replay_sample_index = None

    def update(self, agents, t, index=None):
        """maddpg.py:161-196.  Returns None while warming up / off-period, else the six statistics."""
        if len(self.replay_buffer) < self.max_replay_buffer_len:
            return
        if not t % UPDATE_PERIOD == 0:
            return
        core = self.core
        if not core.ring.aligned():
            raise NotImplementedError("update() needs index-aligned replay buffers (every agent inserts every step, "
                                      "maddpg.py:173-178 relies on the same)")
        B = self.args.batch_size
        if index is None and not getattr(self.args, "python_index_stream", False):
            # ReplayBuffer.make_index (replay_buffer.py:46-47) on the device: B uniform draws in [0, len) from the core's
            # Philox stream, no host round trip.  args.python_index_stream = True keeps the reference's python `random`
            # stream (bit-identical index lists for a seeded `random`); an explicit ``index`` is always honoured.
            if self._idx_dev is None or self._idx_dev.shape[0] != B:
                self._idx_dev = torch.zeros(B, dtype=torch.int64, device=core.device)
            core.make_index(self._idx_dev, for_update=(self.agent_index, 1))
            self.replay_sample_index = idx = self._idx_dev
        else:
            self.replay_sample_index = self.replay_buffer.make_index(B) if index is None else index
            idx = core.ring.index_tensor(self.replay_sample_index)
        ut, ua = self._noise.get("u_target"), self._noise.get("u_actor")
        self._noise = {}
        if ut is not None:
            ut_j = torch.zeros((idx.shape[0], core.act_stride), dtype=torch.float32, device=core.device)
            ut_j[:, :core.act_sum] = self._dev(ut)
            ut = ut_j
        if ua is not None:
            ua_j = torch.zeros((idx.shape[0], core.act_stride), dtype=torch.float32, device=core.device)
            o = core.act_off[self.agent_index]
            ua_j[:, o:o + core.act_dims[self.agent_index]] = self._dev(ua)
            ua = ua_j
        core.update_agent(self.agent_index, core.ring.ring, ut, ua, idx=idx)  # gather fused into the kernels
        return core.snapshot_stats(self.agent_index, idx.shape[0])
