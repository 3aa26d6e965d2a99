"""Batched MPE environment: the reference-facing mirror of ``multiagent.environment.MultiAgentEnv``.

The reference builds its env at experiments/train.py:48-61 (``scenarios.load(name + ".py").Scenario()``,
``MultiAgentEnv(world, reset_world, reward, observation)``) and drives it with ``env.reset()``
(:104,:128) and ``env.step(action_n)`` (:114).  ``BatchedMultiAgentEnv`` keeps that surface --
``n``, ``observation_space[i].shape``, ``action_space[i]``, ``reset()``, ``step(action_n)`` -- over
``num_envs`` lockstep instances whose physics runs in one fused CUDA kernel (csrc/mdp_env.cu).

* ``num_envs == 1`` (the reference's shape): numpy in, numpy out, shapes ``(D_i,)`` / scalars, so
  experiments/train.py's loop runs unmodified apart from its imports.
* ``num_envs > 1``: every per-agent array gains a leading env axis and is a CUDA ``torch.Tensor``
  (views into the joint device arrays described in include/maddpg_b200.h).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .spaces import Box, Discrete, MultiDiscrete

SCENARIOS = ("simple", "simple_spread", "simple_tag", "simple_world_comm",
             "simple_adversary", "simple_push", "simple_speaker_listener", "simple_crypto", "simple_reference")


def _dims_for(scenario, num_agents=None, state_f64=False):
    cfg = _lib.EnvCfg(_lib.SCENARIO_IDS[scenario], int(num_agents or 0), int(bool(state_f64)))
    h = C.c_void_p()
    _lib.check(_lib.lib.mdp_env_create(C.byref(cfg), C.byref(h)), "mdp_env_create")
    d = _lib.EnvDims()
    _lib.check(_lib.lib.mdp_env_get_dims(h, C.byref(d)), "mdp_env_get_dims")
    return h, d


class EnvBatchFloat(float):
    """A per-agent reward of E lockstep env instances as the reference's loop needs it (train.py:123-125 adds it to python
    floats): the float value is the mean over env instances, ``values`` the per-instance (E,) vector for ``experience``."""

    def __new__(cls, mean, values):
        x = float.__new__(cls, mean)
        x.values = values
        return x


class EnvBatchFlag(object):
    """A per-agent ``done`` of E env instances: truth value "every instance is done" (train.py:116 ``all(done_n)``; MPE
    installs no done callback, so it is identically False), ``values`` the per-instance (E,) vector."""

    def __init__(self, values):
        self.values = values

    def __bool__(self):
        return False


class BatchedMultiAgentEnv(object):
    def __init__(self, scenario, num_envs=1, num_agents=None, device="cuda", state_dtype=torch.float32, seed=0,
                 squeeze=None, benchmark=False):
        if scenario.endswith(".py"):
            scenario = scenario[:-3]
        if scenario not in _lib.SCENARIO_IDS:
            raise NotImplementedError("scenario %r is not implemented (have %s)" % (scenario, SCENARIOS))
        assert state_dtype in (torch.float32, torch.float64)
        self.scenario_name = scenario
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.state_dtype = state_dtype
        self.seed = int(seed)
        self.squeeze = (self.num_envs == 1) if squeeze is None else bool(squeeze)
        self.benchmark = bool(benchmark)  # make_env(..., benchmark=True): step() fills info_n (train.py:56-58)
        self.reference_loop = False  # True: batched step() results are wrapped for the reference's unmodified loop
        self._h, d = _dims_for(scenario, num_agents, state_dtype == torch.float64)
        self.dims = d
        self.n = int(d.n_agents)
        self.obs_dims = [int(d.obs_dim[i]) for i in range(self.n)]
        self.act_dims = [int(d.act_dim[i]) for i in range(self.n)]
        self.obs_off = [int(d.obs_off[i]) for i in range(self.n)]
        self.act_off = [int(d.act_off[i]) for i in range(self.n)]
        self.obs_stride, self.act_stride = int(d.obs_stride), int(d.act_stride)
        self.state_comps = int(d.state_comps)
        self.comm_dim = int(d.comm_dim)
        self.n_goal = int(d.n_goal)
        self.comm_off = [int(d.comm_off[i]) for i in range(self.n)]
        self.comm_len = [int(d.comm_len[i]) for i in range(self.n)]
        self.movable = [bool(d.movable[i]) for i in range(self.n)]
        self.n_landmarks = int(d.n_landmarks)
        self.env_bytes_per_step = int(d.env_bytes_per_step)
        self.observation_space = [Box(-np.inf, np.inf, (D,)) for D in self.obs_dims]
        self.action_space = []
        for i in range(self.n):
            if d.n_heads[i] == 1:
                self.action_space.append(Discrete(d.head_dim[i][0]))
            else:
                self.action_space.append(MultiDiscrete([[0, int(d.head_dim[i][h]) - 1] for h in range(d.n_heads[i])]))
        E = self.num_envs
        dev = self.device
        self.state = torch.zeros((self.state_comps, E), dtype=state_dtype, device=dev)
        self._obs = [torch.zeros((E, self.obs_stride), dtype=torch.float32, device=dev) for _ in range(2)]
        self._cur = 0
        self.act = torch.zeros((E, self.act_stride), dtype=torch.float32, device=dev)
        self.rew = torch.zeros((E, self.n), dtype=torch.float32, device=dev)
        self.done = torch.zeros((E, self.n), dtype=torch.uint8, device=dev)
        self.episode = 0
        # pinned staging for the numpy (reference-shaped) path: one H2D and one D2H copy per step
        self._h_act = torch.zeros((E, self.act_stride), dtype=torch.float32).pin_memory() if dev.type == "cuda" else None
        self._h_out = torch.zeros((E, self.obs_stride + self.n), dtype=torch.float32).pin_memory() if dev.type == "cuda" else None
        self._d_out = torch.zeros((E, self.obs_stride + self.n), dtype=torch.float32, device=dev)

    def __del__(self):
        try:
            _lib.lib.mdp_env_destroy(self._h)
        except Exception:
            pass

    # -- joint device views ------------------------------------------------------------------
    @property
    def obs(self):
        """Current joint observation array (E, obs_stride)."""
        return self._obs[self._cur]

    def _split_obs(self, joint):
        return [joint[:, o:o + D] for o, D in zip(self.obs_off, self.obs_dims)]

    def _obs_out(self):
        if not self.squeeze:
            return self._split_obs(self.obs)
        host = self.obs.cpu().numpy()
        return [host[0, o:o + D].copy() for o, D in zip(self.obs_off, self.obs_dims)]

    # -- state injection (parity tests) ------------------------------------------------------
    def state_from_arrays(self, agent_pos, agent_vel, landmark_pos, agent_c=None, goal=None):
        """(E,A,2), (E,A,2), (E,L,2)[, (E,A,dim_c) state.c of every agent][, (E,n_goal) landmark indices drawn by
        reset_world] -> SoA state tensor [comp][E] (row layout: include/maddpg_b200.h, mdp_env_dims)."""
        A, L, E = self.n, self.n_landmarks, self.num_envs
        s = np.zeros((self.state_comps, E), dtype=np.float64)
        ap, av, lp = (np.asarray(x, np.float64) for x in (agent_pos, agent_vel, landmark_pos))
        for i in range(A):
            s[4 * i + 0], s[4 * i + 1] = ap[:, i, 0], ap[:, i, 1]
            s[4 * i + 2], s[4 * i + 3] = av[:, i, 0], av[:, i, 1]
        if self.comm_dim and agent_c is not None:
            c = np.asarray(agent_c, np.float64)
            for i in range(A):  # only speaking agents have state.c rows
                for k in range(self.comm_len[i]):
                    s[4 * A + self.comm_off[i] + k] = c[:, i, k]
        for l in range(L):
            s[4 * A + self.comm_dim + 2 * l + 0] = lp[:, l, 0]
            s[4 * A + self.comm_dim + 2 * l + 1] = lp[:, l, 1]
        if self.n_goal:
            if goal is None:
                raise ValueError("scenario %s needs the goal landmark indices (E, %d)" % (self.scenario_name, self.n_goal))
            g = np.asarray(goal, np.float64).reshape(E, self.n_goal)
            for k in range(self.n_goal):
                s[4 * A + self.comm_dim + 2 * L + k] = g[:, k]
        return torch.from_numpy(s).to(self.state_dtype).to(self.device)

    def state_to_arrays(self):
        A, L = self.n, self.n_landmarks
        s = self.state.detach().cpu().double().numpy()
        ap = np.stack([np.stack([s[4 * i], s[4 * i + 1]], -1) for i in range(A)], 1)
        av = np.stack([np.stack([s[4 * i + 2], s[4 * i + 3]], -1) for i in range(A)], 1)
        b = 4 * A + self.comm_dim
        lp = np.stack([np.stack([s[b + 2 * l], s[b + 2 * l + 1]], -1) for l in range(L)], 1)
        g = s[b + 2 * L:b + 2 * L + self.n_goal].T.astype(np.int64)
        return dict(agent_pos=ap, agent_vel=av, landmark_pos=lp, comm=s[4 * A:b].T.copy(), goal=g)

    # -- reference surface -------------------------------------------------------------------
    def set_ctl(self, ctl):
        """Attach / detach (None) a device control block (include/maddpg_b200.h: mdp_env_set_ctl)."""
        _lib.check(_lib.lib.mdp_env_set_ctl(self._h, _lib.ptr(ctl)), "mdp_env_set_ctl")

    def force_generic_kernel(self, on=True):
        """Always use the table-driven env-step kernel (include/maddpg_b200.h: mdp_env_force_generic)."""
        _lib.check(_lib.lib.mdp_env_force_generic(self._h, int(bool(on))), "mdp_env_force_generic")

    def reset_device(self, init_state=None, episode=None):
        """Device-only reset (no host copy of the observations); ``episode`` overrides the Philox
        episode id (a graph-relative offset while a control block is attached)."""
        if init_state is not None:
            assert init_state.shape == self.state.shape and init_state.dtype == self.state_dtype
            init_state = init_state.contiguous()
        self._cur ^= 1
        _lib.check(_lib.lib.mdp_env_reset(self._h, self.num_envs, _lib.ptr(self.state), _lib.ptr(init_state),
                                          self.seed, self.episode if episode is None else episode, _lib.ptr(self.obs),
                                          _lib.current_stream()), "mdp_env_reset")
        if episode is None:
            self.episode += 1

    def reset(self, init_state=None):
        """``env.reset()`` (train.py:104,128): new positions (Philox on device) or an injected SoA
        state tensor; returns the list of per-agent observations."""
        self.reset_device(init_state)
        return self._obs_out()

    def step_device(self, act_joint=None, ring=None, cursor=None):
        """One lockstep step on device arrays only.  ``act_joint`` defaults to ``self.act``.  When a
        ``JointReplayRing`` is given, the transition rows are inserted by the same call (``cursor``
        overrides the ring's host cursor: a graph-relative offset while a control block is attached)."""
        act = self.act if act_joint is None else act_joint
        prev = self.obs
        self._cur ^= 1
        if ring is None:
            rc = _lib.lib.mdp_env_step(self._h, self.num_envs, _lib.ptr(self.state), _lib.ptr(act), _lib.ptr(self.obs),
                                       _lib.ptr(self.rew), _lib.ptr(self.done), None, None, 0, 0, 0, _lib.current_stream())
        else:
            if cursor is None:
                cursor = ring.reserve_joint(self.num_envs)
            rc = _lib.lib.mdp_env_step(self._h, self.num_envs, _lib.ptr(self.state), _lib.ptr(act), _lib.ptr(self.obs),
                                       _lib.ptr(self.rew), _lib.ptr(self.done), _lib.ptr(prev), _lib.ptr(ring.ring),
                                       ring.capacity, ring.row_stride, cursor, _lib.current_stream())
        _lib.check(rc, "mdp_env_step")

    def step(self, action_n):
        """``env.step(action_n)`` (train.py:114) -> (obs_n, rew_n, done_n, info_n)."""
        assert len(action_n) == self.n
        if self.squeeze:
            h = self._h_act
            for i, a in enumerate(action_n):
                h[0, self.act_off[i]:self.act_off[i] + self.act_dims[i]] = torch.from_numpy(
                    np.asarray(a, dtype=np.float32).reshape(-1))
            self.act.copy_(h, non_blocking=True)
        elif isinstance(action_n[0], torch.Tensor) and action_n[0].is_cuda:
            base = action_n[0].data_ptr() - 4 * self.act_off[0]
            if all(a.dim() == 2 and a.stride(0) == self.act_stride and a.stride(1) == 1
                   and a.data_ptr() == base + 4 * self.act_off[i] for i, a in enumerate(action_n)):
                # the n arrays are the column blocks of ONE joint (E, act_stride) array (MADDPGCore.act_agent returns such
                # views): step on it directly
                joint = torch.as_strided(action_n[0], (self.num_envs, self.act_stride), (self.act_stride, 1),
                                         action_n[0].storage_offset() - self.act_off[0])
                self.step_device(joint)
                return self._device_step_result()
            for i, a in enumerate(action_n):
                dst = self.act[:, self.act_off[i]:self.act_off[i] + self.act_dims[i]]
                if a.data_ptr() != dst.data_ptr():
                    dst.copy_(a)
        else:  # host arrays with a leading env axis: one pinned staging buffer, one H2D copy
            h = self._h_act
            for i, a in enumerate(action_n):
                h[:, self.act_off[i]:self.act_off[i] + self.act_dims[i]] = torch.as_tensor(np.asarray(a, dtype=np.float32))
            self.act.copy_(h, non_blocking=True)
        self.step_device()
        host_io = self.squeeze or not (isinstance(action_n[0], torch.Tensor) and action_n[0].is_cuda)
        if not host_io:
            return self._device_step_result()
        # one packed D2H copy: [obs | rew]; done is identically False in MPE (no done callback)
        self._d_out[:, :self.obs_stride].copy_(self.obs)
        self._d_out[:, self.obs_stride:].copy_(self.rew)
        self._h_out.copy_(self._d_out, non_blocking=True)
        _lib.synchronize_current_stream()
        host = self._h_out.numpy()
        if self.squeeze:
            obs_n = [host[0, o:o + D].copy() for o, D in zip(self.obs_off, self.obs_dims)]
            rew_n = [float(host[0, self.obs_stride + i]) for i in range(self.n)]
            done_n = [False] * self.n
        else:
            obs_n = [host[:, o:o + D].copy() for o, D in zip(self.obs_off, self.obs_dims)]
            rew_n = [host[:, self.obs_stride + i].copy() for i in range(self.n)]
            done_n = [np.zeros(self.num_envs, dtype=bool) for _ in range(self.n)]
        return obs_n, rew_n, done_n, self._info_n()

    def _device_step_result(self):
        obs_n = self._split_obs(self.obs)
        rew_n, done_n = [self.rew[:, i] for i in range(self.n)], [self.done[:, i] for i in range(self.n)]
        if self.reference_loop:  # scalars for the loop's bookkeeping (one (n,) device-to-host copy), vectors for experience()
            m = self.rew.mean(0).cpu().tolist()
            rew_n = [EnvBatchFloat(m[i], rew_n[i]) for i in range(self.n)]
            done_n = [EnvBatchFlag(d) for d in done_n]
        return obs_n, rew_n, done_n, self._info_n()

    def benchmark_data(self):
        """``scenario.benchmark_data(agent, world)`` for every (env instance, agent): device (E, n, 4) float32
        (include/maddpg_b200.h: mdp_env_benchmark)."""
        out = torch.empty((self.num_envs, self.n, 4), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.mdp_env_benchmark(self._h, self.num_envs, _lib.ptr(self.state), _lib.ptr(out),
                                              _lib.current_stream()), "mdp_env_benchmark")
        return out

    def _info_n(self):
        """info_n of ``MultiAgentEnv.step`` in the reference's shapes: {'n': [benchmark_data(agent_i)]}; a tuple of four per
        agent for simple_spread, a collision count for simple_tag / simple_world_comm (arrays over env instances
        unless squeezed)."""
        if not self.benchmark or self.scenario_name == "simple":
            return {"n": [{} for _ in range(self.n)]}
        b = self.benchmark_data().cpu().numpy()
        if self.scenario_name == "simple_spread":
            if self.squeeze:
                return {"n": [(float(b[0, i, 0]), int(b[0, i, 1]), float(b[0, i, 2]), int(b[0, i, 3])) for i in range(self.n)]}
            return {"n": [(b[:, i, 0].copy(), b[:, i, 1].astype(np.int64), b[:, i, 2].copy(), b[:, i, 3].astype(np.int64))
                          for i in range(self.n)]}
        if self.squeeze:
            return {"n": [int(b[0, i, 0]) for i in range(self.n)]}
        return {"n": [b[:, i, 0].astype(np.int64) for i in range(self.n)]}

    def render(self, mode="human"):
        raise NotImplementedError("rendering is out of scope for the batched device environment")


def make_env(scenario_name, arglist=None, benchmark=False, **kw):
    """experiments/train.py:48-61 equivalent."""
    if arglist is not None:
        kw.setdefault("num_envs", getattr(arglist, "num_envs", 1))
        kw.setdefault("seed", getattr(arglist, "seed", 0))
    return BatchedMultiAgentEnv(scenario_name, benchmark=benchmark, **kw)
