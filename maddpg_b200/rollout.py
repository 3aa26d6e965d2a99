"""Device-resident lockstep rollout and update rounds: the batched form of experiments/train.py:110-161
(act -> env.step -> experience -> reset every max_episode_len steps; update rounds) with no host
round trip.

Per lockstep step: one grouped actor+Gumbel kernel (all agents), one fused env-step kernel, one
replay-insert kernel; every ``max_episode_len`` steps a device reset (all env instances share the
episode counter).  With ``use_graph`` a whole episode (and a whole update round) is
captured once into a CUDA graph; the Philox counter, ring cursor, episode id and ring length live
in a 4-word device control block that the graph itself advances (include/maddpg_b200.h), so every
replay draws fresh noise and appends to the ring.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class DeviceCtl(object):
    """{philox_counter, ring_cursor, episode, ring_length} on the device + host mirrors."""

    def __init__(self, env, core):
        self.env, self.core = env, core
        self.t = torch.zeros(4, dtype=torch.int64, device=core.device)
        self._h = torch.zeros(4, dtype=torch.int64)
        self.dirty = True

    def upload(self):
        core, env = self.core, self.env
        self._h[0], self._h[1] = core.counter, core.ring.next_idx[0]
        self._h[2], self._h[3] = (env.episode if env is not None else 0), core.ring.length[0]
        self.t.copy_(self._h)
        self.dirty = False

    def advance(self, d_counter, d_rows, d_episode):
        _lib.check(_lib.lib.mdp_ctl_advance(_lib.ptr(self.t), d_counter, d_rows, self.core.ring.capacity, d_episode,
                                            _lib.current_stream()), "mdp_ctl_advance")


class BatchedRollout(object):
    """mode: "mega" = persistent episode kernel (one launch per episode; falls back to "graph" when the
    scenario does not fit), "graph" = CUDA graph of per-step kernels, "eager" = per-step launches."""

    def __init__(self, env, core, max_episode_len=25, use_graph=True, mode=None):
        assert env.obs_dims == core.obs_dims and env.act_dims == core.act_dims
        self.env, self.core = env, core
        self.max_episode_len = int(max_episode_len)
        self.episode_step = 0
        self.total_steps = 0
        self.mode = mode or ("mega" if use_graph else "eager")
        # float64 state: only the tensor-core episode kernel serves it; run_mega() falls back to "graph" when it declines
        use_graph = self.mode != "eager"
        self.mega_launches = 0
        self.episodes_per_launch = 1  # > 1: consecutive episodes share one launch (mdp_rollout_episodes)
        self.ep_return = None
        self.use_graph = use_graph
        self._graph = None
        self.graph_ok = False
        self.launches_per_graph = 0  # kernels captured in one episode graph
        self.graph_launches = 0      # kernels executed through graph replays (not seen by mdp_launch_count)
        self.ctl = DeviceCtl(env, core)

    # -- eager path ---------------------------------------------------------------------------------
    def step(self):
        """train.py:112-133 for all env instances."""
        env, core = self.env, self.core
        core.act(env.obs, env.act)
        env.step_device(ring=core.ring)
        self.ctl.dirty = True
        self.episode_step += 1
        self.total_steps += 1
        if self.episode_step >= self.max_episode_len:
            env.reset_device()
            self.episode_step = 0

    # -- graph path ---------------------------------------------------------------------------------
    def _episode_body(self):
        env, core, T, E = self.env, self.core, self.max_episode_len, self.env.num_envs
        for s in range(T):
            core.act(env.obs, env.act, counter=s + 1)
            env.step_device(ring=core.ring, cursor=s * E)
        env.reset_device(episode=0)
        self.ctl.advance(T, T * E, 1)

    def _capture(self):
        env, core = self.env, self.core
        assert self.episode_step == 0
        self.ctl.upload()
        env.set_ctl(self.ctl.t)
        core.set_ctl(self.ctl.t)
        cur0 = env._cur
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        try:
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._episode_body()
            self.launches_per_graph = _lib.launch_count() - l0
            assert env._cur == cur0, "an episode must flip the observation double buffer an even number of times"
            self._graph = g
            self.graph_ok = True
        finally:
            env.set_ctl(None)
            core.set_ctl(None)

    def _mirror_episode(self):
        T, E = self.max_episode_len, self.env.num_envs
        self.core.counter += T
        self.core.ring.advance_all(T * E)
        self.env.episode += 1
        self.total_steps += T

    def run_mega(self, steps, reset_after=True):
        """One launch of the persistent episode kernel (include/maddpg_b200.h: mdp_rollout_episode)."""
        env, core = self.env, self.core
        ring = core.ring
        if not ring.aligned():
            raise RuntimeError("the episode kernel needs index-aligned agents")
        rc = _lib.lib.mdp_rollout_episode(env._h, core._h, env.num_envs, _lib.ptr(env.state), _lib.ptr(env.obs),
                                          _lib.ptr(ring.ring), ring.capacity, ring.row_stride, ring.next_idx[0], steps,
                                          core.seed, core.counter, int(reset_after), env.seed, env.episode,
                                          _lib.ptr(self.ep_return), _lib.current_stream())
        if rc == _lib.MDP_ENOTSUP:
            return False
        _lib.check(rc, "mdp_rollout_episode")
        core.counter += steps
        ring.advance_all(env.num_envs * steps)
        if reset_after:
            env.episode += 1
        self.total_steps += steps
        self.mega_launches += 1
        self.ctl.dirty = True
        return True

    def run_mega_episodes(self, n):
        """n episodes (each followed by env.reset()) through mdp_rollout_episodes: ONE launch of the tcgen05 episode
        kernel where it applies, else one launch per episode.  At most ``episodes_per_launch`` per call."""
        env, core = self.env, self.core
        ring, T = core.ring, self.max_episode_len
        if not ring.aligned():
            raise RuntimeError("the episode kernel needs index-aligned agents")
        rc = _lib.lib.mdp_rollout_episodes(env._h, core._h, env.num_envs, _lib.ptr(env.state), _lib.ptr(env.obs),
                                           _lib.ptr(ring.ring), ring.capacity, ring.row_stride, ring.next_idx[0], T, n,
                                           core.seed, core.counter, env.seed, env.episode, _lib.ptr(self.ep_return),
                                           _lib.current_stream())
        if rc == _lib.MDP_ENOTSUP:
            return False
        _lib.check(rc, "mdp_rollout_episodes")
        core.counter += T * n
        ring.advance_all(env.num_envs * T * n)
        env.episode += n
        self.total_steps += T * n
        self.mega_launches += 1
        self.ctl.dirty = True
        return True

    def run_episodes(self, n):
        if self.mode == "mega" and self.episode_step == 0:
            per = max(1, min(self.episodes_per_launch, self.core.ring.capacity // (self.env.num_envs * self.max_episode_len)))
            k = 0
            while k < n:
                m = min(per, n - k)
                ok = self.run_mega_episodes(m) if m > 1 else self.run_mega(self.max_episode_len, True)
                if not ok:
                    self.mode = "graph"  # scenario does not fit the episode kernel
                    return self.run_episodes(n - k)
                k += m
            return
        if not self.use_graph or self.episode_step != 0:
            return self.run_eager(n * self.max_episode_len)
        if self._graph is None:
            self.run_eager(self.max_episode_len)  # warm-up: lazy allocations happen outside the capture
            n -= 1
            self._capture()
        self.ctl.upload()  # host mirrors are the truth between graph launches (eager calls may have moved them)
        for _ in range(n):
            self._graph.replay()
            self.graph_launches += self.launches_per_graph
            self._mirror_episode()

    def run_eager(self, steps):
        for _ in range(steps):
            self.step()

    def run(self, steps):
        T = self.max_episode_len
        if self.use_graph and self.episode_step == 0 and steps % T == 0 and (self.mode == "mega" or (T + 1) % 2 == 0):
            return self.run_episodes(steps // T)
        return self.run_eager(steps)

    @property
    def agent_steps_per_step(self):
        return self.env.num_envs * self.env.n


class HostRollout(object):
    """experiments/train.py:110-133 for E lockstep env instances with HOST (numpy) buffers: one C-ABI call
    per step (include/maddpg_b200.h: mdp_host_step) = one H2D copy of the joint observations, three kernels
    (grouped actors + Gumbel sampling, fused env step, replay insert) and one packed D2H copy.

        host = HostRollout(env, core)
        obs_n = host.reset()                                   # env.reset()                       train.py:104
        action_n, new_obs_n, rew_n, done_n = host.step(obs_n)  # action / env.step / experience    :112-120

    The returned arrays are views into page-locked result buffers that alternate between two slots, so the
    arrays of step t stay valid until step t+2 (train.py rebinds ``obs_n = new_obs_n`` every step)."""

    def __init__(self, env, core, experience=True, chunks=None, use_graph=True, copy_kernels=True):
        """chunks: the env instances are processed in this many ranges whose copies and kernels overlap
        (include/maddpg_b200.h: mdp_host_step_pipelined; default 8 from 2048 instances up with copy kernels, else 1).
        use_graph: replay the whole call (copies, kernels, counter advance) as one CUDA graph per result slot.
        copy_kernels: move the two host buffers with copy kernels instead of the copy engines (mdp_host_copy_mode).
        Measured on B200, 4096 instances of simple_spread (tools/e2e_var.py): engines, eager 104 us/step; engines +
        graph 94; copy kernels + graph 84; + 8 chunks 72 (copy-engine transfers do NOT pipeline: 4 chunks 108 us;
        serialising the ranges' uploads so that the first range can start early: 83 us, worse)."""
        assert env.obs_dims == core.obs_dims and env.act_dims == core.act_dims
        assert env.device.type == "cuda"
        self.env, self.core, self.experience = env, core, bool(experience)
        E = env.num_envs
        if chunks is None:
            chunks = 8 if (copy_kernels and E >= 2048 and E % 8 == 0) else 1
        assert 1 <= chunks <= 8 and E % chunks == 0
        self.chunks, self.use_graph, self.copy_kernels = int(chunks), bool(use_graph), bool(copy_kernels)
        _lib.check(_lib.lib.mdp_host_copy_mode(env._h, 1 if copy_kernels else 0), "mdp_host_copy_mode")
        offs = (C.c_int64 * 4)()
        total = C.c_int64()
        _lib.check(_lib.lib.mdp_host_step_layout(env._h, E, offs, C.byref(total)), "mdp_host_step_layout")
        self.offs, self.total = [int(x) for x in offs], int(total.value)
        self.h_out = [torch.zeros(self.total, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.d_in = torch.zeros((E, env.obs_stride), dtype=torch.float32, device=env.device)
        self.d_out = torch.zeros(self.total, dtype=torch.uint8, device=env.device)
        self._slot = 0
        self._views = [self._make_views(h) for h in self.h_out]
        self.h2d_bytes_per_step = 4 * E * env.obs_stride
        self.d2h_bytes_per_step = self.total
        self.ctl = DeviceCtl(env, core)
        self._graphs = [None, None]   # indexed by the slot the observations are read from
        self._warm = False
        self._expect = None           # (counter, ring cursor) the device control block holds
        self.launches_per_graph = 0
        self.graph_launches = 0

    def _make_views(self, h):
        env, E, o = self.env, self.env.num_envs, self.offs
        a = h.numpy()
        obs = a[o[0]:o[0] + 4 * E * env.obs_stride].view(np.float32).reshape(E, env.obs_stride)
        rew = a[o[1]:o[1] + 4 * E * env.n].view(np.float32).reshape(E, env.n)
        act = a[o[2]:o[2] + 4 * E * env.act_stride].view(np.float32).reshape(E, env.act_stride)
        done = a[o[3]:o[3] + E * env.n].view(np.bool_).reshape(E, env.n)
        return dict(
            joint_obs=obs,
            obs=[obs[:, f:f + D] for f, D in zip(env.obs_off, env.obs_dims)],
            rew=[rew[:, i] for i in range(env.n)],
            act=[act[:, f:f + K] for f, K in zip(env.act_off, env.act_dims)],
            done=[done[:, i] for i in range(env.n)])

    def reset(self):
        """``env.reset()`` (train.py:104,128) -> list of per-agent (E, D_i) host observations."""
        env = self.env
        env.reset_device()
        v = self._views[self._slot]
        t = torch.from_numpy(v["joint_obs"])
        t.copy_(env.obs, non_blocking=True)
        _lib.synchronize_current_stream()
        return v["obs"]

    def _enqueue(self, src, dst, cursor, counter):
        env, core = self.env, self.core
        ring = core.ring if self.experience else None
        _lib.check(_lib.lib.mdp_host_step_pipelined(
            env._h, core._h, env.num_envs, self.chunks, _lib.ptr(env.state),
            C.c_void_p(self.h_out[src].data_ptr() + self.offs[0]), _lib.ptr(self.d_in), _lib.ptr(self.d_out),
            C.c_void_p(self.h_out[dst].data_ptr()), _lib.ptr(ring.ring) if ring is not None else None,
            ring.capacity if ring is not None else 0, ring.row_stride if ring is not None else 0, cursor, core.seed, counter,
            _lib.current_stream()), "mdp_host_step")

    def _capture(self, src, dst):
        env, core = self.env, self.core
        E = env.num_envs
        env.set_ctl(self.ctl.t)
        core.set_ctl(self.ctl.t)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        try:
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._enqueue(src, dst, 0, 1)  # counter / cursor relative to the control block
                self.ctl.advance(1, E if self.experience else 0, 0)
            self.launches_per_graph = _lib.launch_count() - l0
        finally:
            env.set_ctl(None)
            core.set_ctl(None)
        return g

    def step(self, obs_n):
        env, core = self.env, self.core
        E = env.num_envs
        cur = self._views[self._slot]
        if not (obs_n is cur["obs"] or (len(obs_n) == env.n and all(a is b for a, b in zip(obs_n, cur["obs"])))):
            # foreign arrays: stage them in the current slot's page-locked observation block
            h_in = cur["joint_obs"]
            for i, o in enumerate(obs_n):
                h_in[:, env.obs_off[i]:env.obs_off[i] + env.obs_dims[i]] = np.asarray(o, dtype=np.float32)
        src, dst = self._slot, self._slot ^ 1
        self._slot = dst
        nxt = self._views[dst]
        ring = core.ring if self.experience else None
        if self.use_graph and self._warm:
            if self._graphs[src] is None:
                self._graphs[src] = self._capture(src, dst)
                self._expect = None
            here = (core.counter, ring.next_idx[0] if ring is not None else 0)
            if self._expect != here or self.ctl.dirty:
                self.ctl.upload()
            self._graphs[src].replay()
            self.graph_launches += self.launches_per_graph
            core.counter += 1
            if ring is not None:
                ring.advance_all(E)
            self._expect = (core.counter, ring.next_idx[0] if ring is not None else 0)
        else:
            cursor = ring.reserve_joint(E) if ring is not None else 0
            self._enqueue(src, dst, cursor, core.next_counter())
            self._warm = True  # the first call created the library's streams and events; later ones may be captured
        _lib.synchronize_current_stream()
        return nxt["act"], nxt["obs"], nxt["rew"], nxt["done"]


class GraphedUpdateRound(object):
    """One update round (every agent once, sequentially: train.py:160-161 -> maddpg.py:167-194) with
    device-side index draws, captured into a CUDA graph (single GPU; multi-GPU uses DataParallelUpdater)."""

    def __init__(self, core, batch_size, ctl=None, use_graph=True, grouped=False):
        """grouped=False: agents one after the other (the reference's order, parity mode);
        grouped=True: all agents per launch (mdp_update_all, "Jacobi" order -- throughput mode)."""
        self.core, self.B = core, int(batch_size)
        self.grouped = bool(grouped)
        self.idx_all = torch.zeros((core.n, self.B), dtype=torch.int64, device=core.device)
        self.ctl = ctl if ctl is not None else DeviceCtl(None, core)
        self.use_graph = use_graph
        self._graph = None
        self.launches_per_graph = 0
        self.graph_launches = 0
        self.idx = [torch.zeros(self.B, dtype=torch.int64, device=core.device) for _ in range(core.n)]
        _, self.batch = core._scratch(self.B)

    def _body(self, relative):
        core = self.core
        c = 0
        if self.grouped:
            flat = self.idx_all.view(-1)
            # index draw + statistics reset in one launch; the update's first kernel is its programmatic dependent
            if relative:
                core.make_index(flat, length=0, counter=1, for_update=(0, core.n))
                core.update_all(core.ring.ring, idx=self.idx_all, counter=2)
                self.ctl.advance(2, 0, 0)
            else:
                core.make_index(flat, for_update=(0, core.n))
                core.update_all(core.ring.ring, idx=self.idx_all)
            return 2
        for j in range(core.n):
            c += 1
            if relative:
                core.make_index(self.idx[j], length=0, counter=c, for_update=(j, 1))
            else:
                core.make_index(self.idx[j], for_update=(j, 1))
            c += 1
            core.update_agent(j, core.ring.ring, counter=c if relative else None, idx=self.idx[j])  # fused gather
        if relative:
            self.ctl.advance(c, 0, 0)
        return c

    def run(self, rounds=1):
        core = self.core
        if core.ring.length[0] < self.B:  # the reference's gate is batch_size * max_episode_len rows (maddpg.py:162-163)
            raise RuntimeError("GraphedUpdateRound.run: the replay ring holds %d rows, fewer than one batch of %d"
                               % (core.ring.length[0], self.B))
        if not self.use_graph:
            for _ in range(rounds):
                self._body(False)
            self.ctl.dirty = True
            return
        if self._graph is None:
            self._body(False)  # warm-up outside the capture
            rounds -= 1
            self.ctl.upload()
            core.set_ctl(self.ctl.t)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            try:
                l0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    self._per_round = self._body(True)
                self.launches_per_graph = _lib.launch_count() - l0
                self._graph = g
            finally:
                core.set_ctl(None)
        self.ctl.upload()  # ring length / counter may have moved since the last replay
        for _ in range(rounds):
            self._graph.replay()
            self.graph_launches += self.launches_per_graph
            core.counter += self._per_round
