"""The fork's dict-of-agents algorithms on the B200 kernels: ``MaTd3`` and ``Coma``, plus the fork's own
``Maddpg`` and the two inference-only classes built from the same pieces.

Reference surface: maddpg/algorithms/multiagentalgbase.py:22-212 (``predict``, ``compute_values``, ``compute_loss``,
``train_step``, ``learn_generator``, ``learn``, ``save``, ``load``, ``run_updates``), maddpg/algorithms/matd3.py:11-81,
maddpg/algorithms/coma.py:11-63; the graphs they run are maddpg/modules/matd3module.py:46-111 and comamodule.py:58-153 over
``Policy`` (policy.py:63-100), ``Critic`` (critic.py:60-88) and ``LaggingNetwork`` (laggingnetwork.py:15-48).

One policy group or critic group is one ``MADDPGCore`` (flat parameters, gradients and Adam state on the device); a train step is
a handful of grouped launches (grid.y = agent) from csrc/mdp_td3.cu plus the MADDPG path's critic-step and Adam kernels:

    MATD3:  policy_act(target, noisy) -> q_target(min of twin target critics, TD) -> critic_grads x 2 -> [policy_act ->
            policy_grads through the primary critics' target nets] -> Adam;  run_updates(): polyak(5e-3) of every net.
    COMA:   policy_act(worst) / policy_act(best) at o' -> q_target(global) x 2 -> critic_grads(global) -> q_target(personal, reward
            = global value - worst value) -> critic_grads(personal) -> policy_grads(best, +) / policy_grads(worst, -) -> Adam.

Every gradient of a step is taken before any Adam step (one ``session.run`` in the reference).  Agents are ordered by sorted
name (``U.concat_map``, tf_util.py:53-55).  ``shared_policy`` / ``shared_critic`` groups keep one member -- the first name's --
with the reference's one-loss semantics (policygroup.py:129-135, criticgroup.py:94-100) and its assertion that all names have
equal spaces (which ``Coma``, whose global critic group is always shared, therefore demands too); the ``normalize`` option is the
inference-mode BatchNorm the modules apply (a constant gain, see BATCH_NORM_INFERENCE).  All arithmetic runs in libmaddpg_b200.so; there is no PyTorch or CPU fallback.
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _lib
from .spaces import Discrete
from .trainer import MADDPGCore

TrainInfo = namedtuple("TrainInfo", ["observations", "rewards", "dones", "infos", "actor_loss", "critic_loss", "step"])

UNITS = 64               # policy.py:33, critic.py:31
LEARNING_RATE = 1e-4     # policygroup.py:127, criticgroup.py:91
TARGET_POLYAK = 5e-3     # update_targets(5e-3): matd3module.py:104-107, comamodule.py:137-149
NOISE_STD, NOISE_CLIP = 0.2, 0.5   # policy.py:72-73
# ``normalize``: snt.BatchNormV2()(x, is_training=False) (matd3module.py:65-74, comamodule.py:71-80, maddpgmodule.py:67-76).  The
# modules only ever call it in inference mode and hand none of its variables to an optimizer, so the moving mean stays 0, the moving
# variance 1 and the offset 0 (Sonnet's default has no learned scale): y = x * rsqrt(1 + eps), eps = 1e-3.  Applied while the feed
# dicts are packed into the joint rows.
BATCH_NORM_INFERENCE = np.float32(1.0) / np.sqrt(np.float32(1.0) + np.float32(1e-3))


def _spaces(space):
    return space.spaces if hasattr(space, "spaces") else space


class MultiAgentAlgBase(object):
    """multiagentalgbase.py:22-212 without the TF session: spaces are ``Dict``-likes (``.spaces``) or plain dicts of Box."""

    GAMMA = 0.9

    def __init__(self, observation_space, action_space, device="cuda", seed=0, normalize=None):
        self.observation_space, self.action_space = observation_space, action_space
        normalize = normalize or {}
        self._obs_gain = BATCH_NORM_INFERENCE if normalize.get("observation") else None
        self._rew_gain = BATCH_NORM_INFERENCE if normalize.get("reward") else None
        obs_sp, act_sp = _spaces(observation_space), _spaces(action_space)
        self.first = next(iter(obs_sp))      # the key a shared group is named after (criticgroup.py:24)
        self.names = sorted(obs_sp)
        self.n = len(self.names)
        self.obs_dims = [int(np.prod(obs_sp[k].shape)) for k in self.names]
        self.act_dims = [int(np.prod(act_sp[k].shape)) for k in self.names]
        scale, shift = [], []
        for k in self.names:   # policy.py:76-84
            low, high = float(np.min(act_sp[k].low)), float(np.max(act_sp[k].high))
            interval = (high - low) / 2
            scale.append(interval)
            shift.append(interval + low)
        self._scale = (C.c_float * self.n)(*scale)
        self._shift = (C.c_float * self.n)(*shift)
        self.device = torch.device(device)
        self.seed = int(seed)
        self._counter = 0
        self.sp = self.sc = -1     # index of the name a shared policy / critic group is built on, -1: one member per name
        self._cores = []
        self._buf = {}
        # a train step is 12-19 dependent launches of 5-40 us: from its second call on, every (batch, step kind) replays as one
        # CUDA graph.  The graph bakes the Philox counter of its capture; a device control word (mdp_core_set_ctl) advanced
        # inside the graph keeps every replay on a fresh noise stream.
        self.use_graphs = True
        self._graphs, self._graph_seen = {}, {}
        self._ctl = torch.zeros(4, dtype=torch.int64, device=self.device)

    # -- groups ---------------------------------------------------------------------------------
    def _group(self, seed_offset):
        """A PolicyGroup / CriticGroup: running + target nets for every name, Adam(1e-4), no gradient clipping."""
        core = MADDPGCore(self.obs_dims, [Discrete(k) for k in self.act_dims], [False] * self.n, num_units=UNITS, lr=LEARNING_RATE,
                          gamma=self.GAMMA, device=self.device, seed=self.seed + seed_offset, replay_capacity=1,
                          polyak=TARGET_POLYAK, grad_clip=0.0, actor_reg=0.0)
        core.set_tensor_cores(-1)   # the critic step of these algorithms runs on the fp32 SIMT tiles (batch 1024, 64 units)
        core.set_ctl(self._ctl)
        # snt.Linear's defaults (snt.nets.MLP, laggingnetwork.py:22-23): truncated normal, stddev 1/sqrt(fan_in), zero biases;
        # running and target nets are initialised independently
        rng = np.random.RandomState(self.seed + 7919 * (seed_offset + 1))
        for a in range(self.n):
            for net in (_lib.NET_P, _lib.NET_TARGET_P, _lib.NET_Q, _lib.NET_TARGET_Q):
                arrs = []
                for shp in core.net_shapes(a, net):
                    if len(shp) == 2:
                        w = rng.standard_normal(shp)
                        bad = np.abs(w) > 2.0
                        while bad.any():
                            w[bad] = rng.standard_normal(int(bad.sum()))
                            bad = np.abs(w) > 2.0
                        arrs.append((w / np.sqrt(shp[0])).astype(np.float32))
                    else:
                        arrs.append(np.zeros(shp, np.float32))
                core.set_weights(a, net, arrs)
        self._cores.append(core)
        return core

    @property
    def layout(self):
        return self._cores[0].ring.layout

    def _scratch(self, key, shape, dtype=torch.float32):
        t = self._buf.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self._buf[key] = torch.zeros(shape, dtype=dtype, device=self.device)
        return t

    # -- host <-> device ------------------------------------------------------------------------
    def _joint(self, d, dims, stride, gain=None, key=None):
        """dict name -> (B, dim) host arrays  ->  (B, stride) device array in sorted-name column order (``key``: into the
        persistent buffer of that name, whose address a captured train step keeps reading)."""
        B = int(np.reshape(d[self.names[0]], (-1, dims[0])).shape[0])
        host = np.zeros((B, stride), np.float32)
        o = 0
        for k, dim in zip(self.names, dims):
            host[:, o:o + dim] = np.reshape(d[k], (-1, dim))
            o += dim
        if gain is not None:
            host *= gain
        if key is None:
            return torch.from_numpy(host).to(self.device, non_blocking=False)
        dev = self._scratch((key, B), (B, stride))
        dev.copy_(torch.from_numpy(host))
        return dev

    def _rows(self, observations, actions, rewards, observations_n, dones):
        """The five feed dicts as (B, row_stride) joint rows (the replay ring's row layout, include/maddpg_b200.h)."""
        L = self.layout
        B = int(np.reshape(observations[self.names[0]], (-1, self.obs_dims[0])).shape[0])
        host = np.zeros((B, int(L.row_stride)), np.float32)
        for j, k in enumerate(self.names):
            D, K, o, a = self.obs_dims[j], self.act_dims[j], int(L.obs_off[j]), int(L.act_off[j])
            host[:, o:o + D] = np.reshape(observations[k], (-1, D))
            host[:, int(L.nx_off) + o:int(L.nx_off) + o + D] = np.reshape(observations_n[k], (-1, D))
            host[:, int(L.obs_sum) + a:int(L.obs_sum) + a + K] = np.reshape(actions[k], (-1, K))
            host[:, int(L.rw_off) + j] = np.reshape(rewards[k], -1)
            host[:, int(L.dn_off) + j] = np.reshape(dones[k], -1)
        if self._obs_gain is not None:
            host[:, :int(L.obs_sum)] *= self._obs_gain
            host[:, int(L.nx_off):int(L.nx_off) + int(L.obs_sum)] *= self._obs_gain
        if self._rew_gain is not None:
            host[:, int(L.rw_off):int(L.rw_off) + self.n] *= self._rew_gain
        dev = self._scratch(("rows_in", B), (B, int(L.row_stride)))     # persistent: a captured train step reads this address
        dev.copy_(torch.from_numpy(host))
        return dev

    def _split(self, joint, dims):
        out, o = {}, 0
        host = joint.cpu().numpy()
        for k, dim in zip(self.names, dims):
            out[k] = host[:, o:o + dim].copy()
            o += dim
        return out

    # -- kernels --------------------------------------------------------------------------------
    def _policy_act(self, policies, x, x_stride, out, use_target=False, noise_std=0.0, noise=None, shared=-1):
        B = out.shape[0]
        self._counter += 1
        _lib.check(_lib.lib.mdp_td3_policy_act(policies._h, int(use_target), B, _lib.ptr(x), int(x_stride), _lib.ptr(noise),
                                               0 if noise is None else int(noise.stride(0)), float(noise_std), NOISE_CLIP,
                                               self.seed, self._counter, self._scale, self._shift, _lib.ptr(out),
                                               int(out.stride(0)), int(shared), _lib.current_stream()), "mdp_td3_policy_act")
        return out

    def _q_target(self, critics_a, critics_b, batch, obs_field, act, q_out=None, y_out=None, use_target=True, rew_override=None,
                  rew_minus=None, shared_agent=-1):
        B = batch.shape[0]
        _lib.check(_lib.lib.mdp_td3_q_target(critics_a._h, None if critics_b is None else critics_b._h, int(use_target),
                                             C.byref(self.layout), B, _lib.ptr(batch), int(obs_field), _lib.ptr(act),
                                             int(act.stride(0)), _lib.ptr(rew_override), _lib.ptr(rew_minus), int(shared_agent),
                                             float(self.GAMMA), _lib.ptr(q_out), _lib.ptr(y_out), _lib.current_stream()),
                   "mdp_td3_q_target")

    def _policy_grads(self, policies, critics, batch, act_all, sign=1.0, critic_use_target=True, shared_policy=-1, critic_agent=-1):
        _lib.check(_lib.lib.mdp_td3_policy_grads(policies._h, critics._h, int(critic_use_target), float(sign), C.byref(self.layout),
                                                 batch.shape[0], _lib.ptr(batch), _lib.ptr(act_all), int(act_all.stride(0)),
                                                 self._scale, self._shift, int(shared_policy), int(critic_agent),
                                                 _lib.current_stream()), "mdp_td3_policy_grads")

    def _critic_step(self, critics, shared, batch, y):
        """CriticGroup.create_optimizers (criticgroup.py:87-106): every name's critic, or -- shared -- the one critic on the first
        name's targets."""
        if shared >= 0:
            critics.critic_grads(shared, batch, y[shared])
        else:
            self._critic_grads_all(critics, batch, y)

    def _adam(self, core, which, shared):
        if shared >= 0:
            core.clip_adam_polyak(shared, which, do_polyak=False)
        else:
            self._adam_all(core, which)

    def _critic_grads_all(self, critics, batch, y):
        _lib.check(_lib.lib.mdp_critic_grads_all(critics._h, C.byref(self.layout), batch.shape[0], _lib.ptr(batch), None, 0,
                                                 _lib.ptr(y), _lib.current_stream()), "mdp_critic_grads_all")

    @staticmethod
    def _adam_all(core, which):
        _lib.check(_lib.lib.mdp_clip_adam_polyak_all(core._h, int(which), 1.0, 0, _lib.current_stream()), "mdp_clip_adam_polyak_all")

    def _read_stats(self, cores):
        """One device-to-host read of every group's loss accumulators -> list of (n, 8) float64 arrays."""
        host = torch.cat([c.stats for c in cores]).cpu().numpy()
        return [host[8 * self.n * i:8 * self.n * (i + 1)].reshape(self.n, 8) for i in range(len(cores))]

    @staticmethod
    def _polyak(core, mask):
        _lib.check(_lib.lib.mdp_td3_polyak(core._h, int(mask), TARGET_POLYAK, _lib.current_stream()), "mdp_td3_polyak")

    def _act_buf(self, key, B):
        return self._scratch((key, B), (B, self._cores[0].act_stride))   # one buffer per batch size: captured graphs keep the address

    # -- reference surface ----------------------------------------------------------------------
    def predict(self, observations, noisy=True):
        """multiagentalgbase.py:50-67: the running policies' actions, plus N(0, 0.2) exploration noise drawn on the host
        (``npr.normal``: numpy's global stream, exactly as the reference consumes it)."""
        obs = self._joint(observations, self.obs_dims, self._cores[0].obs_stride, self._obs_gain)
        act = self._policy_act(self._predict_policies(), obs, obs.stride(0), self._act_buf("predict", obs.shape[0]),
                               shared=self.sp)
        actions = self._split(act, self.act_dims)
        if noisy:
            return {k: np.squeeze(a + np.random.normal(scale=0.2, size=a.shape)) for k, a in actions.items()}
        return {k: np.squeeze(a) for k, a in actions.items()}

    def compute_values(self, observations):
        """multiagentalgbase.py:69-78: ``critic_predict`` = the value critics' TARGET nets at (obs, predicted actions)."""
        obs = self._joint(observations, self.obs_dims, self._cores[0].obs_stride, self._obs_gain)
        B = obs.shape[0]
        act = self._policy_act(self._predict_policies(), obs, obs.stride(0), self._act_buf("predict", B), shared=self.sp)
        rows = self._scratch(("rows", B), (B, int(self.layout.row_stride)))
        rows[:, :obs.shape[1]].copy_(obs)    # only the obs columns are read (obs_field = 0, no TD combine)
        q = self._scratch(("values", B), (self.n, B))
        self._q_target(self._value_critics(), None, rows, 0, act, q_out=q, shared_agent=self._value_shared())
        host = q.cpu().numpy()
        return {k: host[j][:, None].copy() for j, k in enumerate(self.names)}

    def train_step(self, observations, actions, rewards, observations_n, dones, step=None, noise=None):
        """multiagentalgbase.py:92-104.  -> {'actor': {name: loss}, 'critic': {name: loss}} (``unflatten_map`` of the outputs).
        ``noise``: optional {name: (B, K)} N(0, 1) draws behind the noisy target's ``tf.random.normal`` (parity runs)."""
        rows = self._rows(observations, actions, rewards, observations_n, dones)
        z = None if noise is None else self._joint(noise, self.act_dims, self._cores[0].act_stride, key="z_in")
        return self._train_step(rows, step, z, update=True)

    def compute_loss(self, observations, actions, rewards, observations_n, dones, noise=None):
        """multiagentalgbase.py:80-90: the losses of a policy step without the optimizer."""
        rows = self._rows(observations, actions, rewards, observations_n, dones)
        z = None if noise is None else self._joint(noise, self.act_dims, self._cores[0].act_stride, key="z_in")
        return self._train_step(rows, 2, z, update=False)

    def update_targets(self):
        self.run_updates()

    def _is_policy_step(self, step):
        return True

    def _train_step(self, rows, step, z, update):
        """All launches of one train step (``_launch``), eagerly the first time a (batch, step kind) is seen and as one CUDA graph
        afterwards, then one read of the loss accumulators (``_losses``)."""
        B, policy_step = rows.shape[0], self._is_policy_step(step)
        if not update:     # compute_loss: gradients only, then put the state back
            self._save_adam_t()
            self._launch(rows, z, policy_step, False)
            out = self._losses(B, policy_step)
            self._discard_grads()
            return out
        key = (B, policy_step, z is not None)
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if not self.use_graphs or seen == 0:
            self._launch(rows, z, policy_step, True)
        else:
            graph = self._graphs.get(key)
            if graph is None:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    _lib.check(_lib.lib.mdp_ctl_advance(_lib.ptr(self._ctl), 1 << 32, 0, 1, 0, _lib.current_stream()),
                               "mdp_ctl_advance")
                    self._launch(rows, z, policy_step, True)
                self._graphs[key] = graph
            graph.replay()
        return self._losses(B, policy_step)

    def _check_shared(self, always_shared_critic=False):
        """A shared group asserts that every name has the first name's spaces (policygroup.py:32-34, criticgroup.py:28-30)."""
        if (self.sp >= 0 or self.sc >= 0 or always_shared_critic) and (
                len(set(self.obs_dims)) > 1 or len(set(self.act_dims)) > 1 or len(set(self._scale)) > 1 or len(set(self._shift)) > 1):
            raise AssertionError("a shared policy / critic group needs equal observation and action spaces "
                                 "(policygroup.py:32-34, criticgroup.py:28-30)")

    def _zero_stats(self):
        for c in self._cores:
            c.stats.zero_()

    def _discard_grads(self):
        for c in self._cores:   # compute_loss: drop the accumulated gradients and the Adam step the gradient kernels counted
            c.grads.zero_()
            c.adam_t.zero_().add_(self._adam_t_saved[id(c)])

    def _save_adam_t(self):
        self._adam_t_saved = {id(c): c.adam_t.clone() for c in self._cores}

    def learn_generator(self, env, timesteps=10 ** 6, replay=None):
        """multiagentalgbase.py:106-132, line for line; ``replay`` defaults to a host ``DictReplayBuffer`` of the same size."""
        if replay is None:
            replay = DictReplayBuffer(timesteps // 100)
        done = True
        for step in range(timesteps):
            if done:
                observations_last = env.reset()
            actions = self.predict(observations_last)
            observations, reward, done, infos = env.step(actions)
            rewards = {key: reward for key in observations}
            dones = {key: done for key in observations}
            replay.add(observations_last, actions, rewards, observations, dones)
            observations_last = observations
            train_info = {}
            if step > 1024 and step % 5000 == 0:
                train_info = self.train_step(*replay.sample(1024), step)
                self.run_updates()
            yield TrainInfo(observations, rewards, dones, infos, train_info.get("actor"), train_info.get("critic"), step)

    def learn(self, env, timesteps=10 ** 6, replay=None, verbose=True):
        """multiagentalgbase.py:134-165 without the tqdm bar.  Returns None like the reference; the running episode reward it
        prints is kept in ``self.running_reward``."""
        ep_reward, total_reward = None, 0
        for info in self.learn_generator(env, timesteps, replay):
            total_reward += float(np.mean(list(info.rewards.values())))
            if any(info.dones.values()):
                ep_reward = total_reward if not ep_reward else ep_reward * .99 + total_reward * .01
                total_reward = 0
            if info.actor_loss and verbose:
                print("Training Step:", info.step, "Running Reward: {:+6.6f}".format(ep_reward or 0.0),
                      "Actor Loss:", float(np.mean(list(info.actor_loss.values()))),
                      "Critic Loss:", float(np.mean(list(info.critic_loss.values()))))
            self.running_reward = ep_reward

    def save(self, path):
        """multiagentalgbase.py:167-173: every group's parameters and Adam state."""
        blob = {}
        for i, c in enumerate(self._cores):
            for name in ("params", "adam_m", "adam_v", "adam_t"):
                blob["g%d_%s" % (i, name)] = getattr(c, name).cpu().numpy()
        np.savez(str(path), **blob)

    def load(self, path):
        """multiagentalgbase.py:175-181."""
        path = str(path)
        blob = np.load(path if path.endswith(".npz") else path + ".npz")
        for i, c in enumerate(self._cores):
            for name in ("params", "adam_m", "adam_v", "adam_t"):
                getattr(c, name).copy_(torch.from_numpy(blob["g%d_%s" % (i, name)]))


class DictReplayBuffer(object):
    """maddpg/common/replaybuffer.py:10-103: the fork's host-side dict replay (tuples of per-name dicts, ``random.randint``
    index draws).  Pinned to the real class: tests/golden/dict_replay_ref.npz (tests/test_dict_replay.py).  ``learn_generator`` keeps it on the host like the reference; the sampled batch crosses to the device once per
    train step (every 5000 env steps, multiagentalgbase.py:126)."""

    def __init__(self, size):
        import random
        self._random = random
        self._storage, self._maxsize, self._next_idx = [], int(size), 0

    def __len__(self):
        return len(self._storage)

    def clear(self):
        self._storage, self._next_idx = [], 0

    def add(self, obs_t, action, reward, obs_tp1, done):
        data = (obs_t, action, reward, obs_tp1, done)
        if self._next_idx >= len(self._storage):
            self._storage.append(data)
        else:
            self._storage[self._next_idx] = data
        self._next_idx = (self._next_idx + 1) % self._maxsize

    def make_index(self, batch_size):
        return [self._random.randint(0, len(self._storage) - 1) for _ in range(batch_size)]

    def make_latest_index(self, batch_size):
        idx = [(self._next_idx - 1 - i) % self._maxsize for i in range(batch_size)]
        np.random.shuffle(idx)
        return idx

    def sample_index(self, idxes):
        out = tuple({} for _ in range(5))
        for i in idxes:
            for field, d in zip(self._storage[i], out):
                for key, value in field.items():
                    d.setdefault(key, []).append(value)
        return out

    def sample(self, batch_size):
        return self.sample_index(self.make_index(batch_size) if batch_size > 0 else range(len(self._storage)))

    def collect(self):
        return self.sample(-1)


class MaTd3(MultiAgentAlgBase):
    """maddpg/algorithms/matd3.py:11-81 over MaTD3Module (matd3module.py:20-134).  Two deliberate differences from the reference
    as written: it can be constructed (the reference's module passes ``create_optimizers`` an argument it does not take,
    matd3module.py:98-99 -- dropped), and a critic-only step returns ``{'critic': {name: loss}}`` where the reference pushes the
    un-prefixed dict through ``unflatten_map`` and returns the names split at "_" (matd3.py:71-72)."""

    GAMMA = 0.9   # MaTD3Module._build(..., gamma=0.9), matd3module.py:47

    def __init__(self, observation_space, action_space, shared_policy=False, shared_critic=False, normalize=None, **kw):
        super().__init__(observation_space, action_space, normalize=normalize, **kw)
        self.policies = self._group(0)                       # uses the P nets
        self.critics = [self._group(1), self._group(2)]      # twin critic groups: the Q nets
        first = self.names.index(self.first)
        self.sp = first if shared_policy else -1             # PolicyGroup(shared=...): the first name's member serves all
        self.sc = first if shared_critic else -1             # CriticGroup(shared=...)
        self._check_shared()

    def _predict_policies(self):
        return self.policies

    def _value_critics(self):
        return self.critics[0]

    def _value_shared(self):
        return self.sc

    def run_updates(self):
        self._polyak(self.policies, 1)
        self._polyak(self.critics[0], 2)
        self._polyak(self.critics[1], 2)

    def _is_policy_step(self, step):
        return bool(step) and step % 2 == 0     # matd3.py:69

    def _launch(self, rows, z, policy_step, update):
        B, L = rows.shape[0], self.layout
        self._zero_stats()
        nx = rows[:, int(L.nx_off):]
        # noisy target actions at o', min of the twin target critics, TD combine (matd3module.py:76-83, 113-123)
        sp, sc = self.sp, self.sc
        a_n = self._policy_act(self.policies, nx, rows.stride(0), self._act_buf("a_next", B), use_target=True,
                               noise_std=NOISE_STD, noise=z, shared=sp)
        y = self._scratch(("y", B), (self.n, B))
        self._q_target(self.critics[0], self.critics[1], rows, 1, a_n, y_out=y, shared_agent=sc)
        for cr in self.critics:      # both critic groups regress on the same targets (:88-95)
            self._critic_step(cr, sc, rows, y)
        if policy_step:
            a = self._policy_act(self.policies, rows, rows.stride(0), self._act_buf("a_now", B), shared=sp)
            # a shared group's one loss is the first name's value (policygroup.py:129-135)
            self._policy_grads(self.policies, self.critics[0], rows, a, shared_policy=sp, critic_agent=sc if sc >= 0 else sp)
        if update:
            for cr in self.critics:
                self._adam(cr, 1, sc)
            if policy_step:
                self._adam(self.policies, 0, sp)

    def _losses(self, B, policy_step):
        sp, sc = self.sp, self.sc
        stats = self._read_stats([self.policies] + self.critics)
        out = {"critic": {k: np.float32(np.mean(np.asarray([stats[1][sc if sc >= 0 else j, 0] / B,
                                                             stats[2][sc if sc >= 0 else j, 0] / B], np.float32)))
                          for j, k in enumerate(self.names)}}
        if policy_step:
            out["actor"] = {k: np.float32(stats[0][sp if sp >= 0 else j, 1] / B) for j, k in enumerate(self.names)}
        return out


class Coma(MultiAgentAlgBase):
    """maddpg/algorithms/coma.py:11-63 over ComaModule (comamodule.py:20-171): best / worst policy groups, one shared global
    critic (trained on the first name's reward), per-name personal critics fed the reward ``global value - worst value``."""

    GAMMA = 0.95   # ComaModule._build(..., gamma=0.95), comamodule.py:59

    def __init__(self, observation_space, action_space, shared_policy=False, shared_critic=False, normalize=None, **kw):
        super().__init__(observation_space, action_space, normalize=normalize, **kw)   # shared_critic: ignored by ComaModule too (:36-43)
        self.best = self._group(0)
        self.worst = self._group(1)
        self.global_critic = self._group(2)    # CriticGroup(shared=True): only the first name's critic exists
        self.personal = self._group(3)
        self.shared = self.names.index(self.first)
        self.sp = self.shared if shared_policy else -1
        self._check_shared(always_shared_critic=True)   # ComaModule's global critic group is CriticGroup(shared=True), always

    def _predict_policies(self):
        return self.best

    def _value_critics(self):
        return self.personal

    def _value_shared(self):
        return -1

    def run_updates(self):
        self._polyak(self.global_critic, 2)
        self._polyak(self.personal, 2)
        self._polyak(self.worst, 1)
        self._polyak(self.best, 1)

    def _launch(self, rows, z, policy_step, update):
        B, L, s = rows.shape[0], self.layout, self.shared
        self._zero_stats()
        nx = rows[:, int(L.nx_off):]
        sp = self.sp
        worst_n = self._policy_act(self.worst, nx, rows.stride(0), self._act_buf("worst_next", B), shared=sp)
        best_n = self._policy_act(self.best, nx, rows.stride(0), self._act_buf("best_next", B), shared=sp)
        worst_q = self._scratch(("worst_q", B), (self.n, B))
        y_global = self._scratch(("y_global", B), (self.n, B))
        self._q_target(self.global_critic, None, rows, 1, worst_n, q_out=worst_q, shared_agent=s)     # comamodule.py:82-86
        self._q_target(self.global_critic, None, rows, 1, best_n, y_out=y_global, shared_agent=s)     # :88-92, 155-162
        global_q = self._scratch(("global_q", B), (self.n, B))
        _lib.check(_lib.lib.mdp_critic_grads(self.global_critic._h, s, C.byref(L), B, _lib.ptr(rows), None, _lib.ptr(y_global[s]),
                                             _lib.ptr(global_q[s]), _lib.current_stream()), "mdp_critic_grads")   # :98-102
        if self.n > 1:   # every name sees the shared critic's value (criticgroup.py:58-63)
            for j in range(self.n):
                if j != s:
                    global_q[j].copy_(global_q[s])
        y_personal = self._scratch(("y_personal", B), (self.n, B))
        self._q_target(self.personal, None, rows, 1, best_n, y_out=y_personal, rew_override=global_q, rew_minus=worst_q)  # :104-114
        self._critic_grads_all(self.personal, rows, y_personal)                                       # :115-116
        best_a = self._policy_act(self.best, rows, rows.stride(0), self._act_buf("best_now", B), shared=sp)
        self._policy_grads(self.best, self.personal, rows, best_a, sign=1.0, shared_policy=sp, critic_agent=sp)    # :118-121, 129
        worst_a = self._policy_act(self.worst, rows, rows.stride(0), self._act_buf("worst_now", B), shared=sp)
        self._policy_grads(self.worst, self.personal, rows, worst_a, sign=-1.0, shared_policy=sp, critic_agent=sp)  # :123-127, 130
        if update:
            self.global_critic.clip_adam_polyak(s, 1, do_polyak=False)
            self._adam_all(self.personal, 1)
            self._adam(self.best, 0, sp)
            self._adam(self.worst, 0, sp)

    def _losses(self, B, policy_step):
        s, sp = self.shared, self.sp
        st = dict(zip(("best", "worst", "global_critic", "personal"),
                      self._read_stats([self.best, self.worst, self.global_critic, self.personal])))
        gl = np.float32(st["global_critic"][s, 0] / B)
        out = {"critic": {k: np.float32(np.mean(np.asarray([gl, st["personal"][j, 0] / B], np.float32)))
                          for j, k in enumerate(self.names)},
               "actor": {k: np.float32(np.std(np.asarray([st["best"][sp if sp >= 0 else j, 1] / B,
                                                           st["worst"][sp if sp >= 0 else j, 1] / B], np.float32)))
                         for j, k in enumerate(self.names)}}
        return out


class Maddpg(MultiAgentAlgBase):
    """maddpg/algorithms/maddpg.py:11-76 over MaddpgModule (maddpgmodule.py:19-127): the fork's own MADDPG -- tanh policies, one
    critic group, TD actions from the target policies (no noise), policy loss through the critics' target nets, every group
    steps on every train step.  ``hyperparameters`` is accepted and, like the reference (maddpg.py:19 replaces any given dict by
    ``{}``), has no effect: gamma 0.95, learning rates 1e-4."""

    GAMMA = 0.95

    def __init__(self, observation_space, action_space, shared_policy=False, shared_critic=False, hyperparameters=None, **kw):
        super().__init__(observation_space, action_space, **kw)   # no normalize: it would come from the discarded hyperparameters
        self.policies = self._group(0)
        self.critics = self._group(1)
        first = self.names.index(self.first)
        self.sp = first if shared_policy else -1
        self.sc = first if shared_critic else -1
        self._check_shared()

    def _predict_policies(self):
        return self.policies

    def _value_critics(self):
        return self.critics

    def _value_shared(self):
        return self.sc

    def run_updates(self):
        self._polyak(self.policies, 1)
        self._polyak(self.critics, 2)

    def _launch(self, rows, z, policy_step, update):
        B, L, sp, sc = rows.shape[0], self.layout, self.sp, self.sc
        self._zero_stats()
        nx = rows[:, int(L.nx_off):]
        a_n = self._policy_act(self.policies, nx, rows.stride(0), self._act_buf("a_next", B), use_target=True, shared=sp)  # :77
        y = self._scratch(("y", B), (self.n, B))
        self._q_target(self.critics, None, rows, 1, a_n, y_out=y, shared_agent=sc)                                       # :82-83, 113-118
        self._critic_step(self.critics, sc, rows, y)                                                                    # :89-93
        a = self._policy_act(self.policies, rows, rows.stride(0), self._act_buf("a_now", B), shared=sp)
        self._policy_grads(self.policies, self.critics, rows, a, shared_policy=sp, critic_agent=sc if sc >= 0 else sp)   # :94-98
        if update:
            self._adam(self.critics, 1, sc)
            self._adam(self.policies, 0, sp)

    def _losses(self, B, policy_step):
        sp, sc = self.sp, self.sc
        st = self._read_stats([self.policies, self.critics])
        out = {"actor": {k: np.float32(st[0][sp if sp >= 0 else j, 1] / B) for j, k in enumerate(self.names)},
               "critic": {k: np.float32(st[1][sc if sc >= 0 else j, 0] / B) for j, k in enumerate(self.names)}}
        return out


class _Inference(MultiAgentAlgBase):
    """MaddpgInference (maddpg/algorithms/maddpg.py:79-118) / ComaInference (coma.py:73-112): a policy group and ``predict`` only;
    the other methods are the reference's ``...`` bodies (they return None)."""

    def __init__(self, observation_space, action_space, shared_policy=False, normalize=None, **kw):
        if normalize and normalize.get("reward"):
            normalize = {"observation": normalize.get("observation")}   # the inference modules only normalise observations
        super().__init__(observation_space, action_space, normalize=normalize, **kw)
        self.policies = self._group(0)
        self.sp = self.names.index(self.first) if shared_policy else -1
        self._check_shared()

    def _predict_policies(self):
        return self.policies

    def compute_values(self, observations):
        return None

    def compute_loss(self, *args, **kw):
        return None

    def train_step(self, *args, **kw):
        return None

    def run_updates(self):
        return None


class MaddpgInference(_Inference):
    pass


class ComaInference(_Inference):
    pass
