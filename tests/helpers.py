"""Shared builders for the parity tests: seeded env cases and trainer cases driven through the
oracle (oracle/) -- the CUDA path is checked against these on the same inputs."""
import argparse

import numpy as np

from oracle import maddpg as omaddpg
from oracle import mpe as ompe

ENV_CASES = {
    # name: (scenario, num_agents, E, T)
    "simple": ("simple", None, 16, 25),
    "simple_spread": ("simple_spread", 3, 16, 25),
    "simple_tag": ("simple_tag", None, 16, 25),
    "simple_world_comm": ("simple_world_comm", None, 8, 25),
    "simple_spread_24": ("simple_spread", 24, 2, 4),
    # SURVEY 8(f) rank 2
    "simple_adversary": ("simple_adversary", None, 16, 25),
    "simple_push": ("simple_push", None, 16, 25),
    "simple_speaker_listener": ("simple_speaker_listener", None, 16, 25),
    "simple_crypto": ("simple_crypto", None, 16, 25),
    "simple_reference": ("simple_reference", None, 16, 25),
}


def soft_actions(rng, E, K, heads=None, sharp=2.0):
    """(E, K) float32 rows that look like Gumbel-softmax samples (each head sums to 1)."""
    z = rng.randn(E, K) * sharp
    out = np.empty((E, K), np.float32)
    o = 0
    for h in (heads or [K]):
        zz = z[:, o:o + h]
        e = np.exp(zz - zz.max(1, keepdims=True))
        out[:, o:o + h] = (e / e.sum(1, keepdims=True)).astype(np.float32)
        o += h
    return out


def env_case(name, seed=0, crowd=0.45):
    """Seeded initial state (crowded so that contacts and collisions happen) + action tape."""
    scenario, na, E, T = ENV_CASES[name]
    rng = np.random.RandomState(seed)
    env = ompe.BatchedOracleEnv(scenario, E, na, seed=seed)
    A, L = env.n, len(env.envs[0].world.landmarks)
    agent_pos = rng.uniform(-crowd, crowd, size=(E, A, 2))
    agent_vel = rng.uniform(-0.3, 0.3, size=(E, A, 2))
    landmark_pos = rng.uniform(-crowd, crowd, size=(E, L, 2))
    dim_c = env.envs[0].world.dim_c
    agent_c = np.zeros((E, A, dim_c))
    n_goal = getattr(env.envs[0].scenario, "n_goal", 0)
    goal = rng.randint(L, size=(E, n_goal)) if n_goal else None
    heads = [omaddpg.act_heads(s) for s in env.action_space]
    tape = [[soft_actions(rng, E, env.act_dims[i], heads[i]) for i in range(A)] for _ in range(T)]
    return dict(scenario=scenario, num_agents=na, E=E, T=T, env=env, agent_pos=agent_pos, agent_vel=agent_vel,
                landmark_pos=landmark_pos, agent_c=agent_c, goal=goal, tape=tape, heads=heads)


def run_oracle_rollout(case):
    env = case["env"]
    env.set_state(case["agent_pos"], case["agent_vel"], case["landmark_pos"], case["agent_c"], case.get("goal"))
    obs0 = env.observe()
    obs_t, rew_t = [], []
    for acts in case["tape"]:
        o, r, d = env.step(acts)
        assert not d.any()
        obs_t.append(np.concatenate(o, axis=1))
        rew_t.append(r)
    return dict(obs0=np.concatenate(obs0, axis=1), obs=np.stack(obs_t), rew=np.stack(rew_t), final=env.get_state())


# ------------------------------------------------------------------------------------------------
TRAINER_CASES = {
    # name: (scenario, num_agents, units, B, local_q per agent or None)
    "simple": ("simple", None, 64, 96, None),
    "simple_spread": ("simple_spread", 3, 64, 160, None),
    "simple_tag": ("simple_tag", None, 64, 100, None),
    "simple_world_comm": ("simple_world_comm", None, 128, 64, None),
    "simple_tag_ddpg_adv": ("simple_tag", None, 64, 64, [True, True, True, False]),
    "simple_spread_6": ("simple_spread", 6, 64, 48, None),
    # BASELINE.json batch sizes: configs[1] (batch 1024), configs[2] (batch 4096), configs[3] (num_units 128, batch 1024)
    "simple_spread_b1024": ("simple_spread", 3, 64, 1024, None),
    "simple_tag_b4096": ("simple_tag", None, 64, 4096, None),
    "simple_world_comm_b1024": ("simple_world_comm", None, 128, 1024, None),
    # SURVEY 8(f) rank 2 shapes: a 3-wide observation and a Discrete(3) / Discrete(4) communication head
    "simple_speaker_listener": ("simple_speaker_listener", None, 64, 96, None),
    "simple_crypto": ("simple_crypto", None, 64, 80, None),
    "simple_adversary_ddpg_good": ("simple_adversary", None, 64, 64, [False, True, True]),
    "simple_reference": ("simple_reference", None, 64, 72, None),   # MultiDiscrete([5, 10]) heads: 15 action columns
}


def make_args(units, batch_size, lr=1e-2, gamma=0.95, max_episode_len=25):
    return argparse.Namespace(lr=lr, gamma=gamma, batch_size=batch_size, num_units=units,
                              max_episode_len=max_episode_len)


def trainer_case(name, seed=0, rows=None):
    """Oracle trainers with seeded weights + a seeded pool of transitions + noise tapes."""
    scenario, na, units, B, local_q = TRAINER_CASES[name]
    rng = np.random.RandomState(1000 + seed)
    env = ompe.make_env(scenario, np.random.RandomState(seed), na)
    n = env.n
    obs_shape_n = [s.shape for s in env.observation_space]
    args = make_args(units, B)
    local_q = local_q or [False] * n
    trainers = [omaddpg.OracleAgentTrainer("agent_%d" % i, None, obs_shape_n, env.action_space, i, args,
                                           local_q_func=local_q[i], rng=np.random.RandomState(seed * 100 + i))
                for i in range(n)]
    for tr in trainers:  # non-zero biases so that bias paths are exercised
        for net in (tr.q, tr.target_q, tr.p, tr.target_p):
            for k in (1, 3, 5):
                net.p[k][...] = rng.uniform(-0.1, 0.1, size=net.p[k].shape).astype(np.float32)
    rows = rows or (B * 2 + 7)
    obs_dims = [s[0] for s in obs_shape_n]
    act_dims = trainers[0].act_dims
    heads = trainers[0].heads_n
    pool = dict(
        obs=[rng.randn(rows, D).astype(np.float32) for D in obs_dims],
        act=[soft_actions(rng, rows, K, h) for K, h in zip(act_dims, heads)],
        rew=[rng.randn(rows).astype(np.float32) for _ in range(n)],
        nobs=[rng.randn(rows, D).astype(np.float32) for D in obs_dims],
        done=[(rng.rand(rows) < 0.15).astype(np.float32) for _ in range(n)],
    )
    idx = [rng.randint(0, rows, size=B).tolist() for _ in range(n)]
    u_target = [rng.uniform(1e-6, 1.0, size=(B, sum(act_dims))).astype(np.float32) for _ in range(n)]
    u_actor = [rng.uniform(1e-6, 1.0, size=(B, act_dims[j])).astype(np.float32) for j in range(n)]
    return dict(name=name, n=n, units=units, B=B, local_q=local_q, env=env, args=args, trainers=trainers,
                obs_dims=obs_dims, act_dims=act_dims, heads=heads, pool=pool, idx=idx, u_target=u_target,
                u_actor=u_actor, rows=rows, obs_shape_n=obs_shape_n)


def fill_oracle_replay(case):
    p = case["pool"]
    for i, tr in enumerate(case["trainers"]):
        for r in range(case["rows"]):
            tr.replay_buffer.add(p["obs"][i][r].astype(np.float64), p["act"][i][r], float(p["rew"][i][r]),
                                 p["nobs"][i][r].astype(np.float64), float(p["done"][i][r]))


class NoiseTape(object):
    """Feeds the oracle's ``noise(shape)`` calls from pre-drawn arrays, in call order."""

    def __init__(self):
        self.queue = []

    def push(self, arr):
        self.queue.append(np.asarray(arr, np.float32))

    def __call__(self, shape):
        a = self.queue.pop(0)
        assert tuple(a.shape) == tuple(shape), (a.shape, shape)
        return a


def oracle_update_round(case):
    """Sequential update of every agent (train.py:160-161) with injected indices / noise.
    Returns per-agent dicts of everything the CUDA path is compared against."""
    trainers, n = case["trainers"], case["n"]
    fill_oracle_replay(case)
    tape = NoiseTape()
    for tr in trainers:
        tr.noise = tape
        tr.max_replay_buffer_len = 0
    out = []
    off = np.concatenate([[0], np.cumsum(case["act_dims"])]).astype(int)
    for j, tr in enumerate(trainers):
        ut = case["u_target"][j]
        for i in range(n):  # target_act of every agent, in agent order (maddpg.py:184)
            tape.push(ut[:, off[i]:off[i + 1]])
        tape.push(case["u_actor"][j])  # p_train's sample (maddpg.py:49)
        stats = tr.update(trainers, 100, index=case["idx"][j])
        assert not tape.queue
        out.append(dict(stats=stats, y=tr.last_target_q.astype(np.float32),
                        q_grads=tr.last_grads["q"], p_grads=tr.last_grads["p"],
                        q=[x.copy() for x in tr.q.p], p=[x.copy() for x in tr.p.p],
                        target_q=[x.copy() for x in tr.target_q.p], target_p=[x.copy() for x in tr.target_p.p]))
    return out


# ------------------------------------------------------------------------------------------------
def philox_uniform(seed, counter, tag, row0, nrows, ncols):
    """The U[0,1) draws of the device kernels' Gumbel noise, on the host (include/maddpg_b200.h: mdp_philox_uniform)."""
    import ctypes as C
    from maddpg_b200 import _lib
    out = np.empty((nrows, ncols), np.float32)
    _lib.check(_lib.lib.mdp_philox_uniform(int(seed), int(counter), int(tag), int(row0), int(nrows), int(ncols),
                                           out.ctypes.data_as(C.c_void_p)), "mdp_philox_uniform")
    return out


def oracle_free_rollout(scenario, num_agents, state, weights, seed, counter0, T):
    """Free-running CPU replay of a device rollout: oracle MPE (float64) stepped with the oracle actor
    (numpy float32 mlp_model + SoftCategoricalPd.sample) on the device's own Philox uniforms.
    state: dict(agent_pos, agent_vel, landmark_pos) (E, ., 2) float64; weights[i] = [W1, b1, W2, b2, W3, b3].
    Returns obs (T + 1, E, sum D) float64, act (T, E, sum K) float32, rew (T, E, A) float64."""
    E = state["agent_pos"].shape[0]
    env = ompe.BatchedOracleEnv(scenario, E, num_agents, seed=0)
    env.set_state(state["agent_pos"], state["agent_vel"], state["landmark_pos"], None)
    nets = []
    for i, w in enumerate(weights):
        m = omaddpg.MLP(w[0].shape[0], w[0].shape[1], w[4].shape[1], np.random.RandomState(0))
        m.p = [np.asarray(x, np.float32) for x in w]
        nets.append(m)
    heads = [omaddpg.act_heads(s) for s in env.action_space]
    obs_n = env.observe()
    obs_t, act_t, rew_t = [np.concatenate(obs_n, axis=1)], [], []
    for t in range(T):
        acts = []
        for i, m in enumerate(nets):
            logits, _ = m.forward(obs_n[i].astype(np.float32))  # BatchInput feed cast (tf_util.py:98-112)
            u = philox_uniform(seed, counter0 + t + 1, i, 0, E, env.act_dims[i])
            acts.append(omaddpg.gumbel_softmax(logits, u, heads[i]))
        obs_n, r, d = env.step(acts)
        assert not d.any()
        obs_t.append(np.concatenate(obs_n, axis=1))
        act_t.append(np.concatenate(acts, axis=1))
        rew_t.append(r)
    return dict(obs=np.stack(obs_t), act=np.stack(act_t), rew=np.stack(rew_t))
