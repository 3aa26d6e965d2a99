"""CPU oracle: the reference's rollout/update loop (experiments/train.py:110-161) on the numpy
restatements of MPE (oracle/mpe.py), the trainer (oracle/maddpg.py) and the replay list
(oracle/replay.py).  "Restated reference path (TF-free)": the real train.py needs TensorFlow 1.8,
gym and MPE, none of which can be installed here.

TEST INFRASTRUCTURE ONLY -- used as the timed CPU baseline by bench.py (``cpu_baseline`` and
``--impl reference``), never by the product path.

PINNED (the loop, not the TF graph): the REAL experiments/train.py was executed unmodified in the
build container on these oracle classes -- its own ``train(arglist)`` loop, the REAL
MADDPGAgentTrainer.action / experience / preupdate / update methods and the REAL ReplayBuffer, with
oracle/mpe.py behind ``multiagent.*`` and oracle/maddpg.py's graph callables where TensorFlow would
be (tests/golden/make_train_loop_golden.py -> train_loop_ref.npz) -- and ``run_training`` below
reproduces the learning-curve lists it pickled bit for bit
(tests/test_oracle_golden.py::test_train_loop_matches_the_reference_script).  ``time_rollout`` and
``time_updates`` are the same iteration without the reward bookkeeping.
"""
import argparse
import time

import numpy as np

from oracle import maddpg as omaddpg
from oracle import mpe as ompe


def make_arglist(scenario="simple_spread", batch_size=1024, num_units=64, max_episode_len=25, lr=1e-2, gamma=0.95):
    # experiments/train.py:11-37 defaults
    return argparse.Namespace(scenario=scenario, max_episode_len=max_episode_len, lr=lr, gamma=gamma,
                              batch_size=batch_size, num_units=num_units, num_adversaries=0,
                              good_policy="maddpg", adv_policy="maddpg")


def get_trainers(env, num_adversaries, obs_shape_n, arglist, seed=0, noise=None):
    # experiments/train.py:63-75
    trainers = []
    for i in range(env.n):
        policy = arglist.adv_policy if i < num_adversaries else arglist.good_policy
        trainers.append(omaddpg.OracleAgentTrainer("agent_%d" % i, None, obs_shape_n, env.action_space, i, arglist,
                                                   local_q_func=(policy == "ddpg"),
                                                   rng=np.random.RandomState(seed * 1000 + i), noise=noise))
    return trainers


def time_rollout(scenario, num_agents=None, steps=2000, seed=0, arglist=None):
    """Times ``steps`` iterations of train.py:110-136 while the update is still gated off by the
    warm-up (maddpg.py:162-163): n batch-1 actor calls, one python MPE step, n replay appends, reset
    every max_episode_len steps.  Returns (agent_env_steps_per_sec, env_steps_per_sec, seconds)."""
    arglist = arglist or make_arglist(scenario)
    env = ompe.make_env(scenario, np.random.RandomState(seed), num_agents)
    obs_shape_n = [env.observation_space[i].shape for i in range(env.n)]
    trainers = get_trainers(env, 0, obs_shape_n, arglist, seed)
    obs_n = env.reset()
    episode_step = 0
    train_step = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        action_n = [agent.action(obs) for agent, obs in zip(trainers, obs_n)]
        new_obs_n, rew_n, done_n, info_n = env.step(action_n)
        episode_step += 1
        done = all(done_n)
        terminal = (episode_step >= arglist.max_episode_len)
        for i, agent in enumerate(trainers):
            agent.experience(obs_n[i], action_n[i], rew_n[i], new_obs_n[i], done_n[i], terminal)
        obs_n = new_obs_n
        if done or terminal:
            obs_n = env.reset()
            episode_step = 0
        train_step += 1
        for agent in trainers:
            agent.preupdate()
        for agent in trainers:
            agent.update(trainers, train_step)  # returns at the warm-up gate
    dt = time.perf_counter() - t0
    return steps * env.n / dt, steps / dt, dt


def run_training(scenario, num_episodes, arglist, seed=0, save_rate=1000, num_agents=None, noise=None):
    """The whole loop of ``train(arglist)`` (experiments/train.py:76-197, training mode: no display / benchmark / restore) with
    its bookkeeping: -> (final_ep_rewards, final_ep_ag_rewards, train_step), the two lists the reference pickles for its
    learning curves (:176-178, :181-187).  PINNED: tests/golden/train_loop_ref.npz holds what the REAL train.py computed when it
    was executed unmodified on these same oracle classes (tests/golden/make_train_loop_golden.py); the test
    test_oracle_golden.py::test_train_loop_matches_the_reference_script holds this function to it bit for bit."""
    env = ompe.make_env(scenario, np.random.RandomState(seed), num_agents)
    obs_shape_n = [env.observation_space[i].shape for i in range(env.n)]
    num_adversaries = min(env.n, arglist.num_adversaries)
    trainers = get_trainers(env, num_adversaries, obs_shape_n, arglist, seed, noise)   # noise: one shared U[0,1) stream, or per agent
    episode_rewards = [0.0]
    agent_rewards = [[0.0] for _ in range(env.n)]
    final_ep_rewards, final_ep_ag_rewards = [], []
    obs_n = env.reset()
    episode_step = 0
    train_step = 0
    while True:
        action_n = [agent.action(obs) for agent, obs in zip(trainers, obs_n)]
        new_obs_n, rew_n, done_n, info_n = env.step(action_n)
        episode_step += 1
        done = all(done_n)
        terminal = (episode_step >= arglist.max_episode_len)
        for i, agent in enumerate(trainers):
            agent.experience(obs_n[i], action_n[i], rew_n[i], new_obs_n[i], done_n[i], terminal)
        obs_n = new_obs_n
        for i, rew in enumerate(rew_n):
            episode_rewards[-1] += rew
            agent_rewards[i][-1] += rew
        if done or terminal:
            obs_n = env.reset()
            episode_step = 0
            episode_rewards.append(0)
            for a in agent_rewards:
                a.append(0)
        train_step += 1
        for agent in trainers:
            agent.preupdate()
        for agent in trainers:
            agent.update(trainers, train_step)
        if terminal and (len(episode_rewards) % save_rate == 0):
            final_ep_rewards.append(np.mean(episode_rewards[-save_rate:]))
            for rew in agent_rewards:
                final_ep_ag_rewards.append(np.mean(rew[-save_rate:]))
        if len(episode_rewards) > num_episodes:
            return final_ep_rewards, final_ep_ag_rewards, train_step


def time_updates(scenario, num_agents=None, rounds=5, seed=0, arglist=None, prefill=None):
    """Times forced back-to-back update rounds (every agent once, sequentially: train.py:160-161 ->
    maddpg.py:167-194) on a pre-filled buffer.  Returns (critic_updates_per_sec, seconds)."""
    arglist = arglist or make_arglist(scenario)
    env = ompe.make_env(scenario, np.random.RandomState(seed), num_agents)
    obs_shape_n = [env.observation_space[i].shape for i in range(env.n)]
    trainers = get_trainers(env, 0, obs_shape_n, arglist, seed)
    rng = np.random.RandomState(seed + 1)
    rows = prefill or arglist.batch_size * 4
    for i, tr in enumerate(trainers):
        D, K = obs_shape_n[i][0], tr.act_dims[i]
        obs = rng.randn(rows, D)
        nobs = rng.randn(rows, D)
        act = rng.dirichlet(np.ones(K), size=rows).astype(np.float32)
        rew = rng.randn(rows)
        for r in range(rows):
            tr.replay_buffer.add(obs[r], act[r], float(rew[r]), nobs[r], 0.0)
        tr.max_replay_buffer_len = rows
    t0 = time.perf_counter()
    for _ in range(rounds):
        for agent in trainers:
            agent.preupdate()
        for agent in trainers:
            out = agent.update(trainers, 100)
            assert out is not None
    dt = time.perf_counter() - t0
    return rounds * env.n / dt, dt


def _worker(args):
    kind, scenario, num_agents, amount, seed, batch, units = args
    import os
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    arglist = make_arglist(scenario, batch_size=batch, num_units=units)
    if kind == "rollout":
        a, e, dt = time_rollout(scenario, num_agents, amount, seed, arglist)
        return a, dt
    u, dt = time_updates(scenario, num_agents, amount, seed, arglist)
    return u, dt


def time_parallel(kind, scenario, num_agents, amount, procs, batch=1024, units=64):
    """``procs`` independent single-threaded copies of the reference loop (the reference itself is
    single-threaded, tf_util.py:202-204; replicas are the only way it can use more host cores).
    Returns (aggregate units/s, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(kind, scenario, num_agents, amount, s, batch, units) for s in range(procs)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall
