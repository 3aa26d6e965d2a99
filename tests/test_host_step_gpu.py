"""mdp_host_step (the host-buffer loop body, experiments/train.py:112-120 in one C-ABI call) against the
per-call device API on the same seeds: sampled actions, next observations, rewards, done flags and the
inserted replay rows must be bit-identical (same kernels, same Philox counters), including across an
episode reset and a ring wrap-around."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(scenario, E, capacity):
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGAgentTrainer
    env = BatchedMultiAgentEnv(scenario, num_envs=E, seed=3, squeeze=False)
    arglist = argparse.Namespace(lr=1e-2, gamma=0.95, batch_size=64, num_units=64, max_episode_len=25, seed=7,
                                 replay_capacity=capacity)
    obs_shape_n = [env.observation_space[i].shape for i in range(env.n)]
    trainers = [MADDPGAgentTrainer("agent_%d" % i, None, obs_shape_n, env.action_space, i, arglist) for i in range(env.n)]
    return env, trainers[0].core


@pytest.mark.parametrize("scenario,E,chunks,graph,copy_kernels", [
    ("simple_spread", 300, 1, False, False), ("simple_spread", 300, 1, True, False), ("simple_spread", 300, 4, False, False),
    ("simple_spread", 300, 4, True, False), ("simple_spread", 300, 1, False, True), ("simple_spread", 300, 4, True, True),
    ("simple_spread", 2048, None, True, True),
    ("simple_tag", 64, 2, True, True), ("simple_world_comm", 33, 3, True, False), ("simple", 1, 1, True, True),
    ("simple_speaker_listener", 66, 2, True, True), ("simple_reference", 39, 3, True, False), ("simple_crypto", 33, 1, False, True),
    ("simple_push", 96, 4, True, True), ("simple_adversary", 50, 2, False, False)])
def test_host_step_equals_device_api(scenario, E, chunks, graph, copy_kernels):
    """chunks > 1: the pipelined call (ranges of env instances on separate streams); graph: the call replayed as a CUDA
    graph with device-side counters; copy_kernels: host buffers moved by copy kernels instead of the copy engines.
    Every variant must reproduce the per-call device API bit for bit."""
    from maddpg_b200.rollout import HostRollout
    cap = 5 * E + 7  # forces a wrap-around inside the test
    env_a, core_a = _build(scenario, E, cap)
    env_b, core_b = _build(scenario, E, cap)
    core_b.params.copy_(core_a.params)
    host = HostRollout(env_a, core_a, chunks=chunks, use_graph=graph, copy_kernels=copy_kernels)
    assert host.chunks == (chunks or 8)
    obs_n = host.reset()
    env_b.reset_device()
    np.testing.assert_array_equal(np.concatenate(obs_n, 1), env_b.obs[:, :sum(env_b.obs_dims)].cpu().numpy())
    for t in range(9):
        if t == 4:  # episode boundary
            obs_n = host.reset()
            env_b.reset_device()
        act_n, new_obs_n, rew_n, done_n = host.step(obs_n)
        core_b.act(env_b.obs, env_b.act)
        env_b.step_device(ring=core_b.ring)
        for i in range(env_b.n):
            ao, K = env_b.act_off[i], env_b.act_dims[i]
            oo, D = env_b.obs_off[i], env_b.obs_dims[i]
            np.testing.assert_array_equal(act_n[i], env_b.act[:, ao:ao + K].cpu().numpy(), err_msg="act t=%d agent %d" % (t, i))
            np.testing.assert_array_equal(new_obs_n[i], env_b.obs[:, oo:oo + D].cpu().numpy(), err_msg="obs t=%d agent %d" % (t, i))
            np.testing.assert_array_equal(rew_n[i], env_b.rew[:, i].cpu().numpy(), err_msg="rew t=%d agent %d" % (t, i))
            assert done_n[i].dtype == np.bool_ and not done_n[i].any()
        obs_n = new_obs_n
    assert core_a.ring.next_idx == core_b.ring.next_idx and core_a.ring.length == core_b.ring.length
    assert core_a.counter == core_b.counter
    assert (host.graph_launches > 0) == graph
    n = core_a.ring.length[0]
    for i in range(env_b.n):  # every field of every agent (the padding columns of a row are never written)
        for name, cols in core_a.ring.cols(i).items():
            sl = slice(*cols) if isinstance(cols, tuple) else slice(cols, cols + 1)
            torch.testing.assert_close(core_a.ring.ring[:n, sl], core_b.ring.ring[:n, sl], rtol=0, atol=0, msg="%s agent %d" % (name, i))


def test_host_step_accepts_foreign_arrays():
    """Observations that are not the object's own page-locked views go through the staging copy."""
    from maddpg_b200.rollout import HostRollout
    env_a, core_a = _build("simple_spread", 50, 1000)
    env_b, core_b = _build("simple_spread", 50, 1000)
    core_b.params.copy_(core_a.params)
    ha, hb = HostRollout(env_a, core_a), HostRollout(env_b, core_b)
    oa, ob = ha.reset(), hb.reset()
    ra = ha.step(oa)
    rb = hb.step([np.array(o, dtype=np.float64) for o in ob])  # float64 copies, like the reference's numpy obs
    for x, y in zip(ra, rb):
        for u, v in zip(x, y):
            np.testing.assert_array_equal(u, v)
