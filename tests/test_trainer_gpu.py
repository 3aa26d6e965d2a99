"""CUDA trainer kernels (through the C ABI and the reference-shaped MADDPGAgentTrainer surface)
against the numpy oracle on the same weights, replay rows, index sets and uniform draws.
Tolerance from BASELINE.json north_star: Q-values and losses within 1e-4 relative after a
fixed-seed update."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.helpers import TRAINER_CASES, oracle_update_round, trainer_case

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _build(case):
    """MADDPGAgentTrainer objects built exactly like get_trainers (train.py:63-75), then loaded with
    the oracle's weights and replay rows."""
    from maddpg_b200 import MADDPGAgentTrainer, _lib
    env = case["env"]
    act_space_n = list(env.action_space)
    trainers = [MADDPGAgentTrainer("agent_%d" % i, None, case["obs_shape_n"], act_space_n, i, case["args"],
                                   local_q_func=case["local_q"][i]) for i in range(case["n"])]
    core = trainers[0].core
    for i, o in enumerate(case["trainers"]):
        core.set_weights(i, _lib.NET_P, o.p.p)
        core.set_weights(i, _lib.NET_TARGET_P, o.target_p.p)
        core.set_weights(i, _lib.NET_Q, o.q.p)
        core.set_weights(i, _lib.NET_TARGET_Q, o.target_q.p)
    p = case["pool"]
    if case["rows"] > 1000:  # BASELINE batch sizes: all rows of an agent through ONE batched experience() call
        dev = lambda x: torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        for i, tr in enumerate(trainers):
            tr.experience(dev(p["obs"][i]), dev(p["act"][i]), dev(p["rew"][i]), dev(p["nobs"][i]), dev(p["done"][i]), False)
    else:
        for r in range(case["rows"]):  # every agent inserts every step, like train.py:119-120
            for i, tr in enumerate(trainers):
                tr.experience(p["obs"][i][r], p["act"][i][r], float(p["rew"][i][r]), p["nobs"][i][r], bool(p["done"][i][r]), False)
    for tr in trainers:
        tr.max_replay_buffer_len = 0
    return trainers, core


def _close(a, b, rtol=RTOL, atol=1e-6, msg=""):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol, err_msg=msg)


@pytest.mark.parametrize("name", list(TRAINER_CASES))
def test_forward_surfaces_match_oracle(name):
    case = trainer_case(name, seed=1)
    trainers, core = _build(case)
    n, B = case["n"], case["B"]
    p = case["pool"]
    rows = np.arange(B)
    obs_n = [p["obs"][i][rows] for i in range(n)]
    act_n = [p["act"][i][rows] for i in range(n)]
    off = np.concatenate([[0], np.cumsum(case["act_dims"])]).astype(int)
    for j, (tr, o) in enumerate(zip(trainers, case["trainers"])):
        u = case["u_actor"][j]
        o.noise = lambda shape, u=u: u
        _close(tr.p_debug["p_values"](obs_n[j]), o.p_values(obs_n[j]), msg="p_values %d" % j)
        _close(tr.act(obs_n[j], u=u), o.act(obs_n[j]), msg="act %d" % j)
        _close(tr.p_debug["target_act"](obs_n[j], u=u), o.target_act(obs_n[j]), msg="target_act %d" % j)
        _close(tr.q_debug["q_values"](*(obs_n + act_n)), o.q_values(*(obs_n + act_n)), atol=2e-6, msg="q %d" % j)
        _close(tr.q_debug["target_q_values"](*(obs_n + act_n)), o.target_q_values(*(obs_n + act_n)), atol=2e-6, msg="tq %d" % j)
        a1 = tr.action(obs_n[j][0])  # reference shape: (D,) in -> (K,) float32 out, rows sum to one per head
        assert a1.shape == (case["act_dims"][j],) and a1.dtype == np.float32
        assert abs(float(a1[:case["heads"][j][0]].sum()) - 1.0) < 1e-5
    # ring rows agree bit-exactly with what was inserted (per-agent sample_index surface)
    idx = case["idx"][0]
    for i, tr in enumerate(trainers):
        o_, a_, r_, n_, d_ = tr.replay_buffer.sample_index(idx)
        assert np.array_equal(o_, p["obs"][i][idx]) and np.array_equal(a_, p["act"][i][idx])
        assert np.array_equal(r_, p["rew"][i][idx]) and np.array_equal(n_, p["nobs"][i][idx])
        assert np.array_equal(d_, p["done"][i][idx])


@pytest.mark.parametrize("name", list(TRAINER_CASES))
def test_gradients_match_oracle(name):
    """td_target / critic_grads / actor_grads entry points, agent 0 from the pre-update weights."""
    case = trainer_case(name, seed=2)
    ref = oracle_update_round(trainer_case(name, seed=2))
    trainers, core = _build(case)
    j = 0
    idx = core.ring.index_tensor(case["idx"][j])
    batch = core.ring.gather(idx)
    B = case["B"]
    ut = torch.zeros((B, core.act_stride), device="cuda")
    ut[:, :core.act_sum] = torch.from_numpy(case["u_target"][j]).cuda()
    y = core.td_target(j, batch, ut)
    _close(y.cpu().numpy(), ref[j]["y"], atol=2e-6, msg="td target")
    q = core.critic_grads(j, batch, y, want_q=True)
    got = [g.cpu().numpy() for g in core.train_view(core.grads, j, 1)]
    for k, (g, r) in enumerate(zip(got, ref[j]["q_grads"])):
        _close(g, r, rtol=1e-3, atol=1e-6 + 1e-4 * np.abs(r).max(), msg="critic grad %d" % k)
    core.clip_adam_polyak(j, 1)
    assert float(core.grads.abs().max()) == 0.0  # bucket re-zeroed
    for k, (w, r) in enumerate(zip(core.get_weights(j, 2), ref[j]["q"])):
        _close(w, r, rtol=1e-3, atol=2e-4, msg="critic param %d after Adam" % k)
    ua = torch.zeros((B, core.act_stride), device="cuda")
    o = core.act_off[j]
    ua[:, o:o + core.act_dims[j]] = torch.from_numpy(case["u_actor"][j]).cuda()
    core.actor_grads(j, batch, ua)
    got = [g.cpu().numpy() for g in core.train_view(core.grads, j, 0)]
    for k, (g, r) in enumerate(zip(got, ref[j]["p_grads"])):
        _close(g, r, rtol=2e-3, atol=1e-7 + 2e-4 * np.abs(r).max(), msg="actor grad %d" % k)
    assert core.adam_t.cpu().tolist()[:2] == [1, 1]


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", list(TRAINER_CASES))
def test_sequential_update_round_matches_oracle(name, fused):
    """agent.update(trainers, t) for every agent in order (train.py:160-161), injected index sets and
    uniforms: returned statistics, post-update Q-values and parameters.  ``fused``: TD target and critic
    step in one launch (the default) or two (mdp_core_set_fused_update)."""
    case = trainer_case(name, seed=4)
    ref = oracle_update_round(trainer_case(name, seed=4))
    oracle_after = trainer_case(name, seed=4)
    trainers, core = _build(case)
    core.set_fused_update(fused)
    n = case["n"]
    assert trainers[0].update(trainers, 99, index=case["idx"][0]) is None  # off-period gate (maddpg.py:164)
    for j, tr in enumerate(trainers):
        tr.preupdate()
        tr.inject_noise(u_target=case["u_target"][j], u_actor=case["u_actor"][j])
        stats = tr.update(trainers, 100, index=case["idx"][j])
        rs = ref[j]["stats"]
        assert len(stats) == 6
        for k, nm in enumerate(["q_loss", "p_loss", "mean_target_q", "mean_rew", "mean_target_q_next", "std_target_q"]):
            _close(stats[k], rs[k], rtol=RTOL, atol=2e-6, msg="%s agent %d" % (nm, j))
    # post-update networks: parameters and Q-values on a fresh batch
    p = case["pool"]
    rows = np.arange(case["B"], 2 * case["B"])
    obs_n = [p["obs"][i][rows] for i in range(n)]
    act_n = [p["act"][i][rows] for i in range(n)]
    from oracle import maddpg as om
    for j, tr in enumerate(trainers):
        for net, key in ((2, "q"), (3, "target_q"), (0, "p"), (1, "target_p")):
            for k, (w, r) in enumerate(zip(core.get_weights(j, net), ref[j][key])):
                _close(w, r, rtol=1e-3, atol=2e-4, msg="agent %d %s[%d]" % (j, key, k))
        o = oracle_after["trainers"][j]
        o.q.p, o.p.p = ref[j]["q"], ref[j]["p"]
        q_ref = o.q_values(*(obs_n + act_n))
        q_got = tr.q_debug["q_values"](*(obs_n + act_n))
        scale = np.abs(q_ref).mean()
        _close(q_got, q_ref, rtol=RTOL, atol=RTOL * scale, msg="post-update Q agent %d" % j)


def test_warmup_gate_and_index_stream():
    """update() returns None until batch_size*max_episode_len rows exist (maddpg.py:162-163) and draws
    its indices from python ``random`` like replay_buffer.py:46-47."""
    import random
    case = trainer_case("simple", seed=0)
    case["args"].python_index_stream = True  # the reference's python `random` stream (default: device-side Philox draws)
    trainers, core = _build(case)
    tr = trainers[0]
    tr.max_replay_buffer_len = case["rows"] + 1
    assert tr.update(trainers, 100) is None
    tr.max_replay_buffer_len = case["rows"]
    random.seed(77)
    expect = [random.randint(0, case["rows"] - 1) for _ in range(case["B"])]
    random.seed(77)
    out = tr.update(trainers, 100)
    assert out is not None and tr.replay_sample_index == expect
    assert len(out) == 6 and all(np.isfinite(out))
    # default: ReplayBuffer.make_index on the device -- in range, fresh per update, no host round trip; the statistics are
    # materialised lazily (train.py:161 discards them)
    case["args"].python_index_stream = False
    a = tr.update(trainers, 100)
    i1 = tr.replay_sample_index.clone()
    b = tr.update(trainers, 100)
    i2 = tr.replay_sample_index
    assert i1.is_cuda and i1.shape == (case["B"],) and int(i1.min()) >= 0 and int(i1.max()) < case["rows"]
    assert not torch.equal(i1, i2) and i1.unique().numel() > case["B"] // 2
    assert a._v is None and np.isfinite(np.asarray(a, np.float64)).all() and np.isfinite(list(b)).all()


def test_polyak_tau_invariants():
    """Reference invariant (tests/test_policy.py:71-86): polyak 0 => target == running; 1 => unchanged."""
    from maddpg_b200 import MADDPGCore
    from maddpg_b200.spaces import Discrete
    for pol in (0.0, 1.0):
        core = MADDPGCore([6, 6], [Discrete(5), Discrete(5)], [False, False], polyak=pol, replay_capacity=64, seed=1)
        before_t = core.get_weights(0, 3)
        core.grads.normal_()
        core.adam_t += 1
        core.clip_adam_polyak(0, 1)
        run, tgt = core.get_weights(0, 2), core.get_weights(0, 3)
        for r, t, b in zip(run, tgt, before_t):
            assert np.array_equal(t, r) if pol == 0.0 else np.array_equal(t, b)


def test_philox_actions_are_valid_and_reproducible():
    from maddpg_b200 import MADDPGCore
    from maddpg_b200.spaces import Discrete, MultiDiscrete
    spaces = [MultiDiscrete([[0, 4], [0, 3]]), Discrete(5)]
    outs = []
    for rep in range(2):
        core = MADDPGCore([34, 28], spaces, [False, False], num_units=128, replay_capacity=64, seed=9)
        obs = torch.randn(4096, core.obs_stride, generator=torch.Generator().manual_seed(0)).cuda()
        act = torch.zeros(4096, core.act_stride, device="cuda")
        core.act(obs, act)
        a = act.cpu().numpy()
        outs.append(a)
        assert np.allclose(a[:, 0:5].sum(1), 1, atol=1e-5) and np.allclose(a[:, 5:9].sum(1), 1, atol=1e-5)
        assert np.allclose(a[:, 9:14].sum(1), 1, atol=1e-5) and np.all(a >= 0)
        act2 = torch.zeros_like(act)
        core.act(obs, act2)  # the Philox counter advanced: a fresh sample
        assert not torch.equal(act, act2)
    assert np.array_equal(outs[0], outs[1])  # same seed, same counter -> same draws
    # argmax frequencies follow the Gumbel-max law: compare with softmax(logits) means
    logits = torch.zeros_like(act)
    core2 = MADDPGCore([34, 28], spaces, [False, False], num_units=128, replay_capacity=64, seed=9)
    core2.act(obs, act, logits_out=logits)
    pl = torch.softmax(logits[:, 9:14], 1).mean(0).cpu().numpy()
    freq = np.bincount(act[:, 9:14].argmax(1).cpu().numpy(), minlength=5) / 4096.0
    assert np.abs(freq - pl).max() < 0.04


def test_graph_rollout_and_update_round_match_eager_semantics():
    """CUDA-graph episode/update-round replay: ring cursor, length, Philox counter and episode id advance
    on device exactly as the eager path advances the host mirrors; noise is fresh on every replay."""
    import argparse
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
    from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound
    E, T = 256, 25
    env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False, seed=5)
    core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=E * T * 5, seed=5)
    roll = BatchedRollout(env, core, T, mode="graph")
    env.reset_device()
    roll.run(T * 4)  # 1 eager warm-up episode + 3 replays
    assert roll.graph_ok and core.ring.length == [E * T * 4] * 3 and roll.graph_launches == 3 * roll.launches_per_graph
    ctl = roll.ctl.t.cpu().tolist()
    assert ctl[0] == core.counter and ctl[1] == core.ring.next_idx[0] and ctl[2] == env.episode and ctl[3] == core.ring.length[0]
    L = core.ring.layout
    ring = core.ring.ring.cpu()
    n = core.ring.length[0]
    # every inserted row is a valid transition: actions are per-agent simplices, obs finite, done == 0
    act = ring[:n, L.obs_sum:L.x_dim].view(n, 3, 5)
    assert torch.allclose(act.sum(-1), torch.ones(n, 3), atol=1e-5)
    assert torch.isfinite(ring[:n, :L.x_dim]).all() and (ring[:n, L.dn_off:L.dn_off + 3] == 0).all()
    # consecutive steps of the same env chain: next_obs of step s == obs of step s+1 inside an episode
    for ep in range(4):
        base = ep * T * E
        for s in (0, 7, 23):
            a = ring[base + s * E: base + (s + 1) * E, L.nx_off:L.nx_off + L.obs_sum]
            b = ring[base + (s + 1) * E: base + (s + 2) * E, :L.obs_sum]
            assert torch.equal(a, b), (ep, s)
    # replays draw fresh noise and fresh reset positions
    a1 = ring[1 * T * E: 1 * T * E + E, L.obs_sum:L.x_dim]
    a2 = ring[2 * T * E: 2 * T * E + E, L.obs_sum:L.x_dim]
    assert not torch.equal(a1, a2)
    o1, o2 = ring[1 * T * E: 1 * T * E + E, :4], ring[2 * T * E: 2 * T * E + E, :4]
    assert not torch.equal(o1, o2)
    # graphed update rounds: Adam counters advance by one per round per net, parameters move, stay finite
    upd = GraphedUpdateRound(core, 128, ctl=roll.ctl, use_graph=True)
    p0 = core.params.clone()
    upd.run(3)
    torch.cuda.synchronize()
    assert core.adam_t.cpu().tolist() == [3] * 6
    assert torch.isfinite(core.params).all() and not torch.equal(p0, core.params)
    idx = upd.idx[0].cpu()
    assert idx.min() >= 0 and idx.max() < core.ring.length[0] and idx.unique().numel() > 100
    p1 = core.params.clone()
    upd.run(1)
    assert not torch.equal(p1, core.params)


@pytest.mark.parametrize("scenario,units,generic", [("simple_spread", 64, True), ("simple_spread", 64, False),
                                                    ("simple_tag", 64, True), ("simple", 64, True),
                                                    ("simple_world_comm", 128, True), ("simple_adversary", 64, True),
                                                    ("simple_push", 64, True), ("simple_speaker_listener", 64, True),
                                                    ("simple_crypto", 64, True), ("simple_reference", 64, True)])
def test_episode_kernel_matches_per_step_kernels(scenario, units, generic):
    """mdp_rollout_episode (persistent episode kernel, fp32 SIMT actor tiles) against the per-step path on the same
    seeds and Philox counters: identical replay rows, final state and observations (incl. the device reset)."""
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
    from maddpg_b200.rollout import BatchedRollout
    E, T = 80, 25  # 80 = 2.5 tiles of 32 env instances: exercises the ragged last CTA
    rings, finals = [], []
    for mode in ("eager", "mega"):
        env = BatchedMultiAgentEnv(scenario, num_envs=E, squeeze=False, seed=11)
        # generic: both paths run the table-driven physics; not generic (simple_spread): both run the register-resident
        # spread_step / spread_obs.  Register vs table code is compared step by step in test_env_gpu.py (last-bit
        # differences grow over 75 free steps).
        env.force_generic_kernel(generic)
        core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=units,
                          replay_capacity=E * T * 2 + 13, seed=3)
        core.set_tensor_cores(-1)  # the fp32 SIMT episode kernel; the tcgen05 one is covered in tests/test_rollout_tc_gpu.py
        roll = BatchedRollout(env, core, T, mode=mode)
        if mode == "mega":
            roll.ep_return = torch.zeros((E, env.n), device="cuda")
        env.reset_device()
        roll.run(T * 3)  # 3 episodes into a ring of 2 episodes + 13 rows: wraps
        torch.cuda.synchronize()
        assert roll.mode == mode and core.ring.length == [core.ring.capacity] * env.n
        assert core.counter == 3 * T and env.episode == 4
        rings.append((core.ring.ring.cpu(), list(core.ring.next_idx)))
        finals.append((env.state.cpu(), env.obs.cpu(), roll.ep_return))
    (r0, n0), (r1, n1) = rings
    assert n0 == n1
    L = core.ring.layout
    used = list(range(0, L.x_dim)) + list(range(L.nx_off, L.nx_off + L.obs_sum)) + list(range(L.rw_off, L.dn_off + env.n))
    torch.testing.assert_close(r1[:, used], r0[:, used], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(finals[1][0], finals[0][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(finals[1][1], finals[0][1], rtol=1e-5, atol=1e-6)
    # episode returns accumulated by the kernel == sum of the rewards it wrote (last 2 episodes are in the ring)
    ret = finals[1][2].cpu()
    assert torch.isfinite(ret).all() and ret.abs().sum() > 0


@pytest.mark.parametrize("num_envs", [1, 64])
def test_train_loop_drop_in(tmp_path, num_envs):
    """The reference's OWN experiments/train.py, unmodified, executed through maddpg_b200.train's stand-in modules
    (tensorflow / maddpg.common.tf_util / maddpg.trainer.maddpg / multiagent.*): reference shapes (1 env, numpy) and the
    batched superset; passes the warm-up gate, performs updates at t % 100 == 0, saves and restores parameters.
    The file is read from /root/reference (build container) or baseline/_ref/experiments (the copy __graft_entry__.build()
    leaves for the GPU box; git-ignored, never part of the repo's history)."""
    import pickle
    from maddpg_b200 import train as T
    try:
        path = T.reference_train_path()
    except FileNotFoundError:
        pytest.skip("the reference's experiments/train.py is not available on this machine")
    argv = ["--scenario", "simple_tag", "--num-adversaries", "3", "--adv-policy", "ddpg", "--num-episodes", "12",
            "--batch-size", "8", "--save-rate", "4", "--save-dir", str(tmp_path) + "/", "--plots-dir", str(tmp_path) + "/",
            "--exp-name", "t"]
    trainers = T.run_reference_train(argv, path, num_envs=num_envs, replay_capacity=50000, seed=0)
    assert len(trainers) == 4 and "maddpg.trainer.maddpg" not in sys.modules  # the stand-ins are gone again
    core = trainers[0].core
    ep_rewards = pickle.load(open(os.path.join(str(tmp_path), "t_rewards.pkl"), "rb"))
    assert len(ep_rewards) == 3 and all(np.isfinite(ep_rewards))  # mean episode reward at episodes 4, 8, 12 (train.py:172-176)
    assert core.local_q == [True, True, True, False]  # get_trainers: adversaries first, ddpg -> local_q_func (train.py:63-75)
    assert len(trainers[0].replay_buffer) == 300 * num_envs
    # warm-up gate needs 200 rows: reached at t = 200 with one env (updates at t = 200, 300), at t = 4 with 64
    assert core.adam_t.cpu().tolist() == [2 if num_envs == 1 else 3] * 8
    assert os.path.exists(os.path.join(str(tmp_path), "maddpg_b200.pt")) and os.path.exists(os.path.join(str(tmp_path), "t_agrewards.pkl"))
    saved = torch.load(os.path.join(str(tmp_path), "maddpg_b200.pt"))
    assert saved["counter"] > 0
    core.params.zero_()
    core.counter = 0
    T.load_state(str(tmp_path), trainers)
    assert torch.equal(core.params.cpu(), saved["params"]) and core.counter == saved["counter"]


@pytest.mark.parametrize("scenario,num_envs", [("simple_speaker_listener", 1), ("simple_crypto", 32), ("simple_push", 32),
                                               ("simple_adversary", 1), ("simple_reference", 16)])
def test_train_loop_drop_in_goal_scenarios(tmp_path, scenario, num_envs):
    """the reference's own experiments/train.py on the other MPE scenarios (Discrete(3) / Discrete(4)
    communication heads, immovable agents, goal landmarks drawn by reset_world), reference shapes and the batched superset."""
    import pickle
    from maddpg_b200 import train as T
    try:
        path = T.reference_train_path()
    except FileNotFoundError:
        pytest.skip("the reference's experiments/train.py is not available on this machine")
    argv = ["--scenario", scenario, "--num-adversaries", "1", "--num-episodes", "12", "--batch-size", "8", "--save-rate", "4",
            "--save-dir", str(tmp_path) + "/", "--plots-dir", str(tmp_path) + "/", "--exp-name", "g"]
    trainers = T.run_reference_train(argv, path, num_envs=num_envs, replay_capacity=20000, seed=1)
    core = trainers[0].core
    ep_rewards = pickle.load(open(os.path.join(str(tmp_path), "g_rewards.pkl"), "rb"))
    assert len(ep_rewards) == 3 and all(np.isfinite(ep_rewards))
    assert len(trainers[0].replay_buffer) == 300 * num_envs
    assert min(core.adam_t.cpu().tolist()) >= 2 and torch.isfinite(core.params).all()


def test_grouped_update_all_equals_jacobi_order_of_per_agent_kernels():
    """mdp_update_all (all agents per launch) == the per-agent entry points called in Jacobi order with the
    same index sets and Philox counter; and it differs from the sequential order only slightly."""
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
    from maddpg_b200.rollout import BatchedRollout
    E, B = 512, 256
    cores = []
    for _ in range(3):
        env = BatchedMultiAgentEnv("simple_tag", num_envs=E, squeeze=False, seed=2)
        core = MADDPGCore(env.obs_dims, env.action_space, [False] * 4, replay_capacity=E * 30, seed=4)
        core.ring.ring.zero_()  # padding columns are never written: make whole-row comparisons meaningful
        roll = BatchedRollout(env, core, 25, mode="mega")
        env.reset_device()
        roll.run(25)
        cores.append(core)
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, E * 25, (4, B), generator=g).cuda()
    a, b, c = cores
    assert torch.equal(a.params, b.params) and torch.equal(a.ring.ring[:E * 25], b.ring.ring[:E * 25])
    a.update_all(a.ring.ring, idx=idx, counter=77)
    # the per-agent kernels draw their own counters; replay the Jacobi order with the grouped call's counter instead
    b2 = cores[2]
    b2.counter = 76
    ys = []
    for j in range(4):
        b2.counter = 76
        ys.append(b2.td_target(j, b2.ring.ring, idx=idx[j]).clone())
    for j in range(4):
        b2.critic_grads(j, b2.ring.ring, ys[j], idx=idx[j])
    for j in range(4):
        b2.clip_adam_polyak(j, 1)
    for j in range(4):
        b2.counter = 76
        b2.actor_grads(j, b2.ring.ring, idx=idx[j])
    for j in range(4):
        b2.clip_adam_polyak(j, 0)
    torch.cuda.synchronize()
    assert a.adam_t.cpu().tolist() == [1] * 8 and b2.adam_t.cpu().tolist() == [1] * 8
    torch.testing.assert_close(a.params, b2.params, rtol=1e-4, atol=2e-5)
    assert float(a.grads.abs().max()) == 0.0
