"""CUDA env kernel (through the C ABI) against the MPE oracle on the same injected states/actions.
Tolerance from BASELINE.json north_star: observations and rewards within 1e-5 relative after a
fixed-seed rollout (absolute floor 1e-5 for entries that pass through zero)."""
import numpy as np
import pytest
import torch

from tests.helpers import ENV_CASES, env_case, run_oracle_rollout, soft_actions

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-5


def _make(case, dtype):
    from maddpg_b200 import BatchedMultiAgentEnv
    env = BatchedMultiAgentEnv(case["scenario"], num_envs=case["E"], num_agents=case["num_agents"],
                               state_dtype=dtype, squeeze=False)
    assert env.obs_dims == case["env"].obs_dims and env.act_dims == case["env"].act_dims
    return env


def _joint_act(env, acts):
    a = torch.zeros((env.num_envs, env.act_stride), dtype=torch.float32)
    for i, x in enumerate(acts):
        a[:, env.act_off[i]:env.act_off[i] + env.act_dims[i]] = torch.from_numpy(x)
    return a.cuda()


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_rollout_f64_state_matches_oracle(name):
    case = env_case(name, seed=11)
    ref = run_oracle_rollout(case)
    env = _make(case, torch.float64)
    init = env.state_from_arrays(case["agent_pos"], case["agent_vel"], case["landmark_pos"], case["agent_c"], case.get("goal"))
    obs0 = torch.cat(env.reset(init_state=init), dim=1).cpu().numpy()
    np.testing.assert_allclose(obs0, ref["obs0"], rtol=RTOL, atol=ATOL)
    for t, acts in enumerate(case["tape"]):
        env.step_device(_joint_act(env, acts))
        obs = env.obs[:, :sum(env.obs_dims)].cpu().numpy()
        np.testing.assert_allclose(obs, ref["obs"][t], rtol=RTOL, atol=ATOL, err_msg="obs step %d" % t)
        np.testing.assert_allclose(env.rew.cpu().numpy(), ref["rew"][t], rtol=RTOL, atol=ATOL, err_msg="rew step %d" % t)
        assert int(env.done.sum()) == 0
    st = env.state_to_arrays()
    np.testing.assert_allclose(st["agent_pos"], ref["final"]["agent_pos"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(st["agent_vel"], ref["final"]["agent_vel"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_single_steps_f32_state_match_oracle(name):
    """float32 state (throughput mode): every step restarts from the oracle's float64 state, so the
    comparison isolates one step of float32 arithmetic."""
    case = env_case(name, seed=5)
    oenv = case["env"]
    oenv.set_state(case["agent_pos"], case["agent_vel"], case["landmark_pos"], case["agent_c"], case.get("goal"))
    env = _make(case, torch.float32)
    for t, acts in enumerate(case["tape"][:10]):
        st = oenv.get_state()
        init = env.state_from_arrays(st["agent_pos"], st["agent_vel"], st["landmark_pos"], st["agent_c"], st["goal"])
        env.reset(init_state=init)
        o, r, d = oenv.step(acts)
        env.step_device(_joint_act(env, acts))
        obs = env.obs[:, :sum(env.obs_dims)].cpu().numpy()
        np.testing.assert_allclose(obs, np.concatenate(o, 1), rtol=1e-4, atol=2e-5, err_msg="obs step %d" % t)
        # rewards jump by +-1/5/10 at contact thresholds: a float32 state could only differ there by flipping an exact tie,
        # which these seeded cases do not contain
        np.testing.assert_allclose(env.rew.cpu().numpy(), r, rtol=1e-4, atol=1e-4, err_msg="rew step %d" % t)


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_rollout_f32_state_free_running_matches_oracle(name):
    """The benchmarked precision mode (float32 state), FREE-RUNNING for the whole action tape (25 steps; 4 for
    simple_spread N = 24) from one injected state: observations and rewards against the float64 oracle at the north-star
    tolerance 1e-5 relative (|x| < 1 entries: 1e-5 absolute).  Measured drift on B200: observations <= 3e-6, rewards
    <= 5e-6 for the four scenarios; the 24-agent case reaches 2e-5 on a few relative positions inside its 276-pair
    contact cluster and gets 5e-5."""
    tol = 5e-5 if name == "simple_spread_24" else 1e-5
    for seed in (11, 5):
        case = env_case(name, seed=seed)
        ref = run_oracle_rollout(case)
        env = _make(case, torch.float32)
        env.reset(init_state=env.state_from_arrays(case["agent_pos"], case["agent_vel"], case["landmark_pos"], case["agent_c"], case.get("goal")))
        for t, acts in enumerate(case["tape"]):
            env.step_device(_joint_act(env, acts))
            obs = env.obs[:, :sum(env.obs_dims)].cpu().numpy()
            np.testing.assert_allclose(obs, ref["obs"][t], rtol=tol, atol=tol, err_msg="obs step %d" % t)
            np.testing.assert_allclose(env.rew.cpu().numpy(), ref["rew"][t], rtol=1e-5, atol=1e-5, err_msg="rew step %d" % t)


@pytest.mark.parametrize("name", ["simple_spread", "simple_tag", "simple_world_comm"])
def test_device_reset_ranges_and_obs(name):
    from maddpg_b200 import BatchedMultiAgentEnv
    from oracle import mpe
    scenario, na, _, _ = ENV_CASES[name]
    E = 64
    env = BatchedMultiAgentEnv(scenario, num_envs=E, num_agents=na, state_dtype=torch.float64, squeeze=False, seed=3)
    obs = torch.cat(env.reset(), 1).cpu().numpy()
    st = env.state_to_arrays()
    lo, hi = (-1.0, 1.0) if name == "simple_spread" else (-0.9, 0.9)
    assert st["agent_pos"].min() >= -1 and st["agent_pos"].max() < 1
    assert st["landmark_pos"].min() >= lo and st["landmark_pos"].max() < hi
    assert np.all(st["agent_vel"] == 0)
    assert st["agent_pos"].std() > 0.4  # actually random
    oenv = mpe.BatchedOracleEnv(scenario, E, na)
    oenv.set_state(st["agent_pos"], st["agent_vel"], st["landmark_pos"])
    np.testing.assert_allclose(obs, np.concatenate(oenv.observe(), 1), rtol=RTOL, atol=ATOL)
    obs2 = torch.cat(env.reset(), 1).cpu().numpy()  # next episode draws new positions
    assert not np.allclose(obs, obs2)


def test_full_size_properties_spread_4096():
    """BASELINE config 2 size: translation equivariance of observations, shared reward equality,
    run-to-run determinism."""
    from maddpg_b200 import BatchedMultiAgentEnv
    E = 4096
    rng = np.random.RandomState(0)
    # float64 state: a float32 shift re-rounds positions and the stiff contact model (k = 1e-3)
    # amplifies that by ~100x per step in contact, which is not what this property is about
    env = BatchedMultiAgentEnv("simple_spread", num_envs=E, state_dtype=torch.float64, squeeze=False)
    ap, av, lp = rng.uniform(-1, 1, (E, 3, 2)), rng.uniform(-.2, .2, (E, 3, 2)), rng.uniform(-1, 1, (E, 3, 2))
    acts = [soft_actions(rng, E, 5) for _ in range(3)]
    a = _joint_act(env, acts)

    def run(shift):
        env.reset(init_state=env.state_from_arrays(ap + shift, av, lp + shift))
        env.step_device(a)
        return env.obs.cpu().numpy().copy(), env.rew.cpu().numpy().copy()

    o0, r0 = run(0.0)
    o1, r1 = run(0.0)
    assert np.array_equal(o0, o1) and np.array_equal(r0, r1)
    o2, r2 = run(0.25)
    mask = np.ones(56, bool)
    for i in range(3):
        mask[18 * i + 2:18 * i + 4] = False  # absolute p_pos columns shift, everything else is relative
    np.testing.assert_allclose(o2[:, mask], o0[:, mask], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r2, r0, rtol=1e-5, atol=1e-5)
    assert np.all(r0[:, 0] == r0[:, 1]) and np.all(r0[:, 1] == r0[:, 2])
    assert np.all(o0[:, 54:] == 0)


def test_numpy_single_env_surface_matches_oracle():
    """num_envs=1: numpy in / numpy out with the reference's shapes (train.py:104-120)."""
    from maddpg_b200 import make_env
    case = env_case("simple_tag", seed=2)
    env = make_env("simple_tag", state_dtype=torch.float64)
    assert env.n == 4 and [s.shape for s in env.observation_space] == [(16,), (16,), (16,), (14,)]
    assert all(s.n == 5 for s in env.action_space)
    init = env.state_from_arrays(case["agent_pos"][:1], case["agent_vel"][:1], case["landmark_pos"][:1])
    obs_n = env.reset(init_state=init)
    assert isinstance(obs_n[0], np.ndarray) and obs_n[0].shape == (16,) and obs_n[3].shape == (14,)
    from oracle import mpe
    o = mpe.BatchedOracleEnv("simple_tag", 1)
    o.set_state(case["agent_pos"][:1], case["agent_vel"][:1], case["landmark_pos"][:1])
    for t in range(5):
        acts = [case["tape"][t][i][0] for i in range(4)]
        obs_n, rew_n, done_n, info_n = env.step(acts)
        oo, rr, dd = o.step([a[None] for a in acts])
        assert isinstance(rew_n[0], float) and done_n == [False] * 4 and "n" in info_n
        for i in range(4):
            np.testing.assert_allclose(obs_n[i], oo[i][0], rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(rew_n[i], rr[0][i], rtol=RTOL, atol=ATOL)


def test_step_with_fused_ring_insert():
    from maddpg_b200 import BatchedMultiAgentEnv, JointReplayRing
    E = 96
    rng = np.random.RandomState(1)
    env = BatchedMultiAgentEnv("simple_tag", num_envs=E, squeeze=False)
    ring = JointReplayRing(env.obs_dims, env.act_dims, capacity=250)
    env.reset()
    rows = []
    for t in range(4):  # 4 * 96 = 384 rows into 250: wraps
        prev = env.obs.clone()
        a = _joint_act(env, [soft_actions(rng, E, 5) for _ in range(4)])
        env.step_device(a, ring=ring)
        row = torch.zeros((E, ring.row_stride))
        L = ring.layout
        row[:, :L.obs_sum] = prev[:, :L.obs_sum].cpu()
        row[:, L.obs_sum:L.x_dim] = a[:, :L.act_sum].cpu()
        row[:, L.nx_off:L.nx_off + L.obs_sum] = env.obs[:, :L.obs_sum].cpu()
        row[:, L.rw_off:L.rw_off + 4] = env.rew.cpu()
        rows.append(row)
    assert ring.length == [250] * 4 and ring.next_idx == [384 % 250] * 4
    allrows = torch.cat(rows)
    got = ring.ring.cpu()
    L = ring.layout
    used = list(range(0, L.x_dim)) + list(range(L.nx_off, L.nx_off + L.obs_sum)) + list(range(L.rw_off, L.dn_off + 4))
    for k in range(384 - 250, 384):
        assert torch.equal(got[k % 250, used], allrows[k, used]), k


@pytest.mark.parametrize("A,E", [(2, 77), (3, 4096 + 5), (4, 130), (5, 64), (6, 129), (7, 67), (24, 263), (32, 40)])
def test_spread_register_kernel_matches_table_driven_kernel(A, E):
    """simple_spread fast paths (A <= 6: one thread per env, registers; A = 7..32: one warp per env, lane = agent, shuffle
    loops over the entities) vs the table-driven kernel on the same Philox
    resets and action tape, 30 steps with contacts.  Both kernels start every step from the same state (the stiff
    contact model amplifies last-bit differences ~100x per contact step), so the comparison
    isolates one step of arithmetic."""
    from maddpg_b200 import BatchedMultiAgentEnv
    envs = [BatchedMultiAgentEnv("simple_spread", num_envs=E, num_agents=A, squeeze=False, seed=5) for _ in range(2)]
    envs[1].force_generic_kernel(True)
    g = torch.Generator(device="cuda").manual_seed(0)
    for env in envs:
        env.reset_device()
    for t in range(30):
        act = torch.softmax(3.0 * torch.randn((E, envs[0].act_stride), device="cuda", generator=g), -1)
        envs[0].state.copy_(envs[1].state)
        for env in envs:
            env.step_device(act)
        torch.testing.assert_close(envs[0].obs, envs[1].obs, rtol=1e-5, atol=1e-5, msg=lambda m: "obs t=%d %s" % (t, m))
        torch.testing.assert_close(envs[0].rew, envs[1].rew, rtol=1e-5, atol=1e-5, msg=lambda m: "rew t=%d %s" % (t, m))
        torch.testing.assert_close(envs[0].state, envs[1].state, rtol=1e-5, atol=1e-5, msg=lambda m: "state t=%d %s" % (t, m))
        assert int(envs[0].done.sum()) == 0


@pytest.mark.parametrize("name", ["simple", "simple_spread", "simple_tag", "simple_world_comm"])
def test_benchmark_data_matches_oracle(name):
    """scenario.benchmark_data (the info_n tape of train.py --benchmark, train.py:139-148) for every (env, agent) after a
    few crowded steps, float64 state: exact counts, distances to 1e-9."""
    case = env_case(name, seed=5, crowd=0.3)
    oenv = case["env"]
    oenv.set_state(case["agent_pos"], case["agent_vel"], case["landmark_pos"], case["agent_c"], case.get("goal"))
    env = _make(case, torch.float64)
    env.reset(init_state=env.state_from_arrays(case["agent_pos"], case["agent_vel"], case["landmark_pos"], case["agent_c"], case.get("goal")))
    saw_collision = False
    for t, acts in enumerate(case["tape"][:6]):
        oenv.step(acts)
        env.step_device(_joint_act(env, acts))
        got = env.benchmark_data().cpu().numpy()
        assert got.shape == (case["E"], env.n, 4)
        for e, oe in enumerate(oenv.envs):
            for i, ag in enumerate(oe.agents):
                ref = oe.scenario.benchmark_data(ag, oe.world)
                if name == "simple":
                    assert ref == {} and not got[e, i].any()
                elif name == "simple_spread":
                    np.testing.assert_allclose(got[e, i, [0, 2]], [ref[0], ref[2]], rtol=1e-6, atol=1e-6)
                    assert int(got[e, i, 1]) == ref[1] and int(got[e, i, 3]) == ref[3]
                    saw_collision |= ref[1] > 1
                else:
                    assert int(got[e, i, 0]) == ref and not got[e, i, 1:].any()
                    saw_collision |= ref > 0
    assert saw_collision or name == "simple"


def test_benchmark_env_fills_info_n():
    """make_env(..., benchmark=True) (train.py:56-58): step() returns the reference's info_n shapes."""
    from maddpg_b200.env import make_env
    env = make_env("simple_spread", benchmark=True, num_envs=1)
    obs_n = env.reset()
    act_n = [np.eye(5, dtype=np.float32)[1] for _ in range(env.n)]
    obs_n, rew_n, done_n, info_n = env.step(act_n)
    assert len(info_n["n"]) == env.n
    rew, collisions, min_dists, occupied = info_n["n"][0]
    assert isinstance(collisions, int) and collisions >= 1 and isinstance(occupied, int) and min_dists > 0
    # shared reward = sum over agents of the per-agent benchmark reward (simple_spread is collaborative)
    np.testing.assert_allclose(sum(x[0] for x in info_n["n"]), rew_n[0], rtol=1e-5)
    env2 = make_env("simple_tag", benchmark=True, num_envs=4, squeeze=False)
    env2.reset()
    _, _, _, info2 = env2.step([np.tile(np.eye(5, dtype=np.float32)[0], (4, 1)) for _ in range(env2.n)])
    assert all(x.shape == (4,) and x.dtype == np.int64 for x in info2["n"])
    assert make_env("simple_spread", num_envs=1).step(act_n)[3] == {"n": [{} for _ in range(3)]}


@pytest.mark.parametrize("A,E", [(24, 101), (8, 300)])
def test_spread_warp_kernel_fused_ring_insert_matches_separate_insert(A, E):
    """simple_spread, one warp per env instance: the step kernel writes the joint replay rows itself (ReplayBuffer.add fused
    into env.step); the table-driven kernel + the separate insert kernel must produce the same ring, incl. the wrap-around."""
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
    rings = []
    for generic in (False, True):
        env = BatchedMultiAgentEnv("simple_spread", num_envs=E, num_agents=A, squeeze=False, seed=5)
        env.force_generic_kernel(generic)
        core = MADDPGCore(env.obs_dims, env.action_space, [False] * A, replay_capacity=2 * E + 7, seed=1)
        core.ring.ring.zero_()
        env.reset_device()
        g = torch.Generator(device="cuda").manual_seed(3)
        states = []
        for t in range(3):  # 3 E rows into a ring of 2 E + 7: wraps
            act = torch.softmax(3.0 * torch.randn((E, env.act_stride), device="cuda", generator=g), -1)
            if rings:
                env.state.copy_(rings[0][2][t])  # same pre-step state as the fast path (isolates one step of arithmetic)
            states.append(env.state.clone())
            env.step_device(act, ring=core.ring)
        rings.append((core.ring.ring.cpu(), list(core.ring.next_idx), states))
    (r0, n0, _), (r1, n1, _) = rings
    assert n0 == n1
    L = core.ring.layout
    used = list(range(0, L.x_dim)) + list(range(L.nx_off, L.nx_off + L.obs_sum)) + list(range(L.rw_off, L.dn_off + A))
    torch.testing.assert_close(r0[:, used], r1[:, used], rtol=1e-5, atol=1e-5)
    assert torch.equal(r0[:, L.obs_sum:L.x_dim], r1[:, L.obs_sum:L.x_dim])  # act_t is a copy of the same tape: bit-identical


@pytest.mark.parametrize("name", ["simple_adversary", "simple_push", "simple_speaker_listener", "simple_crypto", "simple_reference"])
def test_goal_scenarios_device_reset(name):
    """reset_world of the goal scenarios: the goal landmark index (np.random.choice(world.landmarks)) is drawn on the device per
    env instance and episode, uniformly; observations of the fresh state equal the oracle's with the same goals injected."""
    from maddpg_b200 import BatchedMultiAgentEnv
    from oracle import mpe
    E = 512
    env = BatchedMultiAgentEnv(name, num_envs=E, state_dtype=torch.float64, squeeze=False, seed=9)
    obs = torch.cat(env.reset(), 1).cpu().numpy()
    st = env.state_to_arrays()
    L = env.n_landmarks
    assert st["goal"].shape == (E, env.n_goal) and st["goal"].min() == 0 and st["goal"].max() == L - 1
    counts = np.bincount(st["goal"][:, 0], minlength=L)
    assert counts.min() > E / L * 0.7, counts
    assert np.all(st["agent_vel"] == 0) and np.all(st["comm"] == 0)
    oenv = mpe.BatchedOracleEnv(name, E)
    oenv.set_state(st["agent_pos"], st["agent_vel"], st["landmark_pos"], None, st["goal"])
    np.testing.assert_allclose(obs[:, :sum(env.obs_dims)], np.concatenate(oenv.observe(), 1), rtol=RTOL, atol=ATOL)
    g2 = None
    for _ in range(3):
        env.reset()
        g2 = env.state_to_arrays()["goal"]
        if not np.array_equal(g2, st["goal"]):
            break
    assert not np.array_equal(g2, st["goal"])  # a new episode draws new goals
    # immovable agents stay where reset put them; speakers publish their action as state.c
    env.act.copy_(torch.softmax(torch.randn_like(env.act), -1))
    before = env.state_to_arrays()
    env.step_device()
    after = env.state_to_arrays()
    for i in range(env.n):
        if not env.movable[i]:
            assert np.array_equal(before["agent_pos"][:, i], after["agent_pos"][:, i])
        if env.comm_len[i]:
            lo = env.act_off[i] + (5 if env.movable[i] else 0)
            np.testing.assert_array_equal(after["comm"][:, env.comm_off[i]:env.comm_off[i] + env.comm_len[i]],
                                          env.act[:, lo:lo + env.comm_len[i]].double().cpu().numpy())


def test_goal_scenarios_numpy_surface_and_spaces():
    """The reference-shaped surface of the new scenarios: Discrete(3) speaker / Discrete(5) listener, Discrete(4) crypto agents."""
    from maddpg_b200 import make_env
    env = make_env("simple_speaker_listener")
    assert [s.n for s in env.action_space] == [3, 5] and [s.shape for s in env.observation_space] == [(3,), (11,)]
    obs_n = env.reset()
    assert [o.shape for o in obs_n] == [(3,), (11,)] and sorted(np.round(obs_n[0].astype(np.float64), 2).tolist()) == [0.15, 0.15, 0.65]
    obs_n, rew_n, done_n, info_n = env.step([np.asarray([0.2, 0.5, 0.3], np.float32), np.eye(5, dtype=np.float32)[1]])
    np.testing.assert_allclose(obs_n[1][-3:], [0.2, 0.5, 0.3], rtol=1e-6)   # the listener hears the speaker's message
    assert rew_n[0] == rew_n[1] and done_n == [False, False]                # collaborative: shared reward
    env = make_env("simple_crypto")
    assert [s.n for s in env.action_space] == [4, 4, 4] and [s.shape for s in env.observation_space] == [(4,), (8,), (8,)]
    env = make_env("simple_reference")
    assert [(int(s.low[0]), int(s.high[0]), int(s.low[1]), int(s.high[1])) for s in env.action_space] == [(0, 4, 0, 9)] * 2
    assert [s.shape for s in env.observation_space] == [(21,), (21,)]
    with pytest.raises(NotImplementedError):
        make_env("simple_football")
