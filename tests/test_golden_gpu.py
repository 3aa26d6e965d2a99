"""CUDA path against the COMMITTED golden fixtures (tests/golden/*.npz) -- no oracle execution needed here:
env rollouts (fp64 state) and one sequential update round per trainer case."""
import os

import numpy as np
import pytest
import torch

from tests.golden.make_oracle_golden import ENV_SEED, UPD_SEED
from tests.helpers import ENV_CASES, TRAINER_CASES, trainer_case

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_env_rollout_matches_golden(name):
    from maddpg_b200 import BatchedMultiAgentEnv
    g = np.load(os.path.join(G, "env_%s.npz" % name))
    scenario, na, E, T = ENV_CASES[name]
    env = BatchedMultiAgentEnv(scenario, num_envs=E, num_agents=na, state_dtype=torch.float64, squeeze=False)
    init = env.state_from_arrays(g["agent_pos"], g["agent_vel"], g["landmark_pos"], goal=g["goal"] if "goal" in g.files else None)
    obs0 = torch.cat(env.reset(init_state=init), dim=1).cpu().numpy()
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-5, atol=1e-5)
    nobs = sum(env.obs_dims)
    for t in range(T):
        a = torch.zeros((E, env.act_stride))
        a[:, :g["tape"].shape[2]] = torch.from_numpy(g["tape"][t])
        env.step_device(a.cuda())
        np.testing.assert_allclose(env.obs[:, :nobs].cpu().numpy(), g["obs"][t], rtol=1e-5, atol=1e-5, err_msg="obs step %d" % t)
        np.testing.assert_allclose(env.rew.cpu().numpy(), g["rew"][t], rtol=1e-5, atol=1e-5, err_msg="rew step %d" % t)
    st = env.state_to_arrays()
    np.testing.assert_allclose(st["agent_pos"], g["final_pos"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(st["agent_vel"], g["final_vel"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", list(TRAINER_CASES))
def test_update_round_matches_golden(name):
    from tests.test_trainer_gpu import _build
    g = np.load(os.path.join(G, "update_%s.npz" % name))
    case = trainer_case(name, seed=UPD_SEED)  # seeded inputs only; the oracle is not run
    trainers, core = _build(case)
    for j, tr in enumerate(trainers):
        tr.inject_noise(u_target=case["u_target"][j], u_actor=case["u_actor"][j])
        stats = tr.update(trainers, 100, index=case["idx"][j])
        np.testing.assert_allclose(np.asarray(stats, np.float64), g["stats_%d" % j], rtol=1e-4, atol=2e-6, err_msg="stats agent %d" % j)
    for j in range(case["n"]):
        for net, key in ((2, "q"), (0, "p"), (3, "target_q"), (1, "target_p")):
            w = core.get_weights(j, net)
            s = np.asarray([float(np.sum(x, dtype=np.float64)) for x in w])
            # element-wise agreement is ~2e-4 after one Adam step (tests/test_trainer_gpu.py); a checksum over n
            # elements of magnitude ~|x| may drift by about 2e-4 * n: bound it with the tensor's L1 mass
            tol = np.maximum(2e-3 * g["%s_abs_%d" % (key, j)], 5e-3)
            bad = np.abs(s - g["%s_sum_%d" % (key, j)]) > tol
            assert not bad.any(), "agent %d %s checksums: got %s want %s" % (j, key, s, g["%s_sum_%d" % (key, j)])
