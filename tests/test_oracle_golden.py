"""The oracles against their committed fixtures (tests/golden/make_oracle_golden.py): CPU-only regression pin."""
import os

import numpy as np
import pytest

from tests.golden.make_oracle_golden import ENV_SEED, UPD_SEED
from tests.helpers import ENV_CASES, TRAINER_CASES, env_case, oracle_update_round, run_oracle_rollout, trainer_case

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", [n for n in ENV_CASES if n != "simple_spread_24"] + ["simple_spread_24"])
def test_env_oracle_matches_golden(name):
    g = np.load(os.path.join(G, "env_%s.npz" % name))
    case = env_case(name, seed=ENV_SEED)
    assert np.array_equal(case["agent_pos"], g["agent_pos"]) and np.array_equal(case["landmark_pos"], g["landmark_pos"])
    ref = run_oracle_rollout(case)
    np.testing.assert_allclose(ref["obs0"], g["obs0"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ref["obs"], g["obs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ref["rew"], g["rew"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ref["final"]["agent_pos"], g["final_pos"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", list(TRAINER_CASES))
def test_update_oracle_matches_golden(name):
    g = np.load(os.path.join(G, "update_%s.npz" % name))
    ref = oracle_update_round(trainer_case(name, seed=UPD_SEED))
    for j, r in enumerate(ref):
        np.testing.assert_allclose(np.asarray(r["stats"], np.float64), g["stats_%d" % j], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(r["y"], g["y_%d" % j], rtol=1e-5, atol=1e-6)
        for key in ("q", "p", "target_q", "target_p"):
            s = np.asarray([float(np.sum(x, dtype=np.float64)) for x in r[key]])
            np.testing.assert_allclose(s, g["%s_sum_%d" % (key, j)], rtol=1e-4, atol=1e-4)
