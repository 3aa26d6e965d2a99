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


@pytest.mark.parametrize("case", ["spread", "tag_ddpg_adv"])
def test_train_loop_matches_the_reference_script(case):
    """tests/golden/train_loop_ref.npz: the learning-curve lists the REAL experiments/train.py pickled (:181-187) when it was
    executed unmodified -- its own loop, the REAL MADDPGAgentTrainer methods and ReplayBuffer -- on the oracle's MPE and graph
    callables (tests/golden/make_train_loop_golden.py).  oracle/train_loop.py::run_training, the loop bench.py's reference arm
    times, must reproduce them bit for bit."""
    import argparse
    import random
    from oracle import train_loop
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_loop_ref.npz"))
    argv = dict(zip(gold[case + "_argv"][0::2], gold[case + "_argv"][1::2]))
    arglist = argparse.Namespace(scenario=str(argv["--scenario"]), max_episode_len=int(argv["--max-episode-len"]), lr=1e-2, gamma=0.95,
                                 batch_size=int(argv["--batch-size"]), num_units=int(argv["--num-units"]),
                                 num_adversaries=int(argv.get("--num-adversaries", 0)), good_policy="maddpg",
                                 adv_policy=str(argv.get("--adv-policy", "maddpg")))
    random.seed(3)
    rewards, agrewards, steps = train_loop.run_training(arglist.scenario, int(argv["--num-episodes"]), arglist, seed=3,
                                                        save_rate=int(argv["--save-rate"]))
    assert np.array_equal(np.asarray(rewards, np.float64), gold[case + "_rewards"])
    assert np.array_equal(np.asarray(agrewards, np.float64), gold[case + "_agrewards"])
    assert steps >= 200      # update rounds ran (t = 100, 200 past the warm-up gate)


def test_whole_program_matches_the_oracle_loop():
    """tests/golden/whole_program_ref.npz: the reference's whole training program -- experiments/train.py, the real
    MADDPGAgentTrainer with its real graph-building code (q_train, p_train, make_update_exp), distributions.py, tf_util.py and
    replay_buffer.py, all unmodified -- executed on two stand-ins only (tests/tf_shim.py for TensorFlow, oracle/mpe.py for the
    un-vendored MPE package; tests/golden/make_whole_program_golden.py).  The all-oracle loop started from the same weights and
    fed the same noise stream must produce the same learning curve: 300 environment steps, 61 episodes, update rounds at
    t = 100, 200, 300 feeding back into the actions (float32 sums in another order: measured 4.5e-9 relative, held to 1e-6)."""
    import argparse
    import random
    from oracle import train_loop
    from tests.update_case import shared_noise
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "whole_program_ref.npz"))
    argv = dict(zip(gold["argv"][0::2], gold["argv"][1::2]))
    arglist = argparse.Namespace(scenario=str(argv["--scenario"]), max_episode_len=int(argv["--max-episode-len"]), lr=1e-2, gamma=0.95,
                                 batch_size=int(argv["--batch-size"]), num_units=int(argv["--num-units"]), num_adversaries=0,
                                 good_policy="maddpg", adv_policy="maddpg")
    random.seed(3)
    rewards, agrewards, steps = train_loop.run_training(arglist.scenario, int(argv["--num-episodes"]), arglist, seed=3,
                                                        save_rate=int(argv["--save-rate"]), noise=shared_noise(3))
    assert len(rewards) == len(gold["rewards"]) == 15 and steps == 300
    np.testing.assert_allclose(np.asarray(rewards, np.float64), gold["rewards"], rtol=1e-6)
    np.testing.assert_allclose(np.asarray(agrewards, np.float64), gold["agrewards"], rtol=1e-6)
    print("worst relative difference of the learning curve: %.2e" % np.max(np.abs(np.asarray(rewards) - gold["rewards"]) / np.abs(gold["rewards"])))


def test_every_golden_file_has_its_generating_script():
    """Each fixture under tests/golden/ is written by a committed script in the same directory (the rule for vectors produced from
    the reference or from the oracle): the script names the file it writes."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    scripts = "".join(open(os.path.join(here, f)).read() for f in sorted(os.listdir(here)) if f.startswith("make_") and f.endswith(".py"))
    missing = []
    for f in sorted(os.listdir(here)):
        if not f.endswith(".npz"):
            continue
        stem = f[:-4]
        family = stem.split("_")[0] + "_"            # env_<case>.npz / update_<case>.npz are written by one loop over the cases
        if f not in scripts and not (family in ("env_", "update_") and ('"%s%%s.npz"' % family in scripts or "'%s%%s.npz'" % family in scripts
                                                                       or family + "%s.npz" in scripts)):
            missing.append(f)
    assert not missing, missing
