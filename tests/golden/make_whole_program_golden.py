"""Generates tests/golden/whole_program_ref.npz: the reference's WHOLE training program executed in the build container --

    experiments/train.py (runpy, unmodified)  ->  maddpg/trainer/maddpg.py (MADDPGAgentTrainer, q_train, p_train, make_update_exp)
    ->  maddpg/common/distributions.py, maddpg/common/tf_util.py, maddpg/trainer/replay_buffer.py     -- all the REAL files --

on two stand-ins only: tests/tf_shim.py for the TensorFlow calls and oracle/mpe.py for the (un-vendored) MPE package.

    python tests/golden/make_whole_program_golden.py

The trainer class is a subclass of the REAL one whose only addition is to overwrite the freshly initialised variables with the
oracle's initial weights (so that both sides start equal) after the real ``__init__`` has built the real graph.  Gumbel noise:
one shared U[0,1) stream behind ``tf.random_uniform``.  Recorded: the learning-curve lists train.py pickles.  The all-oracle loop
(oracle/train_loop.py::run_training with the same shared noise stream) must reproduce them
(tests/test_oracle_golden.py::test_whole_program_matches_the_oracle_loop).
"""
import os
import pickle
import random
import runpy
import sys
import tempfile
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import tf_shim  # noqa: E402

tf = tf_shim.install()
gym, spaces = types.ModuleType("gym"), types.ModuleType("gym.spaces")
sys.modules.update({"gym": gym, "gym.spaces": spaces, "tqdm": types.ModuleType("tqdm")})
sys.modules["tqdm"].tqdm, sys.modules["tqdm"].trange = (lambda it, **k: it), range
gym.spaces = spaces
sys.path.insert(0, "/root/reference")
from oracle import maddpg as omaddpg  # noqa: E402
from oracle import mpe as ompe  # noqa: E402

# make_pdtype (distributions.py:408-422) tells action spaces apart by gym class: the oracle MPE's Discrete IS gym's for this run
spaces.Discrete, spaces.Box, spaces.MultiBinary, spaces.Dict = ompe.Discrete, ompe.Box, type("MultiBinary", (), {}), type("Dict", (), {})
import maddpg.common.tf_util as U  # noqa: E402  (REAL)
import maddpg.trainer.maddpg as real_trainer_module  # noqa: E402  (REAL)

from tests.update_case import shared_noise  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 3
ARGV = ["--scenario", "simple_spread", "--num-episodes", "60", "--max-episode-len", "5", "--batch-size", "8", "--num-units", "16",
        "--save-rate", "4"]
RealTrainer = real_trainer_module.MADDPGAgentTrainer
NETS = (("q_func", "q"), ("target_q_func", "target_q"), ("p_func", "p"), ("target_p_func", "target_p"))


class SeededTrainer(RealTrainer):
    def __init__(self, name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=False):
        super().__init__(name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=local_q_func)   # the REAL graph
        o = omaddpg.OracleAgentTrainer(name, None, obs_shape_n, act_space_n, agent_index, args, local_q_func=local_q_func,
                                       rng=np.random.RandomState(SEED * 1000 + agent_index))
        byname = {v.op.name: v for v in tf_shim._VARIABLES}
        for scope, attr in NETS:
            w = getattr(o, attr).p
            for k, layer in enumerate(("fully_connected", "fully_connected_1", "fully_connected_2")):
                byname["%s/%s/%s/weights" % (name, scope, layer)].load(w[2 * k])
                byname["%s/%s/%s/biases" % (name, scope, layer)].load(w[2 * k + 1])


def install_mpe(rng):
    env_mod, scen_mod, pkg = types.ModuleType("multiagent.environment"), types.ModuleType("multiagent.scenarios"), types.ModuleType("multiagent")
    env_mod.MultiAgentEnv = ompe.MultiAgentEnv

    def load(fname):
        m = types.ModuleType("scenario")
        m.Scenario = lambda: ompe.make_scenario(fname[:-3], rng, None)
        return m
    scen_mod.load = load
    pkg.environment, pkg.scenarios = env_mod, scen_mod
    sys.modules.update({"multiagent": pkg, "multiagent.environment": env_mod, "multiagent.scenarios": scen_mod})


def main():
    U.save_state = lambda *a, **k: None          # no checkpoint files (tf.train.Saver is a token in the stand-in)
    real_trainer_module.MADDPGAgentTrainer = SeededTrainer
    install_mpe(np.random.RandomState(SEED))
    random.seed(SEED)
    tf_shim.NOISE[0] = shared_noise(SEED)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        sys.argv = ["train.py"] + ARGV + ["--exp-name", "whole", "--plots-dir", tmp + "/", "--save-dir", tmp + "/"]
        runpy.run_path("/root/reference/experiments/train.py", run_name="__main__")
        out["rewards"] = np.asarray(pickle.load(open(os.path.join(tmp, "whole_rewards.pkl"), "rb")), np.float64)
        out["agrewards"] = np.asarray(pickle.load(open(os.path.join(tmp, "whole_agrewards.pkl"), "rb")), np.float64)
    out["argv"] = np.asarray(ARGV)
    np.savez_compressed(os.path.join(HERE, "whole_program_ref.npz"), **out)
    print("wrote whole_program_ref.npz:", out["rewards"])


if __name__ == "__main__":
    main()
