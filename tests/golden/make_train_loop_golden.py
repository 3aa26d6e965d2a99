"""Generates tests/golden/train_loop_ref.npz by executing the REAL ``experiments/train.py`` (/root/reference, unmodified, via
runpy) on the oracle classes -- build container only:

    python tests/golden/make_train_loop_golden.py

Stand-ins: permissive ``tensorflow`` / ``gym`` / ``tqdm`` modules for the imports; ``multiagent.environment`` / ``multiagent.scenarios``
-> oracle/mpe.py's restated MPE (same constructor and ``scenarios.load(name).Scenario()`` protocol); ``MADDPGAgentTrainer`` -> a
subclass of the REAL class whose ``__init__`` installs oracle/maddpg.py's graph callables instead of building a TF graph (so the
REAL ``action`` / ``experience`` / ``preupdate`` / ``update`` methods and the REAL ReplayBuffer run); ``U.initialize`` /
``U.save_state`` no-ops.  Recorded: the two learning-curve lists train.py pickles at the end (:181-187) for two cases.
oracle/train_loop.py::run_training -- the loop bench.py's ``--impl reference`` arm and ``cpu_baseline`` time -- must reproduce
them bit for bit (tests/test_oracle_golden.py::test_train_loop_matches_the_reference_script).
"""
import os
import pickle
import random
import runpy
import sys
import tempfile
import types

import numpy as np


class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


for name in ("tensorflow", "tensorflow.python", "tensorflow.python.ops", "tensorflow.contrib", "tensorflow.contrib.layers", "gym",
             "gym.spaces", "tqdm"):
    sys.modules[name] = _Any(name)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import maddpg.common.tf_util as U  # noqa: E402  (the REAL module, on the stand-in tensorflow)
import maddpg.trainer.maddpg as real_trainer_module  # noqa: E402
from maddpg.trainer.replay_buffer import ReplayBuffer  # noqa: E402

from oracle import maddpg as omaddpg  # noqa: E402
from oracle import mpe as ompe  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {  # name -> train.py command line (small batch / short episodes so that update rounds happen within a few hundred steps)
    "spread": ["--scenario", "simple_spread", "--num-episodes", "60", "--max-episode-len", "5", "--batch-size", "8", "--num-units",
               "16", "--save-rate", "4"],
    "tag_ddpg_adv": ["--scenario", "simple_tag", "--num-episodes", "45", "--max-episode-len", "6", "--batch-size", "6", "--num-units",
                     "16", "--save-rate", "5", "--num-adversaries", "3", "--adv-policy", "ddpg"],
}
SEED = 3
RealTrainer = real_trainer_module.MADDPGAgentTrainer


class ShimTrainer(RealTrainer):
    """The REAL trainer class with the TF graph replaced by the oracle's restated graph callables."""

    def __init__(self, name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=False):
        o = omaddpg.OracleAgentTrainer(name, None, obs_shape_n, act_space_n, agent_index, args, local_q_func=local_q_func,
                                       rng=np.random.RandomState(SEED * 1000 + agent_index))
        n = o.n
        self.name, self.n, self.agent_index, self.args = name, n, agent_index, args
        self.q_train = lambda *a: o.q_train(list(a[:n]), list(a[n:2 * n]), a[2 * n])
        self.p_train = lambda *a: o.p_train(list(a[:n]), list(a[n:2 * n]))
        self.p_update, self.q_update = o.p_update, o.q_update
        self.p_debug, self.q_debug, self.act = o.p_debug, o.q_debug, o.act
        self.replay_buffer = ReplayBuffer(1e6)
        self.max_replay_buffer_len = args.batch_size * args.max_episode_len
        self.replay_sample_index = None


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
    U.initialize = lambda: None
    U.save_state = lambda *a, **k: None
    real_trainer_module.MADDPGAgentTrainer = ShimTrainer
    out = {}
    for case, argv in CASES.items():
        install_mpe(np.random.RandomState(SEED))
        random.seed(SEED)
        with tempfile.TemporaryDirectory() as tmp:
            sys.argv = ["train.py"] + argv + ["--exp-name", case, "--plots-dir", tmp + "/", "--save-dir", tmp + "/"]
            runpy.run_path("/root/reference/experiments/train.py", run_name="__main__")
            out[case + "_rewards"] = np.asarray(pickle.load(open(os.path.join(tmp, case + "_rewards.pkl"), "rb")), np.float64)
            out[case + "_agrewards"] = np.asarray(pickle.load(open(os.path.join(tmp, case + "_agrewards.pkl"), "rb")), np.float64)
        out[case + "_argv"] = np.asarray(argv)
    np.savez_compressed(os.path.join(HERE, "train_loop_ref.npz"), **out)
    print("wrote train_loop_ref.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
