"""Runs the reference's OWN ``experiments/train.py`` -- unmodified, straight from the reference checkout -- on the B200
kernels.  Nothing of that file is restated here: this module only installs stand-ins for the four imports the file makes and
then executes it with ``runpy``:

    reference import (experiments/train.py)                      stand-in
    -----------------------------------------------------------  -------------------------------------------------------
    import tensorflow as tf                    (:3)              inert module: tf.train.Saver() (:101) is a token object;
    import tensorflow.contrib.layers as layers (:9)              mlp_model (:39-46) is never called -- the MLP is in the kernels
    import maddpg.common.tf_util as U          (:7)              single_threaded_session / initialize (:79,89) no-ops,
                                                                 save_state / load_state (:95,164) -> the .pt checkpoint below
    from maddpg.trainer.maddpg import MADDPGAgentTrainer (:8)    maddpg_b200.MADDPGAgentTrainer
    from multiagent.environment import MultiAgentEnv (:49)       maddpg_b200.BatchedMultiAgentEnv behind the same constructor
    import multiagent.scenarios as scenarios   (:50)             scenarios.load(name + ".py").Scenario() -> a scenario token

    python -m maddpg_b200.train --reference-train /path/to/maddpg/experiments/train.py -- --scenario simple_spread --num-episodes 1000
    python -m maddpg_b200.train --num-envs 4096 -- --scenario simple_spread --num-episodes 40960      # batched superset

Options of this shim come BEFORE ``--``; everything after it is the reference's own command line (train.py:11-37).  With
``--num-envs 1`` every call in the loop has the reference's shapes (numpy in / numpy out).  With ``--num-envs E > 1`` one
loop iteration is E lockstep transitions: rewards reach the loop as floats (the mean over env instances) that carry the
per-instance vector for ``experience``, ``done`` flags as objects whose truth value is "all instances done" (SURVEY H9).
"""
import argparse
import os
import runpy
import sys
import types

import numpy as np
import torch

from .env import BatchedMultiAgentEnv, SCENARIOS
from .trainer import MADDPGAgentTrainer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference/experiments/train.py", os.path.join(ROOT, "baseline", "_ref", "experiments", "train.py"))

# shim options (set by main() / run_reference_train); trainers and envs created by the reference's code read them
OPTIONS = {"num_envs": 1, "num_agents": None, "device": "cuda", "seed": 0, "replay_capacity": int(1e6)}
_LIVE = []  # trainers constructed by the running reference script (for U.save_state / U.load_state)


def reference_train_path(path=None):
    for p in ([path] if path else []) + [os.environ.get("MADDPG_REFERENCE_TRAIN", "")] + list(CANDIDATES):
        if p and os.path.isfile(p):
            return p
    raise FileNotFoundError("the reference's experiments/train.py was not found; pass --reference-train PATH or set "
                            "MADDPG_REFERENCE_TRAIN (looked in %s)" % (CANDIDATES,))


# -- checkpoints: U.save_state / U.load_state (tf_util.py:259-273) ---------------------------------------------------------------
def save_state(save_dir, trainers=None, saver=None):
    """All variables incl. Adam slots and the Philox stream position (the replay ring is not saved, like the reference)."""
    core = (trainers or _LIVE)[0].core
    os.makedirs(save_dir, exist_ok=True)
    torch.save({"params": core.params.cpu(), "adam_m": core.adam_m.cpu(), "adam_v": core.adam_v.cpu(),
                "adam_t": core.adam_t.cpu(), "obs_dims": core.obs_dims, "act_dims": core.act_dims,
                "num_units": core.num_units, "counter": core.counter, "seed": core.seed},
               os.path.join(save_dir, "maddpg_b200.pt"))


def load_state(load_dir, trainers=None, saver=None):
    core = (trainers or _LIVE)[0].core
    st = torch.load(os.path.join(load_dir, "maddpg_b200.pt"), map_location="cpu")
    assert st["obs_dims"] == core.obs_dims and st["act_dims"] == core.act_dims and st["num_units"] == core.num_units
    core.params.copy_(st["params"])
    core.adam_m.copy_(st["adam_m"])
    core.adam_v.copy_(st["adam_v"])
    core.adam_t.copy_(st["adam_t"])
    core.counter = int(st.get("counter", 0))  # a restored run continues the noise stream instead of replaying it
    core.seed = int(st.get("seed", core.seed))


# -- stand-in modules ---------------------------------------------------------------------------------------------------------
class _ShimTrainer(MADDPGAgentTrainer):
    """maddpg.trainer.maddpg.MADDPGAgentTrainer as the reference constructs it (train.py:63-75); the shim's device options
    are attached to the reference's ``arglist`` namespace, which has no such flags."""

    def __init__(self, name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=False):
        for k in ("device", "seed", "replay_capacity"):
            if not hasattr(args, k):
                setattr(args, k, OPTIONS[k])
        MADDPGAgentTrainer.__init__(self, name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=local_q_func)
        _LIVE.append(self)


class _World(object):
    def __init__(self, scenario_name):
        self.scenario_name = scenario_name


class _Scenario(object):
    """multiagent.scenarios.<name>.Scenario as train.py:53-60 uses it: the callbacks are tokens, the physics, rewards
    and observations they stand for run inside the env kernels (csrc/mdp_env_dev.cuh)."""

    def __init__(self, scenario_name):
        self.scenario_name = scenario_name

    def make_world(self):
        return _World(self.scenario_name)

    def reset_world(self, world):
        raise NotImplementedError("evaluated on the device by BatchedMultiAgentEnv.reset")

    reward = observation = benchmark_data = reset_world


def _multi_agent_env(world, reset_callback=None, reward_callback=None, observation_callback=None, info_callback=None,
                     done_callback=None, shared_viewer=True):
    """multiagent.environment.MultiAgentEnv(world, reset, reward, observation[, benchmark_data]) (train.py:57-60)."""
    env = BatchedMultiAgentEnv(world.scenario_name, num_envs=OPTIONS["num_envs"], num_agents=OPTIONS["num_agents"],
                               device=OPTIONS["device"], seed=OPTIONS["seed"], benchmark=info_callback is not None)
    env.reference_loop = True
    return env


class _Session(object):
    def __enter__(self):
        del _LIVE[:]
        return self

    def __exit__(self, *exc):
        return False


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # importable as a package
    return m


def install_stubs():
    """Puts the stand-ins into sys.modules (replacing a real tensorflow / maddpg / multiagent if one is importable).
    Returns the previous entries so that remove_stubs can restore them."""
    def unavailable(*a, **k):
        raise NotImplementedError("TensorFlow graph construction is not part of the B200 path: the MLP of train.py:39-46 "
                                  "is evaluated by libmaddpg_b200")
    layers = _module("tensorflow.contrib.layers", fully_connected=unavailable)
    contrib = _module("tensorflow.contrib", layers=layers)
    tf = _module("tensorflow", contrib=contrib, variable_scope=unavailable,
                 nn=types.SimpleNamespace(relu=unavailable), train=types.SimpleNamespace(Saver=lambda *a, **k: object()))
    tf_util = _module("maddpg.common.tf_util", single_threaded_session=_Session, initialize=lambda: None,
                      save_state=lambda fname, saver=None: save_state(fname, None, saver),
                      load_state=lambda fname, saver=None: load_state(fname, None, saver))
    common = _module("maddpg.common", tf_util=tf_util)
    trainer_mod = _module("maddpg.trainer.maddpg", MADDPGAgentTrainer=_ShimTrainer)
    trainer_pkg = _module("maddpg.trainer", maddpg=trainer_mod)
    maddpg = _module("maddpg", common=common, trainer=trainer_pkg)
    scen = _module("multiagent.scenarios")

    def load(name):  # scenarios.load(scenario_name + ".py") (train.py:53)
        base = name[:-3] if name.endswith(".py") else name
        if base not in SCENARIOS:
            raise NotImplementedError("scenario %r is not implemented (have %s)" % (base, SCENARIOS))
        return types.SimpleNamespace(Scenario=lambda: _Scenario(base))
    scen.load = load
    environment = _module("multiagent.environment", MultiAgentEnv=_multi_agent_env)
    multiagent = _module("multiagent", scenarios=scen, environment=environment)
    mods = {"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers, "maddpg": maddpg,
            "maddpg.common": common, "maddpg.common.tf_util": tf_util, "maddpg.trainer": trainer_pkg,
            "maddpg.trainer.maddpg": trainer_mod, "multiagent": multiagent, "multiagent.scenarios": scen,
            "multiagent.environment": environment}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    return saved


def remove_stubs(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def run_reference_train(ref_argv, path=None, **options):
    """Executes the reference's experiments/train.py as ``__main__`` with ``ref_argv`` as its command line.
    options: num_envs, num_agents, device, seed, replay_capacity.  Returns the trainers the script constructed."""
    path = reference_train_path(path)
    OPTIONS.update({k: v for k, v in options.items() if v is not None})
    saved = install_stubs()
    argv0 = sys.argv
    sys.argv = [path] + list(ref_argv)
    try:
        runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = argv0
        remove_stubs(saved)
    return list(_LIVE)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    own, ref = (argv[:argv.index("--")], argv[argv.index("--") + 1:]) if "--" in argv else ([], argv)
    ap = argparse.ArgumentParser("maddpg_b200.train", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference-train", default=None, help="path of the reference's experiments/train.py")
    ap.add_argument("--num-envs", type=int, default=1, help="lockstep env instances on the GPU")
    ap.add_argument("--num-agents", type=int, default=None, help="simple_spread only: N agents = N landmarks")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--replay-capacity", type=int, default=int(1e6))
    o = ap.parse_args(own)
    run_reference_train(ref, o.reference_train, num_envs=o.num_envs, num_agents=o.num_agents, device=o.device, seed=o.seed,
                        replay_capacity=o.replay_capacity)


if __name__ == "__main__":
    main()
