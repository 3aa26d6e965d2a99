"""Generates tests/golden/fork_graph_ref.npz by executing the fork's OWN algorithm and module code -- ``Coma`` / ``Maddpg`` /
``MaTd3`` (maddpg/algorithms/*.py), ``ComaModule`` / ``MaddpgModule`` / ``MaTD3Module``, ``PolicyGroup`` / ``CriticGroup``,
``Policy`` / ``Critic`` / ``LaggingNetwork`` (maddpg/modules/*.py), ``TfFunction`` / ``create_default`` / ``map_to_batch``
(maddpg/common/utils_common.py), all unmodified -- on tests/tf_shim.py, the torch-backed stand-in for the TensorFlow 1.x and
Sonnet 1.x calls those files make.  Build container only:

    python tests/golden/make_fork_graph_golden.py

As with make_graph_golden.py the primitive ops are the stand-in's (restated), everything the fork's code decides is executed for
real: which policies / critics / targets feed which loss, the shared global critic on the first name's reward, the personal
reward, the sign of the worst policy's loss, which variables each optimizer owns, the 5e-3 "polyak" target update, what
``train_step`` returns.  Facts about the reference recorded as well: ``MaTd3(...)`` cannot be constructed (TypeError, the
two-argument ``create_optimizers`` call) -- it is then run with that one call made tolerant of its extra argument, the only
modification of reference code here --, ``Maddpg(...)`` needs a truthy ``hyperparameters`` that it then discards, ``Coma`` asserts
equal spaces, and MaTd3's critic-only steps return their per-name losses split at the "_" of the names.
oracle/matd3.py must reproduce the recorded losses, predictions, values and variables
(tests/test_oracle_matd3.py::test_oracle_matches_the_fork_graph_code).
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import tf_shim  # noqa: E402

tf = tf_shim.install()
tf_shim.install_sonnet()
gym, spaces = types.ModuleType("gym"), types.ModuleType("gym.spaces")


class Box(object):
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = np.full(shape, low, np.float32), np.full(shape, high, np.float32), tuple(shape)

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)


class Dict(object):
    def __init__(self, spaces):
        self.spaces = dict(spaces)


spaces.Box, spaces.Dict, spaces.Discrete = Box, Dict, type("Discrete", (), {})
gym.spaces = spaces
sys.modules.update({"gym": gym, "gym.spaces": spaces, "tqdm": types.ModuleType("tqdm")})
sys.modules["tqdm"].tqdm, sys.modules["tqdm"].trange = (lambda it, **k: it), range
sys.path.insert(0, "/root/reference")
from maddpg.algorithms import Coma, MaTd3, Maddpg  # noqa: E402  (REAL)

from oracle.matd3 import ComaOracle, MaddpgOracle  # noqa: E402
from tests.test_oracle_matd3 import ACT, EQ_ACT, EQ_HIGH, EQ_LOW, EQ_OBS, HIGH, LOW, NAMES, OBS, make_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def lagging(k):
    """The k-th LaggingNetwork built since the last reset: its running and target MLP variables [W1, b1, W2, b2, W3, b3]."""
    byname = {v.op.name: v for v in tf_shim._VARIABLES}
    sfx = "" if k == 0 else "_%d" % k
    return tuple([byname["%s%s/linear_%d/%s" % (net, sfx, layer, wb)] for layer in range(3) for wb in ("w", "b")]
                 for net in ("running", "target"))


def load(k, member):
    run, tgt = lagging(k)
    for v, w in zip(run, member.running.p):
        v.load(w)
    for v, w in zip(tgt, member.target.p):
        v.load(w)


def dump(out, key, k):
    run, tgt = lagging(k)
    for i, v in enumerate(run):
        out["%s_running_%d" % (key, i)] = v.numpy()
    for i, v in enumerate(tgt):
        out["%s_target_%d" % (key, i)] = v.numpy()


def drive(alg, out, prefix, steps, dims=(OBS, ACT, LOW, HIGH)):
    for step in steps:
        obs, act, rew, obs_n, done, _ = make_batch(48, 1000 + step, *dims)
        res = alg.train_step(obs, act, rew, obs_n, done, step)
        alg.run_updates()
        for kind in ("actor", "critic"):
            out["%s_s%d_%s" % (prefix, step, kind)] = np.asarray([res[kind][n] for n in NAMES], np.float64)
    obs = make_batch(16, 2000, *dims)[0]
    pred, val = alg.predict(obs, noisy=False), alg.compute_values(obs)
    for n in NAMES:
        out["%s_predict_%s" % (prefix, n)] = np.asarray(pred[n], np.float64).reshape(16, -1)
        out["%s_values_%s" % (prefix, n)] = np.asarray(val[n], np.float64).reshape(16)


def main():
    obs_space = Dict({n: Box(-10.0, 10.0, (OBS[n],)) for n in NAMES})
    act_space = Dict({n: Box(LOW[n], HIGH[n], (ACT[n],)) for n in NAMES})
    out = {}
    n = len(NAMES)
    # ---- COMA: its always-shared global critic group asserts equal spaces (criticgroup.py:28-30): unequal ones cannot be built
    tf_shim.reset()
    try:
        Coma(obs_space, act_space)
        out["coma_unequal_spaces_error"] = np.asarray("")
    except AssertionError:
        out["coma_unequal_spaces_error"] = np.asarray("AssertionError")
    eq = (EQ_OBS, EQ_ACT, EQ_LOW, EQ_HIGH)
    eq_obs_space = Dict({n: Box(-10.0, 10.0, (EQ_OBS[n],)) for n in NAMES})
    eq_act_space = Dict({n: Box(EQ_LOW[n], EQ_HIGH[n], (EQ_ACT[n],)) for n in NAMES})
    # LaggingNetworks are built best[names], worst[names], shared global critic, personal[names]
    tf_shim.reset()
    alg = Coma(eq_obs_space, eq_act_space)
    _ = alg.session
    assert len(tf_shim._VARIABLES) == (3 * n + 1) * 12, len(tf_shim._VARIABLES)
    o = ComaOracle(*eq, seed=61, first=NAMES[0])
    for i, name in enumerate(NAMES):
        load(i, o.best[name])
        load(n + i, o.worst[name])
        load(2 * n + 1 + i, o.personal[name])
    load(2 * n, o.global_critic)
    drive(alg, out, "coma", (1, 2, 3), eq)
    for i, name in enumerate(NAMES):
        dump(out, "coma_best_" + name, i)
        dump(out, "coma_worst_" + name, n + i)
        dump(out, "coma_personal_" + name, 2 * n + 1 + i)
    dump(out, "coma_global", 2 * n)
    # ---- the fork's Maddpg: policies[names], critics[names]; hyperparameters=None dies, a given dict is replaced by {}
    tf_shim.reset()
    try:
        Maddpg(obs_space, act_space)
        out["maddpg_none_hyperparameters_error"] = np.asarray("")
    except AttributeError as e:
        out["maddpg_none_hyperparameters_error"] = np.asarray(str(e))
    tf_shim.reset()
    alg = Maddpg(obs_space, act_space, hyperparameters={"gamma": 0.5})
    _ = alg.session
    assert len(tf_shim._VARIABLES) == 2 * n * 12
    o = MaddpgOracle(OBS, ACT, LOW, HIGH, seed=62, first=NAMES[0])
    for i, name in enumerate(NAMES):
        load(i, o.policies[name])
        load(n + i, o.critics[name])
    drive(alg, out, "maddpg", (1, 2, 3))
    for i, name in enumerate(NAMES):
        dump(out, "maddpg_policy_" + name, i)
        dump(out, "maddpg_critic_" + name, n + i)
    # ---- shared groups (PolicyGroup / CriticGroup(shared=True)): ONE policy and ONE critic LaggingNetwork, equal spaces
    tf_shim.reset()
    alg = Maddpg(eq_obs_space, eq_act_space, shared_policy=True, shared_critic=True, hyperparameters={"x": 1})
    _ = alg.session
    assert len(tf_shim._VARIABLES) == 2 * 12, len(tf_shim._VARIABLES)
    o = MaddpgOracle(*eq, seed=64, shared_policy=True, shared_critic=True, first=NAMES[0])
    load(0, o.policies[NAMES[0]])
    load(1, o.critics[NAMES[0]])
    drive(alg, out, "maddpg_shared", (1, 2, 3), eq)
    dump(out, "maddpg_shared_policy", 0)
    dump(out, "maddpg_shared_critic", 1)
    tf_shim.reset()
    alg = Coma(eq_obs_space, eq_act_space, shared_policy=True)      # best, worst: one policy each; global critic; personal[names]
    _ = alg.session
    assert len(tf_shim._VARIABLES) == (3 + n) * 12, len(tf_shim._VARIABLES)
    o = ComaOracle(*eq, seed=65, first=NAMES[0], shared_policy=True)
    load(0, o.best[NAMES[0]])
    load(1, o.worst[NAMES[0]])
    load(2, o.global_critic)
    for i, name in enumerate(NAMES):
        load(3 + i, o.personal[name])
    drive(alg, out, "coma_shared", (1, 2), eq)
    dump(out, "coma_shared_best", 0)
    dump(out, "coma_shared_worst", 1)
    # ---- MaTd3: the reference's graph cannot be built ...
    tf_shim.reset()
    try:
        MaTd3(obs_space, act_space)
        out["matd3_error"] = np.asarray("")
    except TypeError as e:
        out["matd3_error"] = np.asarray(str(e))
    # ... unless the one call that kills it is made to tolerate its extra argument: PolicyGroup.create_optimizers(values, entropy)
    # (matd3module.py:98-99 against policygroup.py:123).  The ONLY modification of reference code in this file: the argument is
    # dropped, everything else of MaTD3Module / MaTd3 runs as written.  LaggingNetworks: policies[names], critics 1, critics 2.
    import maddpg.modules.policygroup as policygroup_module
    original = policygroup_module.PolicyGroup.create_optimizers
    policygroup_module.PolicyGroup.create_optimizers = lambda self, values, *dropped: original(self, values)
    from oracle.matd3 import MaTd3Oracle
    tf_shim.reset()
    alg = MaTd3(obs_space, act_space)
    _ = alg.session
    assert len(tf_shim._VARIABLES) == 3 * n * 12
    o = MaTd3Oracle(OBS, ACT, LOW, HIGH, seed=63)
    for i, name in enumerate(NAMES):
        load(i, o.policies[name])
        load(n + i, o.critics[0][name])
        load(2 * n + i, o.critics[1][name])
    queue = []

    def normal(shape):      # tf.random.normal of the noisy targets: one (B, K_name) draw per policy, sorted-name order
        z = queue.pop(0)
        assert tuple(shape) == z.shape, (shape, z.shape)
        return z
    tf_shim.NORMAL[0] = normal
    for step in (1, 2, 3, 4):
        obs, act, rew, obs_n, done, z = make_batch(48, 3000 + step)
        queue[:] = [z[name] for name in sorted(NAMES)]
        res = alg.train_step(obs, act, rew, obs_n, done, step)
        assert not queue
        alg.run_updates()
        if step % 2 == 0:
            out["matd3_s%d_actor" % step] = np.asarray([res["actor"][nm] for nm in NAMES], np.float64)
            out["matd3_s%d_critic" % step] = np.asarray([res["critic"][nm] for nm in NAMES], np.float64)
        else:   # critic-only steps: unflatten_map splits the un-prefixed names at "_" (matd3.py:71-72, utils_common.py:69-77)
            out["matd3_s%d_raw_keys" % step] = np.asarray(sorted(res))
            out["matd3_s%d_critic" % step] = np.asarray([res[nm.split("_", 1)[0]][nm.split("_", 1)[1]] for nm in NAMES], np.float64)
    obs = make_batch(16, 2000)[0]
    pred, val = alg.predict(obs, noisy=False), alg.compute_values(obs)
    for nm in NAMES:
        out["matd3_predict_%s" % nm] = np.asarray(pred[nm], np.float64).reshape(16, -1)
        out["matd3_values_%s" % nm] = np.asarray(val[nm], np.float64).reshape(16)
    for i, name in enumerate(NAMES):
        dump(out, "matd3_policy_" + name, i)
        dump(out, "matd3_critic0_" + name, n + i)
        dump(out, "matd3_critic1_" + name, 2 * n + i)
    policygroup_module.PolicyGroup.create_optimizers = original
    np.savez_compressed(os.path.join(HERE, "fork_graph_ref.npz"), **out)
    print("wrote fork_graph_ref.npz: %d arrays" % len(out))
    print("coma step 1 actor", out["coma_s1_actor"], "critic", out["coma_s1_critic"])
    print("MaTd3():", out["matd3_error"], "| Maddpg(hyperparameters=None):", out["maddpg_none_hyperparameters_error"],
          "| Coma(unequal spaces):", out["coma_unequal_spaces_error"])


if __name__ == "__main__":
    main()
