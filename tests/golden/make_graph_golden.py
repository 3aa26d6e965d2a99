"""Generates tests/golden/graph_ref.npz by executing the reference's OWN graph-building code -- ``MADDPGAgentTrainer.__init__``,
``q_train``, ``p_train``, ``make_update_exp`` (maddpg/trainer/maddpg.py:20-149), ``SoftCategoricalPd`` / ``make_pdtype``
(maddpg/common/distributions.py), ``U.function`` / ``U.scope_vars`` / ``U.minimize_and_clip`` / ``U.BatchInput`` / sessions
(maddpg/common/tf_util.py) and ``mlp_model`` (experiments/train.py:39-46), all unmodified -- on tests/tf_shim.py, a torch-backed
stand-in for the slice of TensorFlow 1.x those files use.  Build container only:

    python tests/golden/make_graph_golden.py

TensorFlow itself cannot run here, so this is not a run of the reference on TensorFlow: the primitive ops (dense layer, softmax,
clip_by_norm, Adam, autograd) are the shim's, restated from TensorFlow's documentation.  Everything the REFERENCE's code decides is
executed for real: the graph wiring of the centralized / local critic inputs, the loss expressions, which variables each optimizer
owns, where the clip sits, the sorted-name pairing of the polyak update, what every ``U.function`` feeds and fetches, and -- on top
-- the real ``update`` method and ReplayBuffer.  oracle/maddpg.py's restated trainer must reproduce the recorded statistics,
debug outputs and post-update variables (tests/test_oracle_maddpg.py::test_oracle_matches_the_reference_graph_code).
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import tf_shim  # noqa: E402

tf = tf_shim.install()
gym, spaces = types.ModuleType("gym"), types.ModuleType("gym.spaces")


class Discrete(object):
    def __init__(self, n):
        self.n = n


spaces.Discrete, spaces.Box, spaces.MultiBinary, spaces.Dict = Discrete, type("Box", (), {}), type("MultiBinary", (), {}), type("Dict", (), {})
gym.spaces = spaces
sys.modules.update({"gym": gym, "gym.spaces": spaces, "tqdm": types.ModuleType("tqdm")})
sys.modules["tqdm"].tqdm, sys.modules["tqdm"].trange = (lambda it, **k: it), range
sys.path.insert(0, "/root/reference")
import maddpg.common.tf_util as U  # noqa: E402  (REAL)
from maddpg.trainer.maddpg import MADDPGAgentTrainer  # noqa: E402  (REAL)

spec = importlib.util.spec_from_file_location("reference_train", "/root/reference/experiments/train.py")
reference_train = importlib.util.module_from_spec(spec)
spec.loader.exec_module(reference_train)     # REAL mlp_model (train.py:39-46)

from tests.update_case import N, OBS_DIMS, build_oracle_trainers, make_args, shared_noise, transition  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NETS = (("q_func", "q"), ("target_q_func", "target_q"), ("p_func", "p"), ("target_p_func", "target_p"))
LAYERS = ("fully_connected", "fully_connected_1", "fully_connected_2")


def variables_of(agent, scope):
    byname = {v.op.name: v for v in tf_shim._VARIABLES}
    out = []
    for layer in LAYERS:
        out += [byname["agent_%d/%s/%s/weights" % (agent, scope, layer)], byname["agent_%d/%s/%s/biases" % (agent, scope, layer)]]
    return out


def main():
    act_space_n = [Discrete(5)] * N
    oracles = build_oracle_trainers(act_space_n=act_space_n)      # only as the source of the initial weights
    args = make_args()
    obs_shape_n = [(d,) for d in OBS_DIMS]
    out = {}
    with U.single_threaded_session():
        agents = [MADDPGAgentTrainer("agent_%d" % i, reference_train.mlp_model, obs_shape_n, act_space_n, i, args,
                                     local_q_func=(i == 2)) for i in range(N)]
        U.initialize()
        assert len(tf_shim._VARIABLES) == N * 4 * 6, [v.name for v in tf_shim._VARIABLES]
        for i, o in enumerate(oracles):
            for scope, attr in NETS:
                for v, w in zip(variables_of(i, scope), getattr(o, attr).p):
                    v.load(w)
        tf_shim.NOISE[0] = shared_noise()
        random.seed(11)
        for k in range(120):
            tr = transition(k)
            for i, a in enumerate(agents):
                a.experience(tr["obs"][i], tr["act"][i], tr["rew"][i], tr["obs2"][i], tr["done"][i], False)
        # debug surfaces on a fixed batch (no optimizer step): p_values, q_values, target_q_values, act, target_act
        batch = [transition(500 + k) for k in range(10)]
        obs_n = [np.asarray([b["obs"][i] for b in batch]) for i in range(N)]
        act_n = [np.asarray([b["act"][i] for b in batch]) for i in range(N)]
        for i, a in enumerate(agents):
            out["a%d_p_values" % i] = np.asarray(a.p_debug["p_values"](obs_n[i]), np.float64)
            out["a%d_q_values" % i] = np.asarray(a.q_debug["q_values"](*(obs_n + act_n)), np.float64)
            out["a%d_target_q_values" % i] = np.asarray(a.q_debug["target_q_values"](*(obs_n + act_n)), np.float64)
            out["a%d_act" % i] = np.asarray(a.act(obs_n[i]), np.float64)
            out["a%d_target_act" % i] = np.asarray(a.p_debug["target_act"](obs_n[i]), np.float64)
        for rnd, t in enumerate((100, 200)):
            for a in agents:
                a.preupdate()
            for i, a in enumerate(agents):
                out["r%d_a%d_stats" % (rnd, i)] = np.asarray(a.update(agents, t), np.float64)
            for i in range(N):
                for scope, attr in NETS:
                    for k, v in enumerate(variables_of(i, scope)):
                        out["r%d_a%d_%s_%d" % (rnd, i, attr, k)] = v.numpy()
        out["variable_names"] = np.asarray([v.name for v in tf_shim._VARIABLES])
        # the multi-head sample (simple_world_comm's leader): SoftMultiCategoricalPd (distributions.py:305-336), the class this
        # fork's make_pdtype no longer reaches (:416-418) -- one random_uniform per head, in head order
        from maddpg.common.distributions import SoftMultiCategoricalPdType, make_pdtype
        rng = np.random.RandomState(9)
        logits, u = rng.randn(12, 9).astype(np.float32), rng.uniform(size=(12, 9)).astype(np.float32)
        pdtype = SoftMultiCategoricalPdType(np.array([0, 0]), np.array([4, 3]))
        flat = pdtype.param_placeholder([None])
        sample = pdtype.pdfromflat(flat).sample()
        pieces = [u[:, :5], u[:, 5:]]

        def per_head(shape):
            z = pieces.pop(0)
            assert tuple(shape) == z.shape
            return z
        tf_shim.NOISE[0] = per_head
        out["multi_logits"], out["multi_u"] = logits, u
        out["multi_sample"] = np.asarray(tf.get_default_session().run(sample, feed_dict={flat: logits}), np.float64)
        try:
            make_pdtype(types.SimpleNamespace(low=np.array([0, 0]), high=np.array([4, 3])))
            out["make_pdtype_multidiscrete"] = np.asarray("")
        except NotImplementedError:
            out["make_pdtype_multidiscrete"] = np.asarray("NotImplementedError")
    np.savez_compressed(os.path.join(HERE, "graph_ref.npz"), **out)
    print("wrote graph_ref.npz: %d arrays; stats of round 0: %s" % (len(out), out["r0_a0_stats"]))


if __name__ == "__main__":
    main()
