"""Generates tests/golden/update_orchestration_ref.npz by running the REAL ``MADDPGAgentTrainer.update`` / ``experience`` /
``preupdate`` / ``action`` methods (/root/reference/maddpg/trainer/maddpg.py:151-196) -- build container only:

    python tests/golden/make_update_golden.py

What the reference computes in TensorFlow (the six graph callables q_train, p_train, p_update, q_update, p_debug['target_act'],
q_debug['target_q_values'], and act) is supplied by oracle/maddpg.py's restated networks; everything AROUND them is the reference's
own code, executed unmodified: the warm-up and every-100-steps gates, the index draw through the real ReplayBuffer (python
``random``), the per-agent gathers, the float64 numpy TD combine ``rew + gamma * (1 - done) * target_q_next``, the call order
(q_train, p_train, p_update, q_update) and the six returned statistics.  ``__init__`` (graph construction) is bypassed with
``object.__new__``; permissive stand-in ``tensorflow`` / ``gym`` / ``tqdm`` modules satisfy the imports.

tests/test_oracle_maddpg.py::test_update_orchestration_matches_the_reference_method holds ``OracleAgentTrainer.update`` (the method
every GPU update-round test is compared with) to this file bit for bit.
"""
import os
import random
import sys
import types

import numpy as np


class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return self


for name in ("tensorflow", "tensorflow.python", "tensorflow.python.ops", "tensorflow.contrib", "tensorflow.contrib.layers", "gym",
             "gym.spaces", "tqdm"):
    sys.modules[name] = _Any(name)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from maddpg.trainer.maddpg import MADDPGAgentTrainer  # noqa: E402  (the REAL class)
from maddpg.trainer.replay_buffer import ReplayBuffer  # noqa: E402  (the REAL class)

from tests.update_case import N, T_SEQUENCE, build_oracle_trainers, transition  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def shim(o):
    """A reference trainer object without its TF graph: the REAL methods, oracle-backed graph callables."""
    s = object.__new__(MADDPGAgentTrainer)
    s.name, s.n, s.agent_index, s.args = o.name, o.n, o.agent_index, o.args
    s.replay_buffer = ReplayBuffer(1e6)
    s.max_replay_buffer_len = o.args.batch_size * o.args.max_episode_len
    s.replay_sample_index = None
    n = o.n
    s.q_train = lambda *a: o.q_train(list(a[:n]), list(a[n:2 * n]), a[2 * n])
    s.p_train = lambda *a: o.p_train(list(a[:n]), list(a[n:2 * n]))
    s.p_update, s.q_update = o.p_update, o.q_update
    s.p_debug, s.q_debug, s.act = o.p_debug, o.q_debug, o.act
    return s


def main():
    oracles = build_oracle_trainers()
    agents = [shim(o) for o in oracles]
    random.seed(11)
    out, rows = {}, 0
    for step, t in enumerate(T_SEQUENCE):
        for _ in range(40):      # 40 more transitions before every update attempt
            tr = transition(rows)
            for i, a in enumerate(agents):
                a.experience(tr["obs"][i], tr["act"][i], tr["rew"][i], tr["obs2"][i], tr["done"][i], False)
            rows += 1
        for a in agents:
            a.preupdate()
        for i, a in enumerate(agents):
            res = a.update(agents, t)
            key = "s%d_a%d" % (step, i)
            out[key + "_stats"] = np.full(6, np.nan) if res is None else np.asarray(res, np.float64)
            out[key + "_index"] = np.asarray([] if a.replay_sample_index is None else a.replay_sample_index, np.int64)
        out["s%d_params" % step] = np.asarray([float(np.sum([np.sum(p.astype(np.float64)) for net in (o.q, o.target_q, o.p, o.target_p)
                                                             for p in net.p])) for o in oracles])
    obs = transition(999)["obs"]
    out["action"] = np.concatenate([np.asarray(a.action(obs[i]), np.float64) for i, a in enumerate(agents)])
    np.savez_compressed(os.path.join(HERE, "update_orchestration_ref.npz"), **out)
    done = [k for k in out if k.endswith("_stats") and not np.isnan(out[k][0])]
    print("wrote update_orchestration_ref.npz: %d arrays, %d updates ran (%s)" % (len(out), len(done), done[:N]))


if __name__ == "__main__":
    main()
