"""maddpg_b200.algorithms.DictReplayBuffer against the REAL reference class (maddpg/common/replaybuffer.py:10-103): the golden
tests/golden/dict_replay_ref.npz was recorded by tests/golden/make_dict_replay_golden.py driving the reference's own code under
fixed ``random`` / ``numpy.random`` seeds; the same script replayed here must reproduce every sample, index set, length and
cursor exactly."""
import os
import random

import numpy as np

NAMES = ["b", "a"]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dict_replay_ref.npz")


def transition(t):
    return ({n: np.array([t, 10 * t + i], np.float32) for i, n in enumerate(NAMES)},
            {n: np.array([-t - i], np.float32) for i, n in enumerate(NAMES)},
            {n: float(t) for n in NAMES},
            {n: np.array([t + 1, 10 * (t + 1) + i], np.float32) for i, n in enumerate(NAMES)},
            {n: bool(t % 5 == 4) for n in NAMES})


def test_dict_replay_matches_the_reference_class():
    from maddpg_b200.algorithms import DictReplayBuffer
    gold = np.load(GOLD)
    ops = {0: "add", 1: "sample", 2: "latest", 3: "collect", 4: "clear"}
    rb = DictReplayBuffer(7)
    random.seed(123)
    np.random.seed(321)
    t = 0
    checked = 0
    for step, (code, arg) in enumerate(gold["script"]):
        op, key = ops[int(code)], "s%d" % step
        if op == "add":
            for _ in range(int(arg)):
                rb.add(*transition(t))
                t += 1
        elif op in ("sample", "collect"):
            res = rb.sample(int(arg)) if op == "sample" else rb.collect()
            for f, d in zip(("obs", "act", "rew", "obs2", "done"), res):
                assert list(d) == NAMES      # key order of the stored dicts, like zip_map over the first field
                for n in NAMES:
                    assert np.array_equal(np.asarray(d[n], np.float64), gold["%s_%s_%s" % (key, f, n)]), (key, f, n)
                    checked += 1
        elif op == "latest":
            assert np.array_equal(np.asarray(rb.make_latest_index(int(arg)), np.int64), gold[key + "_latest"])
        elif op == "clear":
            rb.clear()
        assert len(rb) == int(gold[key + "_len"]) and rb._next_idx == int(gold[key + "_next"]), key
    assert checked == 50
