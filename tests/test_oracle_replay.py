"""oracle/replay.py against the golden vectors produced by the REAL reference ReplayBuffer
(tests/golden/make_replay_golden.py, generated in the build container from /root/reference)."""
import os
import random

import numpy as np

from oracle.replay import ReplayBuffer

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "replay_ref.npz"))


def _drive(rb_add, rb_len, rb_next, make_index, sample_index):
    random.seed(1234)
    N = G["in_obs"].shape[0]
    for t in range(N):
        rb_add(G["in_obs"][t], G["in_act"][t], float(G["in_rew"][t]), G["in_nobs"][t], float(G["in_done"][t]))
        assert rb_len() == G["lens"][t]
        assert rb_next() == G["nexts"][t]
        if "idx_%d" % t in G.files:
            idx = make_index(int(G["batch"]))
            assert np.array_equal(np.asarray(idx), G["idx_%d" % t])  # same python MT19937 stream
            o, a, r, n2, d = sample_index(idx)
            for got, key in ((o, "obs"), (a, "act"), (r, "rew"), (n2, "nobs"), (d, "done")):
                ref = G["%s_%d" % (key, t)]
                assert got.shape == ref.shape
                assert np.array_equal(np.asarray(got, np.float64), ref), key


def test_oracle_matches_reference_golden():
    rb = ReplayBuffer(int(G["cap"]))
    _drive(rb.add, lambda: len(rb), lambda: rb._next_idx, rb.make_index, rb.sample_index)
    np.random.seed(5)
    assert np.array_equal(np.asarray(rb.make_latest_index(16)), G["latest_idx"])
    o, a, r, n2, d = rb.collect()
    assert np.array_equal(o, G["collect_obs"]) and np.array_equal(r, G["collect_rew"])


def test_append_until_full_then_wrap():
    rb = ReplayBuffer(3)
    for t in range(5):
        rb.add(np.full(2, t), np.full(1, t, np.float32), float(t), np.full(2, t + 1), 0.0)
    assert len(rb) == 3 and rb._next_idx == 2
    o, _, r, _, _ = rb.sample_index([0, 1, 2])
    assert r.tolist() == [3.0, 4.0, 2.0]
    rb.clear()
    assert len(rb) == 0 and rb._next_idx == 0
