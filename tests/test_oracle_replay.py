"""oracle/replay.py against the golden vectors produced by the REAL reference ReplayBuffer
(tests/golden/make_replay_golden.py, generated in the build container from /root/reference)."""
import os
import random

import numpy as np
import pytest

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


REF_FILE = "/root/reference/maddpg/trainer/replay_buffer.py"


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("seed", range(10))
def test_oracle_matches_real_class_on_random_scripts(seed):
    """Beyond the committed goldens: the REAL reference ReplayBuffer executed side by side with the restatement on random
    capacities and random add / make_index / sample / make_latest_index / clear scripts -- same python `random` and numpy
    streams, same cursors, lengths, index lists and sampled arrays.  (Build container only.)"""
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from maddpg.trainer.replay_buffer import ReplayBuffer as Real
    finally:
        sys.path.remove("/root/reference")
    rng = np.random.RandomState(seed)
    cap = int(rng.choice([1, 2, 3, 7, 16, 33, 100]))
    ref, orc = Real(cap), ReplayBuffer(cap)
    t = 0
    for step in range(60):
        op = rng.choice(["add", "index", "sample", "latest", "clear"], p=[0.5, 0.2, 0.15, 0.1, 0.05])
        if op == "add":
            for _ in range(int(rng.randint(1, 2 * cap + 2))):
                row = (rng.randn(3), rng.rand(2).astype(np.float32), float(t), rng.randn(3), float(t % 2))
                ref.add(*row)
                orc.add(*row)
                t += 1
        elif len(ref) > 0 and op == "index":
            B = int(rng.randint(1, 20))
            sd = int(rng.randint(1 << 30))
            random.seed(sd)
            i0 = ref.make_index(B)
            random.seed(sd)
            i1 = orc.make_index(B)
            assert list(i0) == list(i1)
            for a, b in zip(ref.sample_index(i0), orc.sample_index(i1)):
                assert np.array_equal(a, b) and a.dtype == b.dtype and a.shape == b.shape
        elif len(ref) > 0 and op == "sample":
            sd = int(rng.randint(1 << 30))
            arg = int(rng.choice([-1, 1, 5]))  # -1: collect() semantics (every stored row)
            random.seed(sd)
            s0 = ref.sample(arg)
            random.seed(sd)
            s1 = orc.sample(arg)
            for a, b in zip(s0, s1):
                assert np.array_equal(a, b)
        elif len(ref) > 0 and op == "latest":
            B = int(rng.randint(1, cap + 3))
            sd = int(rng.randint(1 << 30))
            np.random.seed(sd)
            l0 = ref.make_latest_index(B)
            np.random.seed(sd)
            l1 = orc.make_latest_index(B)
            assert list(l0) == list(l1)
        elif op == "clear":
            ref.clear()
            orc.clear()
        assert len(ref) == len(orc) and ref._next_idx == orc._next_idx
