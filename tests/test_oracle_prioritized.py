"""oracle/prioritized.py against the outputs of the REAL reference classes (tests/golden/prioritized_ref.npz, produced by
tests/golden/make_prioritized_golden.py from /root/reference/maddpg/trainer/prioritized_replay_buffer.py): every tree array,
tree index, data slot, IS weight and beta, bit for bit, including the samples the reference aborts with IndexError."""
import os

import numpy as np
import pytest

from oracle.prioritized import PrioritizedReplayOracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "prioritized_ref.npz")


def replay_script(gold, cap, make, add, sample, update, tree_of, isw_rtol=0.0):
    """Drives an implementation through the golden script of one capacity; shared with the GPU test."""
    script = gold["c%d_script" % cap]
    mem = make(cap)
    serial = 0
    slot_serial = {}
    ptr = 0
    for step, (kind, arg) in enumerate(script):
        key = "c%d_s%d" % (cap, step)
        if kind == 0:
            n = int(arg)
            add(mem, serial, n)
            for j in range(n):
                slot_serial[(ptr + j) % cap] = serial + j
            ptr = (ptr + n) % cap
            serial += n
        elif kind == 1:
            n = int(arg)
            u = gold[key + "_u"]
            if int(gold[key + "_err"]):
                with pytest.raises(IndexError):
                    sample(mem, n, u)
            else:
                b_idx, b_data, isw, beta = sample(mem, n, u)
                assert np.array_equal(np.asarray(b_idx, np.int64), gold[key + "_idx"]), key
                got_serial = np.asarray([slot_serial[int(d)] for d in b_data], np.float64)
                assert np.array_equal(got_serial, gold[key + "_serial"]), key
                if isw_rtol:
                    np.testing.assert_allclose(np.asarray(isw, np.float64), gold[key + "_isw"], rtol=isw_rtol, atol=0, err_msg=key)
                else:
                    np.testing.assert_array_equal(np.asarray(isw, np.float64), gold[key + "_isw"], err_msg=key)
                assert beta == float(gold[key + "_beta"]), key
        else:
            update(mem, gold[key + "_tidx"], gold[key + "_abs"])
        tree, pending = tree_of(mem)
        assert int(gold[key + "_ptr"]) == ptr
        if not pending:  # the reference's tree lags behind its dirty list; compare whenever nothing is pending
            assert int(gold[key + "_ndirty"]) == 0
        if int(gold[key + "_ndirty"]) == 0:
            assert not pending
            assert np.array_equal(tree, gold[key + "_tree"]), key
        else:
            assert np.array_equal(tree, gold[key + "_tree"]), key   # lazily flushed implementations match the stale tree too


@pytest.mark.parametrize("cap", [37, 64, 1000])
def test_oracle_matches_real_reference_class(cap):
    gold = np.load(GOLD)

    def add(mem, serial, n):
        for j in range(n):
            mem.add(serial + j)

    def sample(mem, n, u):
        b_idx, b_data, isw = mem.sample(n, u)
        return b_idx, b_data, isw, float(mem.beta)

    replay_script(gold, cap, PrioritizedReplayOracle, add, sample,
                  lambda mem, ti, ae: mem.batch_update(ti, ae.copy()),
                  lambda mem: (mem.tree.tree, mem.tree.dirty_count > 0))


def test_update_all_equals_sequential_updates_when_exact():
    """With exactly representable priorities the deferred flush and one update() per leaf give the same tree."""
    from oracle.prioritized import SumTreeOracle
    a, b = SumTreeOracle(100), SumTreeOracle(100)
    a.add(1e6, 130)
    a.update_all()
    for d in list(range(0, 100)):
        b.update(b.leaf_of(d), 1e6)
    assert np.array_equal(a.tree, b.tree)
