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


@pytest.mark.parametrize("cap", [37, 64, 100, 5, 1000])
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
    """With exactly representable priorities the deferred flush and one update() per leaf give the same tree -- for EVEN k.  For odd
    k the reference's flush hands the root slot 0's delta twice and drops the right subtree's (SumTreeOracle.update_all)."""
    from oracle.prioritized import SumTreeOracle
    a, b = SumTreeOracle(50), SumTreeOracle(50)  # k = 6
    a.add(1e6, 70)
    a.update_all()
    for d in range(50):
        b.update(b.leaf_of(d), 1e6)
    assert np.array_equal(a.tree, b.tree)
    a, b = SumTreeOracle(100), SumTreeOracle(100)  # k = 7: slots 1..64 hang under node 1, slots 65..99 and slot 0 under node 2
    a.add(1e6, 100)
    a.update_all()
    for d in range(100):
        b.update(b.leaf_of(d), 1e6)
    assert b.tree[0] == 100e6 and a.tree[0] == 1e6 + (1e6 + 64e6)
    assert np.array_equal(a.tree[1:], b.tree[1:])


REF_FILE = "/root/reference/maddpg/trainer/prioritized_replay_buffer.py"


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("seed", range(40))
def test_oracle_matches_real_class_on_random_scripts(seed):
    """Beyond the committed goldens: random capacities (incl. powers of two and capacity 3) and random add / sample /
    batch_update scripts, the REAL reference class executed side by side with the restatement -- tree arrays, tree indices,
    IS weights and IndexError behaviour bit for bit.  (Build container only: the GPU box has no /root/reference.)"""
    import sys
    import types
    stub = types.ModuleType("tensorflow")
    stub.set_random_seed = lambda s: None
    had = "tensorflow" in sys.modules
    sys.modules.setdefault("tensorflow", stub)
    sys.path.insert(0, "/root/reference")
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from maddpg.trainer.prioritized_replay_buffer import PrioritizedReplayMemory
    finally:
        sys.path.remove("/root/reference")
        if not had:
            sys.modules.pop("tensorflow", None)
    rng = np.random.RandomState(100 + seed)
    cap = int(rng.choice([3, 4, 7, 8, 16, 33, 64, 100, 129, 200]))
    ref, orc = PrioritizedReplayMemory(cap), PrioritizedReplayOracle(cap)
    serial, last = 0, None
    with np.errstate(all="ignore"):
        for step in range(40):
            op = rng.choice(["add", "sample", "update"], p=[0.4, 0.35, 0.25])
            if op == "add":
                for _ in range(int(rng.randint(1, 2 * cap))):
                    ref.add(serial, 0, float(serial), serial + 1, 0.0)
                    orc.add(serial)
                    serial += 1
            elif op == "sample" and serial > 0:
                n = int(rng.randint(1, 3 * cap))
                sd = int(rng.randint(1 << 30))
                np.random.seed(sd)
                u = np.random.random_sample(n)
                np.random.seed(sd)
                try:
                    b_idx, b_mem, isw = ref.sample(n)
                    err = False
                except (IndexError, TypeError):  # slot 0's node: data index out of range (or a None row: zip(*) fails)
                    err = True
                if err:
                    try:
                        orc.sample(n, u)
                        # the restatement raises IndexError where data_idx >= capacity; a None row (never-written slot)
                        # makes the reference fail later, in zip(*b_memory): also an abort
                        got_rows = True
                    except IndexError:
                        got_rows = False
                    assert (not got_rows) or any(r is None for r in ref.tree.data), step
                else:
                    o_idx, o_data, o_isw = orc.sample(n, u)
                    assert list(b_idx) == list(o_idx), step
                    np.testing.assert_array_equal(np.asarray(isw, np.float64), np.asarray(o_isw, np.float64))
                    last = np.asarray(b_idx, np.int64)
                assert float(ref.beta) == float(orc.beta)
            elif op == "update" and last is not None:
                errs = np.abs(rng.randn(last.size)) * float(rng.choice([0.05, 0.5, 3.0]))
                ref.batch_update(last, errs.copy())
                orc.batch_update(last, errs.copy())
            if not ref.tree.dirty:
                orc.tree.update_all()
            if orc.tree.dirty_count == 0:
                assert np.array_equal(np.asarray(ref.tree.tree, np.float64), orc.tree.tree), (cap, step, op)
