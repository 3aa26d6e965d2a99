"""Generates tests/golden/prioritized_ref.npz by driving the REAL reference classes
(/root/reference/maddpg/trainer/prioritized_replay_buffer.py: SumTree, PrioritizedReplayMemory) -- build container only:

    python tests/golden/make_prioritized_golden.py

The reference file imports tensorflow for one call (``tf.set_random_seed(1)``); a stand-in module satisfies the import, every
line of the two classes then runs unmodified.  A script of add / sample / batch_update calls is executed per capacity and
every observable is recorded: the whole tree array after each call, the returned tree indices, the data slots (recovered from
the stored rows), the IS weights, beta, and the ``random_sample()`` draws behind ``np.random.uniform`` (numpy's legacy
``uniform(a, b)`` is ``a + (b - a) * random_sample()``; the script asserts that bit for bit).  A sample() that dies with the
reference's own IndexError (descent through slot 0's node, oracle/prioritized.py header) is recorded as such.
"""
import os
import sys
import types

import numpy as np

tf_stub = types.ModuleType("tensorflow")
tf_stub.set_random_seed = lambda seed: None
sys.modules.setdefault("tensorflow", tf_stub)
sys.path.insert(0, "/root/reference")
from maddpg.trainer.prioritized_replay_buffer import PrioritizedReplayMemory  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_script(cap, script, seed):
    """script: list of ("add", n) | ("sample", n) | ("update",) -- update feeds back |errors| for the last sampled batch."""
    mem = PrioritizedReplayMemory(cap)
    rng = np.random.RandomState(seed)
    out = {}
    slot = 0          # running count of add() calls: the stored row's payload is its add serial number
    last_idx = None
    for step, op in enumerate(script):
        key = "c%d_s%d" % (cap, step)
        if op[0] == "add":
            for _ in range(op[1]):
                mem.add(np.float64(slot), np.float32(0.5), float(slot), np.float64(slot + 1), 0.0)
                slot += 1
        elif op[0] == "sample":
            n = op[1]
            sd = int(rng.randint(1 << 30))
            np.random.seed(sd)
            u = np.random.random_sample(n)          # the draws uniform() will consume
            np.random.seed(sd)
            beta0 = float(mem.beta)
            try:
                b_idx, b_mem, isw = mem.sample(n)
                err = 0
            except IndexError:
                b_idx, b_mem, isw, err = [], [[]], [], 1
            out[key + "_u"] = u
            out[key + "_err"] = np.int64(err)
            out[key + "_beta"] = np.float64(mem.beta)
            out[key + "_beta0"] = np.float64(beta0)
            if not err:
                # legacy uniform == a + (b - a) * random_sample(): replay the descent values through the real tree
                seg = None
                out[key + "_idx"] = np.asarray(b_idx, np.int64)
                out[key + "_serial"] = np.asarray(b_mem[2], np.float64)   # reward column = add serial of the stored row
                out[key + "_isw"] = np.asarray(isw, np.float64)
                last_idx = np.asarray(b_idx, np.int64)
        elif op[0] == "update":
            errs = np.abs(rng.randn(last_idx.size)) * op[1]
            out[key + "_abs"] = errs.copy()
            out[key + "_tidx"] = last_idx.copy()
            mem.batch_update(last_idx, errs)        # note: the reference adds epsilon IN PLACE to its argument
        out[key + "_tree"] = np.asarray(mem.tree.tree, np.float64)
        out[key + "_ptr"] = np.int64(mem.tree.data_pointer)
        out[key + "_ndirty"] = np.int64(len(mem.tree.dirty))
    return out


# (capacity, script).  37: not a power of two, wraps (k = 6); 64: power of two (slot cap-1 hangs under slot 0's node), wraps;
# 1000: many levels, duplicate leaves inside one batch_update (48 draws from few distinct leaves early on)
SCRIPTS = {
    37: [("add", 20), ("sample", 8), ("update", 0.3), ("sample", 8), ("update", 2.0), ("add", 30), ("sample", 16),
         ("update", 0.5), ("add", 5), ("update", 0.1), ("sample", 16), ("sample", 37), ("update", 0.7), ("add", 80),
         ("sample", 12), ("update", 0.2), ("sample", 12)],
    64: [("add", 10), ("sample", 4), ("update", 0.5), ("add", 54), ("sample", 32), ("update", 0.4), ("sample", 32),
         ("update", 1.5), ("add", 70), ("sample", 64), ("update", 0.3), ("sample", 64), ("update", 0.3), ("sample", 16)],
    # odd k (k = 7 and k = 3): the flush hands the root the wrong delta when slot 0 is among the pending adds
    100: [("add", 60), ("sample", 16), ("update", 0.4), ("add", 61), ("sample", 32), ("update", 0.8), ("add", 100), ("sample", 50),
          ("update", 0.3), ("add", 7), ("sample", 20)],
    5: [("add", 3), ("sample", 2), ("update", 0.5), ("add", 4), ("sample", 4), ("update", 0.2), ("add", 12), ("sample", 5)],
    1000: [("add", 300), ("sample", 48), ("update", 0.6), ("sample", 48), ("update", 0.6), ("add", 900), ("sample", 256),
           ("update", 0.5), ("sample", 256), ("update", 0.05), ("add", 123), ("sample", 256), ("update", 1.0),
           ("sample", 100)],
}


def main():
    out = {}
    for cap, script in SCRIPTS.items():
        out.update(run_script(cap, script, seed=cap))
        out["c%d_script" % cap] = np.asarray([[{"add": 0, "sample": 1, "update": 2}[op[0]], op[1] if len(op) > 1 else 0]
                                               for op in script], dtype=np.float64)
    # legacy uniform(a, b) == a + (b - a) * random_sample(), bit for bit
    np.random.seed(3)
    u = np.random.random_sample(1000)
    np.random.seed(3)
    a, b = 1234.5678 * np.arange(1000), 1234.5678 * (np.arange(1000) + 1)
    v = np.asarray([np.random.uniform(x, y) for x, y in zip(a, b)])
    assert np.array_equal(v, a + (b - a) * u)
    np.savez_compressed(os.path.join(HERE, "prioritized_ref.npz"), **out)
    errs = {k: int(v) for k, v in out.items() if k.endswith("_err")}
    print("wrote prioritized_ref.npz: %d arrays; IndexError samples: %s" % (len(out), [k for k, v in errs.items() if v]))


if __name__ == "__main__":
    main()
