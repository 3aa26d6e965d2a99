"""Generates tests/golden/dict_replay_ref.npz by driving the REAL reference class (/root/reference/maddpg/common/replaybuffer.py:
ReplayBuffer, the fork's dict-of-agents replay) -- build container only:

    python tests/golden/make_dict_replay_golden.py

The class's module imports maddpg.common.utils_common, which imports tensorflow, tqdm and gym.spaces at module level without the
replay using any of them; stand-in modules satisfy those imports, every line of the class then runs unmodified.  A script of
add / sample / make_latest_index / collect calls is executed under fixed ``random`` / ``numpy.random`` seeds and every observable
recorded.
"""
import os
import random
import sys
import types

import numpy as np

for name in ("tensorflow", "tqdm", "gym", "gym.spaces"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["tqdm"].tqdm = object
sys.modules["tqdm"].trange = range
for cls in ("Box", "Discrete", "Dict"):
    setattr(sys.modules["gym.spaces"], cls, type(cls, (), {}))
sys.modules["gym"].spaces = sys.modules["gym.spaces"]
sys.path.insert(0, "/root/reference")
import importlib.util  # noqa: E402

pkg = types.ModuleType("maddpg")
pkg.__path__ = ["/root/reference/maddpg"]
sys.modules["maddpg"] = pkg
common = types.ModuleType("maddpg.common")
common.__path__ = ["/root/reference/maddpg/common"]
sys.modules["maddpg.common"] = common
for mod in ("utils_common", "replaybuffer"):
    spec = importlib.util.spec_from_file_location("maddpg.common." + mod, "/root/reference/maddpg/common/%s.py" % mod)
    m = importlib.util.module_from_spec(spec)
    sys.modules["maddpg.common." + mod] = m
    spec.loader.exec_module(m)
ReplayBuffer = sys.modules["maddpg.common.replaybuffer"].ReplayBuffer

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["b", "a"]


def transition(t):
    return ({n: np.array([t, 10 * t + i], np.float32) for i, n in enumerate(NAMES)},
            {n: np.array([-t - i], np.float32) for i, n in enumerate(NAMES)},
            {n: float(t) for n in NAMES},
            {n: np.array([t + 1, 10 * (t + 1) + i], np.float32) for i, n in enumerate(NAMES)},
            {n: bool(t % 5 == 4) for n in NAMES})


def main():
    out = {}
    rb = ReplayBuffer(7)
    random.seed(123)
    np.random.seed(321)
    t = 0
    script = [("add", 4), ("sample", 6), ("add", 5), ("sample", 9), ("latest", 3), ("add", 11), ("sample", 8), ("collect",),
              ("clear",), ("add", 2), ("sample", 5)]
    for step, op in enumerate(script):
        key = "s%d" % step
        if op[0] == "add":
            for _ in range(op[1]):
                rb.add(*transition(t))
                t += 1
        elif op[0] in ("sample", "collect"):
            res = rb.sample(op[1]) if op[0] == "sample" else rb.collect()
            for f, d in zip(("obs", "act", "rew", "obs2", "done"), res):
                for n in NAMES:
                    out["%s_%s_%s" % (key, f, n)] = np.asarray(d[n], np.float64)
        elif op[0] == "latest":
            out[key + "_latest"] = np.asarray(rb.make_latest_index(op[1]), np.int64)
        elif op[0] == "clear":
            rb.clear()
        out[key + "_len"] = np.int64(len(rb))
        out[key + "_next"] = np.int64(rb._next_idx)
    out["script"] = np.asarray([[{"add": 0, "sample": 1, "latest": 2, "collect": 3, "clear": 4}[op[0]], op[1] if len(op) > 1 else 0]
                                for op in script], np.int64)
    np.savez_compressed(os.path.join(HERE, "dict_replay_ref.npz"), **out)
    print("wrote dict_replay_ref.npz: %d arrays" % len(out))


if __name__ == "__main__":
    main()
