"""Generates tests/golden/replay_ref.npz by driving the REAL reference ReplayBuffer
(/root/reference/maddpg/trainer/replay_buffer.py) -- run in the build container only:

    python tests/golden/make_replay_golden.py

The fixture pins oracle/replay.py (and through it the CUDA ring) to the reference's own outputs:
a ring of capacity 37 receives 100 transitions (so it wraps), then python ``random`` seeded with
1234 draws index sets through the reference's own ``make_index``.
"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from maddpg.trainer.replay_buffer import ReplayBuffer as RefReplayBuffer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.RandomState(7)
    D, K, CAP, N, B = 18, 5, 37, 100, 64
    obs = rng.randn(N, D).astype(np.float32).astype(np.float64)  # float32-representable values
    act = rng.rand(N, K).astype(np.float32)
    rew = rng.randn(N).astype(np.float32).astype(np.float64)
    nobs = rng.randn(N, D).astype(np.float32).astype(np.float64)
    done = (rng.rand(N) < 0.1).astype(np.float64)
    rb = RefReplayBuffer(CAP)
    lens, nexts = [], []
    out = {}
    random.seed(1234)
    for t in range(N):
        rb.add(obs[t], act[t], float(rew[t]), nobs[t], float(done[t]))
        lens.append(len(rb))
        nexts.append(rb._next_idx)
        if t in (10, 36, 37, 60, 99):
            idx = rb.make_index(B)
            o, a, r, n2, d = rb.sample_index(idx)
            out["idx_%d" % t] = np.asarray(idx, np.int64)
            out["obs_%d" % t], out["act_%d" % t], out["rew_%d" % t] = o, a, r
            out["nobs_%d" % t], out["done_%d" % t] = n2, d
    np.random.seed(5)
    out["latest_idx"] = np.asarray(rb.make_latest_index(16), np.int64)
    o, a, r, n2, d = rb.collect()
    out["collect_obs"], out["collect_rew"] = o, r
    np.savez_compressed(os.path.join(HERE, "replay_ref.npz"), in_obs=obs, in_act=act, in_rew=rew, in_nobs=nobs,
                        in_done=done, lens=np.asarray(lens), nexts=np.asarray(nexts), cap=CAP, batch=B, **out)
    print("wrote replay_ref.npz", sorted(out)[:5], "...")


if __name__ == "__main__":
    main()
