"""Times one MATD3 / COMA train step (maddpg_b200/algorithms.py) on the device next to the numpy oracle on the host.
usage: python tools/time_td3.py [B] [n_agents]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maddpg_b200 import _lib  # noqa: E402
from maddpg_b200.algorithms import Coma, MaTd3  # noqa: E402
from maddpg_b200.spaces import Box, Dict  # noqa: E402
from oracle.matd3 import ComaOracle, MaTd3Oracle  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
names = ["agent_%d" % i for i in range(n)]
D, K = 18, 2
obs_sp = Dict({k: Box(-np.inf, np.inf, (D,)) for k in names})
act_sp = Dict({k: Box(-np.ones(K, np.float32), np.ones(K, np.float32), (K,)) for k in names})
rng = np.random.RandomState(0)
feed = [{k: rng.randn(B, D).astype(np.float32) for k in names}, {k: rng.uniform(-1, 1, (B, K)).astype(np.float32) for k in names},
        {k: rng.randn(B, 1).astype(np.float32) for k in names}, {k: rng.randn(B, D).astype(np.float32) for k in names},
        {k: (rng.rand(B, 1) < 0.1).astype(np.float32) for k in names}]
for cls, ocls in ((MaTd3, MaTd3Oracle), (Coma, ComaOracle)):
    alg = cls(obs_sp, act_sp, seed=0)
    rows = alg._rows(*feed)
    for _ in range(5):
        alg._train_step(rows, 2, None, True)
        alg.run_updates()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 100
    e0.record()
    for _ in range(iters):
        alg._train_step(rows, 2, None, True)
        alg.run_updates()
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / iters
    launches = (_lib.launch_count() - l0) / iters
    t0 = time.perf_counter()
    for _ in range(20):
        alg.train_step(*feed, step=2)
        alg.run_updates()
    torch.cuda.synchronize()
    api_ms = (time.perf_counter() - t0) / 20 * 1e3
    o = ocls({k: D for k in names}, {k: K for k in names}, {k: -1.0 for k in names}, {k: 1.0 for k in names}, seed=0)
    z = {k: rng.randn(B, K).astype(np.float32) for k in names}
    kw = {"z": z} if cls is MaTd3 else {}
    o.train_step(*feed, step=2, **kw)
    t0 = time.perf_counter()
    for _ in range(5):
        o.train_step(*feed, step=2, **kw)
        o.run_updates()
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    print("%s B=%d n=%d: device %.3f ms/step (%.1f launches), through train_step() with host dicts %.3f ms, numpy oracle %.1f ms"
          % (cls.__name__, B, n, dev_ms, launches, api_ms, cpu_ms))
