"""The small trainer case shared by tests/golden/make_update_golden.py (which drives the REAL reference methods) and
tests/test_oracle_maddpg.py (which drives oracle/maddpg.py): 3 agents, one with a local critic, deterministic transitions."""
import types

import numpy as np

from oracle.maddpg import OracleAgentTrainer

N = 3
OBS_DIMS = [6, 5, 4]
T_SEQUENCE = [100, 100, 150, 200, 300]   # 40, 80 rows: warm-up gate (needs 96); 120 rows at t = 150: period gate; then two updates


class _Discrete(object):
    def __init__(self, n):
        self.n = n


def make_args():
    return types.SimpleNamespace(lr=1e-2, gamma=0.95, batch_size=24, num_units=16, max_episode_len=4)


def shared_noise(seed=77):
    """One U[0,1) float32 stream for every Gumbel draw of a run (what a single TF graph-level generator would be)."""
    rng = np.random.RandomState(seed)
    return lambda shape: np.minimum(rng.uniform(size=shape).astype(np.float32), np.nextafter(np.float32(1), np.float32(0)))


def build_oracle_trainers(noise=None, act_space_n=None):
    args = make_args()
    obs_shape_n = [(d,) for d in OBS_DIMS]
    act_space_n = act_space_n or [_Discrete(5)] * N
    return [OracleAgentTrainer("agent_%d" % i, None, obs_shape_n, act_space_n, i, args, local_q_func=(i == 2),
                               rng=np.random.RandomState(100 + i), noise=noise) for i in range(N)]


def transition(k):
    """Deterministic transition number k for all agents (float64 python-side values, like env outputs)."""
    rng = np.random.RandomState(5000 + k)
    act = []
    for i in range(N):
        a = rng.rand(5)
        act.append(a / a.sum())
    return {"obs": [rng.randn(d) for d in OBS_DIMS], "act": act, "rew": [float(rng.randn()) for _ in range(N)],
            "obs2": [rng.randn(d) for d in OBS_DIMS], "done": [bool(k % 9 == 8)] * N}
