"""Learning check of the fork's algorithms (maddpg_b200/algorithms.py) on a toy cooperative task with a known optimum: every agent
sees a 2-vector g in [-0.8, 0.8]^2 and should output it; the shared reward is -sum_i |a_i - g_i|^2, episodes last one step.
The loop is the reference's learn_generator (multiagentalgbase.py:106-132) with a train step every env step.
usage: python tools/td3_learning_curve.py [MaTd3|Coma|Maddpg] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maddpg_b200 import algorithms  # noqa: E402
from maddpg_b200.spaces import Box, Dict  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "MaTd3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
names = ["agent_0", "agent_1"]
obs_sp = Dict({k: Box(-1.0, 1.0, (2,)) for k in names})
act_sp = Dict({k: Box(-np.ones(2, np.float32), np.ones(2, np.float32), (2,)) for k in names})
rng = np.random.RandomState(0)
np.random.seed(0)


def draw():
    return {k: rng.uniform(-0.8, 0.8, 2).astype(np.float32) for k in names}


def reward(obs, act):
    return -float(sum(np.sum(np.square(np.asarray(act[k]) - obs[k])) for k in names))


def evaluate(alg, n=256):
    obs = {k: rng.uniform(-0.8, 0.8, (n, 2)).astype(np.float32) for k in names}
    act = alg.predict(obs, noisy=False)
    return -float(np.mean(sum(np.sum(np.square(act[k].reshape(n, 2) - obs[k]), axis=1) for k in names)))


alg = getattr(algorithms, name)(obs_sp, act_sp, seed=0)
replay = algorithms.DictReplayBuffer(20000)
print("%s on the match-the-goal task: greedy reward (optimum 0) every 250 steps" % name)
print("step %5d  reward %+.4f" % (0, evaluate(alg)))
obs = draw()
for step in range(1, steps + 1):
    act = alg.predict(obs)
    act = {k: np.clip(a, -1, 1) for k, a in act.items()}
    r = reward(obs, act)
    nxt = draw()
    replay.add(obs, act, {k: r for k in names}, nxt, {k: True for k in names})
    obs = nxt
    if len(replay) >= 256:
        alg.train_step(*replay.sample(256), step=2 * step)   # an even step: MaTd3's policies step too (matd3.py:69)
        alg.run_updates()
    if step % 250 == 0:
        print("step %5d  reward %+.4f" % (step, evaluate(alg)))
