"""maddpg_b200.algorithms.MultiAgentAlgBase.learn_generator / learn against the REAL reference loop
(maddpg/algorithms/multiagentalgbase.py:106-165): tests/golden/learn_loop_ref.npz is the call log of the reference's own code
(tests/golden/make_learn_loop_golden.py: TF graph construction bypassed, recording stand-ins for predict / train_step /
run_updates, a deterministic toy env).  The same stand-ins on this package's base class must produce the same log: which steps
predict on what, when the batch of 1024 is drawn (python ``random`` stream of the dict replay included), what every TrainInfo
carries, when the targets update, and the running reward ``learn`` prints.  No GPU: the loop is host logic."""
import os
import random
import re

import numpy as np

NAMES = ["scout", "anchor"]
TIMESTEPS = 10003
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "learn_loop_ref.npz")


class ToyEnv(object):
    def __init__(self):
        self.t, self.k = 0, 0

    def _obs(self):
        return {n: np.array([self.t, i], np.float32) for i, n in enumerate(NAMES)}

    def reset(self):
        self.k = 0
        return self._obs()

    def step(self, actions):
        self.t += 1
        self.k += 1
        r = -float(self.k) - 0.25 * float(sum(np.sum(a) for a in actions.values()))
        return self._obs(), r, self.k == 7, {"t": self.t}


def instrument(obj, log):
    def predict(observations, noisy=True):
        log.append(("predict", float(observations[NAMES[0]][0])))
        return {n: np.array([0.5 * (i + 1)], np.float32) for i, n in enumerate(NAMES)}

    def train_step(observations, actions, rewards, observations_n, dones, step=None):
        rows = len(observations[NAMES[0]])
        log.append(("train", float(step), float(rows), float(np.sum(rewards[NAMES[1]])), float(np.sum(dones[NAMES[0]]))))
        return {"actor": {n: 0.125 * step for n in NAMES}, "critic": {n: 2.0 * step for n in NAMES}}

    def run_updates():
        log.append(("update",))
    obj.predict, obj.train_step, obj.run_updates = predict, train_step, run_updates


def test_learn_loop_matches_the_reference_loop(capsys):
    from maddpg_b200.algorithms import MultiAgentAlgBase
    gold = np.load(GOLD)
    random.seed(7)
    log, infos = [], []
    alg = object.__new__(MultiAgentAlgBase)
    instrument(alg, log)
    for info in alg.learn_generator(ToyEnv(), timesteps=TIMESTEPS):
        infos.append((info.step, float(info.rewards[NAMES[0]]), float(info.dones[NAMES[1]]),
                      -1.0 if info.actor_loss is None else float(info.actor_loss[NAMES[0]]),
                      -1.0 if info.critic_loss is None else float(info.critic_loss[NAMES[1]]), float(info.infos["t"]),
                      float(info.observations[NAMES[0]][0])))
    assert np.array_equal(np.asarray(infos, np.float64), gold["infos"])
    trains = np.asarray([e[1:] for e in log if e[0] == "train"], np.float64)
    assert np.array_equal(trains, gold["trains"])          # steps 5000 and 10000, 1024 rows, the same sampled rows
    order = np.asarray([{"predict": 0, "train": 1, "update": 2}[e[0]] for e in log], np.int8)
    assert np.array_equal(order, gold["order"])
    assert np.array_equal(np.asarray([e[1] for e in log if e[0] == "predict"], np.float64), gold["predict_arg"])
    # learn(): returns None, prints the same running reward at the same steps
    random.seed(8)
    alg2 = object.__new__(MultiAgentAlgBase)
    instrument(alg2, [])
    capsys.readouterr()
    assert alg2.learn(ToyEnv(), timesteps=TIMESTEPS, verbose=True) is None and np.isnan(gold["learn_return"][0])
    printed = [float(x) for x in re.findall(r"Running Reward: ([-+0-9.]+)", capsys.readouterr().out)]
    assert len(printed) == len(gold["running_reward"]) == 2
    np.testing.assert_allclose(printed, gold["running_reward"], atol=1e-6)
