"""Generates tests/golden/learn_loop_ref.npz by running the REAL ``MultiAgentAlgBase.learn_generator`` / ``learn``
(/root/reference/maddpg/algorithms/multiagentalgbase.py:106-165) -- build container only:

    python tests/golden/make_learn_loop_golden.py

The loop is plain Python around four methods of the subclass (``predict``, ``train_step``, ``run_updates``) and the fork's dict
replay; the TensorFlow graph only lives in ``__init__``, which is bypassed (``object.__new__``), and stand-in modules satisfy the
module-level imports (tensorflow, tqdm, gym.spaces).  A deterministic toy env and recording stand-ins for the three methods give
the call log the loop produces: which steps predict, which train (and on how many sampled rows), when the targets update, what
every TrainInfo carries, and the running reward ``learn`` prints.
"""
import os
import random
import sys
import types

import numpy as np

for name in ("tensorflow", "tqdm", "gym", "gym.spaces"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["tqdm"].tqdm = lambda it, **kw: it
PRINTED = []
sys.modules["tqdm"].tqdm.write = lambda text, *a, **k: PRINTED.append(text)
sys.modules["tqdm"].trange = range
for cls in ("Box", "Discrete", "Dict"):
    setattr(sys.modules["gym.spaces"], cls, type(cls, (), {}))
sys.modules["gym"].spaces = sys.modules["gym.spaces"]
import importlib.util  # noqa: E402

for pkg_name, path in (("maddpg", "/root/reference/maddpg"), ("maddpg.common", "/root/reference/maddpg/common"),
                       ("maddpg.algorithms", "/root/reference/maddpg/algorithms")):
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [path]
    sys.modules[pkg_name] = pkg
for mod, path in (("maddpg.common.utils_common", "common/utils_common.py"), ("maddpg.common.replaybuffer", "common/replaybuffer.py")):
    spec = importlib.util.spec_from_file_location(mod, "/root/reference/maddpg/" + path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[mod] = m
    spec.loader.exec_module(m)
sys.modules["maddpg.common"].ReplayBuffer = sys.modules["maddpg.common.replaybuffer"].ReplayBuffer
spec = importlib.util.spec_from_file_location("maddpg.algorithms.multiagentalgbase",
                                              "/root/reference/maddpg/algorithms/multiagentalgbase.py")
base_mod = importlib.util.module_from_spec(spec)
sys.modules["maddpg.algorithms.multiagentalgbase"] = base_mod
spec.loader.exec_module(base_mod)

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["scout", "anchor"]
TIMESTEPS = 10003


class ToyEnv(object):
    """Deterministic dict env: episodes of 7 steps, reward = -(step in episode) - 0.25 * sum(actions)."""

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
    """predict / train_step / run_updates stand-ins that record how the loop calls them."""
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


def run(make):
    random.seed(7)
    log, infos = [], []
    alg = make()
    instrument(alg, log)
    for info in alg.learn_generator(ToyEnv(), timesteps=TIMESTEPS):
        infos.append((info.step, float(info.rewards[NAMES[0]]), float(info.dones[NAMES[1]]),
                      -1.0 if info.actor_loss is None else float(info.actor_loss[NAMES[0]]),
                      -1.0 if info.critic_loss is None else float(info.critic_loss[NAMES[1]]), float(info.infos["t"]),
                      float(info.observations[NAMES[0]][0])))
    random.seed(8)
    alg2 = make()
    log2 = []
    instrument(alg2, log2)
    del PRINTED[:]
    ret = alg2.learn(ToyEnv(), timesteps=TIMESTEPS, verbose=True)
    running = [float(line.split(":")[1]) for line in PRINTED if line.startswith("Running Reward")]
    trains = np.asarray([e[1:] for e in log if e[0] == "train"], np.float64)
    order = np.asarray([{"predict": 0, "train": 1, "update": 2}[e[0]] for e in log], np.int8)
    return dict(infos=np.asarray(infos, np.float64), trains=trains, order=order,
                predict_arg=np.asarray([e[1] for e in log if e[0] == "predict"], np.float64), learn_return=np.asarray(
                    [np.nan if ret is None else ret], np.float64), running_reward=np.asarray(running, np.float64))


def main():
    Base = base_mod.MultiAgentAlgBase
    Sub = type("Sub", (Base,), {k: (lambda self, *a, **kw: None) for k in Base.__abstractmethods__})
    out = run(lambda: object.__new__(Sub))
    np.savez_compressed(os.path.join(HERE, "learn_loop_ref.npz"), **out)
    print("wrote learn_loop_ref.npz: %d yields, train steps at %s, learn() returned %s, printed running rewards %s" % (
        len(out["infos"]), out["trains"][:, 0].tolist(), out["learn_return"], out["running_reward"]))


if __name__ == "__main__":
    main()
