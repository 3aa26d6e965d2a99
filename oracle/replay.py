"""CPU oracle: restatement of the reference's per-agent replay ring.

TEST INFRASTRUCTURE ONLY (see oracle/mpe.py header for who may import ``oracle/``).

PARITY PINNED: ``tests/golden/make_replay_golden.py`` imports the *real*
``maddpg.trainer.replay_buffer.ReplayBuffer`` from /root/reference in the build container,
drives it and this restatement with the same insert/index streams and commits the outputs as
``tests/golden/replay_*.npz``; ``tests/test_oracle_replay.py`` replays them on any box.

Follows maddpg/trainer/replay_buffer.py (reference):
  __init__ :5-16   add :25-32   _encode_sample :34-44   make_index :46-47
  make_latest_index :49-53   sample_index :55-56   sample :58-82   collect :84-85
"""
import random

import numpy as np


class ReplayBuffer(object):
    def __init__(self, size):
        # replay_buffer.py:14-16
        self._storage = []
        self._maxsize = int(size)
        self._next_idx = 0

    def __len__(self):
        return len(self._storage)

    def clear(self):
        self._storage = []
        self._next_idx = 0

    def add(self, obs_t, action, reward, obs_tp1, done):
        # replay_buffer.py:25-32 -- append until full, then overwrite the oldest slot
        data = (obs_t, action, reward, obs_tp1, done)
        if self._next_idx >= len(self._storage):
            self._storage.append(data)
        else:
            self._storage[self._next_idx] = data
        self._next_idx = (self._next_idx + 1) % self._maxsize

    def _encode_sample(self, idxes):
        # replay_buffer.py:34-44 -- python gather, then np.array() stacking (float64 for obs/rew/done
        # when the env hands over float64, float32 for actions)
        obses_t, actions, rewards, obses_tp1, dones = [], [], [], [], []
        for i in idxes:
            obs_t, action, reward, obs_tp1, done = self._storage[i]
            obses_t.append(np.asarray(obs_t))
            actions.append(np.asarray(action))
            rewards.append(reward)
            obses_tp1.append(np.asarray(obs_tp1))
            dones.append(done)
        return np.array(obses_t), np.array(actions), np.array(rewards), np.array(obses_tp1), np.array(dones)

    def make_index(self, batch_size):
        # replay_buffer.py:46-47 -- inclusive randint on the python global MT19937 stream
        return [random.randint(0, len(self._storage) - 1) for _ in range(batch_size)]

    def make_latest_index(self, batch_size):
        # replay_buffer.py:49-53
        idx = [(self._next_idx - 1 - i) % self._maxsize for i in range(batch_size)]
        np.random.shuffle(idx)
        return idx

    def sample_index(self, idxes):
        return self._encode_sample(idxes)

    def sample(self, batch_size):
        # replay_buffer.py:58-82
        if batch_size > 0:
            idxes = self.make_index(batch_size)
        else:
            idxes = range(0, len(self._storage))
        return self._encode_sample(idxes)

    def collect(self):
        return self.sample(-1)
