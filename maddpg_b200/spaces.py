"""Minimal stand-ins for the gym / MPE action and observation spaces the reference reads.

gym 0.10.5 and MPE's ``multiagent.multi_discrete`` are not dependencies of this package; only the
attributes the reference touches are provided: ``Discrete.n`` and ``MultiDiscrete.low/high``
(maddpg/common/distributions.py:408-422, the MultiDiscrete branch being the SoftMultiCategorical
one), ``Box.shape`` (experiments/train.py:83).  Real gym spaces are accepted anywhere these are.
"""
import numpy as np


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return hasattr(other, "n") and int(other.n) == self.n


class MultiDiscrete(object):
    """List of [min, max] pairs, like ``multiagent.multi_discrete.MultiDiscrete``."""

    def __init__(self, array_of_param_array):
        self.low = np.array([x[0] for x in array_of_param_array])
        self.high = np.array([x[1] for x in array_of_param_array])
        self.num_discrete_space = self.low.shape[0]

    def __repr__(self):
        return "MultiDiscrete(%s)" % [[int(l), int(h)] for l, h in zip(self.low, self.high)]


class Box(object):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return "Box%s" % (self.shape,)


class Dict(object):
    """``gym.spaces.Dict``: the fork's algorithms take one for observations and one for actions (multiagentalgbase.py:27-35)."""

    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __repr__(self):
        return "Dict(%s)" % ", ".join("%s:%r" % kv for kv in self.spaces.items())


def act_heads(space):
    """Soft one-hot head sizes of an action space (make_pdtype, distributions.py:408-422)."""
    if hasattr(space, "n"):
        return [int(space.n)]
    if hasattr(space, "high") and hasattr(space, "low") and np.ndim(space.high) == 1:
        return [int(h - l + 1) for l, h in zip(space.low, space.high)]
    raise NotImplementedError("unsupported action space %r" % (space,))
