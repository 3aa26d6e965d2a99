"""TEST INFRASTRUCTURE -- CPU restatement of the reference's prioritized replay memory.

Follows /root/reference/maddpg/trainer/prioritized_replay_buffer.py:
  SumTree.__init__ :26-43, add :45-56, update_all :58-100, update :102-109, get_leaf :111-142, total_p :144-146,
  PrioritizedReplayMemory.add :163-169, sample :171-194, batch_update :196-201.

PINNED: tests/golden/prioritized_ref.npz holds the outputs of the REAL reference class (executed in the build container
with a stand-in ``tensorflow`` module: the file's only use of it is ``tf.set_random_seed(1)`` at import), produced by
tests/golden/make_prioritized_golden.py; tests/test_oracle_prioritized.py holds this restatement to them bit for bit.

The restatement is array-shaped on purpose (a float64 numpy tree, a level-synchronous ``update_all``): it is the second
formulation the CUDA kernels (maddpg_b200/csrc/mdp_prio.cu) were designed from, and the goldens prove it equals the
reference's list-popping code exactly -- including the reference's quirks, which are part of "results identical":

* leaf of data slot d is tree index ``d + parent_nodes - 1`` with ``parent_nodes = 2^k - 1`` (:46): slot 0 lives in the LAST
  INTERNAL node ``2^k - 2``; a descent that reaches it continues into its (empty) children and returns data index ``2^k - 1`` or
  ``2^k`` -> ``IndexError`` on ``self.data[...]`` unless capacity == 2^k and it went left (:138-142);
* ``add`` does not touch the tree: (leaf, 1e6) pairs wait in ``dirty`` until the next ``get_leaf`` (:50-51, :124);
* ``update_all`` adds the SUM of the two children's deltas to a parent (one rounding), not one delta after the other;
* ``min_prob`` scans the LAST ``capacity`` entries of the tree array (:181-182), which for capacity < 2^k includes never-used
  leaves: min = 0, ``prob / min_prob`` = inf and every IS weight is 0.0;
* ``sample`` reads ``total_p`` BEFORE the first ``get_leaf`` flushes the pending adds (:175 vs :124);
* ``batch_update`` propagates one leaf after the other (float64 ``+=`` per ancestor, batch order);
* for odd k, ``update_all`` with slot 0 among the pending adds hands the root the wrong delta (``SumTreeOracle.update_all``).

Nothing under maddpg_b200/ imports this module.
"""
from math import ceil, log2

import numpy as np


class SumTreeOracle(object):
    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.k = int(ceil(log2(capacity)))
        self.size = 2 ** (self.k + 1) - 1
        self.parent_nodes = 2 ** self.k - 1
        self.tree = np.zeros(self.size, dtype=np.float64)
        self.data_pointer = 0
        # pending adds: one circular range of data slots [dirty_start, dirty_start + dirty_count); the value is always the same
        self.dirty_start = 0
        self.dirty_count = 0
        self.dirty_value = 0.0

    def leaf_of(self, data_idx):
        return data_idx + self.parent_nodes - 1

    def add(self, p, n=1):
        """n consecutive SumTree.add(p, data) calls (update=False)."""
        if self.dirty_count == 0:
            self.dirty_start = self.data_pointer
        self.dirty_value = float(p)
        self.dirty_count = min(self.capacity, self.dirty_count + n)
        if self.dirty_count == self.capacity:
            self.dirty_start = 0
        self.data_pointer = (self.data_pointer + n) % self.capacity

    def update_all(self):
        """Level-synchronous form of :58-100.  Iteration t touches the ancestors at depth k - t of the dirty true leaves and
        the ancestor at depth k - 1 - t of slot 0's node (one level ahead, always the last node of its level).

        One more quirk of the reference lives here.  Its loop keeps two parallel lists (nodes, deltas) and appends a parent
        for every node except the root (:96-99), but hands the WHOLE delta list on (:95).  Slot 0's chain reaches the root one
        iteration before the true leaves' chain; if the lists are in descending node order at that moment -- they alternate
        direction every iteration, descending at even ones, so this is the case when k is odd -- the root's delta is the last
        list element and the next iteration pops it for the wrong node: the root then receives ``c0 + delta(node 1)`` (or
        ``c0`` alone when only one of its children changed) instead of ``delta(node 1) + delta(node 2)``: slot 0's delta is
        counted twice and the right subtree's is dropped.  With k even (capacity 1e6: k = 20) the lists are ascending there
        and the sum is right."""
        if self.dirty_count == 0:
            return
        slots = (self.dirty_start + np.arange(self.dirty_count)) % self.capacity
        self.dirty_count = 0
        value = self.dirty_value
        slot0 = bool((slots == 0).any())
        c0 = 0.0
        if slot0:  # slot 0's node q and its ancestors: every node gets c0 before the true leaves' sums reach it
            q = self.leaf_of(0)
            c0 = value - self.tree[q]
            self.tree[q] = value
            m = q
            while m != 0:
                m = (m - 1) // 2
                self.tree[m] += c0
        leaves = np.unique(self.leaf_of(slots[slots != 0]))
        delta = value - self.tree[leaves]
        self.tree[leaves] = value
        nodes = leaves
        while nodes.size and not (nodes.size == 1 and nodes[0] == 0):
            parents = (nodes - 1) // 2
            up, first = np.unique(parents, return_index=True)
            cnt = np.diff(np.append(first, parents.size))
            assert cnt.max() <= 2
            summed = delta[first].copy()
            two = cnt == 2
            summed[two] = delta[first[two]] + delta[first[two] + 1]  # left + right in one rounding (float add commutes)
            if up.size == 1 and up[0] == 0 and slot0 and self.k % 2 == 1:
                has_left = bool((nodes == 1).any())
                summed[0] = c0 + delta[np.nonzero(nodes == 1)[0][0]] if (has_left and nodes.size == 2) else c0
            self.tree[up] += summed
            nodes, delta = up, summed

    def update(self, tree_idx, p):
        change = p - self.tree[tree_idx]
        self.tree[tree_idx] = p
        while tree_idx != 0:
            tree_idx = (tree_idx - 1) // 2
            self.tree[tree_idx] += change

    def get_leaf(self, v):
        """-> (leaf_idx, priority, data_idx); the caller decides what an out-of-range data_idx means."""
        self.update_all()
        parent = 0
        while True:
            cl = 2 * parent + 1
            if cl >= self.size:
                leaf = parent
                break
            if v <= self.tree[cl]:
                parent = cl
            else:
                v -= self.tree[cl]
                parent = cl + 1
        return leaf, self.tree[leaf], leaf - self.parent_nodes + 1

    @property
    def total_p(self):
        return self.tree[0]


class PrioritizedReplayOracle(object):
    epsilon = 0.01
    alpha = 0.6
    beta_increment_per_sampling = 0.001
    abs_err_upper = 1.0

    def __init__(self, capacity):
        self.tree = SumTreeOracle(capacity)
        self.beta = 0.4
        self.data = [None] * int(capacity)

    def add(self, *row):
        self.data[self.tree.data_pointer] = row
        self.tree.add(1e6)

    def sample(self, n, uniforms):
        """uniforms: the n draws of numpy's random_sample() that np.random.uniform(a, b) consumes (a + (b - a) * u).
        Raises IndexError exactly where the reference does (a descent through slot 0's node)."""
        t = self.tree
        total0 = t.total_p                      # read before the flush (:175)
        pri_seg = total0 / n
        self.beta = np.min([1.0, self.beta + self.beta_increment_per_sampling])
        with np.errstate(divide="ignore", invalid="ignore"):
            min_prob = np.min(t.tree[-t.capacity:]) / t.total_p    # also before the flush (:181-182)
        b_idx, b_data, isw = [], [], []
        for i in range(n):
            a, b = pri_seg * i, pri_seg * (i + 1)
            v = a + (b - a) * uniforms[i]
            leaf, p, data_idx = t.get_leaf(v)
            if data_idx >= t.capacity:
                raise IndexError("list index out of range")
            prob = p / t.total_p
            with np.errstate(divide="ignore", invalid="ignore"):
                isw.append(np.power(prob / min_prob, -self.beta))
            b_idx.append(leaf)
            b_data.append(data_idx)
        return b_idx, b_data, isw

    def priorities(self, abs_errors):
        e = np.asarray(abs_errors, dtype=np.float64) + self.epsilon
        return np.power(np.minimum(e, self.abs_err_upper), self.alpha)

    def batch_update(self, tree_idx, abs_errors):
        for ti, p in zip(tree_idx, self.priorities(abs_errors)):
            self.tree.update(int(ti), p)
