"""CPU oracle: numpy restatement of the reference's MADDPG trainer math (TensorFlow-free).

TEST INFRASTRUCTURE ONLY (see oracle/mpe.py header for who may import ``oracle/``).

PARITY PARTLY PINNED.  The TensorFlow GRAPH pieces (mlp_model, the Gumbel-softmax sample, the two
losses and their gradients, clip_by_norm, Adam, polyak) are unpinned against the reference's own
execution: maddpg/trainer/maddpg.py builds them with TensorFlow 1.8.0 (reference README.md:16),
which is not installed and cannot be fetched, and the reference has no test or golden vector for
trainer/, distributions.py or train.py.  They are cross-checked against torch autograd
in float64 (tests/test_oracle_maddpg.py) and the polyak step against the one invariant the
reference does test (tests/test_policy.py:71-86: polyak with tau such that target==running).
The graph WIRING is pinned one level down: the reference's own graph-building code
(``MADDPGAgentTrainer.__init__``, ``q_train``, ``p_train``, ``make_update_exp``, ``SoftCategoricalPd``,
``U.function`` / ``scope_vars`` / ``minimize_and_clip``, ``mlp_model``) was executed unmodified on a
torch-backed stand-in for the TensorFlow calls it makes (tests/tf_shim.py: dense layer, softmax,
clip_by_norm, Adam and autograd are the stand-in's, restated from TensorFlow's documentation) and
driven through the real ``update``; this module reproduces its debug surfaces and statistics to
8e-7 relative and its variables after two update rounds to 6e-8
(tests/golden/make_graph_golden.py -> graph_ref.npz,
test_oracle_matches_the_reference_graph_code).
Everything AROUND the graph is pinned: the REAL ``MADDPGAgentTrainer.update`` / ``experience`` /
``preupdate`` / ``action`` methods (maddpg.py:151-196) were executed unmodified in the build
container on top of these restated graph callables and the REAL ReplayBuffer
(tests/golden/make_update_golden.py -> update_orchestration_ref.npz), and
``OracleAgentTrainer``'s own restatement of them reproduces that run bit for bit
(test_update_orchestration_matches_the_reference_method): the warm-up and every-100-steps gates,
the python-``random`` index draws, the per-agent gathers, the float64 numpy TD combine, the call
order q_train, p_train, p_update, q_update and the six returned statistics.

What is restated (reference file:line):
  mlp_model                      experiments/train.py:39-46      (3 x fully_connected, ReLU, ReLU, none)
  SoftCategoricalPd.sample       maddpg/common/distributions.py:264-266
  SoftMultiCategoricalPd.sample  maddpg/common/distributions.py:332-336 (per-head softmax, low = 0)
  p_train                        maddpg/trainer/maddpg.py:28-73
  q_train                        maddpg/trainer/maddpg.py:75-110
  make_update_exp (polyak)       maddpg/trainer/maddpg.py:20-26
  minimize_and_clip              maddpg/common/tf_util.py:166-182 (per-variable clip_by_norm)
  tf.train.AdamOptimizer         TF 1.8 semantics, SURVEY Appendix B.4
  MADDPGAgentTrainer             maddpg/trainer/maddpg.py:112-196
All network arithmetic is float32 (TF placeholders are float32, tf_util.py:98-112); the TD
combine is float64 numpy then cast to float32 on feed (maddpg.py:186-188).
"""
import numpy as np

from oracle.replay import ReplayBuffer

F32 = np.float32


# ------------------------------------------------------------------------------------------
# action-space helpers (make_pdtype, distributions.py:408-422; the MultiDiscrete branch is the
# SoftMultiCategorical one this fork commented out -- needed for simple_world_comm's leader)
# ------------------------------------------------------------------------------------------
def act_heads(space):
    """Sizes of the soft one-hot heads of an action space: Discrete(n) -> [n];
    MultiDiscrete(low, high) -> high - low + 1."""
    if hasattr(space, "n"):
        return [int(space.n)]
    if hasattr(space, "high") and hasattr(space, "low") and np.ndim(space.high) == 1:
        return [int(h - l + 1) for l, h in zip(space.low, space.high)]
    raise NotImplementedError(space)  # distributions.py:422


def gumbel_softmax(logits, u, heads):
    """softmax(logits - log(-log(u))) per head, float32 (distributions.py:264-266, 332-336)."""
    logits = np.asarray(logits, F32)
    u = np.asarray(u, F32)
    z = logits - np.log(-np.log(u)).astype(F32)
    out = np.empty_like(z)
    o = 0
    for h in heads:
        zz = z[:, o:o + h]
        zz = zz - zz.max(axis=1, keepdims=True)
        e = np.exp(zz).astype(F32)
        out[:, o:o + h] = e / e.sum(axis=1, keepdims=True, dtype=F32)
        o += h
    return out


def gumbel_softmax_backward(a, da, heads):
    """d/dlogits of the per-head softmax given its output ``a`` and upstream ``da``."""
    dl = np.empty_like(a)
    o = 0
    for h in heads:
        aa, dd = a[:, o:o + h], da[:, o:o + h]
        dl[:, o:o + h] = aa * (dd - (aa * dd).sum(axis=1, keepdims=True, dtype=F32))
        o += h
    return dl.astype(F32)


# ------------------------------------------------------------------------------------------
# MLP (train.py:39-46; xavier_initializer == uniform(+-sqrt(6/(fan_in+fan_out))), zero biases)
# ------------------------------------------------------------------------------------------
class MLP:
    NAMES = ("W1", "b1", "W2", "b2", "W3", "b3")

    def __init__(self, in_dim, units, out_dim, rng):
        dims = [(in_dim, units), (units, units), (units, out_dim)]
        self.p = []
        for fi, fo in dims:
            lim = np.sqrt(6.0 / (fi + fo))
            self.p.append(rng.uniform(-lim, lim, size=(fi, fo)).astype(F32))
            self.p.append(np.zeros(fo, F32))

    def copy_from(self, other):
        self.p = [x.copy() for x in other.p]

    def forward(self, x):
        W1, b1, W2, b2, W3, b3 = self.p
        x = np.asarray(x, F32)
        z1 = x @ W1 + b1
        h1 = np.maximum(z1, 0)
        z2 = h1 @ W2 + b2
        h2 = np.maximum(z2, 0)
        out = h2 @ W3 + b3
        return out.astype(F32), (x, h1, h2)

    def backward(self, cache, dout, need_dx=False):
        W1, b1, W2, b2, W3, b3 = self.p
        x, h1, h2 = cache
        dout = np.asarray(dout, F32)
        gW3 = h2.T @ dout
        gb3 = dout.sum(axis=0, dtype=F32)
        dh2 = (dout @ W3.T) * (h2 > 0)
        gW2 = h1.T @ dh2
        gb2 = dh2.sum(axis=0, dtype=F32)
        dh1 = (dh2 @ W2.T) * (h1 > 0)
        gW1 = x.T @ dh1
        gb1 = dh1.sum(axis=0, dtype=F32)
        dx = dh1 @ W1.T if need_dx else None
        return [g.astype(F32) for g in (gW1, gb1, gW2, gb2, gW3, gb3)], dx


class Adam:
    """TF-1.8 AdamOptimizer: eps is added to the *uncorrected* sqrt(v)."""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]
        self.t = 0

    def step(self, params, grads):
        self.t += 1
        lr_t = F32(self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m[...] = F32(self.b1) * m + F32(1.0 - self.b1) * g
            v[...] = F32(self.b2) * v + F32(1.0 - self.b2) * g * g
            p -= lr_t * m / (np.sqrt(v) + F32(self.eps))


def clip_by_norm(g, clip):
    """tf.clip_by_norm per variable (tf_util.py:176-180): g * clip / max(||g||, clip)."""
    n = np.sqrt(np.sum(g.astype(F32) * g, dtype=F32))
    return (g * F32(clip) / np.maximum(n, F32(clip))).astype(F32)


def polyak_update(target, running, polyak=1.0 - 1e-2):
    """maddpg.py:20-26 -- target <- polyak*target + (1-polyak)*running, variable by variable."""
    for t, r in zip(target.p, running.p):
        t[...] = F32(polyak) * t + F32(1.0 - polyak) * r


# ------------------------------------------------------------------------------------------
# MADDPGAgentTrainer (maddpg.py:112-196)
# ------------------------------------------------------------------------------------------
class OracleAgentTrainer:
    """Same surface as the reference trainer.  ``model`` is accepted and ignored (the oracle owns
    the MLP).  Randomness is injected: ``noise(shape)`` must return U[0,1) float32 draws (the
    reference uses tf.random_uniform, which cannot be reproduced -- SURVEY H6)."""

    def __init__(self, name, model, obs_shape_n, act_space_n, agent_index, args, local_q_func=False,
                 rng=None, noise=None, replay_size=1e6):
        self.name = name
        self.n = len(obs_shape_n)
        self.agent_index = agent_index
        self.args = args
        self.local_q_func = local_q_func
        rng = rng if rng is not None else np.random.RandomState(agent_index)
        # tf.random_uniform(float32) is in [0, 1): a float64 draw cast to float32 may round up to 1.0 -> clamp below 1
        self.noise = noise if noise is not None else (
            lambda shape: np.minimum(rng.uniform(size=shape).astype(F32), np.nextafter(F32(1), F32(0))))
        self.obs_dims = [int(s[0]) for s in obs_shape_n]
        self.heads_n = [act_heads(s) for s in act_space_n]
        self.act_dims = [sum(h) for h in self.heads_n]
        U = args.num_units
        j = agent_index
        q_in = (self.obs_dims[j] + self.act_dims[j]) if local_q_func else (sum(self.obs_dims) + sum(self.act_dims))
        # variable creation order in the reference graph: q_func, target_q_func (q_train, :123),
        # then p_func, target_p_func (p_train, :134); all four independently initialised (train.py:89)
        self.q = MLP(q_in, U, 1, rng)
        self.target_q = MLP(q_in, U, 1, rng)
        self.p = MLP(self.obs_dims[j], U, self.act_dims[j], rng)
        self.target_p = MLP(self.obs_dims[j], U, self.act_dims[j], rng)
        self.q_opt = Adam(self.q.p, args.lr)
        self.p_opt = Adam(self.p.p, args.lr)
        self.grad_norm_clipping = 0.5
        self.replay_buffer = ReplayBuffer(replay_size)
        self.max_replay_buffer_len = args.batch_size * args.max_episode_len
        self.replay_sample_index = None
        self.p_debug = {"p_values": self.p_values, "target_act": self.target_act}
        self.q_debug = {"q_values": self.q_values, "target_q_values": self.target_q_values}
        self.last_grads = {}

    # -- callable graph pieces ------------------------------------------------------------
    def _q_input(self, obs_n, act_n):
        j = self.agent_index
        if self.local_q_func:
            return np.concatenate([np.asarray(obs_n[j], F32), np.asarray(act_n[j], F32)], axis=1)
        return np.concatenate([np.asarray(o, F32) for o in obs_n] + [np.asarray(a, F32) for a in act_n], axis=1)

    def p_values(self, obs):
        return self.p.forward(obs)[0]

    def act(self, obs):
        logits = self.p.forward(obs)[0]
        return gumbel_softmax(logits, self.noise(logits.shape), self.heads_n[self.agent_index])

    def target_act(self, obs):
        logits = self.target_p.forward(obs)[0]
        return gumbel_softmax(logits, self.noise(logits.shape), self.heads_n[self.agent_index])

    def q_values(self, *args):
        return self.q.forward(self._q_input(args[:self.n], args[self.n:]))[0][:, 0]

    def target_q_values(self, *args):
        return self.target_q.forward(self._q_input(args[:self.n], args[self.n:]))[0][:, 0]

    # -- reference surface ----------------------------------------------------------------
    def action(self, obs):
        return self.act(np.asarray(obs)[None])[0]

    def experience(self, obs, act, rew, new_obs, done, terminal):
        self.replay_buffer.add(obs, act, rew, new_obs, float(done))

    def preupdate(self):
        self.replay_sample_index = None

    def q_train(self, obs_n, act_n, target_q):
        """maddpg.py:75-100: critic MSE step.  Returns the pre-step loss (float32)."""
        x = self._q_input(obs_n, act_n)
        y = np.asarray(target_q, F32)
        out, cache = self.q.forward(x)
        q = out[:, 0]
        diff = q - y
        loss = np.mean(diff * diff, dtype=F32)
        dq = (F32(2.0) * diff / F32(len(y))).astype(F32)
        grads, _ = self.q.backward(cache, dq[:, None])
        self.last_grads["q"] = [g.copy() for g in grads]
        grads = [clip_by_norm(g, self.grad_norm_clipping) for g in grads]
        self.q_opt.step(self.q.p, grads)
        return loss

    def p_train(self, obs_n, act_n):
        """maddpg.py:28-61: actor step through the (already updated) running critic."""
        j = self.agent_index
        heads = self.heads_n[j]
        logits, pcache = self.p.forward(obs_n[j])
        a = gumbel_softmax(logits, self.noise(logits.shape), heads)
        act_in = [np.asarray(x, F32) for x in act_n]
        act_in[j] = a
        x = self._q_input(obs_n, act_in)
        out, qcache = self.q.forward(x)
        q = out[:, 0]
        B = F32(len(q))
        pg_loss = -np.mean(q, dtype=F32)
        p_reg = np.mean(logits * logits, dtype=F32)
        loss = F32(pg_loss + p_reg * F32(1e-3))
        dq = np.full((len(q), 1), -1.0 / B, F32)
        _, dx = self.q.backward(qcache, dq, need_dx=True)
        if self.local_q_func:
            off = self.obs_dims[j]
        else:
            off = sum(self.obs_dims) + sum(self.act_dims[:j])
        da = dx[:, off:off + self.act_dims[j]]
        dlogits = gumbel_softmax_backward(a, da, heads)
        dlogits = dlogits + F32(1e-3) * F32(2.0) * logits / F32(logits.size)
        grads, _ = self.p.backward(pcache, dlogits)
        self.last_grads["p"] = [g.copy() for g in grads]
        grads = [clip_by_norm(g, self.grad_norm_clipping) for g in grads]
        self.p_opt.step(self.p.p, grads)
        return loss

    def p_update(self):
        polyak_update(self.target_p, self.p)

    def q_update(self):
        polyak_update(self.target_q, self.q)

    def update(self, agents, t, index=None):
        """maddpg.py:161-196.  ``index`` overrides make_index for injected index streams."""
        if len(self.replay_buffer) < self.max_replay_buffer_len:
            return
        if not t % 100 == 0:
            return
        self.replay_sample_index = self.replay_buffer.make_index(self.args.batch_size) if index is None else index
        obs_n, obs_next_n, act_n = [], [], []
        index = self.replay_sample_index
        for i in range(self.n):
            obs, act, rew, obs_next, done = agents[i].replay_buffer.sample_index(index)
            obs_n.append(obs)
            obs_next_n.append(obs_next)
            act_n.append(act)
        obs, act, rew, obs_next, done = self.replay_buffer.sample_index(index)
        return self.update_from_batch(agents, obs_n, act_n, obs_next_n, rew, done)

    def update_from_batch(self, agents, obs_n, act_n, obs_next_n, rew, done):
        """maddpg.py:181-196 on an already gathered batch."""
        target_act_next_n = [agents[i].p_debug["target_act"](obs_next_n[i]) for i in range(self.n)]
        target_q_next = self.q_debug["target_q_values"](*(obs_next_n + target_act_next_n))
        target_q = np.asarray(rew, np.float64) + self.args.gamma * (1.0 - np.asarray(done, np.float64)) * target_q_next
        q_loss = self.q_train(obs_n, act_n, target_q)
        p_loss = self.p_train(obs_n, act_n)
        self.p_update()
        self.q_update()
        self.last_target_q = target_q
        return [q_loss, p_loss, np.mean(target_q), np.mean(rew), np.mean(target_q_next), np.std(target_q)]
