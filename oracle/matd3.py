"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float32) of the fork's tanh-policy algorithms: MATD3 and the best/worst-policy
"COMA" variant.

PARITY PARTLY PINNED.  The modules need TensorFlow 1.x and DeepMind Sonnet 1.x (``import sonnet as snt``), neither is installed
nor fetchable, and the reference holds no golden vector for them (its tests/test_policy.py and tests/test_maddpg.py build TF
graphs).  What IS pinned is the wiring: the fork's own ``Coma`` and ``Maddpg`` classes, their modules, groups, policies, critics
and ``TfFunction`` plumbing were executed unmodified in the build container on tests/tf_shim.py, a torch-backed stand-in for the
TensorFlow and Sonnet calls those files make (tests/golden/make_fork_graph_golden.py -> fork_graph_ref.npz), and ComaOracle /
MaddpgOracle reproduce the losses ``train_step`` returned to 1.5e-6 relative and every variable after three train steps +
target updates to 1.3e-7 (tests/test_oracle_matd3.py::test_oracle_matches_the_fork_graph_code).  The same run records that the
reference's ``MaTd3`` cannot be constructed as written (below); with the one crashing call made tolerant of its extra argument --
the generator's only modification of reference code -- MaTD3Module / MaTd3 run too (noisy targets fed the recorded N(0,1) draws,
four steps: critic-only, full, critic-only, full) and MaTd3Oracle reproduces them to the same accuracy.  The primitive op semantics (dense layer, tanh, Adam) are the
stand-in's, restated; every restated gradient is also cross-checked against torch autograd in float64.

What is restated (reference file:line):
  LaggingNetwork            maddpg/modules/laggingnetwork.py:15-48   running + target snt.nets.MLP (ReLU, linear last layer);
                                                                      update_target: t <- polyak*t + (1-polyak)*r
  Policy._build             maddpg/modules/policy.py:63-88           tanh, clip(N(0,.2),-.5,.5) target noise, clip to [-1,1], Box rescale
  Policy.create_optimizer   maddpg/modules/policy.py:90-100          loss = -mean(value); Adam(lr, use_locking), no clipping
  Critic                    maddpg/modules/critic.py:60-88           MLP on concat([obs, act]); loss = mse(values - target)
  Policy/CriticGroup        maddpg/modules/policygroup.py, criticgroup.py   per-name members (lr 1e-4), or one shared member
  MaTD3Module._build        maddpg/modules/matd3module.py:46-111     twin critic groups, min of the target critics, policy loss through
                                                                      the PRIMARY critic group's TARGET network, polyak 5e-3
  MaTd3._train_step         maddpg/algorithms/matd3.py:63-72         policies step only when ``step and step % 2 == 0``
  ComaModule._build         maddpg/modules/comamodule.py:58-153      best / worst policy groups, shared global critic, personal critics
  Coma._train_step          maddpg/algorithms/coma.py:57-63

Things the reference does that look like slips but are what its graph computes (kept, they are part of "same results"):
  * ``update_targets(5e-3)`` passes 5e-3 as POLYAK, so a target keeps 0.5 % of itself and takes 99.5 % of the running net;
  * the MATD3 / COMA policy losses run through the critics' TARGET networks (``.target_values``);
  * TD targets use ``gamma`` = 0.9 (MaTD3Module) / 0.95 (ComaModule): the ``_build`` defaults, never overridden;
  * COMA's shared global critic trains on the FIRST name's reward and TD target only (criticgroup.py:94-100), its "personal
    reward" is Q_global(o, a) - Q_global^target(o', worst(o')) for every name, and the best-policy TD actions come from the
    RUNNING best policies (``.actions``), no target noise.
  * ``CriticGroup(shared=True)`` / ``PolicyGroup(shared=True)`` assert that every name has the first name's spaces
    (criticgroup.py:28-30, policygroup.py:32-34) -- ComaModule's global critic group is always shared, so the reference's Coma
    only accepts equal spaces.
One thing is NOT reproducible: ``MaTD3Module._build`` calls ``PolicyGroup.create_optimizers(target_vals, policies.entropy)``
(matd3module.py:98-99) but the method takes one argument (policygroup.py:123) -- building the reference's MATD3 graph raises
TypeError.  The restatement drops the extra argument (the "entropy" is not used by any loss in the fork).  And one thing is not
reproduced on purpose: on critic-only steps ``MaTd3._train_step`` passes the un-prefixed per-name loss dict through
``unflatten_map`` (matd3.py:71-72), which splits every name at its first "_" ({'a': {'agent': loss}, ...} for "a_agent") and raises
ValueError for a name without one; here, and in maddpg_b200.algorithms, those losses come back as {'critic': {name: loss}}.

All optimizer steps of one ``train_step`` belong to one ``session.run``: every gradient is taken at the pre-step variables.
Agents are ordered by sorted name (``U.concat_map``, tf_util.py:53-55).  Nothing under maddpg_b200/ imports this module.
"""
import numpy as np

from oracle.maddpg import MLP, Adam, polyak_update

F32 = np.float32
UNITS = 64            # policy.py:33, critic.py:31
LR = 1e-4             # policygroup.py:127, criticgroup.py:91
POLYAK = 5e-3         # matd3module.py:104-107, comamodule.py:137-149
NOISE_STD, NOISE_CLIP = 0.2, 0.5   # policy.py:72-73


class Lagging(object):
    def __init__(self, in_dim, out_dim, rng):
        self.running = MLP(in_dim, UNITS, out_dim, rng)
        self.target = MLP(in_dim, UNITS, out_dim, rng)
        self.adam = Adam(self.running.p, LR)

    def update_target(self, polyak=POLYAK):
        polyak_update(self.target, self.running, polyak)


def box_rescale(low, high):
    lo, hi = float(np.min(low)), float(np.max(high))
    interval = (hi - lo) / 2
    return F32(interval), F32(interval + lo)


class PolicyOracle(Lagging):
    def __init__(self, obs_dim, act_dim, low, high, rng):
        super().__init__(obs_dim, act_dim, rng)
        self.scale, self.shift = box_rescale(low, high)

    def act(self, obs, target=False, z=None):
        """-> (action, tanh, cache); z: N(0,1) draws for the noisy target (policy.py:72-75)."""
        net = self.target if target else self.running
        out, cache = net.forward(obs)
        t = np.tanh(out).astype(F32)
        tn = t
        if z is not None:
            noise = np.clip(F32(NOISE_STD) * np.asarray(z, F32), -NOISE_CLIP, NOISE_CLIP).astype(F32)
            tn = np.clip(t + noise, -1, 1).astype(F32)
        return (tn * self.scale + self.shift).astype(F32), t, cache

    def grads(self, cache, t, da):
        """Gradients of the running net given dL/d(action)."""
        dout = (np.asarray(da, F32) * self.scale * (F32(1) - t * t)).astype(F32)
        g, _ = self.running.backward(cache, dout)
        return g


class CriticOracle(Lagging):
    def __init__(self, in_dim, rng):
        super().__init__(in_dim, 1, rng)

    def q(self, x, target=False):
        net = self.target if target else self.running
        out, cache = net.forward(x)
        return out[:, 0].astype(F32), cache

    def mse_grads(self, x, y):
        """loss = mean((Q(x) - y)^2) and its gradients (critic.py:83)."""
        B = x.shape[0]
        q, cache = self.q(x)
        diff = q - np.asarray(y, F32)
        g, _ = self.running.backward(cache, (F32(2.0) * diff / F32(B))[:, None])
        return F32(np.mean(diff.astype(np.float64) ** 2)), g, q

    def dq_dx(self, x, dq, target=True):
        """(q, dL/dx) for upstream dL/dq, through the target (default) or the running net."""
        net = self.target if target else self.running
        out, cache = net.forward(x)
        _, dx = net.backward(cache, np.asarray(dq, F32)[:, None], need_dx=True)
        return out[:, 0].astype(F32), dx.astype(F32)


def td_combine(rew, done, q, gamma):
    """R + gamma * (1 - D) * Q in float32, the graph's left-to-right arithmetic."""
    return (np.asarray(rew, F32) + (F32(gamma) * (F32(1.0) - np.asarray(done, F32))) * np.asarray(q, F32)).astype(F32)


def unique(group):
    """The distinct members of a group (a shared group maps every name to ONE member)."""
    seen, out = set(), []
    for n in sorted(group):
        if id(group[n]) not in seen:
            seen.add(id(group[n]))
            out.append(group[n])
    return out


BN_EPS = 1e-3   # snt.BatchNormV2 default ``eps``


def batch_norm_inference(x):
    """``snt.BatchNormV2()(x, is_training=False)`` as the modules call it (matd3module.py:65-74, comamodule.py:71-80,
    maddpgmodule.py:67-76: ``norm(obs, False)``): never trained, so the moving mean stays 0 and the moving variance 1, there is no
    learned scale (Sonnet's default ``scale=False``) and the offset -- not among the variables any optimizer of the fork is given
    (``get_trainable_variables`` returns the running MLP's only, laggingnetwork.py:50-52) -- stays 0:
    y = (x - 0) * rsqrt(1 + eps) + 0."""
    return (np.asarray(x, F32) * (F32(1.0) / np.sqrt(F32(1.0) + F32(BN_EPS)))).astype(F32)


class _Base(object):
    normalize = None

    def _normalized(self, obs=None, rew=None):
        """The ``normalize`` option: {'observation': bool, 'reward': bool} -> inference-mode BatchNorm of the feeds."""
        nz = self.normalize or {}
        if obs is not None and nz.get("observation"):
            obs = {n: batch_norm_inference(obs[n]) for n in obs}
        if rew is not None and nz.get("reward"):
            rew = {n: batch_norm_inference(rew[n]) for n in rew}
        return obs if rew is None else (rew if obs is None else (obs, rew))

    def __init__(self, obs_dims, act_dims, lows, highs, first=None):
        self.names = sorted(obs_dims)
        self.first = first if first is not None else next(iter(obs_dims))   # a shared group is named after the FIRST key
        self.obs_dims, self.act_dims = dict(obs_dims), dict(act_dims)
        self.lows, self.highs = dict(lows), dict(highs)
        self.x_dim = sum(obs_dims.values()) + sum(act_dims.values())

    def _policies(self, rng, shared=False):
        """PolicyGroup (policygroup.py:22-42): one Policy per name, or ONE for every name (equal spaces asserted, :32-34)."""
        if shared:
            f = self.first
            assert all(self.obs_dims[n] == self.obs_dims[f] and self.act_dims[n] == self.act_dims[f] for n in self.names)
            one = PolicyOracle(self.obs_dims[f], self.act_dims[f], self.lows[f], self.highs[f], rng)
            return {n: one for n in self.names}
        return {n: PolicyOracle(self.obs_dims[n], self.act_dims[n], self.lows[n], self.highs[n], rng) for n in self.names}

    def _critics(self, rng, shared=False):
        """CriticGroup (criticgroup.py:21-41)."""
        if shared:
            one = CriticOracle(self.x_dim, rng)
            return {n: one for n in self.names}
        return {n: CriticOracle(self.x_dim, rng) for n in self.names}

    def _critic_step(self, critics, x, y):
        """CriticGroup.create_optimizers (criticgroup.py:87-106): per-name losses, or -- shared -- the first name's loss for every
        name.  -> (losses by name, [(critic, grads)])."""
        if len(unique(critics)) == 1 and len(self.names) > 1:
            l, g, _ = critics[self.first].mse_grads(x, y[self.first])
            return {n: l for n in self.names}, [(critics[self.first], g)]
        losses, steps = {}, []
        for n in self.names:
            losses[n], g, _ = critics[n].mse_grads(x, y[n])
            steps.append((critics[n], g))
        return losses, steps

    def cat(self, d):
        return np.concatenate([np.asarray(d[n], F32).reshape(len(d[n]), -1) for n in self.names], axis=1)

    def _policy_step(self, policies, critics_of, obs, sign=1.0):
        """Every policy's loss -mean(sign * Q_name^target(o, a_all)) and gradient wrt its own variables; -> (losses, grads)."""
        B = len(next(iter(obs.values())))
        acts, tanhs, caches = {}, {}, {}
        for n in self.names:
            acts[n], tanhs[n], caches[n] = policies[n].act(obs[n])
        x = np.concatenate([self.cat(obs), self.cat(acts)], axis=1)
        losses, grads = {}, {}
        o_sum = sum(self.obs_dims.values())
        if len(unique(policies)) == 1 and len(self.names) > 1:
            # shared group (policygroup.py:129-135): ONE loss, -mean(value of the first name), reaching the shared variables
            # through every name's action
            q, dx = critics_of(self.first).dq_dx(x, np.full(B, -sign / B, F32), target=True)
            loss = F32(-np.mean(sign * q.astype(np.float64)))
            total, col = None, o_sum
            for n in self.names:
                K = self.act_dims[n]
                g = policies[n].grads(caches[n], tanhs[n], dx[:, col:col + K])
                total = g if total is None else [a + b for a, b in zip(total, g)]
                col += K
            return {n: loss for n in self.names}, {self.first: total}
        col = o_sum
        for n in self.names:
            q, dx = critics_of(n).dq_dx(x, np.full(B, -sign / B, F32), target=True)
            losses[n] = F32(-np.mean(sign * q.astype(np.float64)))
            K = self.act_dims[n]
            grads[n] = policies[n].grads(caches[n], tanhs[n], dx[:, col:col + K])
            col += K
        return losses, grads


class MaTd3Oracle(_Base):
    GAMMA = 0.9   # matd3module.py:47

    def __init__(self, obs_dims, act_dims, lows, highs, seed=0, shared_policy=False, shared_critic=False, first=None):
        super().__init__(obs_dims, act_dims, lows, highs, first)
        rng = np.random.RandomState(seed)
        self.policies = self._policies(rng, shared_policy)
        self.critics = [self._critics(rng, shared_critic), self._critics(rng, shared_critic)]

    def predict(self, obs):
        obs = self._normalized(obs=obs)
        return {n: self.policies[n].act(obs[n])[0] for n in self.names}

    def compute_values(self, obs):
        acts = self.predict(obs)
        obs = self._normalized(obs=obs)
        x = np.concatenate([self.cat(obs), self.cat(acts)], axis=1)
        return {n: self.critics[0][n].q(x, target=True)[0] for n in self.names}

    def td_targets(self, rew, obs_n, done, z):
        a_n = {n: self.policies[n].act(obs_n[n], target=True, z=z[n])[0] for n in self.names}
        xn = np.concatenate([self.cat(obs_n), self.cat(a_n)], axis=1)
        y = {}
        for n in self.names:
            qmin = np.minimum(self.critics[0][n].q(xn, target=True)[0], self.critics[1][n].q(xn, target=True)[0])
            y[n] = td_combine(np.ravel(rew[n]), np.ravel(done[n]), qmin, self.GAMMA)
        return y, a_n

    def train_step(self, obs, act, rew, obs_n, done, step=None, z=None):
        """z: {name: (B, K) N(0,1) draws} behind ``tf.random.normal`` of the noisy target.  -> {'actor': {...} (only on policy
        steps), 'critic': {...}} like ``unflatten_map(self._train(feed))``."""
        obs, obs_n, rew = self._normalized(obs=obs), self._normalized(obs=obs_n), self._normalized(rew=rew)
        y, _ = self.td_targets(rew, obs_n, done, z)
        x = np.concatenate([self.cat(obs), self.cat(act)], axis=1)
        closs, csteps = [], []
        for c in range(2):
            l, st = self._critic_step(self.critics[c], x, y)
            closs.append(l)
            csteps += st
        out = {"critic": {n: F32(np.mean(np.asarray([closs[0][n], closs[1][n]], F32))) for n in self.names}}
        policy_step = bool(step) and step % 2 == 0
        if policy_step:
            plosses, pgrads = self._policy_step(self.policies, lambda n: self.critics[0][n], obs)
            out["actor"] = plosses
        for cr, g in csteps:
            cr.adam.step(cr.running.p, g)
        if policy_step:
            for n, g in pgrads.items():
                self.policies[n].adam.step(self.policies[n].running.p, g)
        return out

    def run_updates(self):
        for member in unique(self.policies) + unique(self.critics[0]) + unique(self.critics[1]):
            member.update_target()


class ComaOracle(_Base):
    GAMMA = 0.95   # comamodule.py:59

    def __init__(self, obs_dims, act_dims, lows, highs, seed=0, first=None, shared_policy=False):
        super().__init__(obs_dims, act_dims, lows, highs, first)
        rng = np.random.RandomState(seed)
        self.best = self._policies(rng, shared_policy)
        self.worst = self._policies(rng, shared_policy)
        self.global_critic = CriticOracle(self.x_dim, rng)     # CriticGroup(shared=True): always (comamodule.py:36-39)
        self.personal = self._critics(rng)

    def predict(self, obs):
        return {n: self.best[n].act(obs[n])[0] for n in self.names}

    def compute_values(self, obs):
        acts = self.predict(obs)
        x = np.concatenate([self.cat(obs), self.cat(acts)], axis=1)
        return {n: self.personal[n].q(x, target=True)[0] for n in self.names}

    def train_step(self, obs, act, rew, obs_n, done, step=None):
        f = self.first
        obs, obs_n, rew = self._normalized(obs=obs), self._normalized(obs=obs_n), self._normalized(rew=rew)
        worst_n = {n: self.worst[n].act(obs_n[n])[0] for n in self.names}
        best_n = {n: self.best[n].act(obs_n[n])[0] for n in self.names}
        xn_worst = np.concatenate([self.cat(obs_n), self.cat(worst_n)], axis=1)
        xn_best = np.concatenate([self.cat(obs_n), self.cat(best_n)], axis=1)
        worst_q = self.global_critic.q(xn_worst, target=True)[0]
        best_q = self.global_critic.q(xn_best, target=True)[0]
        x = np.concatenate([self.cat(obs), self.cat(act)], axis=1)
        y_global = td_combine(np.ravel(rew[f]), np.ravel(done[f]), best_q, self.GAMMA)
        gl, gg, gq = self.global_critic.mse_grads(x, y_global)
        personal_reward = (gq - worst_q).astype(F32)
        closs, pcg = {}, {}
        for n in self.names:
            qn = self.personal[n].q(xn_best, target=True)[0]
            y = td_combine(personal_reward, np.ravel(done[n]), qn, self.GAMMA)
            l, g, _ = self.personal[n].mse_grads(x, y)
            closs[n] = F32(np.mean(np.asarray([gl, l], F32)))
            pcg[n] = g
        bl, bg = self._policy_step(self.best, lambda n: self.personal[n], obs, sign=1.0)
        wl, wg = self._policy_step(self.worst, lambda n: self.personal[n], obs, sign=-1.0)
        out = {"critic": closs,
               "actor": {n: F32(np.std(np.asarray([bl[n], wl[n]], F32))) for n in self.names}}
        self.global_critic.adam.step(self.global_critic.running.p, gg)
        for n in self.names:
            self.personal[n].adam.step(self.personal[n].running.p, pcg[n])
        for group, grads in ((self.best, bg), (self.worst, wg)):
            for n, g in grads.items():
                group[n].adam.step(group[n].running.p, g)
        return out

    def run_updates(self):
        self.global_critic.update_target()
        for member in unique(self.personal) + unique(self.best) + unique(self.worst):
            member.update_target()


class MaddpgOracle(_Base):
    """The fork's third spelling of MADDPG (maddpg/modules/maddpgmodule.py:52-111, maddpg/algorithms/maddpg.py:11-76): tanh
    policies, ONE critic group, TD actions from the TARGET policies without noise (``.target_actions``, :77), policy loss
    through the critics' TARGET nets (:96-97), both groups step on every train step, polyak 5e-3.  ``gamma`` is
    ``hyperparameters.get('gamma', 0.95)`` -- and the constructor's ``{} if hyperparameters else hyperparameters`` (maddpg.py:19)
    turns any given dict into ``{}`` (and dies on None), so it is always 0.95."""
    GAMMA = 0.95

    def __init__(self, obs_dims, act_dims, lows, highs, seed=0, shared_policy=False, shared_critic=False, first=None):
        super().__init__(obs_dims, act_dims, lows, highs, first)
        rng = np.random.RandomState(seed)
        self.policies = self._policies(rng, shared_policy)
        self.critics = self._critics(rng, shared_critic)

    def predict(self, obs):
        return {n: self.policies[n].act(obs[n])[0] for n in self.names}

    def compute_values(self, obs):
        x = np.concatenate([self.cat(obs), self.cat(self.predict(obs))], axis=1)
        return {n: self.critics[n].q(x, target=True)[0] for n in self.names}

    def train_step(self, obs, act, rew, obs_n, done, step=None):
        obs, obs_n, rew = self._normalized(obs=obs), self._normalized(obs=obs_n), self._normalized(rew=rew)
        a_n = {n: self.policies[n].act(obs_n[n], target=True)[0] for n in self.names}
        xn = np.concatenate([self.cat(obs_n), self.cat(a_n)], axis=1)
        y = {n: td_combine(np.ravel(rew[n]), np.ravel(done[n]), self.critics[n].q(xn, target=True)[0], self.GAMMA)
             for n in self.names}
        x = np.concatenate([self.cat(obs), self.cat(act)], axis=1)
        closs, csteps = self._critic_step(self.critics, x, y)
        plosses, pgrads = self._policy_step(self.policies, lambda n: self.critics[n], obs)
        for cr, g in csteps:
            cr.adam.step(cr.running.p, g)
        for n, g in pgrads.items():
            self.policies[n].adam.step(self.policies[n].running.p, g)
        return {"actor": plosses, "critic": closs}

    def run_updates(self):
        for member in unique(self.policies) + unique(self.critics):
            member.update_target()
