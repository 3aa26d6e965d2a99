"""oracle/matd3.py against torch autograd in float64: the restated MATD3 / best-worst "COMA" losses and gradients
(maddpg/modules/matd3module.py:46-123, comamodule.py:58-171, policy.py:63-100, critic.py:60-88) are rebuilt as torch graphs from
the oracle's own weights and differentiated by autograd.  Parity with the reference's execution is unpinned (oracle header)."""
import numpy as np
import pytest
import torch

from oracle.matd3 import ComaOracle, MaddpgOracle, MaTd3Oracle, NOISE_CLIP, NOISE_STD, POLYAK

NAMES = ["b_agent", "a_agent", "c_agent"]      # insertion order differs from sorted order on purpose
OBS = {"b_agent": 7, "a_agent": 5, "c_agent": 6}
ACT = {"b_agent": 2, "a_agent": 3, "c_agent": 1}
LOW = {"b_agent": -1.0, "a_agent": -2.0, "c_agent": 0.0}
HIGH = {"b_agent": 1.0, "a_agent": 2.0, "c_agent": 3.0}


# equal spaces: what a shared policy group needs (policygroup.py:32-34)
EQ_OBS, EQ_ACT = {n: 6 for n in NAMES}, {n: 2 for n in NAMES}
EQ_LOW, EQ_HIGH = {n: -1.0 for n in NAMES}, {n: 3.0 for n in NAMES}


def make_batch(B, seed, OBS=OBS, ACT=ACT, LOW=LOW, HIGH=HIGH):
    rng = np.random.RandomState(seed)
    obs = {n: rng.randn(B, OBS[n]).astype(np.float32) for n in NAMES}
    obs_n = {n: rng.randn(B, OBS[n]).astype(np.float32) for n in NAMES}
    act = {n: rng.uniform(LOW[n], HIGH[n], (B, ACT[n])).astype(np.float32) for n in NAMES}
    rew = {n: rng.randn(B, 1).astype(np.float32) for n in NAMES}
    done = {n: (rng.rand(B, 1) < 0.2).astype(np.float32) for n in NAMES}
    z = {n: rng.randn(B, ACT[n]).astype(np.float32) for n in NAMES}
    return obs, act, rew, obs_n, done, z


def tparams(mlp, grad=False):
    return [torch.tensor(p.astype(np.float64), requires_grad=grad) for p in mlp.p]


def tmlp(p, x):
    h = torch.relu(x @ p[0] + p[1])
    h = torch.relu(h @ p[2] + p[3])
    return h @ p[4] + p[5]


def tcat(o, d):
    return torch.cat([torch.tensor(np.asarray(d[n], np.float64)) if not torch.is_tensor(d[n]) else d[n] for n in o.names], dim=1)


def tact(pol, p, obs, z=None):
    t = torch.tanh(tmlp(p, torch.tensor(obs.astype(np.float64))))
    if z is not None:
        t = torch.clamp(t + torch.clamp(NOISE_STD * torch.tensor(z.astype(np.float64)), -NOISE_CLIP, NOISE_CLIP), -1, 1)
    return t * float(pol.scale) + float(pol.shift)


def grads_close(got, want, tol=2e-4):
    for g, w in zip(got, want):
        w = w.grad.numpy()
        assert np.abs(g - w).max() <= tol * max(1e-3, np.abs(w).max()), (np.abs(g - w).max(), np.abs(w).max())


def test_matd3_targets_losses_and_gradients_match_autograd():
    o = MaTd3Oracle(OBS, ACT, LOW, HIGH, seed=3)
    obs, act, rew, obs_n, done, z = make_batch(48, 1)
    # TD targets
    y, _ = o.td_targets(rew, obs_n, done, z)
    a_n = {n: tact(o.policies[n], tparams(o.policies[n].target), obs_n[n], z[n]) for n in o.names}
    xn = torch.cat([tcat(o, obs_n), tcat(o, a_n)], dim=1)
    for n in o.names:
        q = torch.minimum(tmlp(tparams(o.critics[0][n].target), xn), tmlp(tparams(o.critics[1][n].target), xn))[:, 0]
        want = rew[n][:, 0] + 0.9 * (1 - done[n][:, 0]) * q.numpy()
        np.testing.assert_allclose(y[n], want, rtol=2e-5, atol=2e-5)
    # critic gradients
    x = torch.cat([tcat(o, obs), tcat(o, act)], dim=1)
    for c in range(2):
        for n in o.names:
            p = tparams(o.critics[c][n].running, grad=True)
            loss = torch.mean((tmlp(p, x)[:, 0] - torch.tensor(y[n].astype(np.float64))) ** 2)
            loss.backward()
            l, g, _ = o.critics[c][n].mse_grads(x.numpy().astype(np.float32), y[n])
            assert abs(l - loss.item()) <= 2e-5 * max(1.0, abs(loss.item()))
            grads_close(g, p)
    # policy gradients: loss_j = -mean(Q1_j^target(o, a_all)), differentiated wrt policy j only
    losses, pg = o._policy_step(o.policies, lambda n: o.critics[0][n], obs)
    for n in o.names:
        ps = {m: tparams(o.policies[m].running, grad=(m == n)) for m in o.names}
        a = {m: tact(o.policies[m], ps[m], obs[m]) for m in o.names}
        xa = torch.cat([tcat(o, obs), tcat(o, a)], dim=1)
        loss = -torch.mean(tmlp(tparams(o.critics[0][n].target), xa))
        loss.backward()
        assert abs(losses[n] - loss.item()) <= 2e-5 * max(1.0, abs(loss.item()))
        grads_close(pg[n], ps[n])


def test_matd3_step_schedule_and_target_update():
    o = MaTd3Oracle(OBS, ACT, LOW, HIGH, seed=4)
    obs, act, rew, obs_n, done, z = make_batch(32, 2)
    p0 = [x.copy() for x in o.policies["a_agent"].running.p]
    c0 = [x.copy() for x in o.critics[1]["a_agent"].running.p]
    for step in (None, 0, 1, 3):     # matd3.py:69: the policies step only when ``step and step % 2 == 0``
        out = o.train_step(obs, act, rew, obs_n, done, step=step, z=z)
        assert "actor" not in out and set(out["critic"]) == set(NAMES)
        assert all(np.array_equal(a, b) for a, b in zip(p0, o.policies["a_agent"].running.p))
    assert not np.array_equal(c0[0], o.critics[1]["a_agent"].running.p[0])
    out = o.train_step(obs, act, rew, obs_n, done, step=2, z=z)
    assert set(out["actor"]) == set(NAMES)
    assert not np.array_equal(p0[0], o.policies["a_agent"].running.p[0])
    # update_targets(5e-3): the target keeps 0.5 % of itself (matd3module.py:104-107 through laggingnetwork.py:36-48)
    t0 = o.policies["b_agent"].target.p[0].copy()
    o.run_updates()
    r = o.policies["b_agent"].running.p[0]
    np.testing.assert_allclose(o.policies["b_agent"].target.p[0], np.float32(POLYAK) * t0 + np.float32(1 - POLYAK) * r, rtol=1e-6)


def test_coma_losses_and_gradients_match_autograd():
    o = ComaOracle(OBS, ACT, LOW, HIGH, seed=5, first="b_agent")
    obs, act, rew, obs_n, done, _ = make_batch(40, 3)
    f = "b_agent"
    worst_n = {n: tact(o.worst[n], tparams(o.worst[n].running), obs_n[n]) for n in o.names}
    best_n = {n: tact(o.best[n], tparams(o.best[n].running), obs_n[n]) for n in o.names}
    xw = torch.cat([tcat(o, obs_n), tcat(o, worst_n)], dim=1)
    xb = torch.cat([tcat(o, obs_n), tcat(o, best_n)], dim=1)
    gt = tparams(o.global_critic.target)
    worst_q, best_q = tmlp(gt, xw)[:, 0], tmlp(gt, xb)[:, 0]
    x = torch.cat([tcat(o, obs), tcat(o, act)], dim=1)
    R, D = (lambda n: torch.tensor(rew[n][:, 0].astype(np.float64))), (lambda n: torch.tensor(done[n][:, 0].astype(np.float64)))
    gp = tparams(o.global_critic.running, grad=True)
    gq = tmlp(gp, x)[:, 0]
    gloss = torch.mean((gq - (R(f) + 0.95 * (1 - D(f)) * best_q).detach()) ** 2)
    gloss.backward()
    want_critic, want_pgrads = {}, {}
    for n in o.names:
        pp = tparams(o.personal[n].running, grad=True)
        y = (gq - worst_q).detach() + 0.95 * (1 - D(n)) * tmlp(tparams(o.personal[n].target), xb)[:, 0]
        loss = torch.mean((tmlp(pp, x)[:, 0] - y.detach()) ** 2)
        loss.backward()
        want_critic[n] = 0.5 * (gloss.item() + loss.item())
        want_pgrads[n] = pp
    want_actor, want_best, want_worst = {}, {}, {}
    for n in o.names:
        ls = []
        for group, sign, store in ((o.best, 1.0, want_best), (o.worst, -1.0, want_worst)):
            ps = {m: tparams(group[m].running, grad=(m == n)) for m in o.names}
            a = {m: tact(group[m], ps[m], obs[m]) for m in o.names}
            xa = torch.cat([tcat(o, obs), tcat(o, a)], dim=1)
            loss = -torch.mean(sign * tmlp(tparams(o.personal[n].target), xa))
            loss.backward()
            ls.append(loss.item())
            store[n] = ps[n]
        want_actor[n] = float(np.std(ls))
    # the oracle's own gradients, taken before its Adam steps move anything
    bl, bg = o._policy_step(o.best, lambda n: o.personal[n], obs, sign=1.0)
    wl, wg = o._policy_step(o.worst, lambda n: o.personal[n], obs, sign=-1.0)
    for n in o.names:
        grads_close(bg[n], want_best[n])
        grads_close(wg[n], want_worst[n])
    before = [p.copy() for p in o.global_critic.running.p]
    out = o.train_step(obs, act, rew, obs_n, done)
    for n in o.names:
        assert abs(out["critic"][n] - want_critic[n]) <= 5e-5 * max(1.0, abs(want_critic[n]))
        assert abs(out["actor"][n] - want_actor[n]) <= 5e-5 * max(1.0, abs(want_actor[n]))
    # first Adam step: every weight with a non-negligible gradient moved by ~lr against the gradient's sign
    g = gp[0].grad.numpy()
    moved = o.global_critic.running.p[0] - before[0]
    big = np.abs(g) > 1e-4
    assert big.any() and np.all(np.sign(moved[big]) == -np.sign(g[big]))
    np.testing.assert_allclose(np.abs(moved[big]), 1e-4, rtol=1e-2)


def test_shared_groups_have_one_loss_through_every_action():
    """PolicyGroup / CriticGroup with shared=True (policygroup.py:26-37, 54-70, 129-135; criticgroup.py:24-34, 48-66, 94-100)."""
    o = MaTd3Oracle(EQ_OBS, EQ_ACT, EQ_LOW, EQ_HIGH, seed=8, shared_policy=True, shared_critic=True, first="b_agent")
    assert len({id(p) for p in o.policies.values()}) == 1 and len({id(c) for c in o.critics[1].values()}) == 1
    obs, act, rew, obs_n, done, z = make_batch(36, 4, EQ_OBS, EQ_ACT, EQ_LOW, EQ_HIGH)
    f = "b_agent"
    y, _ = o.td_targets(rew, obs_n, done, z)
    x = torch.cat([tcat(o, obs), tcat(o, act)], dim=1)
    # the shared critic regresses on the first name's targets only
    p = tparams(o.critics[0][f].running, grad=True)
    loss = torch.mean((tmlp(p, x)[:, 0] - torch.tensor(y[f].astype(np.float64))) ** 2)
    loss.backward()
    losses, steps = o._critic_step(o.critics[0], x.numpy().astype(np.float32), y)
    assert len(steps) == 1 and all(abs(losses[n] - loss.item()) <= 2e-5 * max(1.0, loss.item()) for n in NAMES)
    grads_close(steps[0][1], p)
    # the shared policy: one loss, gradients through every name's action slice
    ps = tparams(o.policies[f].running, grad=True)
    a = {m: tact(o.policies[m], ps, obs[m]) for m in o.names}
    xa = torch.cat([tcat(o, obs), tcat(o, a)], dim=1)
    ploss = -torch.mean(tmlp(tparams(o.critics[0][f].target), xa))
    ploss.backward()
    pl, pg = o._policy_step(o.policies, lambda n: o.critics[0][n], obs)
    assert list(pg) == [f] and all(abs(pl[n] - ploss.item()) <= 2e-5 * max(1.0, abs(ploss.item())) for n in NAMES)
    grads_close(pg[f], ps)
    t_before = o.policies[f].adam.t
    out = o.train_step(obs, act, rew, obs_n, done, step=2, z=z)
    assert o.policies[f].adam.t == t_before + 1 and o.critics[0][f].adam.t == 1     # ONE optimizer step per shared member
    assert len(set(float(v) for v in out["actor"].values())) == 1


def test_fork_maddpg_targets_and_losses_match_autograd():
    o = MaddpgOracle(OBS, ACT, LOW, HIGH, seed=9)
    obs, act, rew, obs_n, done, _ = make_batch(44, 6)
    a_n = {n: tact(o.policies[n], tparams(o.policies[n].target), obs_n[n]) for n in o.names}     # target policies, no noise
    xn = torch.cat([tcat(o, obs_n), tcat(o, a_n)], dim=1)
    x = torch.cat([tcat(o, obs), tcat(o, act)], dim=1)
    want_c, want_a = {}, {}
    for n in o.names:
        y = torch.tensor(rew[n][:, 0].astype(np.float64)) + 0.95 * (1 - torch.tensor(done[n][:, 0].astype(np.float64))) * \
            tmlp(tparams(o.critics[n].target), xn)[:, 0]
        want_c[n] = torch.mean((tmlp(tparams(o.critics[n].running), x)[:, 0] - y) ** 2).item()
        a = {m: tact(o.policies[m], tparams(o.policies[m].running), obs[m]) for m in o.names}
        xa = torch.cat([tcat(o, obs), tcat(o, a)], dim=1)
        want_a[n] = -torch.mean(tmlp(tparams(o.critics[n].target), xa)).item()
    out = o.train_step(obs, act, rew, obs_n, done)
    for n in o.names:
        assert abs(out["critic"][n] - want_c[n]) <= 5e-5 * max(1.0, abs(want_c[n]))
        assert abs(out["actor"][n] - want_a[n]) <= 5e-5 * max(1.0, abs(want_a[n]))
    assert o.policies["a_agent"].adam.t == 1 and o.critics["a_agent"].adam.t == 1


@pytest.mark.parametrize("cls", [MaTd3Oracle, ComaOracle])
def test_sorted_name_order(cls):
    o = cls(OBS, ACT, LOW, HIGH, seed=0)
    assert o.names == sorted(NAMES)      # U.concat_map sorts by key (tf_util.py:53-55)
    if cls is ComaOracle:
        assert o.first == "b_agent"      # the shared group is named after the FIRST key in insertion order (criticgroup.py:24)


def test_oracle_matches_the_fork_graph_code():
    """tests/golden/fork_graph_ref.npz: the fork's OWN ``Coma`` / ``Maddpg`` algorithm classes, modules, groups, policies, critics
    and ``TfFunction`` plumbing executed unmodified on the torch-backed stand-in for TensorFlow + Sonnet (tests/tf_shim.py,
    tests/golden/make_fork_graph_golden.py).  The restated oracle must reproduce the losses ``train_step`` returned, the
    predictions, the values and every variable after three train steps + target updates -- i.e. the same wiring: shared global
    critic on the first name's reward, personal reward, worst-policy sign, losses through the TARGET critics, per-optimizer
    variable sets, the 5e-3 polyak.  Also recorded there: the reference's ``MaTd3`` cannot be constructed as written; it is run with
    the one crashing call made tolerant of its extra argument."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fork_graph_ref.npz"))
    assert "takes 2 positional arguments but 3 were given" in str(gold["matd3_error"])        # matd3module.py:98-99
    assert "'NoneType' object has no attribute 'get'" in str(gold["maddpg_none_hyperparameters_error"])   # maddpg.py:19-29
    assert str(gold["coma_unequal_spaces_error"]) == "AssertionError"                          # criticgroup.py:28-30
    eq = (EQ_OBS, EQ_ACT, EQ_LOW, EQ_HIGH)
    worst = {"loss": 0.0, "var": 0.0}

    def drive(o, prefix, dims):
        for step in (1, 2, 3):
            obs, act, rew, obs_n, done, _ = make_batch(48, 1000 + step, *dims)
            res = o.train_step(obs, act, rew, obs_n, done, step=step)
            o.run_updates()
            for kind in ("actor", "critic"):
                got, want = np.asarray([res[kind][n] for n in NAMES], np.float64), gold["%s_s%d_%s" % (prefix, step, kind)]
                worst["loss"] = max(worst["loss"], float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-2))))
                np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-7, err_msg="%s step %d %s" % (prefix, step, kind))
        obs = make_batch(16, 2000, *dims)[0]
        pred, val = o.predict(obs), o.compute_values(obs)
        for n in NAMES:
            np.testing.assert_allclose(pred[n], gold["%s_predict_%s" % (prefix, n)], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(val[n], gold["%s_values_%s" % (prefix, n)], rtol=1e-5, atol=1e-6)

    def check(member, key):
        for net, attr in (("running", member.running), ("target", member.target)):
            for i, w in enumerate(attr.p):
                d = float(np.abs(w - gold["%s_%s_%d" % (key, net, i)]).max())
                worst["var"] = max(worst["var"], d)
                assert d <= 2e-6, (key, net, i, d)      # 2 % of one Adam step (lr = 1e-4)

    o = ComaOracle(*eq, seed=61, first=NAMES[0])
    drive(o, "coma", eq)
    for n in NAMES:
        check(o.best[n], "coma_best_" + n)
        check(o.worst[n], "coma_worst_" + n)
        check(o.personal[n], "coma_personal_" + n)
    check(o.global_critic, "coma_global")
    m = MaddpgOracle(OBS, ACT, LOW, HIGH, seed=62, first=NAMES[0])
    drive(m, "maddpg", (OBS, ACT, LOW, HIGH))
    for n in NAMES:
        check(m.policies[n], "maddpg_policy_" + n)
        check(m.critics[n], "maddpg_critic_" + n)
    # shared groups: one policy / one critic for every name, ONE loss (the first name's) -- policygroup.py:129-135, criticgroup.py:94-100
    ms = MaddpgOracle(*eq, seed=64, shared_policy=True, shared_critic=True, first=NAMES[0])
    drive(ms, "maddpg_shared", eq)
    check(ms.policies[NAMES[0]], "maddpg_shared_policy")
    check(ms.critics[NAMES[0]], "maddpg_shared_critic")
    cs = ComaOracle(*eq, seed=65, first=NAMES[0], shared_policy=True)
    for step in (1, 2):
        obs, act, rew, obs_n, done, _ = make_batch(48, 1000 + step, *eq)
        res = cs.train_step(obs, act, rew, obs_n, done, step=step)
        cs.run_updates()
        for kind in ("actor", "critic"):
            np.testing.assert_allclose(np.asarray([res[kind][n] for n in NAMES], np.float64), gold["coma_shared_s%d_%s" % (step, kind)],
                                       rtol=2e-5, atol=2e-7, err_msg="coma shared step %d %s" % (step, kind))
    check(cs.best[NAMES[0]], "coma_shared_best")
    check(cs.worst[NAMES[0]], "coma_shared_worst")
    # MaTd3: the reference's graph only builds once PolicyGroup.create_optimizers tolerates the extra argument MaTD3Module passes
    # (the generator's one modification of reference code); twin critics, min of the targets, noisy target actions, delayed policies
    t = MaTd3Oracle(OBS, ACT, LOW, HIGH, seed=63)
    for step in (1, 2, 3, 4):
        obs, act, rew, obs_n, done, z = make_batch(48, 3000 + step)
        res = t.train_step(obs, act, rew, obs_n, done, step=step, z=z)
        t.run_updates()
        assert ("actor" in res) == (step % 2 == 0)
        for kind in res:
            got, want = np.asarray([res[kind][n] for n in NAMES], np.float64), gold["matd3_s%d_%s" % (step, kind)]
            worst["loss"] = max(worst["loss"], float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-2))))
            np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-7, err_msg="matd3 step %d %s" % (step, kind))
    assert list(gold["matd3_s1_raw_keys"]) == ["a", "b", "c"]     # the reference's critic-only result: names split at "_"
    obs = make_batch(16, 2000)[0]
    pred, val = t.predict(obs), t.compute_values(obs)
    for n in NAMES:
        np.testing.assert_allclose(pred[n], gold["matd3_predict_%s" % n], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(val[n], gold["matd3_values_%s" % n], rtol=1e-5, atol=1e-6)
        check(t.policies[n], "matd3_policy_" + n)
        check(t.critics[0][n], "matd3_critic0_" + n)
        check(t.critics[1][n], "matd3_critic1_" + n)
    print("worst relative loss difference %.2e, worst variable difference %.2e" % (worst["loss"], worst["var"]))
