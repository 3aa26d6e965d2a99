"""The fork's dict-of-agents algorithms (maddpg_b200/algorithms.py: MaTd3, Coma over csrc/mdp_td3.cu) against the numpy oracle
(oracle/matd3.py) on the same weights, batches and noise draws: losses within 1e-4 relative (BASELINE.json north_star's
tolerance for Q-values and losses), parameters after the Adam steps, predictions, values and the target update.
Reference: maddpg/algorithms/matd3.py:11-81, coma.py:11-63, modules/matd3module.py:46-123, comamodule.py:58-171."""
import numpy as np
import pytest
import torch

from oracle.matd3 import ComaOracle, MaddpgOracle, MaTd3Oracle
from tests.test_oracle_matd3 import ACT, EQ_ACT, EQ_HIGH, EQ_LOW, EQ_OBS, HIGH, LOW, NAMES, OBS, make_batch

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def spaces(OBS=OBS, ACT=ACT, LOW=LOW, HIGH=HIGH):
    from maddpg_b200.spaces import Box, Dict
    obs = Dict({n: Box(-np.inf, np.inf, (OBS[n],)) for n in NAMES})
    act = Dict({n: Box(np.full(ACT[n], LOW[n], np.float32), np.full(ACT[n], HIGH[n], np.float32), (ACT[n],)) for n in NAMES})
    return obs, act


EQ = (EQ_OBS, EQ_ACT, EQ_LOW, EQ_HIGH)


def load_policy(core, o_group, names):
    from maddpg_b200 import _lib
    for j, n in enumerate(names):
        core.set_weights(j, _lib.NET_P, o_group[n].running.p)
        core.set_weights(j, _lib.NET_TARGET_P, o_group[n].target.p)


def load_critic(core, o_group, names):
    from maddpg_b200 import _lib
    for j, n in enumerate(names):
        core.set_weights(j, _lib.NET_Q, o_group[n].running.p)
        core.set_weights(j, _lib.NET_TARGET_Q, o_group[n].target.p)


def params_close(core, net, j, want, what, lr=1e-4):
    """Adam's first steps move every weight by ~lr * sign(g): an element whose gradient is rounding noise may step the other way
    (2 lr apart).  Everything else must agree to a few percent of one step."""
    for k, (w, r) in enumerate(zip(core.get_weights(j, net), want)):
        d = np.abs(w - r)
        assert d.max() <= 2.2 * lr * 3, "%s[%d]: max |diff| %g" % (what, k, d.max())
        assert np.mean(d > 0.05 * lr) <= 0.01, "%s[%d]: %.3f%% of the elements differ by more than 5%% of a step" % (
            what, k, 100 * np.mean(d > 0.05 * lr))


def losses_close(got, want, what):
    assert set(got) == set(want), what
    for n in want:
        assert abs(float(got[n]) - float(want[n])) <= RTOL * max(abs(float(want[n])), 1e-2), (what, n, got[n], want[n])


@pytest.mark.parametrize("B", [48, 1024])
def test_matd3_train_steps_match_oracle(B):
    from maddpg_b200 import _lib
    from maddpg_b200.algorithms import MaTd3
    o = MaTd3Oracle(OBS, ACT, LOW, HIGH, seed=11)
    alg = MaTd3(*spaces(), seed=1)
    assert alg.names == o.names
    load_policy(alg.policies, o.policies, o.names)
    for c in range(2):
        load_critic(alg.critics[c], o.critics[c], o.names)
    for step in (1, 2, 3, 4):     # critic-only, full, critic-only, full (matd3.py:69)
        obs, act, rew, obs_n, done, z = make_batch(B, 100 + step)
        want = o.train_step(obs, act, rew, obs_n, done, step=step, z=z)
        got = alg.train_step(obs, act, rew, obs_n, done, step=step, noise=z)
        assert ("actor" in got) == ("actor" in want) == (step % 2 == 0)
        for key in want:
            losses_close(got[key], want[key], "step %d %s" % (step, key))
        o.run_updates()
        alg.run_updates()
    for j, n in enumerate(o.names):
        params_close(alg.policies, _lib.NET_P, j, o.policies[n].running.p, "policy " + n)
        params_close(alg.policies, _lib.NET_TARGET_P, j, o.policies[n].target.p, "target policy " + n)
        for c in range(2):
            params_close(alg.critics[c], _lib.NET_Q, j, o.critics[c][n].running.p, "critic %d %s" % (c, n))
            params_close(alg.critics[c], _lib.NET_TARGET_Q, j, o.critics[c][n].target.p, "target critic %d %s" % (c, n))
    assert alg.policies.adam_t.cpu().tolist() == [2, 0] * 3 and alg.critics[1].adam_t.cpu().tolist() == [0, 4] * 3
    # predictions and values of the trained nets on a fresh batch (the oracle evaluated on the DEVICE's weights: the comparison
    # is of the forward arithmetic, not of four accumulated Adam steps)
    for j, n in enumerate(o.names):
        o.policies[n].running.p = alg.policies.get_weights(j, _lib.NET_P)
        o.critics[0][n].target.p = alg.critics[0].get_weights(j, _lib.NET_TARGET_Q)
    obs = make_batch(B, 999)[0]
    want_a, got_a = o.predict(obs), alg.predict(obs, noisy=False)
    want_v, got_v = o.compute_values(obs), alg.compute_values(obs)
    for n in o.names:
        np.testing.assert_allclose(got_a[n].reshape(B, -1), want_a[n], rtol=RTOL, atol=2e-6)
        np.testing.assert_allclose(got_v[n][:, 0], want_v[n], rtol=RTOL, atol=RTOL * np.abs(want_v[n]).mean())
    noisy = alg.predict(obs, noisy=True)      # + N(0, 0.2) on the host (multiagentalgbase.py:62-64)
    d = np.concatenate([(noisy[n].reshape(B, -1) - got_a[n].reshape(B, -1)).ravel() for n in o.names])
    assert 0.15 < d.std() < 0.25


def test_wide_observations_take_the_streaming_tiles():
    """Nets too large to stay resident in shared memory (critic input 3 x 44 + 6 columns) stream their weights in chunks: the other
    template branch of every csrc/mdp_td3.cu kernel, same arithmetic."""
    from maddpg_b200.algorithms import Coma, MaTd3
    W_OBS, W_ACT = {n: 44 for n in NAMES}, {n: 2 for n in NAMES}
    W = (W_OBS, W_ACT, EQ_LOW, EQ_HIGH)
    B = 80
    o = MaTd3Oracle(*W, seed=31)
    alg = MaTd3(*spaces(*W), seed=9)
    load_policy(alg.policies, o.policies, o.names)
    for c in range(2):
        load_critic(alg.critics[c], o.critics[c], o.names)
    for step in (1, 2):
        obs, act, rew, obs_n, done, z = make_batch(B, 500 + step, *W)
        want = o.train_step(obs, act, rew, obs_n, done, step=step, z=z)
        got = alg.train_step(obs, act, rew, obs_n, done, step=step, noise=z)
        for key in want:
            losses_close(got[key], want[key], "wide step %d %s" % (step, key))
    from maddpg_b200 import _lib
    for j, n in enumerate(o.names):
        params_close(alg.policies, _lib.NET_P, j, o.policies[n].running.p, "wide policy " + n)
        params_close(alg.critics[0], _lib.NET_Q, j, o.critics[0][n].running.p, "wide critic " + n)
    oc = ComaOracle(*W, seed=32, first=NAMES[0])
    ac = Coma(*spaces(*W), seed=10)
    load_policy(ac.best, oc.best, oc.names)
    load_policy(ac.worst, oc.worst, oc.names)
    load_critic(ac.personal, oc.personal, oc.names)
    load_critic(ac.global_critic, {n: oc.global_critic for n in oc.names}, oc.names)
    obs, act, rew, obs_n, done, _ = make_batch(B, 600, *W)
    want = oc.train_step(obs, act, rew, obs_n, done)
    got = ac.train_step(obs, act, rew, obs_n, done)
    losses_close(got["critic"], want["critic"], "wide coma critic")
    losses_close(got["actor"], want["actor"], "wide coma actor")


def test_matd3_compute_loss_leaves_the_state_untouched_and_philox_noise_is_sane():
    from maddpg_b200.algorithms import MaTd3
    alg = MaTd3(*spaces(), seed=2)
    obs, act, rew, obs_n, done, z = make_batch(256, 5)
    before = [c.params.clone() for c in alg._cores]
    l1 = alg.compute_loss(obs, act, rew, obs_n, done, noise=z)
    l2 = alg.compute_loss(obs, act, rew, obs_n, done, noise=z)
    assert set(l1) == {"actor", "critic"}
    for key in l1:
        for n in l1[key]:
            assert float(l1[key][n]) == pytest.approx(float(l2[key][n]), rel=1e-6)
    for c, b in zip(alg._cores, before):
        assert torch.equal(c.params, b) and float(c.grads.abs().max()) == 0.0 and int(c.adam_t.abs().max()) == 0
    # in-kernel target noise (no injected draws): clip(N(0, 0.2), -0.5, 0.5) around the clean target action
    nx = alg._rows(obs, act, rew, obs_n, done)
    L = alg.layout
    view = nx[:, int(L.nx_off):]
    clean = alg._policy_act(alg.policies, view, nx.stride(0), alg._act_buf("t_clean", 256), use_target=True).clone()
    noisy = alg._policy_act(alg.policies, view, nx.stride(0), alg._act_buf("t_noisy", 256), use_target=True, noise_std=0.2).clone()
    scale = np.concatenate([np.full(ACT[n], (HIGH[n] - LOW[n]) / 2) for n in alg.names])
    d = ((noisy - clean)[:, :scale.size].cpu().numpy() / scale)
    inside = np.abs(clean[:, :scale.size].cpu().numpy() / scale) < 10     # every column
    assert np.abs(d).max() <= 0.5 + 1e-6 and 0.12 < d[inside].std() < 0.22 and abs(d.mean()) < 0.03
    alg.train_step(obs, act, rew, obs_n, done, step=2)     # Philox path end to end
    assert all(torch.isfinite(c.params).all() for c in alg._cores)


@pytest.mark.parametrize("B", [40, 1024])
def test_coma_train_steps_match_oracle(B):
    from maddpg_b200 import _lib
    from maddpg_b200.algorithms import Coma
    o = ComaOracle(*EQ, seed=12, first=NAMES[0])      # equal spaces: ComaModule's shared global critic asserts them
    alg = Coma(*spaces(*EQ), seed=3)
    assert alg.names == o.names and alg.first == NAMES[0] and alg.names[alg.shared] == NAMES[0]
    load_policy(alg.best, o.best, o.names)
    load_policy(alg.worst, o.worst, o.names)
    load_critic(alg.personal, o.personal, o.names)
    load_critic(alg.global_critic, {n: o.global_critic for n in o.names}, o.names)
    for step in (1, 2, 3):
        obs, act, rew, obs_n, done, _ = make_batch(B, 200 + step, *EQ)
        want = o.train_step(obs, act, rew, obs_n, done, step=step)
        got = alg.train_step(obs, act, rew, obs_n, done, step=step)
        losses_close(got["critic"], want["critic"], "step %d critic" % step)
        for n in o.names:     # std of two nearly opposite losses: compare at the losses' own scale
            assert abs(float(got["actor"][n]) - float(want["actor"][n])) <= RTOL * max(abs(float(want["actor"][n])), 1e-2)
        o.run_updates()
        alg.run_updates()
    s = alg.shared
    params_close(alg.global_critic, _lib.NET_Q, s, o.global_critic.running.p, "global critic")
    params_close(alg.global_critic, _lib.NET_TARGET_Q, s, o.global_critic.target.p, "global target critic")
    for j, n in enumerate(o.names):
        params_close(alg.best, _lib.NET_P, j, o.best[n].running.p, "best " + n)
        params_close(alg.worst, _lib.NET_P, j, o.worst[n].running.p, "worst " + n)
        params_close(alg.worst, _lib.NET_TARGET_P, j, o.worst[n].target.p, "worst target " + n)
        params_close(alg.personal, _lib.NET_Q, j, o.personal[n].running.p, "personal " + n)
        params_close(alg.personal, _lib.NET_TARGET_Q, j, o.personal[n].target.p, "personal target " + n)


@pytest.mark.parametrize("shared_policy,shared_critic", [(True, False), (False, True), (True, True)])
def test_matd3_shared_groups_match_oracle(shared_policy, shared_critic):
    """PolicyGroup / CriticGroup(shared=True): one member (the first name's), one loss (policygroup.py:129-135,
    criticgroup.py:94-100)."""
    from maddpg_b200 import _lib
    from maddpg_b200.algorithms import MaTd3
    B = 192
    o = MaTd3Oracle(*EQ, seed=21, shared_policy=shared_policy, shared_critic=shared_critic, first=NAMES[0])
    alg = MaTd3(*spaces(*EQ), shared_policy=shared_policy, shared_critic=shared_critic, seed=7)
    f = alg.names.index(NAMES[0])
    assert alg.sp == (f if shared_policy else -1) and alg.sc == (f if shared_critic else -1)
    load_policy(alg.policies, o.policies, o.names)
    for c in range(2):
        load_critic(alg.critics[c], o.critics[c], o.names)
    for step in (2, 3, 4):
        obs, act, rew, obs_n, done, z = make_batch(B, 300 + step, *EQ)
        want = o.train_step(obs, act, rew, obs_n, done, step=step, z=z)
        got = alg.train_step(obs, act, rew, obs_n, done, step=step, noise=z)
        for key in want:
            losses_close(got[key], want[key], "step %d %s" % (step, key))
        o.run_updates()
        alg.run_updates()
    for j, n in enumerate(o.names):
        if not shared_policy or j == f:
            params_close(alg.policies, _lib.NET_P, j, o.policies[n].running.p, "policy " + n)
            params_close(alg.policies, _lib.NET_TARGET_P, j, o.policies[n].target.p, "target policy " + n)
        if not shared_critic or j == f:
            params_close(alg.critics[1], _lib.NET_Q, j, o.critics[1][n].running.p, "critic " + n)
    t = alg.policies.adam_t.cpu().tolist()
    assert [t[2 * j] for j in range(3)] == ([2 if j == f else 0 for j in range(3)] if shared_policy else [2, 2, 2])
    tc = alg.critics[0].adam_t.cpu().tolist()
    assert [tc[2 * j + 1] for j in range(3)] == ([3 if j == f else 0 for j in range(3)] if shared_critic else [3, 3, 3])
    # predictions of a shared policy: the same net on every name's observation
    obs = make_batch(B, 998, *EQ)[0]
    for j, n in enumerate(o.names):
        src = f if shared_policy else j
        o.policies[n].running.p = alg.policies.get_weights(src, _lib.NET_P)
    want_a, got_a = o.predict(obs), alg.predict(obs, noisy=False)
    for n in o.names:
        np.testing.assert_allclose(got_a[n].reshape(B, -1), want_a[n], rtol=RTOL, atol=2e-6)


def test_coma_shared_policy_matches_oracle():
    from maddpg_b200 import _lib
    from maddpg_b200.algorithms import Coma
    B = 160
    o = ComaOracle(*EQ, seed=22, first=NAMES[0], shared_policy=True)
    alg = Coma(*spaces(*EQ), shared_policy=True, seed=8)
    load_policy(alg.best, o.best, o.names)
    load_policy(alg.worst, o.worst, o.names)
    load_critic(alg.personal, o.personal, o.names)
    load_critic(alg.global_critic, {n: o.global_critic for n in o.names}, o.names)
    for step in (1, 2):
        obs, act, rew, obs_n, done, _ = make_batch(B, 400 + step, *EQ)
        want = o.train_step(obs, act, rew, obs_n, done, step=step)
        got = alg.train_step(obs, act, rew, obs_n, done, step=step)
        losses_close(got["critic"], want["critic"], "step %d critic" % step)
        losses_close(got["actor"], want["actor"], "step %d actor" % step)
        o.run_updates()
        alg.run_updates()
    f = alg.shared
    params_close(alg.best, _lib.NET_P, f, o.best[NAMES[0]].running.p, "shared best policy")
    params_close(alg.worst, _lib.NET_P, f, o.worst[NAMES[0]].running.p, "shared worst policy")
    with pytest.raises(AssertionError):
        Coma(*spaces(), shared_policy=True)     # unequal spaces (policygroup.py:32-34)


@pytest.mark.parametrize("shared", [False, True])
def test_fork_maddpg_and_inference_classes_match_oracle(shared):
    """maddpg/algorithms/maddpg.py:11-118, coma.py:73-112."""
    from maddpg_b200 import _lib
    from maddpg_b200.algorithms import ComaInference, Maddpg, MaddpgInference
    B = 128
    dims = EQ if shared else (OBS, ACT, LOW, HIGH)
    o = MaddpgOracle(*dims, seed=41, shared_policy=shared, shared_critic=shared, first=NAMES[0])
    alg = Maddpg(*spaces(*dims), shared_policy=shared, shared_critic=shared, hyperparameters={"gamma": 0.5}, seed=11)
    assert alg.GAMMA == 0.95      # the reference discards the given hyperparameters (maddpg.py:19)
    load_policy(alg.policies, o.policies, o.names)
    load_critic(alg.critics, o.critics, o.names)
    for step in (1, 2, 3):
        obs, act, rew, obs_n, done, _ = make_batch(B, 700 + step, *dims)
        want = o.train_step(obs, act, rew, obs_n, done, step=step)
        got = alg.train_step(obs, act, rew, obs_n, done, step=step)
        losses_close(got["critic"], want["critic"], "step %d critic" % step)
        losses_close(got["actor"], want["actor"], "step %d actor" % step)
        o.run_updates()
        alg.run_updates()
    f = alg.names.index(NAMES[0])
    for j, n in enumerate(o.names):
        if not shared or j == f:
            params_close(alg.policies, _lib.NET_P, j, o.policies[n].running.p, "policy " + n)
            params_close(alg.policies, _lib.NET_TARGET_P, j, o.policies[n].target.p, "target policy " + n)
            params_close(alg.critics, _lib.NET_Q, j, o.critics[n].running.p, "critic " + n)
    # the inference-only classes: predict() of a policy group, everything else returns None like the reference's `...` bodies
    obs = make_batch(B, 997, *dims)[0]
    for cls in (MaddpgInference, ComaInference):
        inf = cls(*spaces(*dims), shared_policy=shared, seed=12)
        for j, n in enumerate(o.names):
            inf.policies.set_weights(j, _lib.NET_P, alg.policies.get_weights(f if shared else j, _lib.NET_P))
        a, b = inf.predict(obs, noisy=False), alg.predict(obs, noisy=False)
        assert all(np.array_equal(a[n], b[n]) for n in o.names)
        assert inf.train_step(obs, obs, obs, obs, obs) is None and inf.compute_values(obs) is None and inf.run_updates() is None


@pytest.mark.parametrize("cls_name", ["MaTd3", "Coma"])
def test_normalize_option_is_inference_mode_batch_norm(cls_name):
    """normalize={'observation': True, 'reward': True}: ``snt.BatchNormV2()(x, False)`` of the feeds (matd3module.py:65-74,
    comamodule.py:71-80) -- with never-updated moving statistics a constant gain rsqrt(1 + 1e-3)."""
    from maddpg_b200 import algorithms
    B = 96
    nz = {"observation": True, "reward": True}
    if cls_name == "MaTd3":
        o = MaTd3Oracle(OBS, ACT, LOW, HIGH, seed=51)
        alg = algorithms.MaTd3(*spaces(), normalize=nz, seed=13)
        load_policy(alg.policies, o.policies, o.names)
        for c in range(2):
            load_critic(alg.critics[c], o.critics[c], o.names)
    else:
        o = ComaOracle(*EQ, seed=52, first=NAMES[0])
        alg = algorithms.Coma(*spaces(*EQ), normalize=nz, seed=14)
        load_policy(alg.best, o.best, o.names)
        load_policy(alg.worst, o.worst, o.names)
        load_critic(alg.personal, o.personal, o.names)
        load_critic(alg.global_critic, {n: o.global_critic for n in o.names}, o.names)
    o.normalize = nz
    dims = (OBS, ACT, LOW, HIGH) if cls_name == "MaTd3" else EQ
    obs, act, rew, obs_n, done, z = make_batch(B, 800, *dims)
    obs = {n: 3.0 * v for n, v in obs.items()}      # make the 0.05 % gain matter at the 1e-4 tolerance
    rew = {n: 5.0 * v for n, v in rew.items()}
    kw_o, kw_a = ({"z": z}, {"noise": z}) if cls_name == "MaTd3" else ({}, {})
    want = o.train_step(obs, act, rew, obs_n, done, step=2, **kw_o)
    got = alg.train_step(obs, act, rew, obs_n, done, step=2, **kw_a)
    for key in want:
        losses_close(got[key], want[key], "normalized %s" % key)
    o.normalize = None                                # and the gain is really there: the un-normalised losses differ
    o2 = o.train_step(obs, act, rew, obs_n, done, step=2, **kw_o)
    assert any(abs(float(o2["critic"][n]) - float(want["critic"][n])) > 3e-4 * abs(float(want["critic"][n])) for n in o.names)
    want_a = o.predict(obs)                           # (oracle weights moved by two steps: compare like with like)
    assert set(alg.predict(obs, noisy=False)) == set(want_a)


@pytest.mark.parametrize("cls_name", ["MaTd3", "Coma", "Maddpg"])
def test_graph_replayed_steps_equal_eager_steps(cls_name):
    """From the second call of a (batch, step kind) on a train step replays as one CUDA graph: same losses and parameters as the
    launch-by-launch path, and the in-kernel target noise of a replay differs from the previous replay's."""
    from maddpg_b200 import algorithms
    cls = getattr(algorithms, cls_name)
    dims = EQ if cls_name == "Coma" else (OBS, ACT, LOW, HIGH)
    a, b = cls(*spaces(*dims), seed=15), cls(*spaces(*dims), seed=15)
    b.use_graphs = False
    for ca, cb in zip(a._cores, b._cores):
        assert torch.equal(ca.params, cb.params)
    B = 128
    kw = lambda z: {"noise": z} if cls_name == "MaTd3" else {}
    for step in range(1, 7):
        obs, act, rew, obs_n, done, z = make_batch(B, 900 + step, *dims)
        la = a.train_step(obs, act, rew, obs_n, done, step=2 * step, **kw(z))
        lb = b.train_step(obs, act, rew, obs_n, done, step=2 * step, **kw(z))
        for key in lb:
            losses_close(la[key], lb[key], "%s step %d %s" % (cls_name, step, key))
        a.run_updates()
        b.run_updates()
    assert len(a._graphs) == 1 and len(b._graphs) == 0
    for ca, cb in zip(a._cores, b._cores):
        d = (ca.params - cb.params).abs()
        assert float(d.max()) <= 6.6e-4 and float((d > 5e-6).float().mean()) <= 0.01
        assert torch.equal(ca.adam_t, cb.adam_t)
    if cls_name == "MaTd3":     # Philox target noise under replay: a fresh stream every time
        obs, act, rew, obs_n, done, z = make_batch(B, 950, *dims)
        seen = []
        for _ in range(4):
            a.train_step(obs, act, rew, obs_n, done, step=2)
            seen.append(a._buf[("a_next", B)].clone())
        assert len(a._graphs) == 2
        for i in range(3):
            assert not torch.equal(seen[i], seen[i + 1])


def test_save_load_round_trip_and_refusals(tmp_path):
    from maddpg_b200.algorithms import Coma, DictReplayBuffer, MaTd3
    a = MaTd3(*spaces(), seed=4)
    obs, act, rew, obs_n, done, z = make_batch(64, 9)
    a.train_step(obs, act, rew, obs_n, done, step=2, noise=z)
    a.save(tmp_path / "ckpt")
    b = MaTd3(*spaces(), seed=5)
    b.load(tmp_path / "ckpt")
    for ca, cb in zip(a._cores, b._cores):
        assert torch.equal(ca.params, cb.params) and torch.equal(ca.adam_m, cb.adam_m) and torch.equal(ca.adam_t, cb.adam_t)
    la = a.train_step(obs, act, rew, obs_n, done, step=4, noise=z)
    lb = b.train_step(obs, act, rew, obs_n, done, step=4, noise=z)
    for n in la["critic"]:
        assert float(la["critic"][n]) == pytest.approx(float(lb["critic"][n]), rel=1e-6)
    with pytest.raises(AssertionError):
        MaTd3(*spaces(), shared_policy=True)      # unequal spaces (policygroup.py:32-34)
    with pytest.raises(AssertionError):
        MaTd3(*spaces(), shared_critic=True)      # (criticgroup.py:28-30)
    with pytest.raises(AssertionError):
        Coma(*spaces())                           # ComaModule's global critic group is always shared
    # the fork's dict replay (common/replaybuffer.py): ring overwrite and dict-of-lists samples
    rb = DictReplayBuffer(5)
    for t in range(8):
        rb.add({"a": [t]}, {"a": [t]}, {"a": t}, {"a": [t + 1]}, {"a": False})
    assert len(rb) == 5 and sorted(x[0]["a"][0] for x in rb._storage) == [3, 4, 5, 6, 7]
    o, ac, r, o2, d = rb.sample(16)
    assert len(o["a"]) == 16 and all(3 <= v <= 7 for v in r["a"])
    assert len(rb.collect()[0]["a"]) == 5


def test_learn_generator_drives_a_dict_env():
    """multiagentalgbase.py:106-132 against a tiny dict env: train steps fire at step > 1024 and step % 5000 == 0."""
    from maddpg_b200.algorithms import MaTd3

    class Env(object):
        def __init__(self):
            self.rng, self.t = np.random.RandomState(0), 0

        def _obs(self):
            return {n: self.rng.randn(OBS[n]).astype(np.float32) for n in NAMES}

        def reset(self):
            self.t = 0
            return self._obs()

        def step(self, actions):
            assert set(actions) == set(NAMES) and all(np.shape(actions[n]) == (ACT[n],) or ACT[n] == 1 for n in NAMES)
            self.t += 1
            return self._obs(), -float(sum(np.sum(np.square(a)) for a in actions.values())), self.t % 25 == 0, {}

    alg = MaTd3(*spaces(), seed=6)
    fired = [info.step for info in alg.learn_generator(Env(), timesteps=5002) if info.critic_loss]
    assert fired == [5000]
    assert alg.critics[0].adam_t.cpu().tolist() == [0, 1] * 3
