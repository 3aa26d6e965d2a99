"""oracle/maddpg.py cross-checked against torch autograd (float64) and the reference's own polyak
invariant (reference tests/test_policy.py:71-86, tests/test_critic.py:67-81)."""
import numpy as np
import torch

from oracle import maddpg as om
from tests.helpers import NoiseTape, fill_oracle_replay, trainer_case


def _t(params):
    return [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in params]


def _mlp(p, x):
    h = torch.relu(x @ p[0] + p[1])
    h = torch.relu(h @ p[2] + p[3])
    return h @ p[4] + p[5]


def _gs(logits, u, heads):
    z = logits - torch.log(-torch.log(u))
    outs, o = [], 0
    for h in heads:
        outs.append(torch.softmax(z[:, o:o + h], dim=1))
        o += h
    return torch.cat(outs, dim=1)


def _check_case(name):
    case = trainer_case(name, seed=3)
    n, B = case["n"], case["B"]
    fill_oracle_replay(case)
    trainers = case["trainers"]
    j = n - 1
    tr = trainers[j]
    idx = case["idx"][j]
    obs_n, act_n, nobs_n = [], [], []
    for i in range(n):
        o, a, r, n2, d = trainers[i].replay_buffer.sample_index(idx)
        obs_n.append(o), act_n.append(a), nobs_n.append(n2)
    _, _, rew, _, done = tr.replay_buffer.sample_index(idx)
    off = np.concatenate([[0], np.cumsum(case["act_dims"])]).astype(int)
    ut, ua = case["u_target"][j], case["u_actor"][j]
    # ---- torch float64 reference of the same graph, from the pre-update weights
    q, tq, p = _t(tr.q.p), _t(tr.target_q.p), _t(tr.p.p)
    tps = [_t(t.target_p.p) for t in trainers]
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    ta = [_gs(_mlp(tps[i], T(nobs_n[i])), T(ut[:, off[i]:off[i + 1]]), case["heads"][i]) for i in range(n)]
    if case["local_q"][j]:
        xq_next = torch.cat([T(nobs_n[j]), ta[j]], 1)
        xq = torch.cat([T(obs_n[j]), T(act_n[j])], 1)
    else:
        xq_next = torch.cat([T(x) for x in nobs_n] + ta, 1)
        xq = torch.cat([T(x) for x in obs_n] + [T(a) for a in act_n], 1)
    qn = _mlp(tq, xq_next)[:, 0]
    y = (T(rew) + 0.95 * (1 - T(done)) * qn).detach()
    q_loss = torch.mean((_mlp(q, xq)[:, 0] - y) ** 2)
    gq = torch.autograd.grad(q_loss, q)
    # ---- oracle step
    tape = NoiseTape()
    for t_ in trainers:
        t_.noise = tape
        t_.max_replay_buffer_len = 0
    for i in range(n):
        tape.push(ut[:, off[i]:off[i + 1]])
    tape.push(ua)
    stats = tr.update(trainers, 100, index=idx)
    assert np.allclose(tr.last_target_q, y.numpy(), rtol=2e-5, atol=2e-6)
    assert np.isclose(stats[0], q_loss.item(), rtol=1e-4)
    for g_o, g_t in zip(tr.last_grads["q"], gq):
        assert np.allclose(g_o, g_t.numpy(), rtol=2e-3, atol=2e-6)
    # actor loss goes through the critic AFTER its Adam step (maddpg.py:188 then :191)
    q2 = _t(tr.q.p)
    logits = _mlp(p, T(obs_n[j]))
    a_hat = _gs(logits, T(ua), case["heads"][j])
    acts = [T(a) for a in act_n]
    acts[j] = a_hat
    xq2 = torch.cat([T(obs_n[j]), a_hat], 1) if case["local_q"][j] else torch.cat([T(x) for x in obs_n] + acts, 1)
    p_loss = -torch.mean(_mlp(q2, xq2)[:, 0]) + 1e-3 * torch.mean(logits ** 2)
    gp = torch.autograd.grad(p_loss, p)
    assert np.isclose(stats[1], p_loss.item(), rtol=1e-4, atol=1e-6)
    for g_o, g_t in zip(tr.last_grads["p"], gp):
        assert np.allclose(g_o, g_t.numpy(), rtol=2e-3, atol=2e-6)


def test_oracle_grads_vs_torch_spread():
    _check_case("simple_spread")


def test_oracle_grads_vs_torch_world_comm_multihead():
    _check_case("simple_world_comm")


def test_oracle_grads_vs_torch_ddpg_local_q():
    _check_case("simple_tag_ddpg_adv")


def test_adam_is_tf_formulation():
    p = [np.array([1.0, -2.0], np.float32)]
    g = [np.array([0.5, -0.25], np.float32)]
    opt = om.Adam(p, lr=1e-2)
    opt.step(p, g)
    # t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1); eps added to sqrt(v) (uncorrected)
    lr_t = 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = np.array([1.0, -2.0]) - lr_t * (0.1 * g[0]) / (np.sqrt(0.001 * g[0] ** 2) + 1e-8)
    assert np.allclose(p[0], exp, rtol=1e-6)


def test_clip_by_norm_per_variable():
    g = np.array([3.0, 4.0], np.float32)
    assert np.allclose(om.clip_by_norm(g, 0.5), g * 0.1)
    small = np.array([0.1, 0.2], np.float32)
    assert np.allclose(om.clip_by_norm(small, 0.5), small)


def test_polyak_invariants():
    rng = np.random.RandomState(0)
    a, b = om.MLP(5, 8, 3, rng), om.MLP(5, 8, 3, rng)
    tgt = [x.copy() for x in b.p]
    om.polyak_update(b, a, polyak=0.0)   # reference tests/test_policy.py:71-86: target == running
    assert all(np.array_equal(x, y) for x, y in zip(b.p, a.p))
    b.p = [x.copy() for x in tgt]
    om.polyak_update(b, a, polyak=1.0)   # unchanged
    assert all(np.array_equal(x, y) for x, y in zip(b.p, tgt))
    om.polyak_update(b, a)               # 0.99 / 0.01 mix
    assert all(np.allclose(x, 0.99 * t + 0.01 * r, rtol=1e-6) for x, t, r in zip(b.p, tgt, a.p))


def test_gumbel_softmax_heads_sum_to_one():
    rng = np.random.RandomState(1)
    a = om.gumbel_softmax(rng.randn(7, 9), rng.uniform(0.01, 1, (7, 9)), [5, 4])
    assert np.allclose(a[:, :5].sum(1), 1, atol=1e-6) and np.allclose(a[:, 5:].sum(1), 1, atol=1e-6)


def test_update_orchestration_matches_the_reference_method():
    """tests/golden/update_orchestration_ref.npz was recorded by the REAL ``MADDPGAgentTrainer.update`` / ``experience`` /
    ``preupdate`` / ``action`` (maddpg/trainer/maddpg.py:151-196, executed unmodified around oracle-backed graph callables and the
    REAL ReplayBuffer; tests/golden/make_update_golden.py).  ``OracleAgentTrainer``'s own restatement of those methods -- the thing
    every GPU update-round test is compared with -- must reproduce it bit for bit: both gates, the index draws, the float64 TD
    combine, the call order and the six statistics."""
    import os
    import random
    from tests.update_case import N, T_SEQUENCE, build_oracle_trainers, transition
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "update_orchestration_ref.npz"))
    agents = build_oracle_trainers()
    random.seed(11)
    rows, ran = 0, 0
    for step, t in enumerate(T_SEQUENCE):
        for _ in range(40):
            tr = transition(rows)
            for i, a in enumerate(agents):
                a.experience(tr["obs"][i], tr["act"][i], tr["rew"][i], tr["obs2"][i], tr["done"][i], False)
            rows += 1
        for a in agents:
            a.preupdate()
        for i, a in enumerate(agents):
            res = a.update(agents, t)
            want = gold["s%d_a%d_stats" % (step, i)]
            if res is None:
                assert np.isnan(want).all(), (step, i)
            else:
                assert np.array_equal(np.asarray(res, np.float64), want), (step, i, res, want)
                ran += 1
            idx = np.asarray([] if a.replay_sample_index is None else a.replay_sample_index, np.int64)
            assert np.array_equal(idx, gold["s%d_a%d_index" % (step, i)]), (step, i)
        sums = np.asarray([float(np.sum([np.sum(p.astype(np.float64)) for net in (o.q, o.target_q, o.p, o.target_p) for p in net.p]))
                           for o in agents])
        assert np.array_equal(sums, gold["s%d_params" % step]), step
    assert ran == 2 * N      # gated at 40 and 80 rows (warm-up) and at t = 150 (period); two rounds ran
    obs = transition(999)["obs"]
    act = np.concatenate([np.asarray(a.action(obs[i]), np.float64) for i, a in enumerate(agents)])
    assert np.array_equal(act, gold["action"])


def test_oracle_matches_the_reference_graph_code():
    """tests/golden/graph_ref.npz: the reference's OWN graph-building code (``MADDPGAgentTrainer.__init__``, ``q_train``,
    ``p_train``, ``make_update_exp``, ``SoftCategoricalPd``, ``U.function`` / ``scope_vars`` / ``minimize_and_clip``, ``mlp_model``)
    executed unmodified on a torch-backed stand-in for TensorFlow (tests/tf_shim.py, tests/golden/make_graph_golden.py), driven
    through the real ``update``.  The restated trainer must give the same debug surfaces, statistics and variables -- i.e. the
    same graph wiring: centralized vs local critic inputs, loss expressions, per-optimizer variable sets, clip placement, polyak
    pairing.  (float32 sums in a different order: statistics agree to 8e-7 relative, variables to 6e-8.)"""
    import os
    import random
    from tests.update_case import N, build_oracle_trainers, shared_noise, transition
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_ref.npz"))
    agents = build_oracle_trainers(noise=shared_noise())
    random.seed(11)
    for k in range(120):
        tr = transition(k)
        for i, a in enumerate(agents):
            a.experience(tr["obs"][i], tr["act"][i], tr["rew"][i], tr["obs2"][i], tr["done"][i], False)
    batch = [transition(500 + k) for k in range(10)]
    obs_n = [np.asarray([b["obs"][i] for b in batch]) for i in range(N)]
    act_n = [np.asarray([b["act"][i] for b in batch]) for i in range(N)]
    for i, a in enumerate(agents):
        for key, got in (("p_values", a.p_debug["p_values"](obs_n[i])), ("q_values", a.q_debug["q_values"](*(obs_n + act_n))),
                         ("target_q_values", a.q_debug["target_q_values"](*(obs_n + act_n))), ("act", a.act(obs_n[i])),
                         ("target_act", a.p_debug["target_act"](obs_n[i]))):
            np.testing.assert_allclose(np.asarray(got, np.float64), gold["a%d_%s" % (i, key)], rtol=2e-5, atol=2e-6,
                                       err_msg="agent %d %s" % (i, key))
    for rnd, t in enumerate((100, 200)):
        for a in agents:
            a.preupdate()
        for i, a in enumerate(agents):
            stats = np.asarray(a.update(agents, t), np.float64)
            np.testing.assert_allclose(stats, gold["r%d_a%d_stats" % (rnd, i)], rtol=1e-5, atol=1e-6, err_msg="round %d agent %d" % (rnd, i))
        for i, a in enumerate(agents):
            for attr in ("q", "target_q", "p", "target_p"):
                for k, w in enumerate(getattr(a, attr).p):
                    d = np.abs(w - gold["r%d_a%d_%s_%d" % (rnd, i, attr, k)])
                    assert d.max() <= 2e-6, (rnd, i, attr, k, d.max())      # measured 6e-8: 1e-4 of one Adam step (lr = 1e-2)
    names = [str(x) for x in gold["variable_names"]]
    assert names[:6] == ["agent_0/q_func/fully_connected/weights:0", "agent_0/q_func/fully_connected/biases:0",
                         "agent_0/q_func/fully_connected_1/weights:0", "agent_0/q_func/fully_connected_1/biases:0",
                         "agent_0/q_func/fully_connected_2/weights:0", "agent_0/q_func/fully_connected_2/biases:0"]
    assert len(names) == N * 24      # q_func, target_q_func, p_func, target_p_func per agent: no second critic from reuse=True


def test_multi_head_sample_matches_the_reference_class():
    """``SoftMultiCategoricalPd.sample`` (distributions.py:305-336: per-head Gumbel-softmax, ``low`` added, concatenated), the
    class behind simple_world_comm's MultiDiscrete leader -- executed on tests/tf_shim.py with one uniform block per head.  The
    fork's ``make_pdtype`` no longer reaches it (:416-418 commented out: NotImplementedError for MultiDiscrete, recorded)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_ref.npz"))
    got = om.gumbel_softmax(gold["multi_logits"], gold["multi_u"], [5, 4])
    np.testing.assert_allclose(got, gold["multi_sample"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(got[:, :5].sum(axis=1), 1.0, atol=1e-6)
    np.testing.assert_allclose(got[:, 5:].sum(axis=1), 1.0, atol=1e-6)
    assert str(gold["make_pdtype_multidiscrete"]) == "NotImplementedError"
