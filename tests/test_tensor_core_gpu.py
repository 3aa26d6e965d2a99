"""tcgen05 (tensor-core) kernels against the numpy oracle and against the fp32 SIMT kernels on identical inputs.
(Gradient comparisons between the two GPU paths allow 1e-4 of the largest gradient entry: both reduce with fp32 RED
atomics in a run-dependent order, and entries that are sums of cancelling terms carry that noise at full size.)
The tensor path issues every GEMM as three kind::tf32 MMAs (hi/lo split), so it must hold the same 1e-4 bar
on TD targets / Q values as the SIMT path (BASELINE.json north_star), and agree with it to ~1e-5."""
import numpy as np
import pytest
import torch

from tests.helpers import TRAINER_CASES, oracle_update_round, trainer_case
from tests.test_trainer_gpu import _build, _close

pytestmark = pytest.mark.gpu
TC_CASES = [n for n, c in TRAINER_CASES.items() if c[2] == 64]


def _ut(core, case, j, B):
    ut = torch.zeros((B, core.act_stride), device="cuda")
    ut[:, :core.act_sum] = torch.from_numpy(case["u_target"][j]).cuda()
    return ut


@pytest.mark.parametrize("name", TC_CASES)
def test_td_target_tensor_cores_match_oracle_and_simt(name):
    case = trainer_case(name, seed=2)
    ref = oracle_update_round(trainer_case(name, seed=2))
    trainers, core = _build(case)
    B = case["B"]
    for j in range(case["n"] if name != "simple_spread_6" else 2):
        if j > 0:
            break  # the oracle round steps agent 0 first; later agents see updated targets
        idx = core.ring.index_tensor(case["idx"][j])
        batch = core.ring.gather(idx)
        ut = _ut(core, case, j, B)
        core.set_tensor_cores(-1)
        y_simt, ta_simt = core.td_target(j, batch, ut, want_target_act=True)
        y_simt, ta_simt = y_simt.clone(), ta_simt.clone()
        st_simt = core.stats.clone()
        core.set_tensor_cores(1)
        y_tc, ta_tc = core.td_target(j, batch, ut, want_target_act=True)
        st_tc = core.stats.clone()
        _close(y_tc.cpu().numpy(), ref[j]["y"], atol=2e-6, msg="td target vs oracle")
        _close(y_tc.cpu().numpy(), y_simt.cpu().numpy(), rtol=2e-5, atol=2e-6, msg="td target vs SIMT")
        _close(ta_tc.cpu().numpy(), ta_simt.cpu().numpy(), rtol=2e-5, atol=1e-6, msg="target actions vs SIMT")
        _close(st_tc[8 * j:8 * j + 8].cpu().numpy(), st_simt[8 * j:8 * j + 8].cpu().numpy(), rtol=1e-5, atol=2e-6 * B, msg="stats")  # sums over B rows, each within 2e-6
        # fused gather: rows addressed through the index set straight from the ring
        y_idx = core.td_target(j, core.ring.ring, ut, idx=idx)
        assert torch.equal(y_idx, y_tc)


@pytest.mark.parametrize("name", ["simple_spread", "simple_tag_ddpg_adv", "simple_spread_6"])
def test_td_target_tensor_cores_many_tiles_philox(name):
    """Several 128-row tiles with a ragged tail, in-kernel Philox noise (same stream as the SIMT kernel)."""
    case = trainer_case(name, seed=3)
    trainers, core = _build(case)
    rows = case["rows"]
    idx = torch.arange(rows, device="cuda", dtype=torch.int64).flip(0).contiguous()
    for j in range(case["n"]):
        c0 = core.counter
        core.set_tensor_cores(-1)
        y_simt = core.td_target(j, core.ring.ring, idx=idx).clone()
        core.counter = c0  # same Philox counter for the second launch
        core.set_tensor_cores(1)
        y_tc = core.td_target(j, core.ring.ring, idx=idx)
        _close(y_tc.cpu().numpy(), y_simt.cpu().numpy(), rtol=2e-5, atol=1e-5, msg="agent %d" % j)


@pytest.mark.parametrize("name", ["simple_spread", "simple_tag", "simple_spread_b1024", "simple_tag_b4096"])
def test_sequential_update_round_with_tensor_cores(name):
    case = trainer_case(name, seed=4)
    ref = oracle_update_round(trainer_case(name, seed=4))
    trainers, core = _build(case)
    core.set_tensor_cores(1)
    for j, tr in enumerate(trainers):
        tr.preupdate()
        tr.inject_noise(u_target=case["u_target"][j], u_actor=case["u_actor"][j])
        stats = tr.update(trainers, 100, index=case["idx"][j])
        for k in range(6):
            _close(stats[k], ref[j]["stats"][k], rtol=1e-4, atol=2e-6, msg="stat %d agent %d" % (k, j))


def test_grouped_td_target_tensor_cores():
    """grid.y = agent (mdp_update_all's launch shape): per-agent index sets, outputs equal the per-agent launches."""
    case = trainer_case("simple_spread", seed=5)
    trainers, core = _build(case)
    B, n = case["B"], case["n"]
    idx = torch.stack([core.ring.index_tensor(case["idx"][j]) for j in range(n)])
    core.set_tensor_cores(1)
    c0 = core.counter
    ys = []
    for j in range(n):
        core.counter = c0
        ys.append(core.td_target(j, core.ring.ring, idx=idx[j]).clone())
    core.counter = c0
    y_all = core.td_target_all(core.ring.ring, idx=idx).clone()
    for j in range(n):
        torch.testing.assert_close(y_all[j], ys[j], rtol=0, atol=0)
    core.counter = c0
    core.update_all(core.ring.ring, idx=idx)  # the Jacobi round starts from the same grouped launch
    torch.testing.assert_close(core._y[("all", B)], y_all, rtol=0, atol=0)


@pytest.mark.parametrize("name", [n for n in TC_CASES if "ddpg" not in n])
def test_critic_grads_tensor_cores_match_oracle_and_simt(name):
    """q_train forward/backward on tcgen05 (5 GEMMs, 3xTF32) vs the oracle's gradients and vs the SIMT kernel."""
    case = trainer_case(name, seed=2)
    ref = oracle_update_round(trainer_case(name, seed=2))
    trainers, core = _build(case)
    j, B = 0, case["B"]
    idx = core.ring.index_tensor(case["idx"][j])
    ut = _ut(core, case, j, B)
    core.set_tensor_cores(-1)
    y = core.td_target(j, core.ring.ring, ut, idx=idx).clone()
    q_simt = core.critic_grads(j, core.ring.ring, y, want_q=True, idx=idx).clone()
    g_simt = [g.clone() for g in core.train_view(core.grads, j, 1)]
    st_simt = core.stats[8 * j].clone()
    core.grads.zero_()
    core.adam_t.zero_()
    core.set_tensor_cores(1)
    q_tc = core.critic_grads(j, core.ring.ring, y, want_q=True, idx=idx)
    g_tc = [g.clone() for g in core.train_view(core.grads, j, 1)]
    assert core.adam_t.cpu().tolist()[1] == 1
    _close(q_tc.cpu().numpy(), q_simt.cpu().numpy(), rtol=2e-5, atol=1e-5, msg="q")
    # the loss accumulator is only cleared by the TD-target launch: the second critic_grads call added its own sum
    _close((core.stats[8 * j] - st_simt).cpu().numpy(), st_simt.cpu().numpy(), rtol=1e-4, atol=1e-7, msg="loss sum")
    names = ["W1", "b1", "W2", "b2", "W3", "b3"]
    for k, (a, b, r) in enumerate(zip(g_tc, g_simt, ref[j]["q_grads"])):
        scale = float(np.abs(r).max())
        _close(a.cpu().numpy(), r, rtol=1e-3, atol=1e-6 + 1e-4 * scale, msg="critic grad %s vs oracle" % names[k])
        _close(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-3, atol=1e-7 + 1e-4 * scale, msg="critic grad %s vs SIMT" % names[k])


def test_critic_grads_tensor_cores_many_tiles_grouped():
    """Several 128-row tiles with a ragged tail, all agents in one launch (grid.y = agent), accumulated gradients."""
    case = trainer_case("simple_spread_6", seed=6)
    trainers, core = _build(case)
    rows, n = case["rows"], case["n"]
    gen = torch.Generator(device="cuda").manual_seed(13)
    idx = torch.stack([torch.randperm(rows, device="cuda", generator=gen)[:rows - 3] for _ in range(n)]).contiguous()
    B = idx.shape[1]
    core.set_tensor_cores(-1)
    c0 = core.counter
    y = core.td_target_all(core.ring.ring, idx=idx).clone()
    outs = []
    for mode in (-1, 1):
        core.grads.zero_()
        core.set_tensor_cores(mode)
        for j in range(n):
            core.critic_grads(j, core.ring.ring, y[j], idx=idx[j])
        outs.append(core.grads.clone())
    scale = float(outs[0].abs().max())
    _close(outs[1].cpu().numpy(), outs[0].cpu().numpy(), rtol=1e-3, atol=1e-4 * scale, msg="grouped critic grads")


@pytest.mark.parametrize("name", [n for n in TC_CASES if "ddpg" not in n])
def test_actor_grads_tensor_cores_match_oracle_and_simt(name):
    """p_train on tcgen05 (actor forward, Gumbel sample, running-critic forward/backward to the action columns, actor
    backward) after the critic's Adam step, vs the oracle's actor gradients and vs the SIMT kernel."""
    case = trainer_case(name, seed=2)
    ref = oracle_update_round(trainer_case(name, seed=2))
    trainers, core = _build(case)
    j, B = 0, case["B"]
    idx = core.ring.index_tensor(case["idx"][j])
    core.set_tensor_cores(-1)
    y = core.td_target(j, core.ring.ring, _ut(core, case, j, B), idx=idx)
    core.critic_grads(j, core.ring.ring, y, idx=idx)
    core.clip_adam_polyak(j, 1)
    ua = torch.zeros((B, core.act_stride), device="cuda")
    o = core.act_off[j]
    ua[:, o:o + core.act_dims[j]] = torch.from_numpy(case["u_actor"][j]).cuda()
    outs, stats = [], []
    for mode in (-1, 1):
        core.grads.zero_()
        s0 = core.stats[8 * j:8 * j + 3].clone()
        core.set_tensor_cores(mode)
        core.actor_grads(j, core.ring.ring, ua, idx=idx)
        outs.append([g.clone() for g in core.train_view(core.grads, j, 0)])
        stats.append((core.stats[8 * j:8 * j + 3] - s0).cpu().numpy())
    _close(stats[1][1:], stats[0][1:], rtol=1e-4, atol=1e-6, msg="sum(-q), sum(logits^2)")
    names = ["W1", "b1", "W2", "b2", "W3", "b3"]
    for k, (a, b, r) in enumerate(zip(outs[1], outs[0], ref[j]["p_grads"])):
        scale = float(np.abs(r).max())
        _close(a.cpu().numpy(), r, rtol=2e-3, atol=1e-7 + 2e-4 * scale, msg="actor grad %s vs oracle" % names[k])
        # both paths reduce the per-tile partials with fp32 atomics (order varies run to run) and 3xTF32 drops the lo*lo terms:
        # elements 30x below the largest one agree to 5e-4 of that largest one
        _close(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-3, atol=1e-7 + 5e-4 * scale, msg="actor grad %s vs SIMT" % names[k])


def test_actor_grads_tensor_cores_many_tiles_philox():
    """Several tiles with a ragged tail, in-kernel Philox noise, every agent of simple_spread N=6 and simple_tag."""
    for name in ("simple_spread_6", "simple_tag"):
        case = trainer_case(name, seed=8)
        trainers, core = _build(case)
        rows, n = case["rows"], case["n"]
        idx = torch.randperm(rows, device="cuda", generator=torch.Generator(device="cuda").manual_seed(14))[:rows - 5].contiguous()
        for j in range(n):
            outs = []
            for mode in (-1, 1):
                core.grads.zero_()
                core.counter = 50
                core.set_tensor_cores(mode)
                core.actor_grads(j, core.ring.ring, idx=idx)
                outs.append(core.grads.clone())
            scale = float(outs[0].abs().max())
            _close(outs[1].cpu().numpy(), outs[0].cpu().numpy(), rtol=1e-3, atol=2e-4 * scale, msg="%s agent %d" % (name, j))


@pytest.mark.parametrize("scenario,na,E", [("simple_tag", None, 4096 + 37), ("simple_spread", 24, 300), ("simple_spread", 3, 129),
                                           ("simple", None, 128)])
def test_actor_act_tensor_cores_match_simt_and_oracle(scenario, na, E):
    """mdp_actor_act: the tcgen05 kernel (128-row tiles, 3xTF32) against the fp32 SIMT kernel on the same Philox counter and
    against the oracle's numpy actor (logits) -- MADDPGAgentTrainer.action, maddpg.py:151-152."""
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore, _lib
    from oracle import maddpg as omaddpg
    env = BatchedMultiAgentEnv(scenario, num_envs=E, num_agents=na, squeeze=False, seed=3)
    core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=64, replay_capacity=64, seed=9)
    rng = np.random.RandomState(1)
    for i in range(env.n):
        w = core.get_weights(i, _lib.NET_P)
        for k in (1, 3, 5):
            w[k] = rng.uniform(-0.2, 0.2, size=w[k].shape).astype(np.float32)
        core.set_weights(i, _lib.NET_P, w)
    obs = torch.zeros((E, core.obs_stride), device="cuda")
    obs[:, :core.obs_sum] = torch.randn((E, core.obs_sum), generator=torch.Generator().manual_seed(0)).cuda()
    out = {}
    for mode in (-1, 1):
        core.set_tensor_cores(mode)
        act, lg = torch.zeros((E, core.act_stride), device="cuda"), torch.zeros((E, core.act_stride), device="cuda")
        core.act(obs, act, logits_out=lg, counter=5)
        out[mode] = (act.cpu().numpy(), lg.cpu().numpy())
    np.testing.assert_allclose(out[1][1], out[-1][1], rtol=2e-5, atol=1e-5, err_msg="logits tcgen05 vs SIMT")
    np.testing.assert_allclose(out[1][0], out[-1][0], rtol=0, atol=3e-6, err_msg="actions tcgen05 vs SIMT (same Philox draws)")
    for i in range(env.n):
        m = omaddpg.MLP(env.obs_dims[i], 64, env.act_dims[i], np.random.RandomState(0))
        m.p = core.get_weights(i, _lib.NET_P)
        o, K = core.obs_off[i], env.act_dims[i]
        ref, _ = m.forward(obs[:, o:o + env.obs_dims[i]].cpu().numpy())
        np.testing.assert_allclose(out[1][1][:, core.act_off[i]:core.act_off[i] + K], ref, rtol=1e-4, atol=1e-5)
    # injected uniforms (parity hook) take the same path
    u = torch.rand((E, core.act_stride), generator=torch.Generator().manual_seed(2)).clamp_(1e-6, 1 - 1e-6).cuda()
    acts = []
    for mode in (-1, 1):
        core.set_tensor_cores(mode)
        a = torch.zeros((E, core.act_stride), device="cuda")
        core.act(obs, a, u=u)
        acts.append(a.cpu().numpy())
    np.testing.assert_allclose(acts[1], acts[0], rtol=0, atol=3e-6)
