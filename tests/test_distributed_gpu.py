"""Two-GPU data-parallel update (NCCL): ranks hold different replay shards, all-reduce the gradient bucket
of the network being stepped, and must end with identical parameters that match a single-process update on
the union batch.  Skipped unless two CUDA devices are visible."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _build_core(dev, B, n=3):
    """n = 3: simple_spread N=3 shapes; n = 10: simple_spread N=10 shapes, whose critic W1 (650 x 64 = 41.6 k floats) takes the
    many-CTA path of the optimizer (mdp_optim.cu: wide_w1_step)."""
    from maddpg_b200 import MADDPGCore
    from maddpg_b200.spaces import Discrete
    core = MADDPGCore([6 * n] * n, [Discrete(5)] * n, [False] * n, device=dev, replay_capacity=4 * B, seed=5)
    return core


def _rows(core, B, seed):
    g = torch.Generator().manual_seed(seed)
    L, n = core.ring.layout, core.n
    batch = torch.randn(B, core.ring.row_stride, generator=g)
    act = torch.softmax(torch.randn(B, n, 5, generator=g), -1).reshape(B, 5 * n)
    batch[:, L.obs_sum:L.x_dim] = act
    batch[:, L.dn_off:L.dn_off + n] = (torch.rand(B, n, generator=g) < 0.1).float()
    return batch


def _worker(rank, world, port, B, out, peer=False, low_latency=None, n=3):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from maddpg_b200.distributed import DataParallelUpdater
    core = _build_core(dev, B, n)
    dp = DataParallelUpdater(core, peer=peer, low_latency=low_latency)
    dp.broadcast_params(core.params)
    full = _rows(core, world * B, 1)
    ut = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(2)).clamp_(1e-6, 1 - 1e-6)
    ua = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(3)).clamp_(1e-6, 1 - 1e-6)
    sl = slice(rank * B, (rank + 1) * B)
    for rnd in range(2 if peer else 1):  # two rounds: the flag epochs and the re-zeroed buckets are exercised
        for j in range(n):
            dp.update_agent(j, full[sl].contiguous().to(dev), ut[sl].contiguous().to(dev), ua[sl].contiguous().to(dev))
    torch.cuda.synchronize()
    out[rank] = core.params.cpu().numpy()
    if peer:
        assert float(core.grads.abs().max()) == 0.0
        dist.barrier()
        dp.peer.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_update_matches_union_batch():
    B, world = 256, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29400 + os.getpid() % 500, B, out), nprocs=world, join=True)
    p0, p1 = out[0], out[1]
    assert np.array_equal(p0, p1), "replicas diverged"
    # single process on the union batch: mean over 2B rows == average of the two rank-local means
    core = _build_core(torch.device("cuda", 0), B)
    full = _rows(core, world * B, 1).cuda()
    ut = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(2)).clamp_(1e-6, 1 - 1e-6).cuda()
    ua = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(3)).clamp_(1e-6, 1 - 1e-6).cuda()
    for j in range(3):
        core.update_agent(j, full, ut, ua)
    ref = core.params.cpu().numpy()
    np.testing.assert_allclose(p0, ref, rtol=2e-3, atol=3e-4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("low_latency", [True, False])
def test_two_rank_peer_exchange_matches_union_batch(low_latency):
    """Fused peer-memory all-reduce inside the clip+Adam+polyak kernel, both protocols (low-latency push of (value, epoch)
    words / flag barriers + peer loads): identical replicas after two rounds (epochs, re-zeroed buckets), matching a
    single-process update on the union batch."""
    B, world = 256, 2
    mgr = mp.Manager()
    out_p, out_n = mgr.dict(), mgr.dict()
    mp.spawn(_worker, args=(world, 29900 + os.getpid() % 500 + int(low_latency), B, out_p, True, low_latency), nprocs=world, join=True)
    assert np.array_equal(out_p[0], out_p[1]), "replicas diverged (peer exchange)"
    # reference: a single process on the union batch (mean over 2B rows == average of the two rank-local means)
    core = _build_core(torch.device("cuda", 0), B)
    full = _rows(core, world * B, 1).cuda()
    ut = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(2)).clamp_(1e-6, 1 - 1e-6).cuda()
    ua = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(3)).clamp_(1e-6, 1 - 1e-6).cuda()
    for rnd in range(2):
        for j in range(3):
            core.update_agent(j, full, ut, ua)
    np.testing.assert_allclose(out_p[0], core.params.cpu().numpy(), rtol=3e-3, atol=5e-4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_peer_exchange_wide_critic_many_cta_path():
    """Barrier protocol with a critic W1 wide enough for the many-CTA optimizer path (k_w1_sqnorm sums the ranks' buckets with
    peer loads into a local scratch, k_w1_adam steps from it): identical replicas after two rounds, matching the union batch."""
    B, world, n = 128, 2, 10
    mgr = mp.Manager()
    out_p = mgr.dict()
    mp.spawn(_worker, args=(world, 28700 + os.getpid() % 500, B, out_p, True, False, n), nprocs=world, join=True)
    assert np.array_equal(out_p[0], out_p[1]), "replicas diverged (peer exchange, wide W1)"
    core = _build_core(torch.device("cuda", 0), B, n)
    full = _rows(core, world * B, 1).cuda()
    ut = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(2)).clamp_(1e-6, 1 - 1e-6).cuda()
    ua = torch.rand(world * B, core.act_stride, generator=torch.Generator().manual_seed(3)).clamp_(1e-6, 1 - 1e-6).cuda()
    for rnd in range(2):
        for j in range(n):
            core.update_agent(j, full, ut, ua)
    np.testing.assert_allclose(out_p[0], core.params.cpu().numpy(), rtol=3e-3, atol=5e-4)


def _oracle_worker(rank, world, port, out):
    """One sequential update round of the seeded simple_spread trainer case, the batch split over the ranks, fused peer exchange."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from maddpg_b200 import MADDPGCore, _lib
    from maddpg_b200.distributed import DataParallelUpdater
    from tests.helpers import trainer_case
    case = trainer_case("simple_spread", seed=4)
    n, B = case["n"], case["B"]
    core = MADDPGCore(case["obs_dims"], list(case["env"].action_space), [False] * n, device=dev, replay_capacity=4 * B, seed=5)
    for i, o in enumerate(case["trainers"]):
        for net, m in ((_lib.NET_P, o.p), (_lib.NET_TARGET_P, o.target_p), (_lib.NET_Q, o.q), (_lib.NET_TARGET_Q, o.target_q)):
            core.set_weights(i, net, m.p)
    dp = DataParallelUpdater(core, peer=True, low_latency=True)
    L, p = core.ring.layout, case["pool"]
    half = B // world
    sl = slice(rank * half, (rank + 1) * half)
    for j in range(n):  # agent j's batch = the pool rows at ITS index set, like the oracle's tr.update(index=idx[j])
        idx = np.asarray(case["idx"][j])
        rows = torch.zeros(B, core.ring.row_stride)
        for i in range(n):
            c = core.ring.cols(i)
            rows[:, c["obs"][0]:c["obs"][1]] = torch.from_numpy(p["obs"][i][idx])
            rows[:, c["act"][0]:c["act"][1]] = torch.from_numpy(p["act"][i][idx])
            rows[:, c["next_obs"][0]:c["next_obs"][1]] = torch.from_numpy(p["nobs"][i][idx])
            rows[:, c["rew"]] = torch.from_numpy(p["rew"][i][idx])
            rows[:, c["done"]] = torch.from_numpy(p["done"][i][idx])
        ut = torch.zeros(B, core.act_stride)
        ut[:, :core.act_sum] = torch.from_numpy(case["u_target"][j])
        ua = torch.zeros(B, core.act_stride)
        ua[:, core.act_off[j]:core.act_off[j] + core.act_dims[j]] = torch.from_numpy(case["u_actor"][j])
        dp.update_agent(j, rows[sl].contiguous().to(dev), ut[sl].contiguous().to(dev), ua[sl].contiguous().to(dev))
    torch.cuda.synchronize()
    out[rank] = [[w.copy() for w in core.get_weights(j, net)] for j in range(n) for net in (_lib.NET_Q, _lib.NET_P, _lib.NET_TARGET_Q, _lib.NET_TARGET_P)]
    dist.barrier()
    dp.peer.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_peer_exchange_matches_oracle_on_the_union_batch():
    """The data-parallel update against the ORACLE (not against the same kernels on one GPU): every rank holds half of each
    agent's batch; after one sequential round with the fused peer exchange the running and target nets of every agent equal the
    numpy oracle's update on the whole batch (mean over B rows == average of the two half-batch means) at the single-GPU test's
    tolerance."""
    from tests.helpers import oracle_update_round, trainer_case
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_oracle_worker, args=(2, 28100 + os.getpid() % 500, out), nprocs=2, join=True)
    ref = oracle_update_round(trainer_case("simple_spread", seed=4))
    k = 0
    for j in range(len(ref)):
        for key in ("q", "p", "target_q", "target_p"):
            for a, b, r in zip(out[0][k], out[1][k], ref[j][key]):
                assert np.array_equal(a, b), "replicas diverged"
                np.testing.assert_allclose(a, r, rtol=1e-3, atol=2e-4, err_msg="agent %d %s" % (j, key))
            k += 1
