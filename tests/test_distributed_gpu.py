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
