"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group (env/replay shards are
rank-local, the only collective is the gradient all-reduce followed by a 1/world scale)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import NoiseTape, fill_oracle_replay, trainer_case


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from maddpg_b200.distributed import DataParallelUpdater, rank_seed, shard_range
    dp = DataParallelUpdater(core=None)
    assert dp.world == world and dp.rank == rank
    # replicas start from rank 0's parameters
    params = torch.full((10,), float(rank + 1))
    dp.broadcast_params(params)
    assert torch.all(params == 1.0)
    # data-parallel gradient of a mean loss == mean over ranks of the local mean-loss gradients:
    # each rank computes the oracle critic gradient on ITS rows, all-reduce sums, kernel scales by 1/world.
    case = trainer_case("simple_spread", seed=9)
    fill_oracle_replay(case)
    tr = case["trainers"][0]
    B = case["B"]
    idx_all = case["idx"][0]
    lo, hi = shard_range(B, rank, world)
    assert (lo, hi) == ((0, B // 2) if rank == 0 else (B // 2, B))

    def grads_on(idx):
        trainers = trainer_case("simple_spread", seed=9)["trainers"]
        c2 = trainer_case("simple_spread", seed=9)
        fill_oracle_replay(c2)
        t0 = c2["trainers"][0]
        obs_n, act_n = [], []
        for i in range(c2["n"]):
            o, a, r, n2, d = c2["trainers"][i].replay_buffer.sample_index(idx)
            obs_n.append(o), act_n.append(a)
        y = np.linspace(-1, 1, len(idx)).astype(np.float32)
        t0.q_train(obs_n, act_n, y)
        return np.concatenate([g.ravel() for g in t0.last_grads["q"]])

    y_all = np.linspace(-1, 1, B)
    # same targets for the same rows on both paths
    def grads_rows(rows, ys):
        c2 = trainer_case("simple_spread", seed=9)
        fill_oracle_replay(c2)
        t0 = c2["trainers"][0]
        obs_n, act_n = [], []
        for i in range(c2["n"]):
            o, a, r, n2, d = c2["trainers"][i].replay_buffer.sample_index(rows)
            obs_n.append(o), act_n.append(a)
        t0.q_train(obs_n, act_n, ys.astype(np.float32))
        return np.concatenate([g.ravel() for g in t0.last_grads["q"]])

    local = torch.from_numpy(grads_rows(idx_all[lo:hi], y_all[lo:hi]))
    bucket = local.clone()
    dp.allreduce_sum(bucket)
    assert dp.allreduce_bytes == bucket.numel() * 4
    reduced = bucket / world
    full = torch.from_numpy(grads_rows(idx_all, y_all))
    torch.testing.assert_close(reduced, full, rtol=2e-4, atol=1e-7)
    stats = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64)
    dp.reduce_stats(stats)
    assert stats.tolist() == [3.0, 4.0]
    assert rank_seed(7, rank) == 7 + rank
    ret[rank] = 1
    dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: 1, 1: 1}


def test_shard_range_partitions_exactly():
    from maddpg_b200.distributed import shard_range
    for total in (7, 4096, 262144):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
