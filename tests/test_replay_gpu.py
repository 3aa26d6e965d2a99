"""CUDA replay ring against the reference's golden vectors and a numpy model: bit-exact."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "replay_ref.npz"))


@pytest.mark.parametrize("mode", [0, 1])
def test_reference_golden_bit_exact(mode):
    """Same insert stream + same python ``random`` index stream as the REAL reference buffer."""
    from maddpg_b200 import DeviceReplayBuffer, JointReplayRing
    ring = JointReplayRing([18], [5], capacity=int(G["cap"]), gather_mode=mode)
    rb = DeviceReplayBuffer(ring, 0)
    random.seed(1234)
    N = G["in_obs"].shape[0]
    for t in range(N):
        rb.add(G["in_obs"][t], G["in_act"][t], float(G["in_rew"][t]), G["in_nobs"][t], float(G["in_done"][t]))
        assert len(rb) == G["lens"][t] and rb._next_idx == G["nexts"][t]
        if "idx_%d" % t in G.files:
            idx = rb.make_index(int(G["batch"]))
            assert np.array_equal(np.asarray(idx), G["idx_%d" % t])
            o, a, r, n2, d = rb.sample_index(idx)
            for got, key in ((o, "obs"), (a, "act"), (r, "rew"), (n2, "nobs"), (d, "done")):
                ref = G["%s_%d" % (key, t)].astype(np.float32)
                assert got.shape == ref.shape and got.dtype == np.float32
                assert np.array_equal(got, ref), key
    np.random.seed(5)
    assert np.array_equal(np.asarray(rb.make_latest_index(16)), G["latest_idx"])
    o, a, r, n2, d = rb.collect()
    assert np.array_equal(o, G["collect_obs"].astype(np.float32)) and np.array_equal(r, G["collect_rew"].astype(np.float32))
    rb.clear()
    assert len(rb) == 0 and rb._next_idx == 0


def _model_insert(model, lay, cur, obs, act, rew, nobs, done):
    E = obs.shape[0]
    cap = model.shape[0]
    for e in range(E):
        r = (cur + e) % cap
        model[r, :lay.obs_sum] = obs[e, :lay.obs_sum]
        model[r, lay.obs_sum:lay.x_dim] = act[e, :lay.act_sum]
        model[r, lay.nx_off:lay.nx_off + lay.obs_sum] = nobs[e, :lay.obs_sum]
        model[r, lay.rw_off:lay.rw_off + lay.n_agents] = rew[e]
        model[r, lay.dn_off:lay.dn_off + lay.n_agents] = done[e]


@pytest.mark.parametrize("dims", [([18, 18, 18], [5, 5, 5]), ([16, 16, 16, 14], [5, 5, 5, 5]),
                                  ([34] * 4 + [28] * 2, [9, 5, 5, 5, 5, 5]), ([144] * 24, [5] * 24), ([4], [5])])
def test_joint_insert_wrap_and_gather_modes(dims):
    from maddpg_b200 import JointReplayRing
    obs_dims, act_dims = dims
    n = len(obs_dims)
    rng = np.random.RandomState(len(obs_dims))
    cap, E = 301, 64
    ring = JointReplayRing(obs_dims, act_dims, capacity=cap)
    ring.ring.fill_(-7.0)
    lay = ring.layout
    OS, AS = (lay.obs_sum + 3) // 4 * 4, (lay.act_sum + 3) // 4 * 4
    model = np.full((cap, ring.row_stride), -7.0, np.float32)
    for step in range(7):  # 448 rows into 301
        obs, nobs = rng.randn(E, OS).astype(np.float32), rng.randn(E, OS).astype(np.float32)
        act = rng.rand(E, AS).astype(np.float32)
        rew = rng.randn(E, n).astype(np.float32)
        done = (rng.rand(E, n) < 0.2).astype(np.uint8)
        cur = ring.next_idx[0]
        ring.insert_joint(*(torch.from_numpy(x).cuda() for x in (obs, act, rew, nobs, done)))
        _model_insert(model, lay, cur, obs, act, rew, nobs, done.astype(np.float32))
        assert ring.length[0] == min(cap, (step + 1) * E)
    used = np.ones(ring.row_stride, bool)
    used[lay.x_dim:lay.nx_off] = False
    used[lay.nx_off + lay.obs_sum:lay.rw_off] = False
    used[lay.dn_off + n:] = False
    got = ring.ring.cpu().numpy()
    assert np.array_equal(got[:, used], model[:, used])
    assert np.all(got[:, ~used] == -7.0)  # padding columns are never written
    idx = torch.from_numpy(rng.randint(0, cap, size=777)).cuda()
    g0 = ring.gather(idx, mode=0).cpu().numpy()
    g1 = ring.gather(idx, mode=1).cpu().numpy()
    assert np.array_equal(g0, got[idx.cpu().numpy()])
    assert np.array_equal(g1, g0)


def test_full_size_gather_checksum():
    """BASELINE config 2 sizes (B=1024 out of 1e5 rows): gather is a permutation-with-replacement of
    rows, so per-row checksums must match a torch index_select of the same ring."""
    from maddpg_b200 import JointReplayRing
    ring = JointReplayRing([18] * 3, [5] * 3, capacity=100000)
    ring.ring.copy_(torch.randn(ring.ring.shape, device="cuda"))
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 100000, (1024,), generator=g).cuda()
    for mode in (0, 1):
        out = ring.gather(idx, mode=mode)
        assert torch.equal(out, ring.ring.index_select(0, idx))


def test_per_agent_device_add_keeps_columns_separate():
    from maddpg_b200 import DeviceReplayBuffer, JointReplayRing
    ring = JointReplayRing([6, 4], [5, 5], capacity=50)
    ring.ring.zero_()
    rbs = [DeviceReplayBuffer(ring, i, numpy_io=False) for i in range(2)]
    E = 8
    data = []
    for i, rb in enumerate(rbs):
        D = ring.obs_dims[i]
        o, a = torch.randn(E, D).cuda(), torch.rand(E, 5).cuda()
        r, n2 = torch.randn(E).cuda(), torch.randn(E, D).cuda()
        d = (torch.rand(E) < 0.5).to(torch.uint8).cuda()
        rb.add(o, a, r, n2, d)
        data.append((o, a, r, n2, d))
    assert ring.aligned() and len(rbs[0]) == 8
    for i, rb in enumerate(rbs):
        o, a, r, n2, d = rb.sample_index(list(range(8)))
        for got, ref in zip((o, a, r, n2), data[i][:4]):
            assert torch.equal(got, ref)
        assert torch.equal(d, data[i][4].float())
