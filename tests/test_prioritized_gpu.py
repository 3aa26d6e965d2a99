"""Device prioritized replay (maddpg_b200.DevicePrioritizedReplayMemory -> C ABI mdp_sumtree_*) against
(1) the outputs of the REAL reference classes (tests/golden/prioritized_ref.npz: every tree array, tree index, stored row and
beta bit for bit, IS weights to 1e-13 -- CUDA's pow -- and the reference's own IndexError samples), and
(2) oracle/prioritized.py (itself pinned to those goldens) at sizes where the multi-launch flush, the wrap-around and batches
of 1024 / 4096 updates with duplicate leaves are exercised."""
import os

import numpy as np
import pytest
import torch

from oracle.prioritized import PrioritizedReplayOracle, SumTreeOracle
from tests.test_oracle_prioritized import GOLD, replay_script

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cap", [37, 64, 100, 5, 1000])
def test_device_memory_matches_real_reference_class(cap):
    from maddpg_b200 import DevicePrioritizedReplayMemory
    gold = np.load(GOLD)

    def add(mem, serial, n):
        if n % 2:  # host scalars, one row per call (the reference's call shape) ...
            for j in range(n):
                s = float(serial + j)
                mem.add(np.float64(s), np.float32(0.5), s, np.float64(s + 1), 0.0)
        else:      # ... or one call with a leading env axis of device rows
            s = torch.arange(serial, serial + n, dtype=torch.float32, device="cuda")
            mem.add(s[:, None].contiguous(), torch.full((n, 1), 0.5, device="cuda"), s.clone(), (s + 1)[:, None].contiguous(),
                    torch.zeros(n, dtype=torch.uint8, device="cuda"))

    def sample(mem, n, u):
        b_idx, b_mem, isw = mem.sample(n, uniforms=u)
        # the stored row's reward column is its add serial: recover the data slot through the rows themselves
        np.testing.assert_array_equal(b_mem[0][:, 0], b_mem[2])
        np.testing.assert_array_equal(b_mem[3][:, 0], b_mem[2] + 1)
        return b_idx, mem.last_data_idx.cpu().numpy(), isw, float(mem.beta)

    def tree_of(mem):
        return mem.tree.cpu().numpy(), mem.dirty_count > 0

    replay_script(gold, cap, DevicePrioritizedReplayMemory, add, sample,
                  lambda mem, ti, ae: mem.batch_update(ti, ae.copy()), tree_of, isw_rtol=1e-13)


def test_rows_come_back_from_the_sampled_slots():
    from maddpg_b200 import DevicePrioritizedReplayMemory
    mem = DevicePrioritizedReplayMemory(500)
    rng = np.random.RandomState(0)
    obs, act = rng.randn(700, 6).astype(np.float32), rng.rand(700, 3).astype(np.float32)
    rew, nobs, done = rng.randn(700).astype(np.float32), rng.randn(700, 6).astype(np.float32), (rng.rand(700) < 0.2)
    for lo in range(0, 700, 100):
        sl = slice(lo, lo + 100)
        mem.add(torch.from_numpy(obs[sl]).cuda(), torch.from_numpy(act[sl]).cuda(), torch.from_numpy(rew[sl]).cuda(),
                torch.from_numpy(nobs[sl]).cuda(), torch.from_numpy(done[sl].astype(np.uint8)).cuda())
    np.random.seed(11)
    mem.strict = False
    b_idx, (o, a, r, n2, d), isw = mem.sample(64)
    slot = b_idx - mem.parent_nodes + 1
    serial = np.where(slot < 200, slot + 500, slot)  # 700 adds into 500 slots: slots [0, 200) hold rows 500..699
    ok = slot < 500
    assert ok.sum() >= 60
    np.testing.assert_array_equal(o[ok], obs[serial[ok]])
    np.testing.assert_array_equal(a[ok], act[serial[ok]])
    np.testing.assert_array_equal(r[ok], rew[serial[ok]])
    np.testing.assert_array_equal(n2[ok], nobs[serial[ok]])
    np.testing.assert_array_equal(d[ok], done[serial[ok]].astype(np.float32))


@pytest.mark.parametrize("cap,E", [(100000, 4096), (131072, 8192), (5000, 5000)])
def test_flush_sample_update_match_oracle_at_size(cap, E):
    """Wide levels (one launch per level), wrap-around, slot 0, B = 1024 / 4096 updates with duplicate leaves: every tree array
    bit-identical to the oracle's, sampled tree indices identical."""
    from maddpg_b200 import DevicePrioritizedReplayMemory
    mem = DevicePrioritizedReplayMemory(cap, strict=False, numpy_io=False)
    orc = PrioritizedReplayOracle(cap)
    rng = np.random.RandomState(cap % 977)
    z = lambda *s: torch.zeros(s, device="cuda")
    n_adds = [3, 1, 2, 30] if cap > 5000 else [1, 1]
    for rnd, reps in enumerate(n_adds):
        for _ in range(reps):
            mem.add(z(E, 2), z(E, 1), z(E), z(E, 2), torch.zeros(E, dtype=torch.uint8, device="cuda"))
            orc.tree.add(1e6, E)
        for B in (1024, 4096):
            u = rng.random_sample(B)
            total0 = orc.tree.total_p
            tidx, _, isw = mem.sample(B, uniforms=u)
            # oracle: same stratified values, no IndexError abort (non-strict device mode clamps instead)
            orc.beta = float(np.min([1.0, orc.beta + orc.beta_increment_per_sampling]))
            seg = total0 / B
            want = []
            for i in range(B):
                a, b = seg * i, seg * (i + 1)
                want.append(orc.tree.get_leaf(a + (b - a) * u[i])[0])
            assert np.array_equal(mem.tree.cpu().numpy(), orc.tree.tree), (rnd, B, "flush")
            assert np.array_equal(tidx.cpu().numpy(), np.asarray(want)), (rnd, B)
            ti = np.asarray(want, np.int64)
            ti = ti[ti >= orc.tree.parent_nodes]            # true leaves only (slot 0's node has its own test)
            ti = np.concatenate([ti, ti[: B - ti.size]])    # refill to B with duplicates
            errs = np.abs(rng.randn(B)) * 0.6
            mem.batch_update(ti, errs.copy())
            orc.batch_update(ti, errs.copy())
            assert int(mem.flag.item()) & 2 == 0
            assert np.array_equal(mem.tree.cpu().numpy(), orc.tree.tree), (rnd, B, "update")


def test_update_of_internal_nodes_takes_the_reference_loop():
    """tree indices that are not true leaves (slot 0's node q, any internal node): the one-thread reference loop runs."""
    from maddpg_b200 import DevicePrioritizedReplayMemory
    cap = 64
    mem = DevicePrioritizedReplayMemory(cap)
    orc = PrioritizedReplayOracle(cap)
    z = lambda *s: torch.zeros(s, device="cuda")
    mem.add(z(cap, 2), z(cap, 1), z(cap), z(cap, 2), torch.zeros(cap, dtype=torch.uint8, device="cuda"))
    orc.tree.add(1e6, cap)
    mem.flush()
    orc.tree.update_all()
    q = 2 ** mem.k - 2
    ti = np.asarray([q + 5, q, 2 * q + 1, q, 3, q + 5, 2 * q + 2], np.int64)  # slot cap-1's leaf hangs under q when cap == 2^k
    errs = np.asarray([0.3, 0.2, 0.9, 0.05, 0.5, 0.01, 0.4])
    mem.batch_update(ti, errs.copy())
    orc.batch_update(ti, errs.copy())
    assert int(mem.flag.item()) & 2
    assert np.array_equal(mem.tree.cpu().numpy(), orc.tree.tree)
    ti2 = np.asarray([q + 1, q + 2, q + 1], np.int64)  # back to the parallel kernels: the flag clears
    mem.batch_update(ti2, errs[:3].copy())
    orc.batch_update(ti2, errs[:3].copy())
    assert int(mem.flag.item()) & 2 == 0
    assert np.array_equal(mem.tree.cpu().numpy(), orc.tree.tree)


def test_device_priorities_use_cuda_pow():
    from maddpg_b200 import DevicePrioritizedReplayMemory
    cap = 3000
    mem = DevicePrioritizedReplayMemory(cap, numpy_io=False, strict=False)
    orc = PrioritizedReplayOracle(cap)
    z = lambda *s: torch.zeros(s, device="cuda")
    mem.add(z(cap, 2), z(cap, 1), z(cap), z(cap, 2), torch.zeros(cap, dtype=torch.uint8, device="cuda"))
    orc.tree.add(1e6, cap)
    mem.flush()
    orc.tree.update_all()
    rng = np.random.RandomState(5)
    ti = rng.randint(orc.tree.parent_nodes, orc.tree.parent_nodes + cap - 1, size=512).astype(np.int64)
    errs = np.abs(rng.randn(512))
    mem.batch_update(torch.from_numpy(ti).cuda(), torch.from_numpy(errs).cuda())
    orc.batch_update(ti, errs.copy())
    np.testing.assert_allclose(mem.tree.cpu().numpy(), orc.tree.tree, rtol=1e-14, atol=0)


@pytest.mark.parametrize("seed", range(12))
def test_random_scripts_match_oracle(seed):
    """Random capacities (odd and even tree depths, powers of two, capacity 3) and random add / sample / batch_update scripts:
    the device tree equals the pinned oracle's bit for bit after every call, sampled tree indices are identical, and the
    device raises IndexError exactly where the oracle (= the reference) does."""
    from maddpg_b200 import DevicePrioritizedReplayMemory
    rng = np.random.RandomState(500 + seed)
    cap = int(rng.choice([3, 4, 7, 8, 16, 33, 64, 100, 129, 200, 511, 513]))
    mem, orc = DevicePrioritizedReplayMemory(cap), PrioritizedReplayOracle(cap)
    last = None
    for step in range(30):
        op = rng.choice(["add", "sample", "update"], p=[0.4, 0.35, 0.25])
        if op == "add":
            n = int(rng.randint(1, 2 * cap))
            s = torch.zeros(n, device="cuda")
            mem.add(s[:, None].contiguous(), s[:, None].contiguous(), s, s[:, None].contiguous(), torch.zeros(n, dtype=torch.uint8, device="cuda"))
            orc.tree.add(1e6, n)
        elif op == "sample" and mem.ring is not None:
            n = int(rng.randint(1, 3 * cap))
            u = rng.random_sample(n)
            try:
                want = orc.sample(n, u)
            except IndexError:
                want = None
            if want is None:
                with pytest.raises(IndexError):
                    mem.sample(n, uniforms=u)
                mem.beta = float(orc.beta)
            else:
                b_idx, _, isw = mem.sample(n, uniforms=u)
                assert np.array_equal(b_idx, np.asarray(want[0], np.int64)), (cap, step)
                np.testing.assert_allclose(isw, np.asarray(want[2], np.float64), rtol=1e-13, atol=0)
                last = np.asarray(want[0], np.int64)
            assert float(mem.beta) == float(orc.beta)
        elif op == "update" and last is not None:
            errs = np.abs(rng.randn(last.size)) * float(rng.choice([0.05, 0.5, 3.0]))
            for lo in range(0, last.size, 4096):
                mem.batch_update(last[lo:lo + 4096], errs[lo:lo + 4096].copy())
            orc.batch_update(last, errs.copy())
        if mem.dirty_count == 0:
            orc.tree.update_all()
            assert np.array_equal(mem.tree.cpu().numpy(), orc.tree.tree), (cap, step, op)
