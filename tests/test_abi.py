"""CPU checks of the C ABI: the library loads, exports every symbol include/maddpg_b200.h declares,
and its host-only entry points (dims / layouts / argument validation) behave."""
import ctypes as C
import os
import re

import numpy as np

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "maddpg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from maddpg_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 20
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert getattr(raw, n) is not None


SURVEY_DIMS = {  # SURVEY.md section 8 table + 8(d) bytes_env
    ("simple", 0): (1, [4], [5], 9, 81),
    ("simple_spread", 3): (3, [18] * 3, [5] * 3, 69, 411),
    ("simple_tag", 0): (4, [16, 16, 16, 14], [5] * 4, 82, 492),
    ("simple_world_comm", 0): (6, [34] * 4 + [28] * 2, [9, 5, 5, 5, 5, 5], 226, 1198),
    ("simple_spread", 24): (24, [144] * 24, [5] * 24, 3576, 15384),
}


@pytest.mark.parametrize("key", list(SURVEY_DIMS))
def test_env_dims_match_survey(key):
    from maddpg_b200.env import _dims_for
    A, D, K, Cdim, nbytes = SURVEY_DIMS[key]
    h, d = _dims_for(key[0], key[1])
    assert d.n_agents == A and list(d.obs_dim[:A]) == D and list(d.act_dim[:A]) == K
    assert d.obs_sum + d.act_sum == Cdim and d.env_bytes_per_step == nbytes
    assert d.obs_stride % 4 == 0 and d.act_stride % 4 == 0


def test_ring_layout_alignment_and_row_bytes():
    from maddpg_b200.replay import make_ring_layout
    for (sc, na), (A, D, K, Cdim, _) in SURVEY_DIMS.items():
        lay = make_ring_layout(D, K)
        assert lay.x_dim == Cdim and lay.nx_off % 4 == 0 and lay.rw_off % 4 == 0 and lay.row_stride % 4 == 0
        algorithmic = sum(2 * d + k + 2 for d, k in zip(D, K))  # 4(2D_i+K_i+2) bytes per agent row
        assert algorithmic <= lay.row_stride <= algorithmic + 12


def test_core_layout_and_flop_model():
    from maddpg_b200 import _lib
    cfg = _lib.CoreCfg()
    cfg.n_agents, cfg.num_units = 3, 64
    for i in range(3):
        cfg.obs_dim[i], cfg.act_dim[i], cfg.n_heads[i] = 18, 5, 1
        cfg.head_dim[i][0] = 5
    h = C.c_void_p()
    _lib.check(_lib.lib.mdp_core_create(C.byref(cfg), C.byref(h)))
    lay = _lib.CoreLayout()
    _lib.check(_lib.lib.mdp_core_get_layout(h, C.byref(lay)))
    # running params per agent: actor 18*64+64+64*64+64+64*5+5, critic 69*64+64+64*64+64+64+1 -> 43 218 total
    assert sum(lay.net_size[i][0] + lay.net_size[i][2] for i in range(3)) == 43218  # SURVEY 8(d) allreduce payload
    B = 1024
    assert abs(lay.update_flops_critic[0] * B - 95.4e6) < 0.1e6
    assert abs(lay.update_flops_actor[0] * B - 58.6e6) < 0.1e6
    _lib.lib.mdp_core_destroy(h)


def test_errors_are_reported_not_thrown():
    from maddpg_b200 import _lib
    bad = _lib.EnvCfg(99, 0, 0)
    h = C.c_void_p()
    rc = _lib.lib.mdp_env_create(C.byref(bad), C.byref(h))
    assert rc == _lib.MDP_ENOTSUP and b"scenario" in _lib.lib.mdp_last_error()
    with pytest.raises(NotImplementedError):
        _lib.check(rc, "mdp_env_create")
    cfg = _lib.CoreCfg()
    cfg.n_agents, cfg.num_units = 1, 96
    assert _lib.lib.mdp_core_create(C.byref(cfg), C.byref(h)) == _lib.MDP_EINVAL


def test_unsupported_action_space_raises_like_make_pdtype():
    from maddpg_b200.spaces import act_heads, Box, Discrete, MultiDiscrete
    assert act_heads(Discrete(5)) == [5] and act_heads(MultiDiscrete([[0, 4], [0, 3]])) == [5, 4]
    with pytest.raises(NotImplementedError):  # distributions.py:422
        act_heads(Box(-1, 1, (2,)))


@pytest.mark.parametrize("scenario", ["simple", "simple_spread", "simple_tag", "simple_world_comm", "simple_adversary", "simple_push",
                                      "simple_speaker_listener", "simple_crypto", "simple_reference"])
def test_scenario_tables_match_oracle_shapes(scenario):
    """mdp_env_create builds the entity / observation-column / action-head tables without a GPU: observation widths, action
    heads, communication rows, goal rows and immovable agents agree with the oracle's MultiAgentEnv for all nine scenarios."""
    from maddpg_b200.env import _dims_for
    from oracle import mpe
    from oracle.maddpg import act_heads
    oenv = mpe.make_env(scenario, np.random.RandomState(0))
    h, d = _dims_for(scenario)
    A = d.n_agents
    assert A == oenv.n and d.n_landmarks == len(oenv.world.landmarks)
    assert list(d.obs_dim[:A]) == [s.shape[0] for s in oenv.observation_space]
    heads = [act_heads(s) for s in oenv.action_space]
    assert [list(d.head_dim[i][:d.n_heads[i]]) for i in range(A)] == heads
    assert list(d.act_dim[:A]) == [sum(hh) for hh in heads]
    assert list(d.movable[:A]) == [int(a.movable) for a in oenv.world.agents]
    speaking = [0 if a.silent else oenv.world.dim_c for a in oenv.world.agents]
    stored = list(d.comm_len[:A])
    # a speaking agent's state.c is stored when somebody's observation reads it (simple_crypto: everybody speaks, everybody's
    # message enters a reward)
    assert all(s in (0, sp) for s, sp in zip(stored, speaking)) and d.comm_dim == sum(stored)
    assert d.n_goal == getattr(oenv.scenario, "n_goal", 0)
    assert d.state_comps == 4 * A + d.comm_dim + 2 * d.n_landmarks + d.n_goal
    assert d.collaborative == int(oenv.world.collaborative)
