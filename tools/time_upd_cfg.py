"""Sequential and grouped update rounds of a BASELINE config through GraphedUpdateRound (CUDA events, graph replays).
   python tools/time_upd_cfg.py <config 2|3|4|5>     (MDP_PDL=0 disables the programmatic dependent launches)"""
import sys, torch
sys.path.insert(0, '.')
from bench import CONFIGS
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound
cfg = CONFIGS[int(sys.argv[1])]
E, B = min(cfg["envs"], 4096), cfg["batch"]
env = BatchedMultiAgentEnv(cfg["scenario"], num_envs=E, num_agents=cfg["agents"], squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=cfg["units"], replay_capacity=E * 30)
roll = BatchedRollout(env, core, 25, mode="eager")
env.reset_device()
roll.run(26)
for grouped in (False, True):
    g = GraphedUpdateRound(core, B, use_graph=True, grouped=grouped)
    g.run(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.run(10); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("config %s %s: %.3f ms per round, %.0f critic updates/s" % (sys.argv[1], "grouped" if grouped else "sequential", ms, env.n / ms * 1e3))
