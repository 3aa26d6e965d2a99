"""A few per-step rollout steps of a BASELINE config (for an ncu launch list: which kernel takes what share of a step).
   python tools/profile_rollout_steps.py <config 3|4|5> [steps]"""
import sys, torch
sys.path.insert(0, '.')
from bench import CONFIGS
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout
cfg = CONFIGS[int(sys.argv[1])]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
E = cfg["envs"]
env = BatchedMultiAgentEnv(cfg["scenario"], num_envs=E, num_agents=cfg["agents"], squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=cfg["units"], replay_capacity=E * 8)
roll = BatchedRollout(env, core, 25, mode="eager"); env.reset_device()
roll.run_eager(steps)
torch.cuda.synchronize()
print("ok")
