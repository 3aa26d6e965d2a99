"""A few launches of the per-step env kernel of a BASELINE config, with and without the fused ring insert (for ncu).
   python tools/profile_env_step.py <config 2|3|4|5>"""
import sys, torch
sys.path.insert(0, '.')
from bench import CONFIGS
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
cfg = CONFIGS[int(sys.argv[1])]
E = cfg["envs"]
env = BatchedMultiAgentEnv(cfg["scenario"], num_envs=E, num_agents=cfg["agents"], squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=cfg["units"], replay_capacity=E * 4)
env.reset_device()
env.act.copy_(torch.softmax(torch.randn_like(env.act), -1))
for _ in range(4): env.step_device()
for _ in range(4): env.step_device(ring=core.ring)
torch.cuda.synchronize()
print("ok")
