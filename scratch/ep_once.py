import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout
E = 4096
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
roll = BatchedRollout(env, core, 25, mode="mega"); env.reset_device()
for _ in range(3): roll.run_mega(25)
torch.cuda.synchronize()
print("ok")
