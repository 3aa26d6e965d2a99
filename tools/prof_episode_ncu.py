"""A few launches of the persistent episode kernel at the bench shape (for ncu: -k regex:k_rollout_episode -s 3 -c 1)."""
import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout
E = 4096
dtype = torch.float64 if "f64" in sys.argv else torch.float32
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False, state_dtype=dtype)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
if "simt" in sys.argv:
    core.set_tensor_cores(-1)
roll = BatchedRollout(env, core, 25, mode="mega"); env.reset_device()
for _ in range(6): roll.run_mega(25)
torch.cuda.synchronize()
print("ok", roll.mega_launches)
