import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout
E = 4096
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
roll = BatchedRollout(env, core, 25, mode="mega"); env.reset_device()
roll.ep_return = torch.zeros(E * 3 + 16, device="cuda")
for _ in range(5): roll.run_mega(25)
torch.cuda.synchronize()
roll.ep_return.zero_()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); roll.run_mega(25); b.record(); torch.cuda.synchronize()
t = roll.ep_return[E * 3:E * 3 + 10].cpu().tolist()
print("launch us", a.elapsed_time(b) * 1e3)
names = ["wait+sync after mlp", "physics", "L1 setup", "fence+sync", "bulk store issue", "L1 G.sync", "layer2", "L1 mma", "head+gumbel", "L1 store"]
tot = sum(t)
for n, v in zip(names, t): print("%-22s %8.0f cycles/step  %5.1f%%" % (n, v / 25, 100 * v / tot))
print("total cycles/step", tot / 25)
