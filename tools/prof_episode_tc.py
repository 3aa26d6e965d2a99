"""clock64 phase profile of the tcgen05 episode kernel (library built by tools/build_prof.sh with -DMDP_EPISODE_PROF):
   MDP_LIB_NAME=libmaddpg_b200_prof.so python tools/prof_episode_tc.py"""
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
t = roll.ep_return[E * 3:E * 3 + 12].cpu().tolist()
k = roll.ep_return[E * 3 + 12:E * 3 + 14].cpu().tolist()
print("launch us", a.elapsed_time(b) * 1e3)
names = ["wait layer-1 acc (MMA)", "epilogue 1 (ld, relu, split, st, fences)", "wait layer-2 acc (MMA)", "epilogue 2 + pair barrier",
         "gumbel-softmax + action store", "physics", "phys_bar + pos write + pos_bar", "obs writes", "pair barrier + issue layer 2",
         "fence.proxy.async", "pair barrier + issue layer 1", "head dot products (part of head)"]
print("kernel: prologue + epilogue %.0f  step loop %.0f cycles" % tuple(k))
tot = sum(t)
for n, v in zip(names, t): print("%-36s %8.0f cycles/step  %5.1f%%" % (n, v / 25, 100 * v / tot))
print("total cycles/step", tot / 25)
