"""Three warm-up + one profiled launch of the per-step env kernel.  Usage: profile_env.py [scenario] [E] [num_agents]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from maddpg_b200 import BatchedMultiAgentEnv  # noqa: E402

scn = sys.argv[1] if len(sys.argv) > 1 else "simple_spread"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
na = int(sys.argv[3]) if len(sys.argv) > 3 else None
env = BatchedMultiAgentEnv(scn, num_envs=E, num_agents=na, squeeze=False)
env.reset_device()
env.act.copy_(torch.softmax(torch.randn_like(env.act), -1))
for _ in range(3):
    env.step_device()
torch.cuda.synchronize()
torch.cuda.profiler.start()
env.step_device()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
