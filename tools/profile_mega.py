"""Two warm-up + one profiled launch of the persistent episode kernel (bench workload)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore  # noqa: E402
from maddpg_b200.rollout import BatchedRollout  # noqa: E402

scenario = sys.argv[1] if len(sys.argv) > 1 else "simple_spread"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
U = int(sys.argv[3]) if len(sys.argv) > 3 else 64
env = BatchedMultiAgentEnv(scenario, num_envs=E, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=U, replay_capacity=max(E * 25 * 4, 200000))
roll = BatchedRollout(env, core, 25, mode="mega")
env.reset_device()
roll.run(50)
torch.cuda.synchronize()
torch.cuda.profiler.start()
roll.run(25)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("mode", roll.mode)
