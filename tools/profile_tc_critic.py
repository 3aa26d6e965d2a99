"""Three launches of the tensor-core critic forward+backward kernel for one agent (for `ncu -k regex:k_critic_grads_tc`).

    python tools/profile_tc_critic.py [scenario] [num_agents] [envs] [batch]
"""
import sys

import torch

sys.path.insert(0, ".")
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore  # noqa: E402
from maddpg_b200.rollout import BatchedRollout  # noqa: E402

scen = sys.argv[1] if len(sys.argv) > 1 else "simple_spread"
na = int(sys.argv[2]) if len(sys.argv) > 2 else 24
E = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
env = BatchedMultiAgentEnv(scen, num_envs=E, num_agents=na or None, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, replay_capacity=E * 30)
roll = BatchedRollout(env, core, 25, mode="eager")
env.reset_device()
roll.run(25)
core.set_tensor_cores(1)
idx = torch.randint(0, core.ring.length[0], (B,), device="cuda")
y = core.td_target(0, core.ring.ring, idx=idx).clone()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    core.critic_grads(0, core.ring.ring, y, idx=idx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
