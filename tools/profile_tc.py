"""Runs the tensor-core TD-target kernel a few times on a filled replay ring (for `ncu -k regex:k_td_target_tc`).

    python tools/profile_tc.py [scenario] [num_agents] [envs] [batch] [grouped 0/1]
"""
import sys

import torch

sys.path.insert(0, ".")
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore  # noqa: E402
from maddpg_b200.rollout import BatchedRollout  # noqa: E402

scen = sys.argv[1] if len(sys.argv) > 1 else "simple_spread"
na = int(sys.argv[2]) if len(sys.argv) > 2 else 3
E = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
grouped = int(sys.argv[5]) if len(sys.argv) > 5 else 0
env = BatchedMultiAgentEnv(scen, num_envs=E, num_agents=na or None, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, replay_capacity=E * 30)
roll = BatchedRollout(env, core, 25, mode="eager")
env.reset_device()
roll.run(25)
core.set_tensor_cores(1)
idx = torch.randint(0, core.ring.length[0], (env.n, B), device="cuda")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    if grouped:
        core.update_all(core.ring.ring, idx=idx)
    else:
        core.td_target(0, core.ring.ring, idx=idx[0])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
