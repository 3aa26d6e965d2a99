"""Short eager (no CUDA graph) run of the bench workload for ncu: 2 warm-up episodes + 2 warm-up update
rounds, then ONE episode (25 x [actor, env step, insert] + reset) and TWO update rounds.
Usage: python tools/profile_workload.py [scenario] [envs] [batch] [units] [num_agents]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore  # noqa: E402
from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound  # noqa: E402

scenario = sys.argv[1] if len(sys.argv) > 1 else "simple_spread"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
U = int(sys.argv[4]) if len(sys.argv) > 4 else 64
NA = int(sys.argv[5]) if len(sys.argv) > 5 else None
env = BatchedMultiAgentEnv(scenario, num_envs=E, num_agents=NA, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=U, replay_capacity=max(E * 25 * 4, 200000))
roll = BatchedRollout(env, core, 25, use_graph=False)
upd = GraphedUpdateRound(core, B, use_graph=False)
env.reset_device()
roll.run(50)
upd.run(2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
roll.run(25)
upd.run(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled region done")
