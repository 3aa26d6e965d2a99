"""Sequential and grouped update rounds of one scenario through GraphedUpdateRound (CUDA events, graph replays).
   python tools/time_upd_scen.py <scenario> [agents] [batch] [units]     (MDP_PLAN_TM=16|32 forces the row-tile height)"""
import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound
scen = sys.argv[1]
agents = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
U = int(sys.argv[4]) if len(sys.argv) > 4 else 64
E = 2048
kw = {"num_agents": agents} if agents else {}
env = BatchedMultiAgentEnv(scen, num_envs=E, squeeze=False, **kw)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=U, replay_capacity=E * 30)
roll = BatchedRollout(env, core, 25, mode="eager")
env.reset_device()
roll.run(26)
out = []
for grouped in (False, True):
    g = GraphedUpdateRound(core, B, use_graph=True, grouped=grouped)
    g.run(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.run(20); b.record(); torch.cuda.synchronize()
    out.append(a.elapsed_time(b) / 20)
print("%-24s n=%d x_dim=%d B=%d U=%d: sequential %.3f ms, grouped %.3f ms" % (scen, env.n, sum(env.obs_dims) + sum(core.act_dims), B, U, out[0], out[1]))
