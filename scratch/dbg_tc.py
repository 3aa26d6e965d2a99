import os, sys, ctypes as C, torch
os.environ["MDP_LIB_NAME"] = "libmaddpg_b200_dbg.so"
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore, _lib
from maddpg_b200.rollout import BatchedRollout
E, B = 2048, 1024
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, num_agents=24, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, replay_capacity=E * 30)
roll = BatchedRollout(env, core, 25, mode="eager"); env.reset_device(); roll.run(25)
core.set_tensor_cores(1)
idx = torch.randint(0, core.ring.length[0], (B,), device="cuda")
y = core.td_target(0, core.ring.ring, idx=idx).clone()
for _ in range(3): core.critic_grads(0, core.ring.ring, y, idx=idx)
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
_lib.lib.__getattr__("mdp_debug_read")(buf)
t = list(buf)[:9]
names = ["fwd L1 loop (112 chunks)", "wait acc L1", "epilogue 1", "wait acc L2", "epilogue 2 (q, dq, dz2)", "wait acc B1 (dW2, dh1)", "epilogue 3 (dW2 out, dz1)", "dW1 loop (112 chunks)"]
for i, n in enumerate(names): print("%-32s %8.1f us" % (n, (t[i + 1] - t[i]) / 1965.0))
print("total %.1f us" % ((t[8] - t[0]) / 1965.0))
