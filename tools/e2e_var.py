import sys, time, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import HostRollout
E = 4096
def wall(fn, n=400):
    for _ in range(30): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
for chunks, graph, ck in [(1, True, True), (2, True, True), (4, True, True), (8, True, True)]:
    env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False)
    core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
    host = HostRollout(env, core, chunks=chunks, use_graph=graph, copy_kernels=ck)
    st = {"obs": host.reset(), "t": 0}
    def step():
        a, o, r, d = host.step(st["obs"]); st["obs"] = o; st["t"] += 1
        if st["t"] % 25 == 0: st["obs"] = host.reset()
    us = wall(step)
    print("chunks %d graph %d copy_kernels %d: %.1f us/step  -> %.1f M agent-env-steps/s" % (chunks, graph, ck, us, E * 3 / us))
