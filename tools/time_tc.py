import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import BatchedRollout

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) * 1e3 / n

def graph_of(fn, reps=20):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    return (lambda: gr.replay()), reps

for scen, na, E, B in [("simple_spread", 3, 4096, 1024), ("simple_tag", None, 16384, 4096), ("simple_spread", 24, 2048, 1024)]:
    env = BatchedMultiAgentEnv(scen, num_envs=E, num_agents=na, squeeze=False)
    core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, replay_capacity=E * 30)
    roll = BatchedRollout(env, core, 25, mode="eager"); env.reset_device(); roll.run(25)
    n = env.n
    idx = torch.randint(0, core.ring.length[0], (n, B), device="cuda")
    ring = core.ring.ring
    key = ("all", B)
    core._y[key] = torch.empty((n, B), dtype=torch.float32, device="cuda")
    import ctypes as C
    from maddpg_b200 import _lib
    def td_all():
        # grouped TD target only: reuse update_all's first launch through td_target per agent is not grouped; call ABI helper
        for j in range(n): core.td_target(j, ring, idx=idx[j])
    for mode in (-1, 1):
        core.set_tensor_cores(mode)
        f, reps = graph_of(lambda: core.td_target(0, ring, idx=idx[0]))
        t1 = timeit(f) / reps
        y0 = core.td_target(0, ring, idx=idx[0]).clone()
        f3, reps3 = graph_of(lambda: core.critic_grads(0, ring, y0, idx=idx[0]))
        t3 = timeit(f3) / reps3
        f4, reps4 = graph_of(lambda: core.actor_grads(0, ring, idx=idx[0]))
        t4 = timeit(f4) / reps4
        print("   critic_grads(agent 0) %.2f us   actor_grads(agent 0) %.2f us" % (t3, t4), flush=True)
        f2, reps2 = graph_of(lambda: core.update_all(ring, idx=idx), 5)
        t2 = timeit(f2) / reps2
        print("%s n=%d B=%d mode=%+d: td_target(agent 0) %.2f us ; grouped update_all round %.2f us" % (scen, n, B, mode, t1, t2), flush=True)
    del core, env, roll
    torch.cuda.empty_cache()
