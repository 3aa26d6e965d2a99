import sys, time, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore, _lib
from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound
E, B = 4096, 1024
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
roll = BatchedRollout(env, core, 25, mode="mega"); env.reset_device(); roll.run(250)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) * 1e3 / n
g = GraphedUpdateRound(core, B, ctl=roll.ctl, use_graph=True)
print("graph round us", timeit(lambda: g.run(1)))
e = GraphedUpdateRound(core, B, use_graph=False)
print("eager round us", timeit(lambda: e.run(1)))
idx = torch.randint(0, core.ring.length[0], (B,), device="cuda")
y, _ = core._scratch(B)
ring = core.ring.ring
print("make_index", timeit(lambda: core.make_index(idx)))
print("td_target", timeit(lambda: core.td_target(0, ring, idx=idx)))
print("critic_grads", timeit(lambda: core.critic_grads(0, ring, y, idx=idx)))
print("adam q", timeit(lambda: core.clip_adam_polyak(0, 1)))
print("actor_grads", timeit(lambda: core.actor_grads(0, ring, idx=idx)))
print("adam p", timeit(lambda: core.clip_adam_polyak(0, 0)))
# graphs of single kernels replayed back-to-back (pure GPU time)
def graph_of(fn, reps=20):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    return lambda: gr.replay(), reps
for name, fn in [("td_target", lambda: core.td_target(0, ring, idx=idx)), ("critic_grads", lambda: core.critic_grads(0, ring, y, idx=idx)),
                 ("adam q", lambda: core.clip_adam_polyak(0, 1)), ("actor_grads", lambda: core.actor_grads(0, ring, idx=idx)),
                 ("make_index", lambda: core.make_index(idx, counter=1)), ("gather", lambda: core.ring.gather(idx))]:
    f, reps = graph_of(fn)
    print("graphed x%d %-14s %.2f us each" % (reps, name, timeit(f, 20) / reps))
