import sys, time, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import HostRollout
E = 4096
for chunks in (1, 4, 8):
    env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False)
    core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
    host = HostRollout(env, core, chunks=chunks, use_graph=True, copy_kernels=True)
    obs = host.reset()
    for _ in range(6): a, obs, r, d = host.step(obs)
    g = host._graphs[0]
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): g.replay()
    b.record(); torch.cuda.synchronize()
    print("chunks %d: GPU-side %.1f us per replay (back to back)" % (chunks, a.elapsed_time(b) * 1e3 / 200))
    t0 = time.perf_counter()
    for _ in range(200): g.replay(); torch.cuda.current_stream().synchronize()
    print("          replay+sync wall %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
    host.ctl.dirty = True
