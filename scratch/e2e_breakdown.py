import sys, time, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
from maddpg_b200.rollout import HostRollout
E = 4096
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=1000000)
host = HostRollout(env, core)
obs = host.reset()
def wall(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
st = {"obs": obs}
def step():
    a, o, r, d = host.step(st["obs"]); st["obs"] = o
print("host.step            %.1f us" % wall(step))
h_in = torch.from_numpy(host._views[0]["joint_obs"])
s = torch.cuda.current_stream()
def h2d(): host.d_in.copy_(h_in, non_blocking=True); s.synchronize()
def d2h(): host.h_out[0].copy_(host.d_out, non_blocking=True); s.synchronize()
def kern():
    core.act(host.d_in, env.act); env.step_device(ring=core.ring); s.synchronize()
def both(): host.d_in.copy_(h_in, non_blocking=True); host.h_out[0].copy_(host.d_out, non_blocking=True); s.synchronize()
def nothing(): s.synchronize()
print("H2D 0.9MB + sync     %.1f us" % wall(h2d))
print("D2H 1.2MB + sync     %.1f us" % wall(d2h))
print("H2D + D2H + sync     %.1f us" % wall(both))
print("3 kernels + sync     %.1f us" % wall(kern))
print("sync only            %.1f us" % wall(nothing))
def mega():
    from maddpg_b200 import _lib
    host_roll.run_mega(1, False); s.synchronize()
from maddpg_b200.rollout import BatchedRollout
host_roll = BatchedRollout(env, core, 25, mode="mega")
print("episode kernel x1 step + sync  %.1f us" % wall(mega))
