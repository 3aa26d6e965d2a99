"""CUDA-event time of the per-step env kernel of a BASELINE config (graph of 24 launches, L2 flushed), with and without the
fused ring insert where the kernel has one.   python tools/time_env_step.py <config 2|3|4|5>"""
import sys, torch
sys.path.insert(0, '.')
from bench import CONFIGS
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
cfg = CONFIGS[int(sys.argv[1])]
E = cfg["envs"]
env = BatchedMultiAgentEnv(cfg["scenario"], num_envs=E, num_agents=cfg["agents"], squeeze=False)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * env.n, num_units=cfg["units"], replay_capacity=E * 4)
env.reset_device()
env.act.copy_(torch.softmax(torch.randn_like(env.act), -1))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for ring in (None, core.ring):
    for _ in range(3): env.step_device(ring=ring)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(24): env.step_device(ring=ring, cursor=(k * E) % (3 * E) if ring is not None else None)
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / 24)
    ts.sort()
    us = ts[len(ts) // 2]
    print("config %s %s: %.2f us per step, %.1f GB/s algorithmic env bytes (%.1f %% of 6542.7)" % (
        sys.argv[1], "step+insert" if ring is not None else "step", us, env.env_bytes_per_step * E / us / 1e3,
        env.env_bytes_per_step * E / us / 1e3 / 65.427))
