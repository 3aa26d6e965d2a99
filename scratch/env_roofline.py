import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv
def run(scn, na, E, dtype=torch.float32, reps=20):
    env = BatchedMultiAgentEnv(scn, num_envs=E, num_agents=na, squeeze=False, state_dtype=dtype)
    env.reset_device()
    env.act.copy_(torch.softmax(torch.randn_like(env.act), -1))
    for _ in range(3): env.step_device()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): env.step_device()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    byt = env.env_bytes_per_step * E
    print("%-18s N=%s E=%8d  %8.1f us/step  %7.1f GB/s algorithmic (%.1f%% of 6542.7)  %.2e agent-steps/s" % (scn, na, E, us, byt / us / 1e3, 100 * byt / us / 1e3 / 6542.7, E * env.n / us * 1e6))
for E in (4096, 65536, 1048576):
    run("simple_spread", 3, E)
run("simple_tag", None, 262144)
run("simple_world_comm", None, 262144)
run("simple_spread", 24, 32768)
run("simple_spread", 3, 1048576, torch.float64)
