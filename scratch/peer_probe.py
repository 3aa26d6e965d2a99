import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore, _lib
from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound
from maddpg_b200.distributed import DataParallelUpdater, PeerGradExchange
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
E, B = 4096, 1024
env = BatchedMultiAgentEnv("simple_spread", num_envs=E, squeeze=False, device=dev, seed=rank)
core = MADDPGCore(env.obs_dims, env.action_space, [False] * 3, replay_capacity=300000, device=dev)
roll = BatchedRollout(env, core, 25, mode="mega"); env.reset_device(); roll.run(250)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) * 1e3 / n
g = GraphedUpdateRound(core, B, ctl=roll.ctl, use_graph=True)
t0 = timeit(lambda: g.run(1))
# (a) gradient bucket re-homed in symmetric memory, peers NOT bound
ex = PeerGradExchange(core)
ex.close()
g = GraphedUpdateRound(core, B, ctl=roll.ctl, use_graph=True)
t1 = timeit(lambda: g.run(1))
# (b) peers bound
ex2 = PeerGradExchange(core)
g = GraphedUpdateRound(core, B, ctl=roll.ctl, use_graph=True)
t2 = timeit(lambda: g.run(1))
# single adam kernel graphs
def graph_of(fn, reps=20):
    fn(); torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    return (lambda: gr.replay()), reps
f, reps = graph_of(lambda: core.clip_adam_polyak(0, 1, grad_scale=0.5))
t3 = timeit(f, 20) / reps
if rank == 0:
    print("round us: plain %.1f | symmetric bucket, no peers %.1f | fused peer exchange %.1f | adam(peer) kernel %.2f us" % (t0, t1, t2, t3), flush=True)
dist.barrier(); ex2.close(); dist.destroy_process_group()
