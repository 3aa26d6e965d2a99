#!/usr/bin/env python
"""Benchmark of the MADDPG hot path on B200 (contract in the task statement; metric from BASELINE.json:
"agent-env-steps/sec + critic updates/sec at 1/2/4/8 B200 vs CPU ref").

Workload (BASELINE.json configs[1], `--config 2`): simple_spread N=3, 4096 lockstep env instances PER GPU (weak
scaling: env instances and replay shards are rank-local), batch 1024, num_units 64, maddpg/maddpg.
`--config 3|4|5` runs configs[2..4] (simple_tag 16384 envs / batch 4096; simple_world_comm 65536 envs / num_units 128;
simple_spread N=24 32768 envs per GPU) through the same sections.

  step            one lockstep rollout step of all env instances of a rank: grouped actor inference +
                  Gumbel-softmax sampling, fused MPE step, replay insert; device reset every 25 steps
                  (experiments/train.py:110-133 batched).  value = E*A*K*n_gpus / max-over-ranks device time.
  critic_updates  sequential per-agent updates (maddpg/trainer/maddpg.py:167-194: gather, TD target, critic
                  step, actor step, polyak), each on a batch of 1024 rows per rank; under N>1 the gradient
                  bucket of the network being stepped is all-reduced over NCCL (2 collectives per agent update).
  e2e             the same loop through the reference-shaped public API (MADDPGAgentTrainer.action /
                  BatchedMultiAgentEnv.step / .experience / .update) with HOST numpy buffers: every H2D/D2H
                  copy is inside the timed region.

Timing protocol: CUDA events on the launching stream.  ONE timed repetition = exactly `--steps` lockstep steps
(whole 25-step episodes, up to --episodes-per-launch of them per launch of the persistent episode kernel, plus one
partial launch for the remainder), bracketed by barrier + synchronize, L2 flushed (256 MB write, untimed) before it;
the repetition is run `reps` times and the MEDIAN is reported (max over ranks per repetition).  Update rounds: L2
flushed before each round, median over rounds.

`--impl reference` times the restated reference loop; one bench step of that arm = 100 env steps per replica.

`--impl reference` times the restated reference loop (oracle/train_loop.py: numpy MPE + numpy trainer, one
env, batch-1 actor calls) as `os.cpu_count()` independent single-threaded replicas on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EP_LEN = 25
# BASELINE.json configs[1..4] (configs[0] is the reference's own 1-env CPU case: the reference arm's shape, not a bench line)
CONFIGS = {
    2: dict(scenario="simple_spread", agents=3, envs=4096, batch=1024, units=64,
            name="simple_spread N=3, 4096 envs per GPU, batch 1024, num_units 64, maddpg/maddpg (BASELINE.json configs[1])"),
    3: dict(scenario="simple_tag", agents=None, envs=16384, batch=4096, units=64,
            name="simple_tag 3 adversaries + 1 good + 2 obstacles, 16384 envs per GPU, batch 4096, num_units 64 (BASELINE.json configs[2])"),
    4: dict(scenario="simple_world_comm", agents=None, envs=65536, batch=1024, units=128,
            name="simple_world_comm 6 agents (leader comm), 65536 envs per GPU, batch 1024, num_units 128 (BASELINE.json configs[3])"),
    5: dict(scenario="simple_spread", agents=24, envs=32768, batch=1024, units=64,
            name="simple_spread N=24, 32768 envs per GPU (262144 over 8 GPUs), batch 1024, num_units 64 (BASELINE.json configs[4])"),
}
METRIC, UNIT = "agent-env-steps/sec", "agent-env-steps/s"
REF_ENV_STEPS_PER_STEP = 100  # reference arm: one bench "step" = this many env steps of the restated train.py loop per replica


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000, help="lockstep rollout steps in ONE timed repetition (honoured exactly)")
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[N-1]; 2 = the headline")
    ap.add_argument("--reps", type=int, default=0, help="timed repetitions of the --steps region (median reported); 0 = auto")
    ap.add_argument("--episodes-per-launch", type=int, default=8)
    ap.add_argument("--update-rounds", type=int, default=100)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--rollout-mode", default="auto", choices=["auto", "mega", "graph", "eager"],
                    help="auto: the persistent episode kernel or the CUDA graph of per-step kernels, whichever is faster on a probe")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--no-tensor-section", action="store_true")
    ap.add_argument("--nccl-allreduce", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# reference arm: the restated reference loop on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    from oracle import train_loop as tl
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per = REF_ENV_STEPS_PER_STEP  # a bench step of this arm = `per` env steps per replica (bounded sample of the workload)
    if warm:
        tl.time_parallel("rollout", cfg["scenario"], cfg["agents"], warm * per, cores, cfg["batch"], cfg["units"])
    value, wall = tl.time_parallel("rollout", cfg["scenario"], cfg["agents"], steps * per, cores, cfg["batch"], cfg["units"])
    rounds = max(2, min(20, steps))
    upd, upd_wall = tl.time_parallel("updates", cfg["scenario"], cfg["agents"], rounds, cores, cfg["batch"], cfg["units"])
    n_ag = {2: 3, 3: 4, 4: 6, 5: 24}[args.config]
    sample = ("%d single-threaded replicas of the restated train.py loop (numpy MPE + numpy trainer, 1 env each; the reference "
              "is single-threaded by construction, tf_util.py:202-204), %d steps x %d env steps each after %d x %d warm-up; "
              "updates: %d forced rounds x %d agents per replica" % (cores, steps, per, warm, per, rounds, n_ag))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 env / f32 nets", "data": "synthetic",
        "config": {"workload": cfg["name"], "env_steps_per_bench_step_per_replica": per,
                   "reference_path": "oracle/train_loop.py (TF-free restatement; real train.py "
                   "needs tensorflow 1.8 + gym + MPE, not installable here)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "critic_updates": {"value": upd, "unit": "critic updates/s", "cores": cores},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first(self, timeout):
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v == "Active":
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of the newest tracked `ncu --set full` summary (profiles/*.txt, written by
    tools/ncu_summary.py) that holds a launch of `kernel_substr`.  Returns (bytes or None, source)."""
    import glob
    import re
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = (None, "no tracked ncu summary holds this kernel")
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*.txt"))):
        try:
            text = open(path).read()
        except OSError:
            continue
        for blk in text.split("\nkernel: ")[1:]:
            if kernel_substr not in blk.splitlines()[0]:
                continue
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                m = re.search(r"^\s*%s\s+([0-9.,]+)\s+(\w+)" % re.escape(key), blk, re.M)
                if not m:
                    tot = None
                    break
                tot += float(m.group(1).replace(",", "")) * mult.get(m.group(2), 1.0)
            if tot is not None:
                best = (tot, "ncu --set full, %s (dram__bytes_read.sum + dram__bytes_write.sum of one launch)" % os.path.relpath(path, ROOT))
    return best


def tensor_core_section(torch, dev):
    """Grouped TD-target launch (maddpg.py:181-187 for all agents) at the simple_spread N=24 update shape, tcgen05 kernel
    vs the fp32 SIMT kernel, CUDA events on the launching stream, L2 flushed before every launch."""
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGCore
    from maddpg_b200.rollout import BatchedRollout
    NA, E, B, U = 24, 2048, 1024, 64
    env = BatchedMultiAgentEnv("simple_spread", num_envs=E, num_agents=NA, device=dev, squeeze=False)
    core = MADDPGCore(env.obs_dims, env.action_space, [False] * NA, num_units=U, device=dev, replay_capacity=E * 26)
    roll = BatchedRollout(env, core, EP_LEN, mode="eager")
    env.reset_device()
    roll.run(EP_LEN)
    g = torch.Generator(device="cpu").manual_seed(7)
    idx = torch.randint(0, core.ring.length[0], (NA, B), generator=g).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    Fpi = sum(d * U + U * U + U * k for d, k in zip(env.obs_dims, env.act_dims))
    Fq = (sum(env.obs_dims) + sum(env.act_dims)) * U + U * U + U
    flops = 2 * (Fpi + Fq) * B * NA
    out = {"workload": "simple_spread N=24 update shape per GPU (BASELINE.json configs[4]): grouped TD target of 24 agents, "
                       "batch 1024, critic input 3576, num_units 64", "algorithmic_flops_per_launch": flops, "bound": "tensor",
           "mma_issue_multiplier": 3,
           "note": "every GEMM is issued as 3 kind::tf32 MMAs (hi/lo split) to hold the 1e-4 parity bar; tensor-pipe "
                   "activity from ncu: profiles/r1_update_tc_cfg5_grouped.txt"}
    for name, mode in (("tcgen05", 1), ("simt_fp32", -1)):
        core.set_tensor_cores(mode)
        for _ in range(3):
            core.td_target_all(core.ring.ring, idx=idx)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in evs:
            flush.fill_(1.0)
            a.record()
            core.td_target_all(core.ring.ring, idx=idx)
            b.record()
        torch.cuda.synchronize()
        us = sum(a.elapsed_time(b) for a, b in evs) * 1e3 / len(evs)
        out[name] = {"avg_launch_us": us, "achieved_tflops": flops / us / 1e6}
    out["speedup_vs_simt"] = out["simt_fp32"]["avg_launch_us"] / out["tcgen05"]["avg_launch_us"]
    # kind::tf32 runs at half the bf16 rate: the roofline is MEASURED_PEAKS.json's cuBLAS bf16 burst figure / 2
    try:
        bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        src = "measured bf16 burst / 2 (MEASURED_PEAKS.json)"
    except Exception:
        bf16, src = 2250.0, "fallback: nominal bf16 / 2 (B200_PROFILING.md)"
    out["peak_tf32_tflops"], out["peak_source"] = bf16 / 2, src
    out["frac_algorithmic"] = out["tcgen05"]["achieved_tflops"] / (bf16 / 2)
    out["frac"] = 3 * out["tcgen05"]["achieved_tflops"] / (bf16 / 2)  # issued MMA work (3 per product) over the tf32 peak
    # the whole grouped (Jacobi) round at the same shape: TD target + critic forward/backward on tcgen05, actor step on SIMT
    rnd = {}
    for name, mode in (("tcgen05", 1), ("simt_fp32", -1)):
        core.set_tensor_cores(mode)
        for _ in range(2):
            core.update_all(core.ring.ring, idx=idx)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in evs:
            flush.fill_(1.0)
            a.record()
            core.update_all(core.ring.ring, idx=idx)
            b.record()
        torch.cuda.synchronize()
        rnd[name + "_us"] = sum(a.elapsed_time(b) for a, b in evs) * 1e3 / len(evs)
    rnd["critic_updates_per_s_tcgen05"] = NA / (rnd["tcgen05_us"] * 1e-6)
    rnd["critic_updates_per_s_simt"] = NA / (rnd["simt_fp32_us"] * 1e-6)
    out["grouped_update_round"] = rnd
    del env, core, roll, flush
    torch.cuda.empty_cache()
    return out


def next_rows_section(torch, batch, with_cpu):
    """prioritized replay at the reference's capacity (1e6 slots: a 21-level float64 tree), one
    sample(batch) + batch_update(batch) round per step; and one env step of every scenario at 4096 env instances."""
    import numpy as np
    from statistics import median
    from maddpg_b200 import BatchedMultiAgentEnv, DevicePrioritizedReplayMemory
    from maddpg_b200.env import SCENARIOS
    cap, E = 1000000, 4096
    mem = DevicePrioritizedReplayMemory(cap, numpy_io=False, strict=False)
    z = lambda *sh: torch.zeros(sh, device="cuda")
    for _ in range(25):  # 25 lockstep steps of 4096 env instances pending, like a rollout between two updates
        mem.add(z(E, 18), z(E, 5), z(E), z(E, 18), torch.zeros(E, dtype=torch.uint8, device="cuda"))
    u = torch.rand(batch, dtype=torch.float64, device="cuda")
    err = torch.rand(batch, dtype=torch.float64, device="cuda")

    def round_():
        tidx, _, _ = mem.sample(batch, uniforms=u)
        mem.batch_update(tidx, err)

    round_()  # flushes the 102 400 pending adds
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record(); round_(); b.record()
    torch.cuda.synchronize()
    round_us = median([a.elapsed_time(b) for a, b in ev]) * 1e3
    a, b = ev[0]
    mem.add(z(E, 18), z(E, 5), z(E), z(E, 18), torch.zeros(E, dtype=torch.uint8, device="cuda"))
    a.record(); mem.flush(); b.record()
    torch.cuda.synchronize()
    prio = {"capacity": cap, "batch": batch, "sample_plus_update_us": round_us, "flush_4096_adds_us": a.elapsed_time(b) * 1e3,
            "api": "DevicePrioritizedReplayMemory.sample(batch) + batch_update(tree_idx, |err|) (mdp_sumtree_sample / _update); "
                   "tree indices bit-identical to the reference class given the same uniforms (tests/test_prioritized_gpu.py)"}
    if with_cpu:
        from oracle.prioritized import PrioritizedReplayOracle
        orc = PrioritizedReplayOracle(cap)
        orc.tree.add(1e6, 25 * E)
        orc.tree.update_all()
        uu, ee = np.random.RandomState(0).random_sample(batch), np.random.RandomState(1).random_sample(batch)
        t0 = time.perf_counter()
        for _ in range(3):
            seg = orc.tree.total_p / batch
            idx = [orc.tree.get_leaf(seg * i + seg * uu[i])[0] for i in range(batch)]
            orc.batch_update(idx, ee.copy())
        prio["cpu_port_us"] = (time.perf_counter() - t0) / 3 * 1e6
    scen = {}
    for name in SCENARIOS:
        env = BatchedMultiAgentEnv(name, num_envs=E, squeeze=False)
        env.reset_device()
        env.act.copy_(torch.softmax(torch.randn_like(env.act), -1))
        g = torch.cuda.CUDAGraph()
        for _ in range(2):
            env.step_device()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(24):
                env.step_device()
        g.replay()
        a, b = ev[1]
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / 24
        scen[name] = {"agents": env.n, "obs_dims": env.obs_dims, "act_dims": env.act_dims, "env_step_us": us,
                      "agent_env_steps_per_s": E * env.n / (us * 1e-6), "bytes_per_env_step": env.env_bytes_per_step}
    return {"prioritized_replay": prio, "env_step_4096_envs": scen, "fork_algorithms": fork_algorithms_section(torch, with_cpu),
            "note": "env step alone (no actor, no insert), graph-replayed"}


def fork_algorithms_section(torch, with_cpu):
    """one MaTd3 / Coma train step + target update at the fork's own training shape (batch 1024,
    multiagentalgbase.py:127; 64 units) for 3 agents with 18-wide observations and 2-wide Box actions."""
    import numpy as np
    from statistics import median
    from maddpg_b200 import _lib
    from maddpg_b200.algorithms import Coma, MaTd3
    from maddpg_b200.spaces import Box, Dict
    B, n, D, K = 1024, 3, 18, 2
    names = ["agent_%d" % i for i in range(n)]
    obs_sp = Dict({k: Box(-np.inf, np.inf, (D,)) for k in names})
    act_sp = Dict({k: Box(-np.ones(K, np.float32), np.ones(K, np.float32), (K,)) for k in names})
    rng = np.random.RandomState(0)
    feed = [{k: rng.randn(B, D).astype(np.float32) for k in names}, {k: rng.uniform(-1, 1, (B, K)).astype(np.float32) for k in names},
            {k: rng.randn(B, 1).astype(np.float32) for k in names}, {k: rng.randn(B, D).astype(np.float32) for k in names},
            {k: (rng.rand(B, 1) < 0.1).astype(np.float32) for k in names}]
    out = {"batch": B, "agents": n, "obs_dim": D, "act_dim": K}
    for cls in (MaTd3, Coma):
        alg = cls(obs_sp, act_sp, seed=0)
        for _ in range(5):
            alg.train_step(*feed, step=2)
            alg.run_updates()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            alg.train_step(*feed, step=2)   # host dicts in, host losses out
            alg.run_updates()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        row = {"train_step_plus_target_update_ms": median(ts) * 1e3, "launches_outside_the_graph": (_lib.launch_count() - l0) / 20,
               "api": "%s.train_step(host dicts, step=2) + run_updates(): the step's kernels replay as one CUDA graph, the "
                      "target updates launch one by one" % cls.__name__}
        if with_cpu:
            from oracle.matd3 import ComaOracle, MaTd3Oracle
            o = (MaTd3Oracle if cls is MaTd3 else ComaOracle)({k: D for k in names}, {k: K for k in names}, {k: -1.0 for k in names},
                                                              {k: 1.0 for k in names}, seed=0)
            kw = {"z": {k: rng.randn(B, K).astype(np.float32) for k in names}} if cls is MaTd3 else {}
            o.train_step(*feed, step=2, **kw)
            t0 = time.perf_counter()
            for _ in range(3):
                o.train_step(*feed, step=2, **kw)
                o.run_updates()
            row["cpu_port_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        out[cls.__name__] = row
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import math
    import numpy as np
    import torch
    import torch.distributed as dist
    from maddpg_b200 import BatchedMultiAgentEnv, MADDPGAgentTrainer, _lib
    from maddpg_b200.distributed import DataParallelUpdater, rank_seed
    from maddpg_b200.rollout import BatchedRollout, GraphedUpdateRound

    cfg = CONFIGS[args.config]
    SCENARIO, N_AGENTS, BATCH, UNITS, WORKLOAD = cfg["scenario"], cfg["agents"], cfg["batch"], cfg["units"], cfg["name"]
    headline = args.config == 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    E = args.envs or cfg["envs"]
    K, W = max(1, args.steps), max(3, args.warmup)
    reps = args.reps or min(50, max(5, int(math.ceil(1000.0 / K))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(xs):
        t = torch.tensor(list(xs), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().tolist()

    def median(xs):
        xs = sorted(xs)
        n = len(xs)
        return xs[n // 2] if n % 2 else 0.5 * (xs[n // 2 - 1] + xs[n // 2])

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def flush_l2():
        flush_buf.fill_(1.0)

    # ---- build the experiment exactly like train.py:80-85 (env, trainers) on this rank's shard -------------
    eps_launch = max(1, args.episodes_per_launch)
    tc_episode = SCENARIO == "simple_spread" and N_AGENTS in (2, 3, 4) and UNITS == 64  # mdp_rollout_episodes loops in-kernel
    if not tc_episode:
        eps_launch = 1
    cap = max(1000000, E * EP_LEN * (eps_launch + 1))  # the reference's 1e6 rows (maddpg.py:147), at least one launch deep
    arglist = argparse.Namespace(lr=1e-2, gamma=0.95, batch_size=BATCH, num_units=UNITS, max_episode_len=EP_LEN,
                                 seed=0, device=str(dev), replay_capacity=cap)
    env = BatchedMultiAgentEnv(SCENARIO, num_envs=E, num_agents=N_AGENTS, device=dev, seed=rank_seed(0, rank), squeeze=False)
    A = env.n
    obs_shape_n = [env.observation_space[i].shape for i in range(env.n)]
    trainers = [MADDPGAgentTrainer("agent_%d" % i, None, obs_shape_n, env.action_space, i, arglist) for i in range(env.n)]
    core = trainers[0].core
    # N > 1: the gradient all-reduce is fused into the clip+Adam+polyak kernel (peer loads over NVLink, PeerGradExchange),
    # so update rounds are kernels only and replay from a CUDA graph; --nccl-allreduce keeps the NCCL collective path
    exchange = "none"
    if world > 1:
        exchange = "nccl"
        if not args.nccl_allreduce:
            try:
                dp = DataParallelUpdater(core, peer=True)
                exchange = "fused-peer"
            except Exception as exc:  # symmetric memory unavailable: fall back to the NCCL collective
                sys.stderr.write("peer exchange unavailable (%r), using NCCL all-reduce\n" % (exc,))
                dp = DataParallelUpdater(core)
        else:
            dp = DataParallelUpdater(core)
    else:
        dp = DataParallelUpdater(core)
    dp.broadcast_params(core.params)

    # ---- (1) device-resident rollout: exactly K lockstep steps per repetition, median over repetitions ---------
    def run_steps(roll, n):
        """n lockstep steps: whole episodes (each ends with env.reset(), train.py:127-129) + a partial episode."""
        full, rem = divmod(n, EP_LEN)
        if roll.mode == "mega":
            if full:
                roll.run_episodes(full)
            if rem and roll.mode == "mega":
                if not roll.run_mega(rem, reset_after=False):
                    roll.mode = "graph"
                    roll.run_eager(rem)
            elif rem:
                roll.run_eager(rem)
        else:
            roll.run(full * EP_LEN)
            if rem:
                roll.run_eager(rem)

    def time_rollout(roll):
        run_steps(roll, max(W, EP_LEN))
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        l0 = _lib.launch_count() + roll.graph_launches
        for e0, e1 in evs:
            barrier()      # every repetition is bracketed by barrier + synchronize (this one and the next iteration's / the final one)
            flush_l2()     # untimed: enqueued ahead of the first event, so the timed launches do not start on an idle queue
            e0.record()
            run_steps(roll, K)
            e1.record()
        barrier()
        launches = (_lib.launch_count() + roll.graph_launches - l0) // reps
        ms = max_over_ranks(a.elapsed_time(b) for a, b in evs)
        return median(ms), ms, launches

    mode, probe = args.rollout_mode, None
    if mode == "auto":
        # probe: 2 episodes through each path after one warm-up episode (device time); the headline config always takes the
        # persistent kernel (it wins by > 3x there), the larger configs may not fit it or may run it from streamed weights
        mode, probe = "mega", {}
        if not tc_episode:
            for m in ("mega", "graph"):
                r = BatchedRollout(env, core, EP_LEN, mode=m)
                env.reset()
                r.run(2 * EP_LEN)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                r.run(2 * EP_LEN)
                b.record()
                torch.cuda.synchronize()
                probe[r.mode if m == "mega" else m] = a.elapsed_time(b) / (2 * EP_LEN)
                del r
            mode = min(probe, key=probe.get)
            t = torch.tensor([0 if mode == "mega" else 1], device=dev)
            if world > 1:  # every rank must run the same path
                dist.broadcast(t, src=0)
            mode = "mega" if int(t.item()) == 0 else "graph"
    roll = BatchedRollout(env, core, EP_LEN, mode=mode)
    roll.episodes_per_launch = eps_launch
    env.reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first(3.0)
    roll_ms, roll_all_ms, launches_roll = time_rollout(roll)
    value = E * A * K * world / (roll_ms * 1e-3)
    ep_kernel = "per-step kernels (CUDA graph)" if roll.mode != "mega" else \
        ("k_rollout_episode_tc<%d,float>" % A if (SCENARIO == "simple_spread" and 2 <= A <= 4 and UNITS == 64) else "k_rollout_episode")

    # ---- (1b) the same rollout with the reference's float64 env state (the mode of the 1e-5 parity contract) ----
    f64 = None
    if headline:
        env64 = BatchedMultiAgentEnv(SCENARIO, num_envs=E, num_agents=N_AGENTS, device=dev, seed=rank_seed(0, rank),
                                     squeeze=False, state_dtype=torch.float64)
        roll64 = BatchedRollout(env64, core, EP_LEN, mode="mega")
        roll64.episodes_per_launch = eps_launch
        env64.reset()
        ms64, all64, l64 = time_rollout(roll64)
        f64 = {"value": E * A * K * world / (ms64 * 1e-3), "unit": UNIT, "ms_per_step": ms64 / K, "rollout_mode": roll64.mode,
               "kernel": "k_rollout_episode_tc<%d,double>" % A if roll64.mode == "mega" else "per-step kernels",
               "note": "positions / velocities / contact forces / rewards in float64 like MPE's numpy state, observations emitted "
                       "as float32 (the reference's placeholder cast); tests/test_rollout_tc_gpu.py replays this mode free-running "
                       "against the float64 oracle"}
        del env64, roll64
        torch.cuda.empty_cache()

    # ---- (2) env-step kernel alone (roofline) ----------------------------------------------------------------
    reps_k, per = 20, EP_LEN - 1  # an even number of launches keeps the observation double buffer in phase
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps_k)]
    env_graph = None
    if not args.no_graph:
        env_graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(env_graph):
            for _ in range(per):
                env.step_device()
    for a, b in kev:
        flush_l2()
        a.record()
        if env_graph is not None:
            env_graph.replay()
        else:
            for _ in range(per):
                env.step_device()
        b.record()
    torch.cuda.synchronize()
    env_us = median([a.elapsed_time(b) for a, b in kev]) * 1e3 / per
    peak, peak_src = measured_peak_hbm()
    env_bytes = env.env_bytes_per_step * E
    achieved = env_bytes / (env_us * 1e-6) / 1e9
    env_kernel = ("k_env_step_spread<%d>" % A if A <= 6 else "k_env_step_spread_warp<%d,false>" % (A if A == 24 else 0)) if SCENARIO == "simple_spread" \
        else "k_env_step<float> (table-driven)"
    # the same per-step kernel where it is actually HBM-sized: 2^20 env instances (431 MB per step)
    big = None
    if headline and rank == 0 and not args.no_graph:
        EL = 1 << 20
        envL = BatchedMultiAgentEnv(SCENARIO, num_envs=EL, num_agents=N_AGENTS, device=dev, squeeze=False)
        envL.reset_device()
        envL.act.copy_(torch.softmax(torch.randn_like(envL.act), -1))
        for _ in range(4):
            envL.step_device()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            envL.step_device()
        b.record()
        torch.cuda.synchronize()
        usL = a.elapsed_time(b) * 1e3 / 10
        bytesL = envL.env_bytes_per_step * EL
        trafficL, trafficL_src = traffic_from_profile("k_env_step_spread")
        big = {"kernel": "k_env_step_spread<3>", "envs": EL, "bound": "hbm", "achieved": bytesL / usL / 1e3, "peak": peak,
               "unit": "GB/s", "frac": bytesL / usL / 1e3 / peak, "algorithmic_bytes_per_launch": bytesL, "avg_launch_us": usL,
               "traffic": trafficL, "traffic_source": trafficL_src,
               "note": "register-resident one-thread-per-env kernel (simple_spread fast path); the table-driven kernel serves the other scenarios"}
        del envL
        torch.cuda.empty_cache()

    # ---- (3) critic updates: sequential agent updates on gathered batches ------------------------------------
    while core.ring.length[0] < BATCH * EP_LEN:  # the reference's warm-up gate (maddpg.py:148,162)
        run_steps(roll, EP_LEN)
    R = max(1, args.update_rounds)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)

    def draw_idx():
        return torch.randint(0, core.ring.length[0], (BATCH,), generator=g).to(dev, non_blocking=False)

    idx_pool = [[draw_idx() for _ in range(A)] for _ in range(8)]
    _, batch = core._scratch(BATCH)

    fused = world == 1 or exchange == "fused-peer"

    def time_updates(make_round, rounds):
        rnd = make_round()
        for r in range(3):
            rnd(r)
        barrier()
        uev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(rounds)]
        l0 = _lib.launch_count()
        for r, (a, b) in enumerate(uev):
            flush_l2()
            a.record()
            rnd(r)
            b.record()
        barrier()
        ms = max_over_ranks(a.elapsed_time(b) for a, b in uev)
        return median(ms), sum(ms)

    gupd = GraphedUpdateRound(core, BATCH, ctl=roll.ctl, use_graph=not args.no_graph) if fused else None

    def seq_round():
        if gupd is not None:  # device-side index draw + gather + update kernels per agent, graph-replayed
            return lambda r: gupd.run(1)
        return lambda r: [dp.update_agent(j, core.ring.ring, idx=idx_pool[r % 8][j]) for j in range(A)]

    upd_med_ms, upd_sum_ms = time_updates(seq_round, R)
    launches_upd = (gupd.launches_per_graph if gupd is not None and gupd.use_graph else 0)
    upd_value = A * world / (upd_med_ms * 1e-3)
    # ---- (3b) grouped ("Jacobi") rounds: all agents per launch -- throughput mode, documented deviation ----------
    grp_value = grp_ms = None
    ggrp = None
    if fused:
        ggrp = GraphedUpdateRound(core, BATCH, ctl=roll.ctl, use_graph=not args.no_graph, grouped=True)
        grp_ms, _ = time_updates(lambda: (lambda r: ggrp.run(1)), R)
        grp_value = A * world / (grp_ms * 1e-3)
    # ---- (3c) N > 1: the same rounds with the exchange switched off (every rank steps on its own gradients): the
    #      single-GPU rate on THIS rank and therefore the efficiency of the gradient exchange --------------------------
    noex = None
    if world > 1 and fused:
        dp.set_exchange(False)
        g1 = GraphedUpdateRound(core, BATCH, ctl=roll.ctl, use_graph=not args.no_graph)
        m1, _ = time_updates(lambda: (lambda r: g1.run(1)), max(10, R // 2))
        g2 = GraphedUpdateRound(core, BATCH, ctl=roll.ctl, use_graph=not args.no_graph, grouped=True)
        m2, _ = time_updates(lambda: (lambda r: g2.run(1)), max(10, R // 2))
        dp.set_exchange(True)
        dp.broadcast_params(core.params)  # the replicas diverged while the exchange was off
        noex = {"sequential_ms_per_round": m1, "grouped_ms_per_round": m2,
                "sequential_efficiency": m1 / upd_med_ms, "grouped_efficiency": m2 / grp_ms,
                "note": "the same update rounds on the same GPUs with the gradient exchange off (every rank steps on its local "
                        "gradients): efficiency = time without / time with the exchange = N-GPU rate / (N x single-GPU rate)"}
    clocks = sampler.stop()

    # ---- (4) end to end through the reference-shaped API with host buffers ------------------------------------
    Ke = max(EP_LEN, args.e2e_steps // EP_LEN * EP_LEN) if headline else EP_LEN
    obs_n = [o.cpu().numpy() for o in env.reset()]
    h2d = d2h = 0

    def e2e_step(obs_n, episode_step):
        action_n = [agent.action(obs) for agent, obs in zip(trainers, obs_n)]               # train.py:112
        new_obs_n, rew_n, done_n, info_n = env.step(action_n)                                 # train.py:114
        for i, agent in enumerate(trainers):                                                  # train.py:119-120
            agent.experience(obs_n[i], action_n[i], rew_n[i], new_obs_n[i], done_n[i], False)
        episode_step += 1
        if episode_step >= EP_LEN:                                                            # train.py:127-129
            new_obs_n = [o.cpu().numpy() for o in env.reset()]
            episode_step = 0
        return new_obs_n, episode_step

    ep = 0
    for _ in range(5):
        obs_n, ep = e2e_step(obs_n, ep)
    obs_n = [o.cpu().numpy() for o in env.reset()]
    ep = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        obs_n, ep = e2e_step(obs_n, ep)
    barrier()
    percall_s = max_over_ranks([time.perf_counter() - t0])[0]
    percall_value = E * A * Ke * world / percall_s
    row_f = sum(2 * d + k + 2 for d, k in zip(env.obs_dims, env.act_dims))
    percall_h2d = 4 * E * (sum(env.obs_dims) + env.act_stride + row_f)   # action() obs, step() actions, experience() rows
    percall_d2h = 4 * E * (sum(env.act_dims) + env.obs_stride + env.n)   # actions, [obs | rew]
    # the same loop body as ONE C-ABI call per lockstep step with host buffers (mdp_host_step): H2D of the joint
    # observations, actors + env step + replay insert, one packed D2H of (actions, next obs, rewards, done)
    from maddpg_b200.rollout import HostRollout
    host = HostRollout(env, core)
    Kh = max(EP_LEN, (4 * args.e2e_steps) // EP_LEN * EP_LEN) if headline else 2 * EP_LEN

    def host_loop(n_steps):
        obs_n = host.reset()
        ep = 0
        for _ in range(n_steps):
            action_n, obs_n, rew_n, done_n = host.step(obs_n)     # train.py:112-120
            ep += 1
            if ep >= EP_LEN:                                        # train.py:127-129
                obs_n = host.reset()
                ep = 0
        return float(rew_n[0][0])                                   # the step's result is read on the host

    host_loop(EP_LEN)
    barrier()
    t0 = time.perf_counter()
    host_loop(Kh)
    barrier()
    e2e_s = max_over_ranks([time.perf_counter() - t0])[0]
    e2e_value = E * A * Kh * world / e2e_s
    h2d, d2h = host.h2d_bytes_per_step, host.d2h_bytes_per_step
    # e2e updates: MADDPGAgentTrainer.update() per agent through the reference-shaped API
    for tr in trainers:
        tr.max_replay_buffer_len = BATCH * EP_LEN
    import random
    random.seed(rank)
    Re = max(3, min(30, R))
    for j, tr in enumerate(trainers):
        tr.update(trainers, 100)
    barrier()
    t0 = time.perf_counter()
    out = None
    for r in range(Re):
        for tr in trainers:
            tr.preupdate()
        for tr in trainers:
            out = tr.update(trainers, 100)
            assert out is not None
    last_stats = [float(x) for x in out]  # the last update's statistics are read on the host (one D2H)
    barrier()
    e2e_upd_s = max_over_ranks([time.perf_counter() - t0])[0]

    # ---- (5) CPU baseline: the restated reference loop on this box's host cores (rank 0, N=1 only) -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[k] = "1"
        from oracle import train_loop as tl
        cpu_steps, cpu_rounds = (40000, 150) if headline else (4000, 10)  # ~10-15 s of single-core work
        a_sps, e_sps, dt = tl.time_rollout(SCENARIO, N_AGENTS, cpu_steps, arglist=tl.make_arglist(SCENARIO, BATCH, UNITS))
        u_ps, udt = tl.time_updates(SCENARIO, N_AGENTS, cpu_rounds, arglist=tl.make_arglist(SCENARIO, BATCH, UNITS))
        cpu = {"value": a_sps, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d env steps of the restated train.py loop (1 env, %.1f s) ; %d forced update rounds x %d agents "
                         "(%.1f s)" % (cpu_steps, dt, cpu_rounds, A, udt),
               "critic_updates_per_sec": u_ps, "host_cores_available": os.cpu_count()}

    # ---- (6) tensor-core TD-target kernel where the GEMMs are large enough to matter: BASELINE.json configs[4]'s
    #          per-GPU update shape (simple_spread N=24: 24 critics of input width 3576, 576 actor passes, batch 1024) ----------
    tensor = None
    if headline and rank == 0 and world == 1 and not args.no_tensor_section:
        del flush_buf
        torch.cuda.empty_cache()
        tensor = tensor_core_section(torch, dev)

    # ---- (7) SURVEY 8(f) rows: device prioritized replay (rank 4) and the other MPE scenarios (rank 2) --------------------
    extras = None
    if headline and rank == 0 and world == 1:
        extras = next_rows_section(torch, BATCH, not args.no_cpu_baseline)

    # dominant kernel of the timed rollout region: the persistent episode kernel
    row_bytes = 4 * sum(2 * d + k + 2 for d, k in zip(env.obs_dims, env.act_dims))
    launch_steps = min(K, EP_LEN * eps_launch) if roll.mode == "mega" else 1
    step_us = roll_ms * 1e3 / K
    ep_bytes = row_bytes * E * launch_steps     # the replay rows are the only HBM traffic the algorithm needs
    actor_flops_step = 2 * sum(d * UNITS + UNITS * UNITS + UNITS * k for d, k in zip(env.obs_dims, env.act_dims)) * E
    sm_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    fp32_peak_tflops = 148 * 113 * 2 * sm_mhz * 1e-6  # 113 FMA/clk/SM measured (tools/fp32_probe.cu), not the nominal 128
    try:
        bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        bf16 = 2250.0
    traffic, traffic_src = traffic_from_profile("k_rollout_episode_tc" if "_tc" in ep_kernel else "k_rollout_episode")
    ep_roof = {"kernel": ep_kernel, "bound": "hbm", "achieved": row_bytes * E / step_us / 1e3, "peak": peak, "unit": "GB/s",
               "frac": row_bytes * E / step_us / 1e3 / peak, "traffic": traffic, "traffic_source": traffic_src,
               "peak_source": peak_src, "algorithmic_bytes_per_launch": ep_bytes, "steps_per_launch": launch_steps,
               "avg_launch_us": step_us * launch_steps,
               "note": "not HBM-bound by construction: state, observations, actions and actor weights never leave the SM "
                       "(shared / tensor memory); the replay rows are the only algorithmic HBM traffic.  The step is a dependent "
                       "chain (MMA -> epilogue -> MMA -> epilogue -> head -> Gumbel-softmax -> World.step -> observation) over 32 "
                       "env instances per SM -- see the clock64 phase table in DESIGN.md",
               "actor_tflops": actor_flops_step / step_us / 1e6,
               "tensor": {"bound": "tensor", "achieved": 3 * actor_flops_step / step_us / 1e6, "peak": bf16 / 2, "unit": "TFLOP/s",
                          "frac": 3 * actor_flops_step / step_us / 1e6 / (bf16 / 2),
                          "note": "issued kind::tf32 MMA work (3 per product, incl. the 64 x 5 head that runs on CUDA cores: upper "
                                  "bound) over MEASURED_PEAKS.json bf16 / 2; tensor-pipe activity from ncu in profiles/"}
               if "_tc" in ep_kernel else None,
               "fp32_peak_tflops": fp32_peak_tflops}
    if rank == 0:
        flops_round = sum(int(core.layout.update_flops_critic[j]) + int(core.layout.update_flops_actor[j]) for j in range(A)) * BATCH
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": roll_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (env state and nets; actor GEMMs as 3xTF32 on tcgen05 with fp32 accumulate) -- `value`; "
                     "`value_f64_state` is the same rollout with the reference's float64 env state",
            "data": "synthetic",
            "value_f64_state": None if f64 is None else f64["value"],
            "f64_state": f64,
            "config": {"workload": WORKLOAD, "bench_config": args.config, "envs_per_gpu": E, "agents": A, "batch": BATCH,
                       "num_units": UNITS, "episode_len": EP_LEN,
                       "replay_capacity_rows": core.ring.capacity, "rollout_mode": roll.mode, "rollout_mode_probe_ms_per_step": probe,
                       "episodes_per_launch": eps_launch if roll.mode == "mega" else None,
                       "cuda_graph_updates": bool(gupd is not None and gupd.use_graph),
                       "timing": "one repetition = exactly `steps` lockstep steps (whole episodes incl. env.reset, + a partial "
                                 "episode), CUDA events, barrier + synchronize on both sides; %d repetitions, median reported, max "
                                 "over ranks per repetition" % reps,
                       "repetitions_ms": roll_all_ms,
                       "l2": "flushed (256 MB write) before every repetition / update round; inside a repetition the env state "
                             "never leaves the SM"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Kh, "ms_per_step": 1e3 * e2e_s / Kh,
                    "api": "maddpg_b200.rollout.HostRollout.step -> C ABI mdp_host_step_pipelined: train.py:112-120 (action, "
                           "env.step, experience for all agents) per call with page-locked host numpy buffers; env.reset every "
                           "25 steps; %d ranges of env instances on forked streams, host buffers moved by %s, the whole call "
                           "replayed as one CUDA graph per result slot (%d kernels per step)"
                           % (host.chunks, "copy kernels (SMs over the unified address space)" if host.copy_kernels
                              else "the copy engines", host.launches_per_graph),
                    "per_call_api": {"value": percall_value, "unit": UNIT, "steps": Ke, "ms_per_step": 1e3 * percall_s / Ke,
                                     "h2d_bytes_per_step": percall_h2d, "d2h_bytes_per_step": percall_d2h,
                                     "api": "MADDPGAgentTrainer.action / BatchedMultiAgentEnv.step / .experience, one call per "
                                            "agent with pageable host numpy arrays (the reference's call granularity)"}},
            "gpu_launches": int(launches_roll),
            "roofline": ep_roof,
            "roofline_env_step_kernel": {"kernel": env_kernel, "bound": "hbm", "achieved": achieved, "peak": peak,
                                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                                         "algorithmic_bytes_per_launch": env_bytes, "avg_launch_us": env_us,
                                         "note": "per-step kernel of the reference-shaped API at the config's size",
                                         "at_1M_envs": big},
            "critic_updates": {"value": upd_value, "unit": "critic updates/s", "rounds": R, "ms_per_round": upd_med_ms,
                               "gpu_launches": int(launches_upd), "flops_per_round": flops_round,
                               "achieved_tflops": flops_round / (upd_med_ms * 1e-3) / 1e12,
                               "allreduce_bytes_per_round": 0 if world == 1 else 4 * int(core.layout.total_train),
                               "gradient_exchange": exchange, "exchange_efficiency": noex,
                               "order": "sequential agents (reference order, parity mode); median over rounds",
                               "grouped": None if grp_value is None else {
                                   "value": grp_value, "unit": "critic updates/s", "ms_per_round": grp_ms,
                                   "order": "all agents per launch (Jacobi order; deviation documented in DESIGN.md)"},
                               "e2e": {"value": Re * A * world / e2e_upd_s, "unit": "critic updates/s", "last_stats": last_stats,
                                       "api": "MADDPGAgentTrainer.update for every agent, %d rounds; statistics stay on the device "
                                              "until read (the last round's are read on the host inside the timed region)" % Re}},
        }
        if tensor is not None:
            line["tensor_core_td_target"] = tensor
        if extras is not None:
            line["next_rows"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
