#!/usr/bin/env python
"""bench.py -- Connect4 6x7x4 random-rollout env-steps/sec (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--games G]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (legal-move generation -> uniform action choice -> transition
-> terminal / reward evaluation, looped to the end of every game) over one batch of 16 Mi synthetic
games per GPU (BASELINE.json configs[1]); each step uses fresh global game ids.  One env-step = one
ply of one game.  Rank 0 prints ONE JSON line.

  value      whole-job env-steps/s, device-timed (CUDA events on the launching stream), max over ranks
  e2e        the same metric through the public API (simulator.batch.HostRollout), per-game results
             and statistics copied device->host into pinned memory inside the timed region
  roofline   integer-issue roofline of connect_rollout_kernel (100 thread-level 32-bit
             integer instructions per env-step against SMs x 128 lanes x f_SM)
  cpu_baseline  the CPU oracle (a port -- the reference's engine cannot be built here) on all
             host cores, bounded sample, rank 0 at N=1 only

--impl reference times that CPU path alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PRODUCT = os.path.join(ROOT, "board-game-simulator-python_b200")
for _p in (ROOT, PRODUCT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CONFIG = (6, 7, 4)
GAMES_PER_GPU = 16 * 2**20
STRONG_GAMES = 64 * 2**20  # strong-scaling pass: this many games in total, whatever the number of GPUs
OPS_PER_STEP = 100  # algorithmic thread-level 32-bit integer instructions per env-step
BOUNCE_OPS_PER_STEP = 1400  # the Bounce budget fixed in round 1, kept for comparability
SM_MAX_MHZ_FALLBACK = 1965.0
METRIC = "connect4_6x7x4_random_rollout_env_steps_per_sec"
UNIT = "env-steps/s"
SEED = 20261018


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_rollout_throughput(budget_s: float, threads: int | None = None, chunk: int = 1 << 15, engine: str = "bitboard"):
    """Times the CPU port of the hot path on `threads` host threads for ~budget_s.

    engine "bitboard": oracle/fast_connect.c -- the same loop written the way a CPU engine would (single-word
    bitboards, shift-and-AND run test, -O3), checked game by game against the oracle; the "best CPU" figure of
    SURVEY.md 8d (iii).  engine "naive": oracle/bgs_oracle.c (bgso_connect_rollout), the brute-force checker.
    ctypes releases the GIL during the call, so plain threads scale across cores."""
    from oracle import binding as o  # the checker, used here only as the timed CPU baseline

    threads = threads or os.cpu_count() or 1
    o.lib()
    H, W, K = CONFIG
    if engine == "bitboard":
        run = lambda gid0: o.fast_connect_rollout(H, W, K, chunk, gid0=gid0, seed=SEED, per_game=False)
        what = "oracle/fast_connect.c (C bitboards, -O3; verified against oracle/bgs_oracle.c)"
    else:
        run = lambda gid0: o.connect_rollout(H, W, K, chunk, gid0=gid0, seed=SEED, want_actions=False, want_grid=False)
        what = "oracle/bgs_oracle.c (C, brute force, -O2)"
    run(0)  # warm up
    done = [0] * threads
    games = [0] * threads
    t_end = time.perf_counter() + budget_s

    def work(tid):
        i = 0
        while time.perf_counter() < t_end:
            res = run((tid * 1_000_003 + i) * chunk)
            done[tid] += int(res["stats"][o.STAT_STEPS])
            games[tid] += chunk
            i += 1

    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return {
        "value": sum(done) / dt,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"{sum(games)} games ({sum(done)} env-steps) of the same Connect(6,7,4) workload in {dt:.1f} s "
                  f"on {threads} threads, {what}",
        "seconds": dt,
        "steps": sum(done),
    }


def python_api_throughput(n_games: int = 300):
    """BASELINE.json configs[0]-style loop (README.md:52-69) through the oracle's object API, 1 thread."""
    import importlib
    import random

    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyapi"))
    saved = {k: v for k, v in sys.modules.items() if k == "simulator" or k.startswith("simulator.")}
    for k in saved:
        del sys.modules[k]
    try:
        connect = importlib.import_module("simulator.game.connect")
        random.seed(0)
        config = connect.Config(*CONFIG)
        steps = 0
        t0 = time.perf_counter()
        for _ in range(n_games):
            state = config.sample_initial_state()
            while not state.has_ended:
                state = random.choice(state.actions).sample_next_state()
                steps += 1
            _ = state.reward
        dt = time.perf_counter() - t0
    finally:
        sys.path.remove(os.path.join(ROOT, "oracle", "pyapi"))
        for k in [k for k in sys.modules if k == "simulator" or k.startswith("simulator.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return steps / dt


_PYTHON_API_WORKER = r"""
import os, random, sys, time
root, n_games, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sys.path.insert(0, os.path.join(root, "oracle", "pyapi"))
sys.path.insert(0, root)
from simulator.game import connect  # the oracle's object-API stand-in
random.seed(seed)
config = connect.Config(6, 7, 4)
steps = 0
t0 = time.perf_counter()
for _ in range(n_games):
    state = config.sample_initial_state()
    while not state.has_ended:
        state = random.choice(state.actions).sample_next_state()
        steps += 1
print(steps, time.perf_counter() - t0)
"""


def python_api_all_cores(n_games: int = 150):
    """The README loop (README.md:52-69) in os.cpu_count() processes with disjoint seeds -- SURVEY.md 8d (ii):
    the reference holds the GIL, so processes are its only way to use more cores.  Aggregate steps divided by
    the slowest worker's time."""
    import subprocess

    procs = os.cpu_count() or 1
    ps = [subprocess.Popen([sys.executable, "-c", _PYTHON_API_WORKER, ROOT, str(n_games), str(1000 + i)],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for i in range(procs)]
    res = []
    for pr in ps:
        out, _ = pr.communicate(timeout=300)
        st, sec = out.split()
        res.append((int(st), float(sec)))
    return sum(r[0] for r in res) / max(r[1] for r in res), procs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = args.ref_seconds or max(1.0, min(15.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_rollout_throughput(min(per_step, 1.0))
    vals, tot_steps, tot_s = [], 0, 0.0
    last = None
    for _ in range(args.steps):
        last = cpu_rollout_throughput(per_step)
        vals.append(last["value"])
        tot_steps += last["steps"]
        tot_s += last["seconds"]
    value = tot_steps / tot_s
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
            "sample": f"each step = a {per_step:.0f} s bounded sample of the workload on {last['cores']} host threads; "
                      "oracle/fast_connect.c (bitboard C loop, -O3, verified against the oracle) -- the reference's own "
                      "engine (jojolebarjos/board-game-simulator@c8f8a07) is not in the reference tree and cannot be "
                      "built offline",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


def workload_config(args):
    return {
        "workload": "Connect4 Config(6,7,4) uniform-random rollouts from the empty board to terminal, "
                    f"{args.games} concurrent games per GPU (BASELINE.json configs[1])",
        "games_per_gpu": args.games,
        "global_games_per_step": args.games * args.gpus,
        "seed": SEED,
        "outputs_per_game": "length u8 + winner i8 (+ int64[256] statistics per step)",
        "parallelism": f"{args.gpus} independent game-id shards; one NCCL all-reduce(sum) of the int64[256] statistics "
                       "vector per step, asynchronous (overlaps the next step's kernel), all waited for before the closing barrier",
        "l2": "the kernel reads no input tensor (the batch starts from the empty board; inputs are launch scalars), so "
              "there is nothing to keep warm in L2; the per-game outputs rotate over 8 buffer sets (8 x 32 MiB > 126 MB L2)",
    }


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "nvmlClocksThrottleReasonSwPowerCap": "sw_power_cap",
            "nvmlClocksThrottleReasonHwSlowdown": "hw_slowdown",
            "nvmlClocksThrottleReasonSwThermalSlowdown": "sw_thermal_slowdown",
            "nvmlClocksThrottleReasonHwThermalSlowdown": "hw_thermal_slowdown",
            "nvmlClocksThrottleReasonHwPowerBrakeSlowdown": "hw_power_brake_slowdown",
        }
        bits = {getattr(nv, k): v for k, v in names.items() if hasattr(nv, k)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.001)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = statistics.median(self.samples) if self.samples else None
        return {
            "sm_mhz": med, "sm_max_mhz": self.max_mhz or SM_MAX_MHZ_FALLBACK, "reasons": sorted(self.reasons),
            "samples": len(self.samples), "sm_mhz_min": min(self.samples) if self.samples else None,
        }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def other_configs(torch, batch, N, dev):
    """Device-timed figures for the other BASELINE.json configurations (configs[2], configs[3]) -- not the
    bench value, reported beside it: median of 5 launches after 2 warm-ups, CUDA events."""
    import statistics

    import numpy as np

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path)).get("hbm_gbs", 6444.4) if os.path.exists(peaks_path) else 6444.4
    out = {}

    def timed(fn, stats):
        ms, steps = [], []
        for i in range(7):
            stats.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            fn(i)
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(a.elapsed_time(b))
                steps.append(int(stats[N.STAT_STEPS]))
        med = statistics.median(ms)
        return med, statistics.mean(steps)

    stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)
    grid = np.zeros((9, 6), dtype=np.int8)
    grid[1] = grid[7] = [1, 2, 3, 3, 2, 1]  # reference src/simulator/textual/bounce.py:66-78
    n = 4 * 2**20
    ms, steps = timed(lambda i: batch.bounce_rollout(grid, n, SEED, i * n, max_plies=512, stats=stats), stats)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_g = sms * 128 * SM_MAX_MHZ_FALLBACK * 1e6 / 1e9
    out["bounce_default_9x6"] = {
        "games": n, "max_plies": 512, "ms_per_launch": ms, "env_steps_per_s": steps / ms * 1e3,
        "kernel": "bounce_rollout_slots_kernel<2, GeoCT<9,6>, 0, 64>",
        "roofline": {
            "bound": "int_issue", "algorithmic_ops_per_env_step": BOUNCE_OPS_PER_STEP,
            "achieved": steps / ms * 1e3 * BOUNCE_OPS_PER_STEP / 1e9, "peak": peak_g, "unit": "G thread-instr/s",
            "frac": steps / ms * 1e3 * BOUNCE_OPS_PER_STEP / 1e9 / peak_g,
            "peak_def": f"{sms} SMs x 4 schedulers x 32 lanes x {SM_MAX_MHZ_FALLBACK:.0f} MHz; budget of 1400 thread-level "
                        "integer instructions per env-step fixed in round 1 (DESIGN.md 4)",
            **bounce_counters(),
        },
    }
    try:  # the same kernel over 8 back-to-back batches alternating over two streams: the straggler tail of a batch
        # (a few ~400-ply games among 4 Mi leave the GPU nearly idle for ~1 ms) overlaps the next batch
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        bstats = torch.zeros((8, N.STATS_LEN), dtype=torch.int64, device=dev)
        for _rep in range(2):
            bstats.zero_()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            done = []
            for st in streams:
                st.wait_event(t0)
            for b in range(8):
                with torch.cuda.stream(streams[b % 2]):
                    batch.bounce_rollout(grid, n, SEED, (200 + b) * n, max_plies=512, stats=bstats[b])
                    e = torch.cuda.Event()
                    e.record()
                    done.append(e)
            for e in done:
                torch.cuda.current_stream().wait_event(e)
            t1.record()
            torch.cuda.synchronize()
        tot_ms, tot_steps = t0.elapsed_time(t1), int(bstats[:, N.STAT_STEPS].sum())
        out["bounce_default_9x6"]["two_streams"] = {
            "batches": 8, "ms_per_batch": tot_ms / 8, "env_steps_per_s": tot_steps / tot_ms * 1e3,
            "frac_of_budget": tot_steps / tot_ms * 1e3 * BOUNCE_OPS_PER_STEP / 1e9 / peak_g,
        }
    except Exception as e:
        out["bounce_default_9x6"]["two_streams"] = {"error": repr(e)}
    try:  # configs[2] end to end: per-game results (2 bytes, packed) + statistics to pinned host memory, pipelined
        bh = batch.HostRollout(grid, n, depth=2, packed=True, game="bounce", max_plies=512)
        for _ in bh.stream(SEED, 0, 2):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bsteps = 0
        for st, _ in bh.stream(SEED, 100 * n, 8):
            bsteps += int(st[N.STAT_STEPS])
        dt = time.perf_counter() - t0
        out["bounce_default_9x6"]["e2e"] = {
            "value": bsteps / dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": bh.d2h_bytes,
            "ms_per_step": 1e3 * dt / 8, "api": "simulator.batch.HostRollout(game='bounce', packed=True).stream",
            "note": "8 batches; the two buffer sets launch on two compute streams, so the straggler tail of a batch "
                    "(a few ~400-ply games) overlaps the next batch -- which is why this exceeds the single-launch figure",
        }
        del bh
    except Exception as e:
        out["bounce_default_9x6"]["e2e"] = {"error": repr(e)}
    for cfg, bytes_per_game in (((8, 9, 5), 154), ((10, 12, 6), 250)):
        res = [None]

        def plain(i, cfg=cfg):
            res[0] = batch.connect_rollout(cfg, n, SEED, i * n, per_game=True, stats=stats, out=res[0])

        ms0, steps0 = timed(plain, stats)
        res[0] = None

        def export(i, cfg=cfg):
            res[0] = batch.connect_rollout(cfg, n, SEED, i * n, per_game=True, actions=True, final_grid=True,
                                           reward=True, stats=stats, out=res[0])

        ms1, steps1 = timed(export, stats)
        out[f"connect_{cfg[0]}x{cfg[1]}x{cfg[2]}"] = {
            "games": n, "ms_per_launch": ms0, "env_steps_per_s": steps0 / ms0 * 1e3,
            "with_trajectory_grid_reward_export": {
                "ms_per_launch": ms1, "env_steps_per_s": steps1 / ms1 * 1e3, "bytes_per_game": bytes_per_game,
                "export_ms": ms1 - ms0, "export_GBps": bytes_per_game * n / (ms1 - ms0) / 1e6,
                "export_frac_of_hbm_copy_peak": bytes_per_game * n / (ms1 - ms0) / 1e6 / hbm,
            },
        }
        res[0] = None
        torch.cuda.empty_cache()
    return out


def bounce_counters():
    path = os.path.join(ROOT, "profiles", "bounce_kernel_counters.json")
    try:
        d = json.load(open(path))
        return {k: d[k] for k in ("measured_inst_per_step", "issue_active", "alu_pipe", "active_lanes", "source") if k in d}
    except Exception:
        return {}


def headline_counters():
    """Measured instruction mix of the dominant kernel from the committed ncu capture (profiles/): thread-level
    instructions executed per env-step, issue-slot and ALU-pipe utilisation.  The 100-ops budget of the roofline is
    an ALGORITHMIC figure; these say what the kernel physically executes."""
    path = os.path.join(ROOT, "profiles", "headline_kernel_counters.json")
    try:
        d = json.load(open(path))
        return {k: d[k] for k in ("measured_inst_per_step", "measured_thread_inst_per_step", "issue_active",
                                  "alu_pipe", "active_lanes", "source") if k in d}
    except Exception:
        return {"measured_inst_per_step": None, "issue_active": None, "alu_pipe": None}


def object_api_throughput(n_games: int = 12):
    """The README loop (README.md:52-69) through the PRODUCT's object API: every property is a CUDA kernel with
    a batch of one plus a synchronisation.  This is the drop-in's single-game cost -- far below the CPU
    reference it replaces on this path; the batched entry points are the product."""
    import random

    from simulator.game import connect

    random.seed(0)
    config = connect.Config(*CONFIG)
    steps = 0
    t0 = None
    for game in range(n_games + 1):
        if game == 1:  # the first game is a warm-up (CUDA context, module load, staging buffers)
            steps, t0 = 0, time.perf_counter()
        state = config.sample_initial_state()
        while not state.has_ended:
            state = random.choice(state.actions).sample_next_state()
            steps += 1
        _ = state.reward
    return steps / (time.perf_counter() - t0)


def run_b200(args):
    import torch
    import torch.distributed as dist

    from simulator import _native as N
    from simulator import batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
        args.gpus = world
    N.require_cuda()  # fails loudly without the CUDA library / a device: there is no fallback
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.games
    total = n * world
    dev = torch.device("cuda", local)
    ROT = 8  # per-game outputs rotate over 8 buffer sets (8 x 32 MiB > the 126 MB L2)
    results = [None] * ROT
    step_id = [0]

    def one_step(stats):
        """One pass of the hot path over this rank's shard of a fresh global batch.  The all-reduce of
        the step's statistics (the path's only collective) is launched asynchronously on NCCL's own
        stream, so it overlaps the next step's rollout kernel; the caller waits for it at the end."""
        slot = step_id[0] % ROT
        base = step_id[0] * total
        step_id[0] += 1
        start, count = batch.shard_range(total, rank, world)
        results[slot] = batch.connect_rollout(CONFIG, count, SEED, base + start, per_game=True, stats=stats,
                                              out=results[slot])
        if world > 1:
            return dist.all_reduce(stats, op=dist.ReduceOp.SUM, async_op=True)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # allocate every rotating output set before anything is timed (allocation is not part of a step)
    scratch_stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)
    for slot in range(ROT):
        results[slot] = batch.connect_rollout(CONFIG, n, SEED, 0, per_game=True, stats=scratch_stats, out=results[slot])
    warm_stats = torch.zeros((args.warmup, N.STATS_LEN), dtype=torch.int64, device=dev)
    for i in range(args.warmup):
        w = one_step(warm_stats[i])
        if w is not None:
            w.wait()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    step_stats = torch.zeros((args.steps, N.STATS_LEN), dtype=torch.int64, device=dev)
    works = []
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        works.append(one_step(step_stats[i]))
    for w in works:
        if w is not None:
            w.wait()  # the launching stream waits for every statistics all-reduce
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    elapsed = torch.tensor(ev0.elapsed_time(ev1) / 1e3, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed_s = float(elapsed)
    job_steps = int(step_stats[:, N.STAT_STEPS].sum())  # after the all-reduces: the whole job's env-steps
    value = job_steps / elapsed_s
    stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)

    # ---- dominant kernel alone (rank-local, no collective): roofline -----------------------------
    # K launches back to back inside ONE event bracket on the launching stream (no host synchronisation between
    # them, so launch latency is hidden behind the previous kernel exactly as in the timed region above).
    barrier()
    stats.zero_()
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ka.record()
    for i in range(args.steps):
        results[i % ROT] = batch.connect_rollout(CONFIG, n, SEED, (10_000 + i) * total + rank * n, per_game=True,
                                                 stats=stats, out=results[i % ROT])
    kb.record()
    torch.cuda.synchronize()
    kernel_s = ka.elapsed_time(kb) / 1e3
    ksteps = int(stats[N.STAT_STEPS])

    # ---- end to end through the public API, pinned host results --------------------------------
    # simulator.batch.HostRollout.stream: every batch's per-game results and statistics are copied
    # device->host into pinned memory inside the timed region as ONE copy per batch (the copy of batch i overlaps
    # the kernel of batch i+1); the loop consumes the host statistics of every batch.  The pinned buffers are
    # allocated on CPUs local to this rank's GPU.
    # per-game results in the dense code: 3 games per 16-bit word (37 outcomes each: 36 lengths with the winner
    # given by the parity, or a draw) = 5.33 bits per game
    host = batch.HostRollout(CONFIG, n, depth=3, packed="dense")
    for i in range(min(args.warmup, 3)):
        host.run(SEED, (20_000 + i) * total + rank * n)
    # the pipelined path is timed 3 times over K steps each (after one untimed pass that touches every
    # buffer set); the median repetition is reported and all three are listed -- a single 20 ms window of
    # wall-clock time on a shared host is too noisy on its own
    for _ in host.stream(SEED, 29_000 * total + rank * n * 4, 4):
        pass
    e2e_reps = []
    for rep in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_steps = 0
        for st, result_h in host.stream(SEED, (30_000 + 100 * rep) * total + rank * n * args.steps, args.steps):
            e2e_steps += int(st[N.STAT_STEPS])
        torch.cuda.synchronize()
        e2e_reps.append((time.perf_counter() - t0, e2e_steps))
    e2e_reps.sort(key=lambda ts: ts[1] / ts[0])
    e2e_t = torch.tensor(e2e_reps[1][0], dtype=torch.float64, device=dev)
    e2e_n = torch.tensor(e2e_reps[1][1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_n, op=dist.ReduceOp.SUM)
    e2e_value = int(e2e_n) / float(e2e_t)
    d2h_gbps_rank = host.d2h_bytes * args.steps / e2e_reps[1][0] / 1e9
    # the same path with separate length / winner arrays (2 bytes per game), pipelined and synchronous
    host2 = batch.HostRollout(CONFIG, n, depth=3)
    host2.run(SEED, 39_000 * total + rank * n)
    for _ in host2.stream(SEED, 39_500 * total + rank * n * 4, 4):  # touch every buffer set before timing
        pass
    barrier()
    t0 = time.perf_counter()
    un_steps = 0
    for st, _, _ in host2.stream(SEED, 40_000 * total + rank * n * args.steps, args.steps):
        un_steps += int(st[N.STAT_STEPS])
    e2e_unpacked_value = un_steps / (time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    sync_steps = 0
    for i in range(args.steps):
        st, _, _ = host2.run(SEED, (45_000 + i) * total + rank * n)
        sync_steps += int(st[N.STAT_STEPS])
    e2e_sync_value = sync_steps / (time.perf_counter() - t0)
    aux = torch.tensor([e2e_unpacked_value, e2e_sync_value], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(aux, op=dist.ReduceOp.MIN)  # the slowest rank bounds the job
    e2e_unpacked_value, e2e_sync_value = float(aux[0]), float(aux[1])
    del host2

    # ---- strong scaling + cross-rank determinism: a FIXED global id range is split
    # over the ranks (shard_range), the statistics are all-reduced, and rank 0 replays the whole range alone
    # (untimed) to check that N ranks give the single-rank answer for the same global ids.
    strong_total = STRONG_GAMES
    s_start, s_count = batch.shard_range(strong_total, rank, world)
    sres = None
    sstats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)
    sres = batch.connect_rollout(CONFIG, s_count, SEED, 60_000 * GAMES_PER_GPU + s_start, per_game=True, stats=sstats, out=sres)
    barrier()
    strong_reps = []
    for rep in range(3):
        sstats.zero_()
        sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sa.record()
        sres = batch.connect_rollout(CONFIG, s_count, SEED, 60_000 * GAMES_PER_GPU + s_start, per_game=True, stats=sstats,
                                     out=sres)
        if world > 1:
            dist.all_reduce(sstats, op=dist.ReduceOp.SUM)
        sb.record()
        torch.cuda.synchronize()
        st_ = torch.tensor(sa.elapsed_time(sb) / 1e3, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(st_, op=dist.ReduceOp.MAX)
        strong_reps.append(float(st_))
    strong_s = sorted(strong_reps)[1]
    strong_steps = int(sstats[N.STAT_STEPS])
    stats_equal = None
    if rank == 0:
        single = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)
        del sres
        batch.connect_rollout(CONFIG, strong_total, SEED, 60_000 * GAMES_PER_GPU, per_game=False, stats=single)
        torch.cuda.synchronize()
        stats_equal = bool(torch.equal(single, sstats))

    # ---- end to end with REAL tensor inputs: rollouts that continue from caller-supplied positions
    # (leaf evaluation of a tree search), simulator.batch.HostLeafRollout.  Every step copies 4 Mi positions
    # host->device from pinned memory as packed records (two bitboards + a meta byte = 17 B each, instead of the
    # 44 B of an int8 grid + player + winner) and the packed results (1 B per game) + statistics device->host;
    # three batches are in flight (H2D of batch i+1, kernel of batch i, D2H of batch i-1 overlap).
    fp = None
    if world == 1:
        try:
            n_pos = 4 * 2**20
            b0 = batch.ConnectBatch.initial(CONFIG, n_pos)
            gen = torch.Generator(device="cuda").manual_seed(1)
            for _ in range(10):  # 10 random plies: mid-game positions
                b0, _ = b0.step(torch.randint(0, CONFIG[1], (n_pos,), device="cuda", generator=gen))
            pk = b0.pack()
            leaf = batch.HostLeafRollout(CONFIG, n_pos, depth=3)
            for slot in range(leaf.depth):
                hp, hm = leaf.host_inputs(slot)
                hp.copy_(pk.packed.cpu())
                hm.copy_(pk.meta.cpu())
            del b0, pk
            reps = max(6, min(args.steps, 20))
            fsteps, t0, tickets = 0, None, []
            for i in range(reps + 3):
                if i == 3:
                    for tk in tickets:
                        leaf.result(tk)
                    tickets = []
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fsteps = 0
                if len(tickets) == leaf.depth:
                    st, _ = leaf.result(tickets.pop(0))
                    fsteps += int(st[N.STAT_STEPS])
                tickets.append(leaf.submit(SEED, 50_000 * total + i * n_pos))
            for tk in tickets:
                st, _ = leaf.result(tk)
                fsteps += int(st[N.STAT_STEPS])
            dt = time.perf_counter() - t0
            fp = {
                "value": fsteps / dt, "unit": UNIT, "positions_per_step": n_pos,
                "h2d_bytes_per_step": leaf.h2d_bytes, "d2h_bytes_per_step": leaf.d2h_bytes,
                "ms_per_step": 1e3 * dt / reps, "h2d_GBps": leaf.h2d_bytes * reps / dt / 1e9,
                "api": "simulator.batch.HostLeafRollout.submit/result -> bgs_connect_rollout_from_packed + "
                       "bgs_connect_pack_results, 3 batches in flight; positions as ConnectBatch.pack() records "
                       "(17 B each; round 1 sent int8 grids, 44 B each: 1.16e10)",
            }
            del leaf
        except Exception as e:  # an auxiliary figure must never break the bench line
            fp = {"error": repr(e)}

    if rank == 0:
        props = torch.cuda.get_device_properties(local)
        sms = props.multi_processor_count
        f_mhz = (clocks or {}).get("sm_mhz") or SM_MAX_MHZ_FALLBACK
        peak = sms * 128 * f_mhz * 1e6 / 1e9  # G thread-instr/s
        achieved = (ksteps / kernel_s) * OPS_PER_STEP / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("connect_rollout_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": workload_config(args),
            "env_steps_per_step": job_steps / args.steps,
            "wall_s_timed_region": t_wall,
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clocks else None,
            "clock_samples": clocks["samples"] if clocks else 0,
            "e2e": {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes * world,
                "d2h_bytes_per_step": host.d2h_bytes * world,
                "api": "simulator.batch.HostRollout(packed='dense').stream -> bgs_connect_rollout + bgs_connect_pack_results_dense; "
                       "every batch's per-game results (length and winner of every game, 3 games per 16-bit word in base 37) "
                       "and statistics copied to pinned host memory as ONE copy per batch (pinned pages allocated on CPUs local to the GPU), the copy of "
                       "batch i overlapping the kernels of the next batches; median of 3 repetitions of K steps (wall "
                       "clock, max over ranks)",
                "repetitions_this_rank": [st_ / t_ for t_, st_ in e2e_reps],
                "d2h_GBps_this_rank": d2h_gbps_rank,
                "pinned_memory_numa_local_cpus": host.numa_cpus,
                "two_arrays_value": e2e_unpacked_value * world,
                "synchronous_call_value": e2e_sync_value * world,
                "from_positions": fp,
            },
            "strong_scaling": {
                "global_games": strong_total, "games_this_rank": s_count, "seconds": strong_s,
                "strong_scaling_value": strong_steps / strong_s, "unit": UNIT,
                "stats_equal_to_single_rank": stats_equal,
                "note": "a fixed range of global game ids split over the ranks (contiguous shards), statistics "
                        "all-reduced inside the timed bracket; rank 0 then replays the whole range alone and compares "
                        "the 256 statistics words (SURVEY.md 8d C5, 4.5 T4); median of 3, max over ranks",
            },
            "strong_scaling_value": strong_steps / strong_s,
            "stats_equal_to_single_rank": stats_equal,
            "gpu_launches": args.steps,
            "gpu_launches_note": "1 connect_rollout_lut_kernel per step in each timed region (value, kernel-only, e2e; "
                             "the e2e region adds 1 pack_results_dense_kernel per step)",
            "roofline": {
                "bound": "int_issue", "achieved": achieved, "peak": peak, "unit": "G thread-instr/s",
                "frac": achieved / peak, "traffic": traffic,
                "kernel": "connect_rollout_lut_kernel<6,7,4,false,false>",
                "kernel_ms_per_launch": 1e3 * kernel_s / args.steps,
                "kernel_env_steps_per_s": ksteps / kernel_s,
                "algorithmic_ops_per_env_step": OPS_PER_STEP,
                "peak_def": f"{sms} SMs x 4 schedulers x 32 lanes x {f_mhz:.0f} MHz (median SM clock sampled by NVML "
                            "during the timed region); the path is register-resident, HBM is not the bound",
                "timing": "K launches back to back in one CUDA-event bracket on the launching stream",
                **headline_counters(),
                "hbm_write_GBps": (2 * n + 2048) * args.steps / kernel_s / 1e9,
                "hbm_peak_GBps_measured": (json.load(open(peaks_path)).get("hbm_gbs") if os.path.exists(peaks_path) else 6650.0),
            },
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_rollout_throughput(args.cpu_seconds)
            cb.pop("seconds"), cb.pop("steps")
            try:
                cb["naive_oracle_value"] = cpu_rollout_throughput(min(4.0, args.cpu_seconds), engine="naive")["value"]
                cb["naive_oracle_note"] = "oracle/bgs_oracle.c (brute-force checker), same threads -- round 1's yardstick"
            except Exception as e:
                cb["naive_oracle_error"] = repr(e)
            try:
                cb["object_api_steps_per_s"] = object_api_throughput()
                cb["object_api_note"] = ("README.md:52-69 loop through simulator.game.connect ON THE GPU (one kernel + one "
                                         "synchronisation per property, batch of one): the drop-in's single-game cost")
            except Exception as e:
                cb["object_api_error"] = repr(e)
            try:
                cb["python_api_1thread_steps_per_s"] = python_api_throughput(200)
                v, procs = python_api_all_cores(150)
                cb["python_api_all_cores_steps_per_s"] = v
                cb["python_api_processes"] = procs
            except Exception as e:  # never let the yardstick break the bench line
                cb["python_api_error"] = repr(e)
            line["cpu_baseline"] = cb
        if world == 1 and not args.no_other:
            try:
                line["other_configs"] = other_configs(torch, batch, N, dev)
            except Exception as e:  # auxiliary figures must never break the bench line
                line["other_configs"] = {"error": repr(e)}
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner to stdout) get
    stderr instead.  Returns a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="concurrent games per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs figures (Bounce, larger boards)")
    ap.add_argument("--ref-seconds", type=float, default=0.0,
                    help="--impl reference: seconds of CPU work per step (default: sized so the run takes ~1 min)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.out = _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
