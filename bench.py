#!/usr/bin/env python
"""bench.py — env-steps/s including observation materialisation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--num-envs E]

A "step" is one lockstep step of the whole batch: every environment advances one turn and its
observation (u8[3,11,11] + food, role, status), reward and done are written to HBM. Workload at
N = 1: BASELINE.json configs[1] — 4,096 lockstep default-grid v1 environments with fused
observation output (weak scaling: 4,096 envs per GPU, global env ids, no data-path collective).

Printed JSON (rank 0, one line):
  value        device-resident throughput: actions for all K steps already in HBM, K steps executed
               by the multi-step kernel in launches of <= --fuse steps, every step's outputs written.
  per_call     the same K steps as K single-step launches replayed from one CUDA graph.
  e2e          K steps (at most 2,048) through the C-ABI host-buffer call (wab_vec_step_host_packed): pinned
               host actions in, every output in pinned host memory when the call returns (up to 16,384 envs
               the kernel streams them there itself, above that H2D + kernel + D2H), stream sync per step.
  roofline     dominant kernel (wab_step_kernel) vs the measured HBM copy bandwidth.
  cpu_baseline the oracle's C restatement of the reference step on all host cores (bounded sample), plus the
               same on one thread.
--impl reference times that C restatement alone (the reference itself is pure Python on pandas, ~16 steps/s per
core, and its sources do not travel to the GPU box); rank 0 only, at least 512 lockstep steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

B_ALG = 436  # algorithmic bytes per env-step, SURVEY.md §8(d): 363 obs + 8 scalars + 1 action + 2x32 state
METRIC = "env-steps/sec incl. obs"
WORKLOAD = "configs[1]: wab_env v1 default 11x11 viewport, 4096 lockstep envs per GPU, fused u8 observation output"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4096)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--fuse", type=int, default=256, help="steps per launch of the multi-step kernel")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--skip-e2e", action="store_true", help="sweeps only: skip the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="sweeps only: skip the cpu_baseline leg")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


_REAL_STDOUT = []     # fd of the process's real stdout once C-level stdout has been pointed at stderr


def emit(text):
    if _REAL_STDOUT:
        sys.stdout.flush()
        os.write(_REAL_STDOUT[0], (text + "\n").encode())
    else:
        print(text, flush=True)


def cpu_reference_run(n_envs, steps, warmup, seed, threads=0, budget_s=None):
    """The oracle's C restatement of the reference step on the host cores (the reference itself is pure
    Python on pandas, ~14-18 steps/s per core in the build container, and cannot travel to this box)."""
    import numpy as np
    from oracle import wab_oracle
    rng = np.random.default_rng(12345)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if threads <= 0:
        threads = cores          # explicit: torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm
    if budget_s is not None:  # bounded sample: size the run from two probes (threads warm on the second)
        rate = 1.0
        for probe_steps in (16, 128):
            a = rng.integers(0, 5, (probe_steps, n_envs)).astype(np.uint8)
            t0 = time.perf_counter()
            wab_oracle.run(None, seed, n_envs, probe_steps, a, threads)
            rate = n_envs * probe_steps / max(time.perf_counter() - t0, 1e-6)
        steps = int(max(8, min(steps, budget_s * rate / n_envs)))
        warmup = 0
    if warmup > 0:
        wab_oracle.run(None, seed, n_envs, warmup, rng.integers(0, 5, (warmup, n_envs)).astype(np.uint8), threads)
    acts = rng.integers(0, 5, (steps, n_envs)).astype(np.uint8)
    t0 = time.perf_counter()
    done_steps, checksum = wab_oracle.run(None, seed, n_envs, steps, acts, threads)
    dt = time.perf_counter() - t0
    return {"value": done_steps / dt, "seconds": dt, "steps": steps, "n_envs": n_envs, "cores": threads,
            "checksum": checksum}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget_env_steps = 2.0e8
    n_envs = args.num_envs
    # A timed run of only a few lockstep steps would mostly measure creating and resetting the envs and waking the
    # OpenMP threads, i.e. understate the CPU arm: time at least 512 lockstep steps after at least 16 warm-up steps.
    steps = max(args.steps, 512)
    warmup = max(args.warmup, 16)
    if n_envs * (steps + warmup) > budget_env_steps:
        steps = max(1, int(budget_env_steps / n_envs) - warmup)
    res = cpu_reference_run(n_envs, steps, warmup, args.seed)
    sample = "%d envs x %d lockstep steps (reset on done) after %d warm-up steps, actions uniform 0-4, obs materialised every step" % (
        n_envs, res["steps"], warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "timed_lockstep_steps": res["steps"],
        "ms_per_step": 1e3 * res["seconds"] / res["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 food / integer rules (CPU)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "num_envs_per_gpu": n_envs},
        "cpu_baseline": {"value": res["value"], "unit": "env-steps/s", "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": res["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle C port of wab_env.py step+obs with OpenMP over envs; the pandas reference itself: 16.6 steps/s single process, 125 steps/s over 8 cores in the build container (profiles/r1_reference_cpu_timing.json)",
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.sharding import reduce_stats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner on fd 1 (NCCL_DEBUG=VERSION may come from
        # the environment or from a conf file), so C-level stdout goes to stderr and the line is written to the real one
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        sys.stdout.flush()
        _REAL_STDOUT.append(os.dup(1))
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    n, K, W, T = args.num_envs, args.steps, max(args.warmup, 3), max(1, min(args.fuse, args.steps))
    env = VecEnv(n, seed=args.seed, device=dev, env_id_base=rank * n)
    env_lpe = env.lanes_per_env
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    actions = torch.randint(0, env.n_actions, (K + W, n), dtype=torch.uint8, device=dev, generator=gen)
    env.reset()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- device-resident, multi-step kernel (value) ----------------
    out = env._alloc(T)
    chunks = [(s, min(T, K - s)) for s in range(0, K, T)]
    for s in range(0, W, T):
        env.step_many(actions[K + s:K + min(W, s + T)], out={k: v[:min(T, W - s)] for k, v in out.items()})
    if world > 1:
        reduce_stats(env.stats_tensor())      # warm-up of the one collective too (NCCL sets its channels up on first use)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in chunks]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        try:                                  # keep the GPU busy ~0.2 ms while the host enqueues the timed launches, so
            torch.cuda._sleep(400_000)        # that a short run (small K) times the K steps and not the enqueue gap
        except Exception:
            pass
        start.record()
        for (s, c), (e0, e1) in zip(chunks, ev):
            e0.record()
            env.step_many(actions[s:s + c], out={k: v[:c] for k, v in out.items()})
            e1.record()
        stats_local = env.stats_tensor()
        stats_all = reduce_stats(stats_local) if world > 1 else None   # the only collective: 64 bytes
        stop.record()
        barrier()
        fused_ms = max_over_ranks(start.elapsed_time(stop))
        kernel_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        if fused_ms < 400.0:   # keep the GPU busy long enough for >= 2 clock samples
            t_end = time.perf_counter() + 0.6
            while time.perf_counter() < t_end:
                env.step_many(actions[:chunks[0][1]], out={k: v[:chunks[0][1]] for k, v in out.items()})
            torch.cuda.synchronize(dev)
    launches = len(chunks)
    del out

    # ---------------- device-resident, one launch per step from a CUDA graph (per_call) ----------------
    G = 64 if K >= 64 else K
    env.reset()
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for t in range(3):
            env.step(actions[K + t % W])
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for t in range(G):
                env.step(actions[t])
    torch.cuda.synchronize(dev)
    reps = max(1, K // G)
    barrier()
    start.record()
    for _ in range(reps):
        graph.replay()
    stop.record()
    barrier()
    percall_ms = max_over_ranks(start.elapsed_time(stop))
    percall = {"value": world * n * reps * G / (percall_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": percall_ms / (reps * G),
               "mode": "1 launch per step, %d-step CUDA graph replayed %d times (actions repeat per replay)" % (G, reps)}

    # ---------------- end to end through the host-buffer C-ABI call (e2e) ----------------
    if args.skip_e2e:
        return finish(args, env, world, rank, n, K, W, T, fused_ms, kernel_ms, launches, clocks, percall, None, stats_all, dist)
    hb = env.alloc_host_buffers(pinned=True)
    host_actions = actions[:min(K, 2048)].cpu().pin_memory()
    Ke = host_actions.shape[0]
    env.reset_host(hb)
    acts_np, hnp = host_actions.numpy(), hb["np"]     # numpy views of the pinned buffers: no per-step torch dispatch
    for t in range(min(W, Ke)):
        hnp["actions"][:] = acts_np[t]
        env.step_host(hb)
    barrier()
    t0 = time.perf_counter()
    acc = 0.0
    for t in range(Ke):
        hnp["actions"][:] = acts_np[t]           # this step's inputs, host memory
        env.step_host(hb)
        acc += float(hnp["reward"][0])           # the step's result is read on the host
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    h2d = n
    d2h = n * (363 + 1 + 1 + 1 + 4 + 1 + 1)
    e2e = {"value": world * n * Ke / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": e2e_ms / Ke,
           "path": ("wab_vec_step_host_packed, <= 16,384 envs: wab_step_kernel reads the pinned host actions and streams grids/food/"
                    "role/status/reward/done/info straight into the pinned host block over PCIe -> stream sync" if n <= 16384 else
                    "wab_vec_step_host_packed: pinned host actions -> H2D -> wab_step_kernel -> one D2H of grids/food/role/status/"
                    "reward/done/info into a pinned block -> stream sync")}

    return finish(args, env, world, rank, n, K, W, T, fused_ms, kernel_ms, launches, clocks, percall, e2e, stats_all, dist)


def finish(args, env, world, rank, n, K, W, T, fused_ms, kernel_ms, launches, clocks, percall, e2e, stats_all, dist):
    stats = env.stats()
    env_lpe = env.lanes_per_env
    env.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_s = (kernel_ms * 1e-3) / launches
        steps_per_launch = K / launches
        achieved = B_ALG * n * steps_per_launch / per_launch_s / 1e9
        line = {
            "metric": METRIC, "value": world * n * K / (fused_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": fused_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 grids / int32 rules / u32 Philox (int food counter proven == f64)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "num_envs_per_gpu": n, "global_envs": world * n, "parallelism": "dp%d" % world,
                       "mode": "wab_vec_step_many, %d steps per launch, every step's obs/reward/done written" % T,
                       "l2": "each launch writes %d MB of distinct output (> L2); state lives in registers, no reuse between steps"
                             % int(T * n * 372 / 1e6),
                       "actions": "uniform 0-4, pre-generated u8[K,N] in HBM"},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "per_call": percall,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (337.2e6 if (n == 4096 and T == 256) else None),   # ncu dram read+write per launch, profiles/r1e_ncu_4096_summary.txt
                         "kernel": "wab_step_kernel<false, LPE=%d>" % env_lpe, "peak_source": peak_src,
                         "bytes_per_env_step": B_ALG, "env_steps_per_launch": n * steps_per_launch,
                         "avg_launch_ms": per_launch_s * 1e3},
            "episode_stats": stats_all or stats,
        }
        try:
            if args.skip_cpu:
                raise RuntimeError("skipped (--skip-cpu)")
            cb = cpu_reference_run(n, 10 ** 9, 4, args.seed, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "env-steps/s", "cores": cb["cores"], "kind": "port",
                                    "sample": "%d envs x %d lockstep steps, %.1f s of %d-thread CPU work" % (
                                        cb["n_envs"], cb["steps"], cb["seconds"], cb["cores"])}
            one = cpu_reference_run(n, 10 ** 9, 4, args.seed, threads=1, budget_s=min(3.0, args.cpu_seconds))
            line["cpu_baseline"]["single_thread"] = {"value": one["value"], "unit": "env-steps/s", "cores": 1,
                                                     "sample": "%d envs x %d lockstep steps, %.1f s" % (one["n_envs"], one["steps"], one["seconds"])}
            line["cpu_baseline"]["pandas_reference"] = ("the pandas reference itself, timed in the build container: 16.6 steps/s "
                                                        "single process, 125 steps/s over 8 cores (profiles/r1_reference_cpu_timing.json)")
        except Exception as exc:  # the oracle is test infrastructure; its absence must not hide the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "port", "sample": "failed: %s" % exc}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
