#!/usr/bin/env python
"""bench.py -- the hot path of DeltaQ's `dq bsdiff` on B200: ISuffixSort.Sort(old) + Diff.Search at every
scan position of new, on BASELINE.json's config #2 (synthetic 16 MiB -> 17 MiB executable-like pair).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one (old, new) pair:
  value  : inputs and outputs resident in HBM (dq_cuda_suffix_sort_device + dq_cuda_bsdiff_search_device)
  e2e    : dq_cuda_bsdiff_streams through the C ABI with pinned HOST buffers: H2D of old and new, sort,
           search, D2H of the (pos, len) table, and the reference's greedy scan/emit loop on the host,
           returning the uncompressed ctrl/diff/extra streams (everything of Diff.Create except bzip2).
Unit: MB/s = 1e6 bytes of `new` per second; whole-job aggregate over ranks.
N > 1: every rank diffs its own pair (independent objects, no data-path collective) -> weak scaling.

--impl reference times the CPU restatement of the reference's path (oracle/: SA-IS sort + Diff.Create's
loop with its inline Search) on one host core, on a bounded sample of the same recipe.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bsdiff Diff.Create hot path (ISuffixSort.Sort + Diff.Search at every scan position) input MB/s"
UNIT = "MB/s"
WORKLOAD = "C2: synthetic 16 MiB -> 17 MiB executable-like pair, ~13% mutated regions (workloads.c2_exe_pair)"
SAMPLE_OLD = 2 << 20
SAMPLE_NEW = (2 << 20) + (1 << 17)


def pass_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one round-0 onesweep launch from the committed ncu --set full
    capture (profiles/r01_onesweep_pass_full_v2.md); per launch, like `achieved`."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_pass_traffic.json")) as f:
            return float(json.load(f)["traffic_bytes_round0_launch"])
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank):
    """CPU arm: the oracle's restatement of the reference's path, one core, bounded sample."""
    if rank != 0:
        return
    import oracle
    from deltaq_b200 import workloads as w
    oracle.build()
    old, new = w.c2_exe_pair(SAMPLE_OLD, SAMPLE_NEW)

    def step():
        sa = oracle.sais(old)
        oracle.bsdiff_streams(old, new, oracle.make_I(sa))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = new.size / dt / 1e6
    sample = (f"C2 recipe at 1/8 scale ({old.size} -> {new.size} bytes); per step: SA-IS sort of old "
              "(oracle/sais.c ~ SAIS.cs) + Diff.Create loop with inline Search (oracle/bsdiff.c ~ Diff.cs:92-298), "
              "no bzip2; gcc -O2, single thread as the reference is single-threaded")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline_once():
    import oracle
    from deltaq_b200 import workloads as w
    oracle.build()
    old, new = w.c2_exe_pair(SAMPLE_OLD, SAMPLE_NEW)
    t0 = time.perf_counter()
    reps = 0
    sort_s = 0.0
    while reps < 2 or (time.perf_counter() - t0 < 10 and reps < 6):
        t1 = time.perf_counter()
        sa = oracle.sais(old)
        sort_s += time.perf_counter() - t1
        oracle.bsdiff_streams(old, new, oracle.make_I(sa))
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": new.size / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sort_only_MBps": old.size / (sort_s / reps) / 1e6,
            "sample": f"C2 recipe at 1/8 scale ({old.size} -> {new.size} bytes), {reps} reps; oracle SA-IS sort + "
                      "Diff.Create loop with inline Search, no bzip2, gcc -O2, 1 thread"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="deltaq_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from deltaq_b200 import CudaSuffixSort, workloads as w

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deltaq_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank diffs its own pair (same recipe, rank-specific seeds)
    old, new = w.c2_exe_pair(seed_old=1 + 1000 * rank, seed_new=2 + 1000 * rank)
    n, m = int(old.size), int(new.size)
    sorter = CudaSuffixSort(device=local_rank)
    ctx = sorter.context

    # ---- device-resident arm -----------------------------------------------------------------------
    d_old = torch.from_numpy(old).cuda()
    d_new = torch.from_numpy(new).cuda()
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    d_pos = torch.empty(m, dtype=torch.int32, device="cuda")
    d_len = torch.empty(m, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    acc = {"launches": 0, "device_ms": 0.0, "pass_ms": 0.0, "pass_pairs": 0, "search_ms": 0.0, "passes": 0,
           "rounds": 0, "alg_bytes": 0}

    def step_device(record):
        ctx.suffix_sort_device(d_old.data_ptr(), n, d_sa.data_ptr())
        ctx.bsdiff_search_device(d_old.data_ptr(), n, None, d_new.data_ptr(), m, 0, m, d_pos.data_ptr(), d_len.data_ptr())
        if record:
            st = ctx.stats()
            acc["launches"] += st["kernel_launches"]
            acc["device_ms"] += st["device_ms"]
            acc["search_ms"] += st["search_ms"]
            acc["pass_ms"] += st["pass_ms"]
            acc["pass_pairs"] += st["pass_pairs"]
            acc["passes"] += st["radix_passes"]
            acc["rounds"] = st["rounds"]
            acc["alg_bytes"] = st["algorithmic_bytes"]

    ctx.set_timing(True)
    for _ in range(args.warmup):
        step_device(False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device(True)
    barrier()
    dt = time.perf_counter() - t0

    # ---- end-to-end arm: host buffers through the C ABI ------------------------------------------------
    ctx.set_timing(False)
    p_old = ctx.pinned(n, np.uint8)
    p_new = ctx.pinned(m, np.uint8)
    p_old.array[:] = old
    p_new.array[:] = new
    streams = None
    for _ in range(args.warmup):
        streams = ctx.bsdiff_streams(p_old.array, p_new.array, copy=False)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        streams = ctx.bsdiff_streams(p_old.array, p_new.array, copy=False)   # the C ABI's own result: pointers + lengths
    barrier()
    dt_e2e = time.perf_counter() - t1
    e2e_stats = ctx.stats()
    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (device arm + e2e arm)

    # max over ranks
    if world > 1:
        tt = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(tt[0]), float(tt[1])

    if rank == 0:
        ms_step = dt / args.steps * 1e3
        value = world * m / (dt / args.steps) / 1e6
        e2e_value = world * m / (dt_e2e / args.steps) / 1e6
        peak, peak_src = measured_peak()
        pass_gbs = acc["pass_pairs"] * 24 / (acc["pass_ms"] * 1e-3) / 1e9 if acc["pass_ms"] > 0 else None
        cpu = None if args.no_cpu_baseline else cpu_baseline_once()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "old_bytes": n, "new_bytes": m, "pairs_per_step": world,
                       "l2": "working set (>= 24 B x 16.7 M pairs per radix pass, 400 MB) exceeds the 126 MB L2; no flush",
                       "parallelism": f"{world} x independent pairs (one process and context per GPU)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": dt_e2e / args.steps * 1e3,
                    "h2d_bytes_per_step": n + m,
                    # the (pos,len) table crosses PCIe coded: 1 B/position + 12 B/match head + 8 B per 1024 positions
                    "d2h_bytes_per_step": (8 * m if e2e_stats["table_fallbacks"] else
                                           m + 12 * e2e_stats["table_heads"] + 8 * ((m + 1023) // 1024)),
                    "includes": "H2D old+new, sort, search, D2H of the coded (pos,len) table in slices, host greedy "
                                "scan/extend/emit loop on host threads overlapped with the slices; result = "
                                "context-owned ctrl/diff/extra buffers; streams "
                                f"ctrl/diff/extra = {len(streams['ctrl'])}/{len(streams['diff'])}/{len(streams['extra'])} B"},
            "gpu_launches": acc["launches"],
            "roofline": {"bound": "hbm", "kernel": "dq::radix::onesweep_pass_kernel",
                         "achieved": pass_gbs, "peak": peak, "unit": "GB/s",
                         "frac": (pass_gbs / peak) if pass_gbs else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_pair": 24, "launches_timed": acc["passes"],
                         "share_of_device_time": acc["pass_ms"] / (acc["device_ms"] + acc["search_ms"])
                         if acc["device_ms"] else None,
                         "traffic": pass_traffic(),
                         "traffic_note": "ncu capture of a round-0 launch (16,777,216 pairs, 402,653,184 algorithmic B); "
                                         "achieved averages all launches of the step (round 0 and the smaller doubling rounds)"},
            "device_ms_per_step": {"sort": acc["device_ms"] / args.steps, "search": acc["search_ms"] / args.steps},
            "sort": {"rounds": acc["rounds"], "algorithmic_bytes": acc["alg_bytes"],
                     "input_MBps_device": n / (acc["device_ms"] / args.steps * 1e-3) / 1e6 if acc["device_ms"] else None,
                     "algorithmic_GBps": acc["alg_bytes"] / (acc["device_ms"] / args.steps * 1e-3) / 1e9
                     if acc["device_ms"] else None},
            "clocks": clocks,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
