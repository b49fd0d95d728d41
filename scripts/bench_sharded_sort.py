"""Suffix sort of ONE text sharded over the ranks of a torchrun job (BASELINE configs #4/#5 shape).

    python -m torch.distributed.run --nproc-per-node N scripts/bench_sharded_sort.py c4 512   # MiB
Prints one JSON line on rank 0: input MB/s through deltaq_b200.parallel.suffix_sort_sharded (host text in,
per-rank SA buckets out; H2D of the text included), max over ranks, plus an O(n) sufcheck of the gathered SA.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402
from deltaq_b200.parallel import suffix_sort_sharded  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
check = (sys.argv[4] != "nocheck") if len(sys.argv) > 4 else True
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = mib << 20
t = {"c4": lambda: w.c4_genome(n), "uniform": lambda: w.c1_uniform(n, 7), "c3": lambda: w.c3_repetitive(n),
     "c2": lambda: w.c2_exe_pair(n, n + 1)[0]}[kind]()
sorter = CudaSuffixSort(device=local)
pin = sorter.context.pinned(t.size, np.int32)   # SA (N=1) or this rank's bucket (N>1) lands in pinned memory
times = []
prof = {}
for it in range(reps + 1):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        prof = {}
        base, mine = suffix_sort_sharded(t, sorter, gather=False, profile=prof, out=pin.array)
    else:
        sorter.context.suffix_sort(t, pin.array)
        mine = pin.array
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if it > 0:
        times.append(float(dt))
ok = None
if check:
    sa = suffix_sort_sharded(t, sorter) if world > 1 else np.asarray(mine)
    if rank == 0:
        ok = int(oracle.sufcheck(t, sa))
if rank == 0:
    best = min(times)
    print(json.dumps({"workload": kind, "n": int(t.size), "n_gpus": world, "best_ms": best * 1e3,
                      "input_MBps": t.size / best / 1e6, "all_ms": [x * 1e3 for x in times], "sufcheck": ok,
                      "phases_ms_rank0": {k: round(v * 1e3, 2) for k, v in prof.items() if k != "rounds"},
                      "rounds": prof.get("rounds")}), flush=True)
if world > 1:
    dist.destroy_process_group()
