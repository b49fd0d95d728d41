"""Suffix sort of ONE text by a device group (BASELINE configs #4/#5 shape), one process driving G GPUs.

    python scripts/bench_sharded_sort.py c4 512 [reps] [nocheck] [G list, e.g. 1,2,4,8] [pageable]
Prints one JSON line per G: input MB/s through dq_cuda_suffix_sort of a group context (host text in, host SA out --
pinned buffers unless `pageable`), best of `reps`, an O(n) sufcheck of the result, and (DQ_TRACE=1) the phase times
the library prints to stderr.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DQ_SHARD_MIN", "1")
import oracle  # noqa: E402
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
    mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    check = (sys.argv[4] != "nocheck") if len(sys.argv) > 4 else True
    ndev = torch.cuda.device_count()
    gs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [g for g in (1, 2, 4, 8) if g <= ndev]
    pageable = len(sys.argv) > 6 and sys.argv[6] == "pageable"
    n = mib << 20
    t = {"c4": lambda: w.c4_genome(n), "uniform": lambda: w.c1_uniform(n, 7), "c3": lambda: w.c3_repetitive(n),
         "c2": lambda: w.c2_exe_pair(n, n + 1)[0], "c5": lambda: w.c5_old(workers=8)}[kind]()
    for G in gs:
        devs = [i % ndev for i in range(G)]
        sorter = CudaSuffixSort(device=devs if G > 1 else devs[0])
        ctx = sorter.context
        if pageable:
            text, sa = t, np.empty(t.size, np.int32)
        else:
            ptext = ctx.pinned(t.size, np.uint8)
            ptext.array[:] = t
            psa = ctx.pinned(t.size, np.int32)
            text, sa = ptext.array, psa.array
        times = []
        for it in range(reps + 1):
            t0 = time.perf_counter()
            ctx.suffix_sort(text, sa)
            dt = time.perf_counter() - t0
            if it > 0:
                times.append(dt)
        st = ctx.stats()
        if check and t.size > (1 << 30):
            # 2 G suffixes: a full sufcheck takes minutes; sample adjacent pairs here (tests/test_big_gpu.py runs the full one)
            idx = np.sort(np.random.default_rng(0).integers(0, t.size - 1, 500_000))
            ok = int(oracle.verify_pairs(t, sa, idx))
        else:
            ok = int(oracle.sufcheck(t, sa)) if check else None
        best = min(times)
        print(json.dumps({"workload": kind, "n": int(t.size), "n_gpus": G, "devices": devs, "best_ms": best * 1e3,
                          "input_MBps": t.size / best / 1e6, "all_ms": [round(x * 1e3, 2) for x in times], "sufcheck": ok,
                          "rounds": st["rounds"], "launches": st["kernel_launches"], "pinned": not pageable}), flush=True)
        sorter.dispose()


if __name__ == "__main__":   # c5_old's generator workers are spawned: they import this file
    main()
