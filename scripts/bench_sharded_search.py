"""Sort + match search of one (old, new) pair by a device group: C5's recipe at a given scale.

    python scripts/bench_sharded_search.py [old MiB] [G list, e.g. 1,2,4,8]
Prints one JSON line per G (wall ms through the host-pointer C ABI with pinned buffers, device ms of the search and of
its index build) and checks every table against the first G's.  DQ_TRACE=1 adds the library's per-shard times."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DQ_SHARD_MIN", "1")
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ndev = torch.cuda.device_count()
    gs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [g for g in (1, 2, 4, 8) if g <= ndev]
    old, new = w.c5_pair(mib << 20, workers=8)
    n, m = int(old.size), int(new.size)
    ref = None
    for G in gs:
        devs = [i % ndev for i in range(G)]
        s = CudaSuffixSort(device=devs if G > 1 else devs[0])
        c = s.context
        p_o = c.pinned(n, np.uint8); p_o.array[:] = old
        p_n = c.pinned(m, np.uint8); p_n.array[:] = new
        p_sa = c.pinned(n, np.int32)
        pos = c.pinned(m, np.int32)
        ln = c.pinned(m, np.int32)
        best = None
        for it in range(3):
            t0 = time.perf_counter()
            c.suffix_sort(p_o.array, p_sa.array)
            t1 = time.perf_counter()
            sort_stats = c.stats()
            c.bsdiff_search(p_o.array, None, p_n.array, 0, m, pos.array, ln.array)
            t2 = time.perf_counter()
            st = c.stats()
            if it and (best is None or t2 - t0 < best[0]):
                best = (t2 - t0, t1 - t0, t2 - t1, st["search_ms"], st["search_index_ms"], sort_stats["rounds"])
        if ref is None:
            ref = (pos.array.copy(), ln.array.copy())
            same = None
        else:
            same = bool(np.array_equal(ref[0], pos.array) and np.array_equal(ref[1], ln.array))
        print(json.dumps({"old_bytes": n, "new_bytes": m, "n_gpus": G, "total_ms": best[0] * 1e3, "sort_ms": best[1] * 1e3,
                          "search_wall_ms": best[2] * 1e3, "search_device_ms": best[3], "search_index_device_ms": best[4],
                          "sort_rounds": best[5], "input_MBps": m / best[0] / 1e6, "table_equals_first": same}), flush=True)
        for p in (p_o, p_n, p_sa, pos, ln):
            p.free()
        s.dispose()


if __name__ == "__main__":   # c5_pair's generator workers are spawned: they import this file
    main()
