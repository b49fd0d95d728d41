"""BASELINE config #5 end to end on a device group: dq_cuda_bsdiff_streams of the C5 pair (2,040,109,466 B old ->
~2.1 GB new) with `old` sorted by all GPUs and the match search sharded over them, round-trip checked.

    python scripts/big_bsdiff_group.py [G list, e.g. 1,8] [old MiB; default: the full C5 size]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DQ_SHARD_MIN", str(32 << 20))
from deltaq_b200 import CudaSuffixSort, bsdiff, workloads as w  # noqa: E402


def main():
    ndev = torch.cuda.device_count()
    gs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, ndev]
    n_old = (int(sys.argv[2]) << 20) if len(sys.argv) > 2 else 2_040_109_466
    t0 = time.time()
    old, new = w.c5_pair(n_old, workers=8)
    if new.size > 2_100_000_000:
        new = np.ascontiguousarray(new[:2_100_000_000])
    print(f"C5 recipe: old={old.size} new={new.size} generated in {time.time() - t0:.0f}s", flush=True)
    ref = None
    for G in gs:
        devs = [i % ndev for i in range(G)]
        s = CudaSuffixSort(device=devs if G > 1 else devs[0])
        ctx = s.context
        p_old = ctx.pinned(old.size, np.uint8)
        p_old.array[:] = old
        p_new = ctx.pinned(new.size, np.uint8)
        p_new.array[:] = new
        best = None
        for it in range(2):
            t1 = time.perf_counter()
            st = ctx.bsdiff_streams(p_old.array, p_new.array, copy=False)
            dt = time.perf_counter() - t1
            stats = ctx.stats()
            best = dt if best is None else min(best, dt)
        sizes = (int(st["ctrl"].size), int(st["diff"].size), int(st["extra"].size))
        digest = (hash(st["ctrl"].tobytes()), hash(st["diff"].tobytes()[:1 << 26]), hash(st["extra"].tobytes()[:1 << 26]))
        if ref is None:
            t2 = time.time()
            rebuilt = bsdiff.apply_streams(old, st["ctrl"], st["diff"], st["extra"], new.size)
            ok = rebuilt == new.tobytes()
            del rebuilt
            ref = (sizes, digest)
            check = f"round trip {'OK' if ok else 'FAILED'} ({time.time() - t2:.0f}s)"
        else:
            ok = ref == (sizes, digest)
            check = f"streams {'identical to' if ok else 'DIFFER from'} the first run's"
        rec = dict(old_bytes=int(old.size), new_bytes=int(new.size), n_gpus=G, e2e_ms=best * 1e3,
                   e2e_MBps=new.size / best / 1e6, sort_ms=stats["device_ms"], search_ms=stats["search_ms"],
                   rounds=stats["rounds"], ctrl_triples=sizes[0] // 24, ok=bool(ok), check=check)
        print(json.dumps(rec), flush=True)
        p_old.free()
        p_new.free()
        s.dispose()


if __name__ == "__main__":   # the generator's workers are spawned: they import this file
    main()
