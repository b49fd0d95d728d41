"""C5-scale bsdiff hot path on ONE B200: sort(old) + search(all positions of new) + host loop, round-trip checked.

    python scripts/big_bsdiff.py [old MiB] -- default 1945 MiB (2,040,109,466 bytes would need the tail; MiB granularity here)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deltaq_b200 import CudaSuffixSort, bsdiff, workloads as w  # noqa: E402

c5 = "--c5" in sys.argv          # BASELINE's C5 recipe itself (workloads._exe_like / _mutate at full size; slower to generate)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
mib = int(args[0]) if args else 1945
n = mib << 20
rng = np.random.default_rng(5)
t0 = time.time()
if c5:
    n = 2_040_109_466 if not args else n
    old = w._exe_like(n, np.random.default_rng(5))
    new = w._mutate(old, np.random.default_rng(105), min(n + n // 16, 2_100_000_000), regions=2000)
    print(f"C5 recipe: old={old.size} new={new.size} in {time.time()-t0:.0f}s", flush=True)
base = None if c5 else w.c2_exe_pair(64 << 20, (64 << 20) + 1)[0]
if not c5:
    parts = []
    total = 0
    while total < n:
        parts.append(base if (len(parts) % 3) else rng.integers(0, 256, base.size, dtype=np.uint8))
        total += base.size
    old = np.concatenate(parts)[:n]
    del parts
    # new = old with ~200 edits (overwrites / inserts / deletes of up to 1 MiB), built by slicing
    cuts = np.sort(rng.integers(0, n, 200))
    out = []
    cur = 0
    for c in cuts:
        c = int(max(c, cur))
        out.append(old[cur:c])
        op = int(rng.integers(0, 3))
        k = int(rng.integers(64, 1 << 20))
        if op == 0:
            out.append(rng.integers(0, 256, k, dtype=np.uint8)); cur = min(n, c + k)
        elif op == 1:
            out.append(rng.integers(0, 256, k, dtype=np.uint8)); cur = c
        else:
            cur = min(n, c + k)
    out.append(old[cur:])
    new = np.concatenate(out)
    if new.size > 2_100_000_000:
        new = new[:2_100_000_000]
    del out
print(f"generated old={old.size} new={new.size} in {time.time()-t0:.0f}s", flush=True)
s = CudaSuffixSort()
ctx = s.context
p_old = ctx.pinned(old.size, np.uint8); p_old.array[:] = old
p_new = ctx.pinned(new.size, np.uint8); p_new.array[:] = new
best = None
for it in range(2):
    t1 = time.perf_counter()
    st = ctx.bsdiff_streams(p_old.array, p_new.array, copy=False)
    dt = time.perf_counter() - t1
    stats = ctx.stats()
    print(it, f"e2e {dt*1e3:.0f} ms  sort_dev {stats['device_ms']:.0f} ms  search_dev {stats['search_ms']:.0f} ms rounds {stats['rounds']} "
              f"visits {st['search_visits']} ctrl {st['ctrl'].size//24} diff {st['diff'].size} extra {st['extra'].size}", flush=True)
    best = dt if best is None else min(best, dt)
t2 = time.time()
rebuilt = bsdiff.apply_streams(old, st["ctrl"].tobytes(), st["diff"], st["extra"].tobytes(), new.size)
ok = rebuilt == new.tobytes()
print(f"round trip {'OK' if ok else 'FAILED'} ({time.time()-t2:.0f}s)", flush=True)
rec = dict(old_bytes=int(old.size), new_bytes=int(new.size), e2e_ms=best * 1e3, e2e_MBps=new.size / best / 1e6,
           sort_device_ms=stats["device_ms"], search_device_ms=stats["search_ms"], rounds=stats["rounds"],
           search_visits=int(st["search_visits"]), ctrl_triples=int(st["ctrl"].size // 24), round_trip_ok=bool(ok))
print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rec, open("gpurun_out/big_bsdiff_c5.json" if c5 else "gpurun_out/big_bsdiff.json", "w"))
