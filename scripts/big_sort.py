"""Large single-GPU sorts (wide look-back descriptors for >= 2^30 pairs): timing + O(n) sufcheck."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
out = {}
s = CudaSuffixSort()
s.context.set_timing(True)
if which == "c4":
    t = w.c4_genome(512 << 20)
elif which == "c5":
    n = 2_040_109_466
    rng = np.random.default_rng(5)
    # C5-scale text: cheap to generate (tiled exe-like blocks with fresh random blocks in between)
    base = w.c2_exe_pair(64 << 20, (64 << 20) + 1)[0]
    parts = []
    total = 0
    while total < n:
        parts.append(base if (len(parts) % 3) else rng.integers(0, 256, base.size, dtype=np.uint8))
        total += base.size
    t = np.concatenate(parts)[:n]
    del parts
else:
    t = w.c1_uniform(int(which), 11)
print("generated", t.size, flush=True)
pin = s.context.pinned(t.size, np.int32)
for it in range(2):
    t0 = time.perf_counter()
    s.context.suffix_sort(t, pin.array)
    dt = time.perf_counter() - t0
    st = s.stats()
    print(it, "e2e_ms", dt * 1e3, st, flush=True)
rec = dict(n=int(t.size), e2e_ms=dt * 1e3, e2e_MBps=t.size / dt / 1e6, device_ms=st["device_ms"], rounds=st["rounds"],
           passes=st["radix_passes"], pass_GBps=st["pass_pairs"] * 24 / st["pass_ms"] / 1e6,
           alg_GBps=st["algorithmic_bytes"] / st["device_ms"] / 1e6)
t0 = time.perf_counter()
rec["sufcheck"] = int(oracle.sufcheck(t, pin.array))
rec["sufcheck_s"] = time.perf_counter() - t0
print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rec, open(f"gpurun_out/big_sort_{which}.json", "w"))
