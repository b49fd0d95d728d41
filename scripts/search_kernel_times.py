"""Per-kernel device times of one bsdiff search via CUDA-event-free means: run under
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file OUT python scripts/search_kernel_times.py MiB
on the big_bsdiff construction (64 MiB blocks, two thirds of them copies of one block) at a chosen size."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
blk = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 20
n = mib << 20
rng = np.random.default_rng(5)
base = w.c2_exe_pair(blk, blk + 1)[0]
parts = []
total = 0
while total < n:
    parts.append(base if (len(parts) % 3) else rng.integers(0, 256, base.size, dtype=np.uint8))
    total += base.size
old = np.concatenate(parts)[:n]
cuts = np.sort(rng.integers(0, n, max(4, 200 * mib // 1945)))
out, cur = [], 0
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
s = CudaSuffixSort()
ctx = s.context
for it in range(2):
    t = time.perf_counter()
    st = ctx.bsdiff_streams(old, new, copy=False)
    dt = time.perf_counter() - t
    stats = ctx.stats()
    print(f"{mib} MiB (blocks of {blk >> 20} MiB): e2e {dt*1e3:.0f} ms sort_dev {stats['device_ms']:.0f} ms search_dev {stats['search_ms']:.0f} ms "
          f"rounds {stats['rounds']} visits {st['search_visits']}", flush=True)
