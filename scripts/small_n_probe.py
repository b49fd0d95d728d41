"""Latency of ISuffixSort.Sort at the reference benchmark's small sizes: C ABI (pinned buffers) vs the CPU oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from deltaq_b200 import CudaSuffixSort  # noqa: E402

s = CudaSuffixSort()
ctx = s.context
rng = np.random.default_rng(670761)
print("| n | C ABI us | device us | launches | CPU SA-IS us |")
print("|---|---|---|---|---|")
for n in [16, 256, 1024, 4096, 16384, 32768, 65536, 262144, 1048576]:
    t = rng.integers(0, 256, n, dtype=np.uint8)
    pt = ctx.pinned(n, np.uint8); pt.array[:] = t
    ps = ctx.pinned(n, np.int32)
    for _ in range(5):
        ctx.suffix_sort(pt.array, ps.array)
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.suffix_sort(pt.array, ps.array)
    gpu = (time.perf_counter() - t0) / reps
    st = ctx.stats()
    t0 = time.perf_counter()
    for _ in range(20):
        ref = oracle.sais(t)
    cpu = (time.perf_counter() - t0) / 20
    assert np.array_equal(ps.array, ref)
    print(f"| {n} | {gpu*1e6:.0f} | {st['device_ms']*1e3:.0f} | {st['kernel_launches']} | {cpu*1e6:.0f} |", flush=True)
