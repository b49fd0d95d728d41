"""Ad-hoc GPU probe: correctness spot checks + timings of the suffix sorter on the BASELINE shapes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402

out = {}
s = CudaSuffixSort()
s.context.set_timing(True)


def run(name, t, check):
    pin = s.context.pinned(t.size, np.int32)
    sa = pin.array
    best = None
    for it in range(4):
        t0 = time.perf_counter()
        s.context.suffix_sort(t, sa)
        dt = time.perf_counter() - t0
        st = s.stats()
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    ok = None
    if check == "exact":
        ok = bool(np.array_equal(sa, oracle.sais(t)))
    elif check == "sufcheck":
        ok = oracle.sufcheck(t, sa) == 0
    rec = dict(n=int(t.size), e2e_ms=dt * 1e3, e2e_MBps=t.size / dt / 1e6, device_ms=st["device_ms"],
               device_MBps=t.size / st["device_ms"] / 1e3, rounds=st["rounds"], passes=st["radix_passes"],
               launches=st["kernel_launches"], active_sum=st["active_sum"], alg_bytes=st["algorithmic_bytes"],
               alg_GBps=st["algorithmic_bytes"] / st["device_ms"] / 1e6,
               pass_ms=st["pass_ms"], pass_GBps=(st["pass_pairs"] * 24 / st["pass_ms"] / 1e6) if st["pass_ms"] else None,
               ok=ok)
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
    pin.free()


run("c1_1MiB_uniform", w.c1_uniform(), "exact")
old, new = w.c2_exe_pair()
run("c2_old_16MiB_exe", old, "sufcheck")
run("uniform_16MiB", w.c1_uniform(16 << 20, 9), "sufcheck")
run("uniform_64MiB", w.c1_uniform(64 << 20, 10), None)
run("c3_64MiB_repetitive", w.c3_repetitive(), "sufcheck")
run("c4_genome_64MiB", w.c4_genome(64 << 20), None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
