"""Per-round view of one suffix sort: active count and onesweep time per doubling round."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
t = {"c2": lambda: w.c2_exe_pair()[0], "c3": lambda: w.c3_repetitive(), "c4": lambda: w.c4_genome(64 << 20)}[name]()
s = CudaSuffixSort()
s.context.set_timing(True)
pin = s.context.pinned(t.size, np.int32)
for _ in range(3):
    s.context.suffix_sort(t, pin.array)
st = s.stats()
rounds = []
for ms, pairs, shift in s.context.pass_times():
    if shift == 0:
        rounds.append([pairs, 0.0, 0])
    rounds[-1][1] += ms
    rounds[-1][2] += 1
print(f"{name}: device {st['device_ms']:.3f} ms, passes {st['pass_ms']:.3f} ms, rounds {st['rounds']}")
for i, (pairs, ms, np_) in enumerate(rounds):
    print(f"round {i:2d}: a={pairs:10d} passes={np_} pass_ms={ms:.3f}  ({pairs*24*np_/ms/1e6:6.0f} GB/s)")
