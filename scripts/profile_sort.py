"""One warm-up + one suffix sort of a named workload, for ncu launch lists."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
t = {"c1": lambda: w.c1_uniform(), "c2": lambda: w.c2_exe_pair()[0], "c3": lambda: w.c3_repetitive(),
     "c4": lambda: w.c4_genome(64 << 20), "u16": lambda: w.c1_uniform(16 << 20, 9)}[name]()
s = CudaSuffixSort()
pin = s.context.pinned(t.size, np.int32)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    s.context.suffix_sort(t, pin.array)
print(s.stats())
