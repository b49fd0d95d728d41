"""Timeline of one dq_cuda_bsdiff_streams call on C2 (DQ_TRACE=1 prints host-clock marks to stderr)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DQ_TRACE"] = "1"
from deltaq_b200 import workloads as w
from deltaq_b200._native import Context

old, new = w.c2_exe_pair()
with Context() as ctx:
    po, pn = ctx.pinned(len(old), "uint8"), ctx.pinned(len(new), "uint8")
    po.array[:] = old
    pn.array[:] = new
    for i in range(4):
        t = time.perf_counter()
        ctx.bsdiff_streams(po.array, pn.array, copy=False)
        print("call %d: %.3f ms" % (i, (time.perf_counter() - t) * 1e3), file=sys.stderr)
print("cpus", os.cpu_count(), file=sys.stderr)
