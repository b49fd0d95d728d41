"""Build variants of the search kernels (positions per chain) on the GPU box and time the C2 search."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deltaq_b200 import _native, build, workloads as w  # noqa: E402

old, new = w.c2_exe_pair()
for chunk, heads in [tuple(int(y) for y in x.split("x")) for x in sys.argv[1:]]:
    out = os.path.join(ROOT, "gpurun_out", f"libdq_chunk{chunk}_{heads}.so")
    cmd = [build.nvcc_path()] + build.NVCC_FLAGS + [f"-DDQ_SEARCH_CHUNK={chunk}", f"-DDQ_SEARCH_HEADS={heads}", "-I", build.INCLUDE, "-o", out,
                                                     os.path.join(build.CSRC, "deltaq_cuda.cu")]
    subprocess.check_call(cmd)
    ctx = _native.Context(lib=_native.Library(out))
    sa = ctx.pinned(old.size, np.int32)
    pos = ctx.pinned(new.size, np.int32)
    ln = ctx.pinned(new.size, np.int32)
    best = None
    for _ in range(4):
        ctx.suffix_sort(old, sa.array)
        ctx.bsdiff_search(old, None, new, 0, new.size, pos.array, ln.array)
        ms = ctx.stats()["search_ms"]
        best = ms if best is None else min(best, ms)
    print(f"chunk={chunk} heads={heads}: search {best:.3f} ms (LCP build + heads + chains), checksum {int(pos.array.sum()) ^ int(ln.array.sum())}", flush=True)
    ctx.close()
