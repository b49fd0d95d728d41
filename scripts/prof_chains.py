"""Where the time of search_chain_kernel / search_heads_kernel goes on C2: builds the library with -DDQ_PROF (per-chain
and per-warp clock counters), runs one search, and relates the slowest chains to what the table says about them."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deltaq_b200 import _native, build, workloads as w  # noqa: E402

out = os.path.join(ROOT, "gpurun_out", "libdq_prof.so")
subprocess.check_call([build.nvcc_path()] + build.NVCC_FLAGS + ["-DDQ_PROF", "-I", build.INCLUDE, "-o", out,
                                                                  os.path.join(build.CSRC, "deltaq_cuda.cu")])
lib = _native.Library(out)
old, new = w.c2_exe_pair()
ctx = _native.Context(lib=lib)
sa = ctx.pinned(old.size, np.int32)
pos = ctx.pinned(new.size, np.int32)
ln = ctx.pinned(new.size, np.int32)
for _ in range(2):
    ctx.suffix_sort(old, sa.array)
    ctx.bsdiff_search(old, None, new, 0, new.size, pos.array, ln.array)
print("search_ms", ctx.stats()["search_ms"])
chains = np.zeros(1 << 21, dtype=np.uint32)
heads = np.zeros(1 << 16, dtype=np.uint32)
lib.L.dq_debug_read_prof.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.L.dq_debug_read_prof(chains.ctypes.data, heads.ctypes.data) == 0
nch = (new.size + 31) // 32
c = chains[:nch].astype(np.float64) * 64 / 1.965e3   # microseconds at 1965 MHz
L = ln.array
short = (L <= 8)
frac_short = np.add.reduceat(short, np.arange(0, new.size, 32)) / 32.0
print("chains: n %d  mean %.1f us  p50 %.1f  p90 %.1f  p99 %.1f  p99.9 %.1f  max %.1f" % (
    nch, c.mean(), *np.percentile(c, [50, 90, 99, 99.9]), c.max()))
for lo, hi in [(0, 0.01), (0.01, 0.5), (0.5, 0.99), (0.99, 1.01)]:
    m = (frac_short >= lo) & (frac_short < hi)
    if m.any():
        print("  chains with short-match fraction in [%.2f, %.2f): %7d  mean %.1f us  p99 %.1f  max %.1f  (sum %.1f chain-ms)" % (
            lo, hi, m.sum(), c[m].mean(), np.percentile(c[m], 99), c[m].max(), c[m].sum() / 1e3))
order = np.argsort(-c)[:25]
print("slowest chains: (us, position, len at head, min/max len in chunk, short fraction, bytes)")
for i in order:
    a = i * 32
    seg = L[a:a + 32]
    print("  %8.1f  %9d  %8d  %6d/%8d  %.2f  %s" % (c[i], a, seg[0], seg.min(), seg.max(), frac_short[i], bytes(new[a:a + 12]).hex()))
nw = (new.size + 2047) // 2048
h = heads[:nw].astype(np.float64) * 64 / 1.965e3
print("heads warps: n %d mean %.1f us p50 %.1f p90 %.1f p99 %.1f max %.1f" % (nw, h.mean(), *np.percentile(h, [50, 90, 99]), h.max()))
fs = np.add.reduceat(short, np.arange(0, new.size, 2048)) / 2048.0
for lo, hi in [(0, 0.01), (0.01, 0.5), (0.5, 0.99), (0.99, 1.01)]:
    m = (fs >= lo) & (fs < hi)
    if m.any():
        print("  warps with short fraction in [%.2f, %.2f): %6d mean %.1f us p99 %.1f max %.1f" % (lo, hi, m.sum(), h[m].mean(), np.percentile(h[m], 99), h[m].max()))
order = np.argsort(-h)[:15]
for i in order:
    a = i * 2048
    seg = L[a:a + 2048]
    print("  %8.1f  %9d  len min/median/max %d/%d/%d  short %.2f  %s" % (h[i], a, seg.min(), int(np.median(seg)), seg.max(), fs[i], bytes(new[a:a + 12]).hex()))
os.remove(out)
