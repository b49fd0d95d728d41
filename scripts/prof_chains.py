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
if len(sys.argv) > 1 and sys.argv[1] == "rep":
    # 16 MiB made of four copies of a 4 MiB block with fresh random blocks in between (long repeats: the seed level runs)
    rng = np.random.default_rng(5)
    base = w.c2_exe_pair(4 << 20, (4 << 20) + 1)[0]
    old = np.concatenate([base if k % 3 else rng.integers(0, 256, base.size, dtype=np.uint8) for k in range(4)])
    new = old.copy()
    for c in rng.integers(0, old.size - 70000, 12):
        new[c:c + int(rng.integers(64, 65536))] = 7
else:
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
clk = np.zeros(1 << 16, dtype=np.uint64)
byt = np.zeros(1 << 16, dtype=np.uint64)
calls = np.zeros(1 << 16, dtype=np.uint32)
lib.L.dq_debug_read_prof_cmp.argtypes = [ctypes.c_void_p] * 3
assert lib.L.dq_debug_read_prof_cmp(clk.ctypes.data, byt.ctypes.data, calls.ctypes.data) == 0
order = np.argsort(-h)[:15]
print("slowest warps: (us total, us inside common_prefix_warp, calls, bytes compared, position, len min/median/max, short fraction)")
for i in order:
    a = i * 2048
    seg = L[a:a + 2048]
    print("  %8.1f  cmp %8.1f us %5d calls %10d B  %9d  len %d/%d/%d  short %.2f  %s" % (
        h[i], clk[i] / 1.965e3, calls[i], byt[i], a, seg.min(), int(np.median(seg)), seg.max(), fs[i], bytes(new[a:a + 12]).hex()))
seeds = np.zeros(1 << 16, dtype=np.uint32)
lib.L.dq_debug_read_prof_seeds.argtypes = [ctypes.c_void_p]
assert lib.L.dq_debug_read_prof_seeds(seeds.ctypes.data) == 0
sd = seeds[seeds > 0].astype(np.float64) * 64 / 1.965e3
if sd.size:
    print("seed-level warps: n %d mean %.1f us p50 %.1f p90 %.1f p99 %.1f max %.1f" % (sd.size, sd.mean(), *np.percentile(sd, [50, 90, 99]), sd.max()))
print("stats", ctx.stats())
os.remove(out)
