"""Same-box A/B of runtime knobs: search time of C2 for each environment setting given as KEY=VALUE[,KEY=VALUE] ("" = default)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from deltaq_b200 import _native, workloads as w
old, new = w.c2_exe_pair()
ctx = _native.Context()
sa = ctx.pinned(old.size, np.int32); pos = ctx.pinned(new.size, np.int32); ln = ctx.pinned(new.size, np.int32)
best = None
for _ in range(6):
    ctx.suffix_sort(old, sa.array)
    ctx.bsdiff_search(old, None, new, 0, new.size, pos.array, ln.array)
    ms = ctx.stats()["search_ms"]; best = ms if best is None else min(best, ms)
print("search %%.3f ms chk %%x" %% (best, int(pos.array.sum()) ^ int(ln.array.sum())))
''' % ROOT
for setting in sys.argv[1:]:
    env = dict(os.environ)
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(f"[{setting or 'default'}] {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
