"""ncu launch-list CSV (--metrics gpu__time_duration.sum --csv) -> markdown table of kernel shares.

    python scripts/launch_summary.py launches.csv "title" "command" > profiles/xxx.md
"""
import collections
import csv
import re
import sys

path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name).replace("dq::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v * scale
total = sum(v for _, v in agg.values())
print(f"# {title}\n")
print(f"Command (1 x B200): `{command}`")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
print(f"Total kernel time: {total:.3f} ms over {sum(c for c, _ in agg.values())} launches\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---|---|---|---|")
for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {c} | {v:.3f} | {100 * v / total:.1f}% | {1e3 * v / c:.1f} |")
