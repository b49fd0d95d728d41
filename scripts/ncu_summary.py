"""ncu --set full report (.ncu-rep) -> markdown: one row per captured launch + the stall samples of one launch.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "title" "command" [bytes_per_unit_field=...] > profiles/xxx.md
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).  The "algorithmic" column is filled for the
onesweep pass kernel only (24 B x pairs, pairs = kernel argument count is not in the report: taken from the grid size x 4096
for full grids, so the last partial tile rounds up).
"""
import csv
import re
import subprocess
import sys

rep, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
col = {name: i for i, name in enumerate(h)}


def val(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


def f(r, name):
    try:
        return float(val(r, name).replace(",", ""))
    except ValueError:
        return float("nan")


print(f"# {title}\n")
print(f"Command: `{command}`\n")
print("| launch | kernel | grid | duration us (under ncu) | dram read MB | dram write MB | dram % of peak | L2 sectors (M) | "
      "warps active % | issue active % | regs | CTAs/SM limit (regs / smem) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
data = rows[2:]
for k, r in enumerate(data):
    name = re.sub(r"\(.*", "", val(r, "Kernel Name")).replace("void ", "")
    print(f"| {k} | `{name}` | {val(r, 'launch__grid_size')} | {f(r, 'gpu__time_duration.sum'):.1f} | "
          f"{f(r, 'dram__bytes_read.sum'):.1f} | {f(r, 'dram__bytes_write.sum'):.1f} | "
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{f(r, 'lts__t_sectors.sum') / 1e6:.1f} | "
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'launch__registers_per_thread')} | "
          f"{f(r, 'launch__occupancy_limit_registers'):.0f} / {f(r, 'launch__occupancy_limit_shared_mem'):.0f} |")
units = {name: rows[1][i] for name, i in col.items()}
# stall samples, per distinct kernel (first launch of each)
seen = set()
for k, r in enumerate(data):
    name = re.sub(r"\(.*", "", val(r, "Kernel Name")).replace("void ", "")
    if name in seen:
        continue
    seen.add(name)
    stalls = []
    for cname, i in col.items():
        m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", cname)
        if m and not cname.endswith("_not_issued"):
            try:
                stalls.append((float(r[i].replace(",", "")), m.group(1)))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    tot = sum(v for v, _ in stalls) or 1.0
    print(f"\n## launch {k} `{name}`: stall samples (smsp__pcsamp_*)\n")
    print("| reason | samples | share |")
    print("|---|---|---|")
    for v, nm in stalls[:9]:
        print(f"| {nm} | {v:.0f} | {100 * v / tot:.0f} % |")
    u = units.get("dram__bytes_read.sum", "")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1e6)
    print(f"\nDRAM traffic of this launch: {(f(r, 'dram__bytes_read.sum') + f(r, 'dram__bytes_write.sum')) * scale:,.0f} B "
          f"in {f(r, 'gpu__time_duration.sum'):.1f} us.")
