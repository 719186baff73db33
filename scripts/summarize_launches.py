"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of one step.
Usage: python scripts/summarize_launches.py gpurun_out/<tag>_launches.csv [step_index] > profiles/<name>.md"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
# steps start at the first gs_cloud_kernel of each group of 4 grid_subsample calls
starts = [i for i, r in enumerate(rows) if r[0].startswith("gs_cloud_kernel")][::4]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
beg, end = starts[k], (starts[k + 1] if k + 1 < len(starts) else len(rows))
step = rows[beg:end]
tot = sum(r[1] for r in step)
agg = OrderedDict()
for n, t, g, b in step:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += t
print(f"# ncu launch list, step {k} of `{sys.argv[1]}`: {len(step)} launches, {tot/1e3:.3f} ms of kernel time (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {t:.1f} | {100*t/tot:.1f}% |")
print("\n## KPConv gather launches of this step\n\n| kernel | grid | block | us |\n|---|---|---|---|")
for n, t, g, b in step:
    if "kpconv_gather" in n:
        print(f"| `{n}` | {g} | {b} | {t:.1f} |")
