"""Steady-state GPU timeline of one bench step (GPU box): kineto/CUPTI kernel records of K steps after warm-up.
Writes gpurun_out/<tag>_timeline.md: per-kernel totals (warm caches, real launch gaps) + busy/idle split per step, and
the largest idle gaps with the kernels around them. Not a bench: a profiler is attached.
Usage: python scripts/timeline.py <tag> [steps]"""
import json
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdmnet_b200 import synthetic  # noqa: E402
from rdmnet_b200.model import create_model  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "tl"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ck = os.path.join(ROOT, "tests/golden/_big/rdmnet_state.pt")
m = create_model()
if os.path.exists(ck):
    m.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True), strict=True)
m = m.cuda().eval()
p = synthetic.make_pair(pair_id=0)
pts = torch.from_numpy(np.concatenate([p["ref_points"], p["src_points"]])).cuda()
lens = torch.tensor([len(p["ref_points"]), len(p["src_points"])], dtype=torch.int64).cuda()
for _ in range(5):
    m({"points": pts, "lengths": lens})
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        m({"points": pts, "lengths": lens})
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"{tag}_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
os.remove(path)
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], ev[-1]["ts"] + ev[-1]["dur"]
busy = sum(e["dur"] for e in ev)
tot = defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"]
    n = n[:n.index("(")] if "(" in n else n
    tot[n][0] += 1
    tot[n][1] += e["dur"]
gaps = []
for a, b in zip(ev, ev[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 0:
        gaps.append((g, a["name"][:60], b["name"][:60]))
idle = sum(g[0] for g in gaps)
with open(os.path.join(ROOT, "gpurun_out", f"{tag}_timeline.md"), "w") as f:
    f.write(f"# steady-state timeline, {steps} steps: span {(t1 - t0) / steps:.1f} us/step, busy {busy / steps:.1f} us/step, "
            f"idle {idle / steps:.1f} us/step, {len(ev) / steps:.1f} GPU ops/step\n\n")
    f.write("| kernel | per step | us/step | share of busy | avg us |\n|---|---|---|---|---|\n")
    for n, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{n[:90]}` | {c / steps:.1f} | {d / steps:.1f} | {100 * d / busy:.1f}% | {d / c:.1f} |\n")
    f.write("\n## gap histogram (us): count, total/step\n\n")
    for lo, hi in ((0, 2), (2, 5), (5, 10), (10, 20), (20, 50), (50, 100), (100, 1e9)):
        sel = [g[0] for g in gaps if lo <= g[0] < hi]
        f.write(f"- [{lo}, {hi}): {len(sel) / steps:.1f}/step, {sum(sel) / steps:.1f} us/step\n")
    f.write("\n## largest gaps\n\n")
    for g in sorted(gaps, key=lambda g: -g[0])[:12 * steps]:
        f.write(f"- {g[0]:.1f} us between `{g[1]}` and `{g[2]}`\n")
print(open(os.path.join(ROOT, "gpurun_out", f"{tag}_timeline.md")).read()[:6000])
