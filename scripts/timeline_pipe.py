"""Stream-level Gantt of the pair pipeline in steady state (GPU box; a profiler is attached: not a bench). For each 100 us
slot: which streams have a kernel running, and the phase markers (first gather / first tf_attend / vote / sinkhorn) of
every pair. Usage: python scripts/timeline_pipe.py <tag> [pairs]"""
import json
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdmnet_b200 import synthetic  # noqa: E402
from rdmnet_b200.model import PairPipeline, create_model  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "tlp"
n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ck = os.path.join(ROOT, "tests/golden/_big/rdmnet_state.pt")
m = create_model()
if os.path.exists(ck):
    m.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True), strict=True)
m = m.cuda().eval()
items = []
for i in range(4):
    p = synthetic.make_pair(pair_id=i)
    items.append((torch.from_numpy(np.concatenate([p["ref_points"], p["src_points"]])).cuda(),
                  torch.tensor([len(p["ref_points"]), len(p["src_points"])], dtype=torch.int64).cuda()))
pipe = PairPipeline(m)
for _ in pipe.run([items[i % 4] for i in range(10)]):
    pass
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import time  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    for _ in pipe.run([items[i % 4] for i in range(n_pairs)]):
        pass
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
path = os.path.join(ROOT, "gpurun_out", f"{tag}_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
os.remove(path)
ev.sort(key=lambda e: e["ts"])
T0 = ev[0]["ts"]
T1 = max(e["ts"] + e["dur"] for e in ev)
streams = sorted({e["args"].get("stream") for e in ev})
busy = defaultdict(float)
for e in ev:
    busy[e["args"].get("stream")] += e["dur"]
slot = 100.0
nslot = int((T1 - T0) / slot) + 1
grid = {s: np.zeros(nslot) for s in streams}
for e in ev:
    a, b = e["ts"] - T0, e["ts"] - T0 + e["dur"]
    for k in range(int(a / slot), min(int(b / slot), nslot - 1) + 1):
        lo, hi = max(a, k * slot), min(b, (k + 1) * slot)
        if hi > lo:
            grid[e["args"].get("stream")][k] += (hi - lo) / slot
marks = defaultdict(list)
for e in ev:
    n = e["name"]
    for key, ch in (("kpconv_gather_c1", "E"), ("tf_attend", "t"), ("vote_finish", "V"), ("sinkhorn128", "S"), ("lgr_refine", "L"),
                    ("gs_cloud", "g"), ("rs_query", "q")):
        if key in n:
            marks[int((e["ts"] - T0) / slot)].append(ch)
out = [f"# pair pipeline, {n_pairs} pairs (profiler attached): wall {1e3 * wall / n_pairs:.2f} ms/pair, GPU span {(T1 - T0) / n_pairs / 1e3:.2f} ms/pair; "
       f"overlap={pipe.overlap}", ""]
for s in streams:
    out.append(f"stream {s}: busy {busy[s] / n_pairs / 1e3:.2f} ms/pair")
union = np.zeros(nslot)
for s in streams:
    union = np.maximum(union, np.minimum(grid[s], 1.0))
out.append(f"any stream busy: {union.sum() * slot / n_pairs / 1e3:.2f} ms/pair of {(T1 - T0) / n_pairs / 1e3:.2f}")
out.append("")
out.append("one column = 100 us; rows = streams (# > 50 % busy, + > 10 %), last row = phase marks (E first gather, t tf_attend, V vote, S sinkhorn, L lgr refine, g grid subsample, q radius query)")
W = 120
for c0 in range(0, nslot, W):
    for s in streams:
        out.append(f"{str(s):>4} " + "".join("#" if v > 0.5 else ("+" if v > 0.1 else ".") for v in grid[s][c0:c0 + W]))
    out.append("     " + "".join((marks[k][0] if marks.get(k) else " ") for k in range(c0, min(c0 + W, nslot))))
    out.append("")
open(os.path.join(ROOT, "gpurun_out", f"{tag}_timeline_pipe.md"), "w").write("\n".join(out))
print("\n".join(out)[:7000])

# detailed list of one steady-state window (between the 4th and the 6th first-gather marks)
firsts = [e["ts"] for e in ev if "kpconv_gather_c1" in e["name"]]
if len(firsts) >= 6:
    a, b = firsts[3], firsts[5]
    lines = []
    for e in ev:
        if a <= e["ts"] < b:
            n = e["name"]
            n = n[:n.index("(")] if "(" in n else n
            lines.append(f"{(e['ts'] - a):9.1f} {e['dur']:7.1f} s{e['args'].get('stream')} {n[-48:]}")
    open(os.path.join(ROOT, "gpurun_out", f"{tag}_window.txt"), "w").write("\n".join(lines))
