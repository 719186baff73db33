"""Debug helper (GPU box): dumps intermediate outputs of the CUDA path for offline comparison with the oracle."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as MO
from oracle import pyramid as OP
from rdmnet_b200.model import create_model

scans = dict(np.load(os.path.join(ROOT, "tests/golden/scans.npz")))
state = torch.load(os.path.join(ROOT, "tests/golden/_big/rdmnet_state.pt"), map_location="cpu", weights_only=True)
m = create_model(); m.load_state_dict(state, strict=True); m = m.cuda().eval()
out_d = {}
keys = ("mask", "ref_feats_c", "src_feats_c", "ref_node_corr_indices", "src_node_corr_indices", "node_corr_scores",
        "matching_scores", "ref_points_c", "src_points_c", "estimated_transform", "ref_corr_points", "src_corr_points",
        "corr_scores", "ref_node_knn_indices", "src_node_knn_indices", "ref_node_knn_masks", "src_node_knn_masks",
        "shifted_ref_points_c", "shifted_src_points_c", "ref_feats_f", "src_feats_f")
for tag, a, b in (("p04", "s000000", "s000004"), ("p07", "s000000", "s000007")):
    pa, pb = scans[a], scans[b]
    pts = torch.from_numpy(np.concatenate([pa, pb])).cuda()
    lens = torch.tensor([len(pa), len(pb)], dtype=torch.int64).cuda()
    out = m({"points": pts, "lengths": lens})
    for k in keys:
        v = out[k].cpu().numpy()
        if k == "matching_scores" and tag != "p04":
            continue
        if k.endswith("feats_f"):
            v = v[:2000]
        out_d[f"{tag}_{k}"] = v
    # timing
    torch.cuda.synchronize()
    for _ in range(3):
        m({"points": pts, "lengths": lens})
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(10):
        m({"points": pts, "lengths": lens})
    torch.cuda.synchronize()
    print(tag, "forward ms/pair", (time.time() - t) * 100)
# p04 with the oracle pyramid (int64 tables)
a, b = scans["s000000"], scans["s000004"]
pyr = OP.precompute_pyramid(np.concatenate([a, b]), [len(a), len(b)], 5, 0.3, 4.25 * 0.3, MO.DEFAULT_LIMITS, "port")
tp = MO.pyramid_to_torch(pyr)
dd = {k: [t.cuda() for t in v] for k, v in tp.items()}
dd["features"] = torch.ones(tp["points"][0].shape[0], 1).cuda()
out = m(dd)
for k in keys:
    v = out[k].cpu().numpy()
    if k.endswith("feats_f"):
        v = v[:2000]
    out_d[f"o04_{k}"] = v
np.savez_compressed(os.path.join(ROOT, "gpurun_out/diag.npz"), **out_d)
# per-section timing with the torch profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m({"points": pts, "lengths": lens}); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
