"""Why does the tf32 mode land 30 cm from the fp32 pose on bundled pair (0,7)? Dumps, for both modes: pose, correspondence count,
how many coarse node pairs / fine correspondences are shared, and the inlier statistics of each mode's correspondences under each
mode's pose (LGR acceptance radius)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rdmnet_b200
from rdmnet_b200.model import create_model

scans = dict(np.load(os.path.join(ROOT, "tests", "golden", "scans.npz")))
gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "pair_outputs.npz")))
state = torch.load(os.path.join(ROOT, "tests", "golden", "_big", "rdmnet_state.pt"), map_location="cpu", weights_only=True)
model = create_model(); model.load_state_dict(state, strict=True); model = model.cuda().eval()
for tag, a, b in (("p04", "s000000", "s000004"), ("p07", "s000000", "s000007")):
    pts = torch.from_numpy(np.concatenate([scans[a], scans[b]])).cuda()
    lens = torch.tensor([len(scans[a]), len(scans[b])], dtype=torch.int64).cuda()
    outs = {}
    for mode in ("fp32", "tf32"):
        rdmnet_b200.set_precision(mode)
        o = model({"points": pts, "lengths": lens})
        outs[mode] = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in o.items() if k in (
            "estimated_transform", "ref_corr_points", "src_corr_points", "corr_scores", "ref_node_corr_indices", "src_node_corr_indices",
            "ref_feats_c", "src_feats_c", "ref_points_c", "src_points_c")}
    rdmnet_b200.set_precision("fp32")
    f, t = outs["fp32"], outs["tf32"]
    print(f"== {tag}")
    print("feats_c rel diff", np.abs(f["ref_feats_c"] - t["ref_feats_c"]).max() / np.abs(f["ref_feats_c"]).max() if f["ref_feats_c"].shape == t["ref_feats_c"].shape else "shape differs")
    nf = set(zip(f["ref_node_corr_indices"].tolist(), f["src_node_corr_indices"].tolist()))
    nt = set(zip(t["ref_node_corr_indices"].tolist(), t["src_node_corr_indices"].tolist()))
    print("coarse node pairs: fp32", len(nf), "tf32", len(nt), "shared", len(nf & nt))
    cf = {tuple(np.round(r, 4)) for r in np.concatenate([f["ref_corr_points"], f["src_corr_points"]], 1).tolist()}
    ct = {tuple(np.round(r, 4)) for r in np.concatenate([t["ref_corr_points"], t["src_corr_points"]], 1).tolist()}
    print("fine correspondences: fp32", len(cf), "tf32", len(ct), "shared", len(cf & ct))
    for cname, c in (("fp32-corr", f), ("tf32-corr", t)):
        for pname, T in (("fp32-pose", f["estimated_transform"]), ("tf32-pose", t["estimated_transform"]), ("golden-pose", gold[tag + "_estimated_transform"])):
            res = np.linalg.norm(c["src_corr_points"] @ T[:3, :3].T + T[:3, 3] - c["ref_corr_points"], axis=1)
            print(f"  {cname} under {pname}: inliers(<1.0 m) {(res < 1.0).sum():4d} (<0.6) {(res < 0.6).sum():4d} (<0.3) {(res < 0.3).sum():4d} median residual {np.median(res):.3f} m mean(inl<1) {res[res<1.0].mean():.3f}")
    print("  pose fp32:", np.round(f["estimated_transform"][:3, 3], 3), "tf32:", np.round(t["estimated_transform"][:3, 3], 3), "golden:", np.round(gold[tag + "_estimated_transform"][:3, 3], 3))
