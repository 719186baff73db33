"""BASELINE config 3 (reduced-precision model, registration-quality parity): rdmnet_b200.set_precision("tf32") switches the dense
contractions to ONE tensor-core product per k-step (gemm_tc_fast.cu, 10-bit mantissa like fp16) while geometry, normalisation,
Sinkhorn and the pose solver stay fp32. Checked here:
  * the tf32 GEMM against fp64 on the backbone's shapes, at the accuracy a 10-bit mantissa allows (and that it really ran);
  * the registration metrics of experiments/eval.py:221-231 (RR with RRE < 5 deg and RTE < 2 m, mean RRE / RTE over the registered
    pairs) on synthetic KITTI-shaped pairs with known ground truth, tf32 mode vs the fp32 mode: same RR, mean RRE within
    0.02 deg, mean RTE within 0.5 cm; the bundled KITTI pairs (no ground truth) must stay registered with
    respect to the reference's fp32 pose."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(300, 128, 64), (23319, 32, 64), (8841, 64, 960), (494, 512, 7680), (494, 2048, 512), (3078, 257, 768), (129, 72, 40),
          (2236, 1024, 1284), (30129, 64, 480)]


@pytest.fixture()
def tf32_mode():
    import rdmnet_b200
    rdmnet_b200.set_precision("tf32")
    try:
        yield
    finally:
        rdmnet_b200.set_precision("fp32")


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_linear_tf32_single_pass_vs_fp64(m, n, k, tf32_mode):
    import rdmnet_b200
    from rdmnet_b200 import ops, _lib
    assert rdmnet_b200.get_precision() == "tf32"
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5
    b = torch.randn(n)
    ref = x.double() @ w.double().t() + b.double()
    n0 = _lib.lib().rdm_tc_gemm_count()
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), act=0)
    torch.cuda.synchronize()
    assert _lib.lib().rdm_tc_gemm_count() > n0, "tensor-core path not taken"
    err = (got.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    # operands truncated to 10 mantissa bits: |rel err per product| <= 2^-10 each side, random signs over K terms
    print(f"[precision] tf32 {m}x{n}x{k}: max err / max|ref| = {err:.2e}")
    assert 1e-6 < err <= 3e-3, err  # > 1e-6: it is NOT the 3-term split
    # with ReLU and without bias, ragged N
    got2 = ops.linear(x.cuda(), w.cuda(), None, act=2).cpu().double()
    ref2 = torch.relu(x.double() @ w.double().t())
    assert (got2 - ref2).abs().max().item() / ref2.abs().max().item() <= 3e-3


def _pose_err(T, Tg):
    T, Tg = np.asarray(T, np.float64), np.asarray(Tg, np.float64)
    rre = np.degrees(np.arccos(np.clip((np.trace(T[:3, :3].T @ Tg[:3, :3]) - 1) / 2, -1, 1)))
    return rre, float(np.linalg.norm(T[:3, 3] - Tg[:3, 3]))


def test_registration_metrics_tf32_vs_fp32(pretrained_state, scans, golden_pairs):
    import rdmnet_b200
    from rdmnet_b200 import synthetic
    from rdmnet_b200.model import create_model
    model = create_model()
    model.load_state_dict(pretrained_state, strict=True)
    model = model.cuda().eval()
    cases = []
    for pid in range(6):
        p = synthetic.make_pair(pair_id=pid)
        cases.append((f"synthetic-{pid}", p["ref_points"], p["src_points"], p["transform"]))
    for name, (ne, na) in synthetic.SIZE_CLASSES.items():
        p = synthetic.make_pair(pair_id=11, n_elev=ne, n_azim=na)
        cases.append((f"synthetic-{name}", p["ref_points"], p["src_points"], p["transform"]))
    for tag, a, b in (("p04", "s000000", "s000004"), ("p07", "s000000", "s000007")):
        cases.append((f"kitti-{tag}", scans[a], scans[b], golden_pairs[f"{tag}_estimated_transform"]))
    res = {}
    try:
        for mode in ("fp32", "tf32"):
            rdmnet_b200.set_precision(mode)
            for name, ref, src, Tg in cases:
                pts = torch.from_numpy(np.concatenate([ref, src])).cuda()
                lens = torch.tensor([len(ref), len(src)], dtype=torch.int64).cuda()
                out = model({"points": pts, "lengths": lens})
                rre, rte = _pose_err(out["estimated_transform"].cpu().numpy(), Tg)
                res[(mode, name)] = (rre, rte, int(out["corr_scores"].shape[0]))
    finally:
        rdmnet_b200.set_precision("fp32")
    synth = [c[0] for c in cases if c[0].startswith("synthetic")]
    summ = {}
    for mode in ("fp32", "tf32"):
        ok = [(res[(mode, n)][0], res[(mode, n)][1]) for n in synth if res[(mode, n)][0] < 5.0 and res[(mode, n)][1] < 2.0]  # config.py:66-67
        summ[mode] = (len(ok) / len(synth), float(np.mean([r for r, _ in ok])), float(np.mean([t for _, t in ok])))
    for name, *_ in cases:
        a, b = res[("fp32", name)], res[("tf32", name)]
        print(f"[precision] {name}: fp32 RRE {a[0]:.4f} deg RTE {100 * a[1]:.2f} cm ({a[2]} corr) | tf32 RRE {b[0]:.4f} deg RTE {100 * b[1]:.2f} cm ({b[2]} corr)")
    print(f"[precision] synthetic pairs (true ground truth), RR / mean RRE / mean RTE: fp32 {summ['fp32'][0]:.3f} / {summ['fp32'][1]:.4f} deg / "
          f"{100 * summ['fp32'][2]:.2f} cm; tf32 {summ['tf32'][0]:.3f} / {summ['tf32'][1]:.4f} deg / {100 * summ['tf32'][2]:.2f} cm")
    assert summ["fp32"][0] == summ["tf32"][0] == 1.0
    assert abs(summ["fp32"][1] - summ["tf32"][1]) <= 0.02
    assert abs(summ["fp32"][2] - summ["tf32"][2]) <= 0.005
    # The two bundled KITTI pairs have no ground truth; their "truth" here is the reference's own fp32 estimate, and that estimate
    # rests on a flat optimum: only 37 / 71 of the 413 / 504 correspondences are within 1 m of ANY of the candidate poses
    # (scripts/diag_precision.py), so a 1e-3 feature perturbation may move the pose by decimetres at an unchanged inlier count
    # (measured: (0,7) moves 31 cm / 1.5 deg with 71 inliers before and after). Gate: still registered w.r.t. the reference's pose.
    for name in ("kitti-p04", "kitti-p07"):
        assert res[("tf32", name)][0] < 5.0 and res[("tf32", name)][1] < 2.0, (name, res[("tf32", name)])
        assert res[("fp32", name)][0] < 0.1 and res[("fp32", name)][1] < 0.01, (name, res[("fp32", name)])
