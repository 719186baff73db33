"""GPU parity: transformer / vote / NMS / partition / coarse matching / Sinkhorn / Procrustes / LGR kernels vs
(a) the reference's own outputs (tests/golden/modules_small.npz) and (b) the torch-fp32 CPU oracle on larger inputs."""
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as MO
from oracle import pyramid as OP

pytestmark = pytest.mark.gpu


def close(got, ref, tol=1e-4):
    got = got.detach().cpu().double() if torch.is_tensor(got) else torch.as_tensor(got).double()
    ref = ref.detach().cpu().double() if torch.is_tensor(ref) else torch.as_tensor(ref).double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    assert err <= tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def sub(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.ascontiguousarray(v)) for k, v in g.items() if k.startswith(prefix)}


def T(g, k):
    return torch.from_numpy(np.ascontiguousarray(g[k])).cuda()


def test_thdroformer_vs_reference_golden(golden_small):
    from rdmnet_b200.modules import ThDRoFormer
    g = golden_small
    tf = ThDRoFormer(48, 40, 32, 4, 2)
    tf.load_state_dict(sub(g, "tf."), strict=True)
    tf = tf.cuda()
    with torch.no_grad():
        ro, so = tf(T(g, "tf_rp")[None], T(g, "tf_sp")[None], T(g, "tf_rf")[None], T(g, "tf_sf")[None])
    close(ro[0], torch.from_numpy(g["tf_ro"]))
    close(so[0], torch.from_numpy(g["tf_so"]))


def test_thdroformer_full_size_vs_oracle():
    from rdmnet_b200.modules import ThDRoFormer
    torch.manual_seed(3)
    tf = ThDRoFormer(2048, 256, 128, 4, 4)
    sd = {k: v.clone() for k, v in tf.state_dict().items()}
    rp, sp = torch.randn(431, 3) * 20, torch.randn(411, 3) * 20
    rf, sf = torch.randn(431, 2048), torch.randn(411, 2048)
    with torch.no_grad():
        ro, so = MO.thdroformer(sd, "", rp, sp, rf, sf)
        go, gs = tf.cuda()(rp.cuda(), sp.cuda(), rf.cuda(), sf.cuda())
    close(go, ro)
    close(gs, so)


def test_vote_nms_vs_reference_golden(golden_small):
    from rdmnet_b200.modules import NMS, Vote_layer
    g = golden_small
    cfg = types.SimpleNamespace(MLPS=[64, 32], MAX_TRANSLATE_RANGE=[3.0, 3.0, 3.0], input_feats_dim=32, NMS_radius=2.4)
    vl = Vote_layer(cfg, 1)
    vl.load_state_dict(sub(g, "vl."), strict=True)
    with torch.no_grad():
        x, f = vl.cuda()(T(g, "vl_xyz"), T(g, "vl_f"))
    close(x, torch.from_numpy(g["vl_oxyz"]))
    close(f, torch.from_numpy(g["vl_of"]))
    nms = NMS(cfg, [9, 9, 9, 9, int(g["nms_limit"])])
    mask = nms(T(g, "nms_nodes"), T(g, "nms_len"))
    assert mask.dtype == torch.bool
    assert np.array_equal(mask.cpu().numpy(), g["nms_mask"])


def test_nms_larger_vs_oracle():
    from rdmnet_b200 import ops
    rng = np.random.default_rng(4)
    nodes = ((rng.random((842, 3)) - 0.5) * [60, 40, 3]).astype(np.float32)
    lens = np.array([431, 411], np.int64)
    nb = OP.radius_search(nodes, nodes, lens, lens, 2.4, 81)
    ref = MO.nms_greedy(nb).numpy()
    got = ops.nms(torch.from_numpy(nb).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
    assert 0 < got.sum() < 842


def test_partition_and_coarse_matching_vs_reference_golden(golden_small):
    from rdmnet_b200 import ops
    g = golden_small
    p2n, nm, knn, km = ops.point_to_node_partition(T(g, "part_pts"), T(g, "part_nodes"), 16)
    assert np.array_equal(p2n.cpu().numpy(), g["part_p2n"]) and np.array_equal(nm.cpu().numpy(), g["part_nm"])
    assert np.array_equal(knn.cpu().numpy(), g["part_knn"]) and np.array_equal(km.cpu().numpy(), g["part_km"])
    ri, si, sc = ops.coarse_matching(T(g, "spm_rf"), T(g, "spm_sf"), T(g, "spm_rm"), T(g, "spm_sm"), 50)
    assert np.array_equal(ri.cpu().numpy(), g["spm_ri"]) and np.array_equal(si.cpu().numpy(), g["spm_si"])
    close(sc, torch.from_numpy(g["spm_sc"]), 1e-5)


def test_partition_full_size_vs_oracle(scans):
    from rdmnet_b200 import ops
    pts = scans["s000004"]
    p1, l1 = OP.grid_subsample(pts, np.array([len(pts)]), 0.6)
    rng = np.random.default_rng(1)
    nodes = p1[rng.choice(len(p1), 215, replace=False)] + rng.normal(0, 0.5, (215, 3)).astype(np.float32)
    ref = MO.point_to_node_partition(torch.from_numpy(p1), torch.from_numpy(nodes), 128)
    got = ops.point_to_node_partition(torch.from_numpy(p1).cuda(), torch.from_numpy(nodes).cuda(), 128)
    assert (got[0].cpu() == ref[0]).float().mean() > 0.9995  # argmin near-ties may flip
    same_rows = (got[2].cpu() == ref[2]).all(1).float().mean().item()
    assert same_rows > 0.98, same_rows
    assert np.array_equal(got[1].cpu().numpy(), ref[1].numpy())


def test_coarse_matching_full_size_vs_oracle():
    from rdmnet_b200 import ops
    torch.manual_seed(0)
    rf = torch.nn.functional.normalize(torch.randn(215, 256), dim=1)
    sf = torch.nn.functional.normalize(rf[torch.randperm(215)[:197]] + 0.3 * torch.randn(197, 256), dim=1)
    rm, sm = torch.rand(215) > 0.05, torch.rand(197) > 0.05
    ri, si, sc = MO.superpoint_matching(rf, sf, rm, sm)
    gi, gj, gs = ops.coarse_matching(rf.cuda(), sf.cuda(), rm.cuda(), sm.cuda(), 256)
    close(gs, sc, 1e-4)
    agree = ((gi.cpu() == ri) & (gj.cpu() == si)).float().mean().item()
    assert agree > 0.98, agree


def test_sinkhorn_vs_reference_golden_and_oracle(golden_small):
    from rdmnet_b200 import ops
    g = golden_small
    alpha = torch.tensor(1.6727).cuda()
    o = ops.sinkhorn(T(g, "ot_in"), T(g, "ot_rm"), T(g, "ot_cm"), alpha, 100).cpu().numpy()
    ref = g["ot_out"]
    live = ref > -1e11
    assert np.array_equal(o > -1e11, live)
    close(o[live], ref[live], 1e-4)
    torch.manual_seed(1)
    s = torch.randn(16, 128, 128) * 3
    rm, cm = torch.rand(16, 128) > 0.3, torch.rand(16, 128) > 0.3
    ref = MO.sinkhorn(s, rm, cm, torch.tensor(1.6727)).numpy()
    o = ops.sinkhorn(s.cuda(), rm.cuda(), cm.cuda(), alpha, 100).cpu().numpy()
    live = ref > -1e11
    close(o[live], ref[live], 1e-4)


def test_patch_scores_vs_torch():
    from rdmnet_b200 import ops
    torch.manual_seed(2)
    fr, fs = torch.randn(900, 256), torch.randn(800, 256)
    rk = torch.randint(0, 901, (40, 128))
    sk = torch.randint(0, 801, (35, 128))
    ri, si = torch.randint(0, 40, (64,)), torch.randint(0, 35, (64,))
    frp, fsp = torch.cat([fr, torch.zeros(1, 256)]), torch.cat([fs, torch.zeros(1, 256)])
    ref = torch.einsum("bnd,bmd->bnm", frp[rk[ri]], fsp[sk[si]]) / 16.0
    got = ops.patch_scores(fr.cuda(), fs.cuda(), rk.cuda(), sk.cuda(), ri.cuda(), si.cuda())
    close(got, ref, 2e-5)


def test_procrustes_lgr_vs_reference_golden(golden_small):
    from rdmnet_b200 import ops
    from rdmnet_b200.modules import LocalGlobalRegistration
    g = golden_small
    Tm = ops.weighted_procrustes(T(g, "wp_src"), T(g, "wp_ref"), T(g, "wp_w"), return_transform=True)
    close(Tm, torch.from_numpy(g["wp_T"]), 1e-4)
    lgr = LocalGlobalRegistration(1, 0.6, mutual=False, confidence_threshold=0, use_dustbin=True, use_global_score=False,
                                  correspondence_threshold=3, correspondence_limit=None, num_refinement_steps=5)
    rc, sc, cs, Te = lgr(T(g, "lgr_rk"), T(g, "lgr_sk"), T(g, "lgr_rm"), T(g, "lgr_sm"), T(g, "lgr_scores"), None)
    assert np.array_equal(rc.cpu().numpy(), g["lgr_rc"]) and np.array_equal(sc.cpu().numpy(), g["lgr_sc"])
    close(cs, torch.from_numpy(g["lgr_cs"]), 1e-5)
    close(Te, torch.from_numpy(g["lgr_T"]), 1e-4)


def test_procrustes_degenerate_and_reflection():
    from rdmnet_b200 import ops
    torch.manual_seed(7)
    src = torch.randn(5, 30, 3)
    src[1, :, 2] = 0  # planar
    src[2] = src[2, :1]  # all points identical -> H = 0
    R = torch.linalg.qr(torch.randn(3, 3))[0]
    if torch.det(R) < 0:
        R[:, 0] = -R[:, 0]
    ref = src @ R.t() + torch.tensor([0.3, -1.0, 2.0])
    ref[3] = src[3] * torch.tensor([1.0, 1.0, -1.0])  # mirrored: the det fix must kick in
    w = torch.rand(5, 30)
    got = ops.weighted_procrustes(src.cuda(), ref.cuda(), w.cuda(), return_transform=True).cpu()
    exp = MO.weighted_procrustes(src, ref, w)
    for b in (0, 1, 3, 4):
        close(got[b], exp[b], 2e-4)
    assert torch.isfinite(got).all()
    for b in range(5):
        assert abs(torch.det(got[b, :3, :3]).item() - 1.0) < 1e-4  # always a proper rotation
