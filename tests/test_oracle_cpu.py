"""CPU tests: pin the oracle (oracle/) against the reference - its own C++ core compiled in place
(oracle/_ref) and the golden fixtures produced by importing the reference's Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as MO
from oracle import pyramid as OP


def sub(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def close(a, b, tol=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


needs_ref = pytest.mark.skipif(not OP.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("seed,n,voxel", [(0, 5000, 0.5), (1, 20000, 0.3), (2, 300, 2.0), (3, 1, 1.0), (4, 40000, 0.11)])
def test_grid_subsample_port_vs_reference(seed, n, voxel):
    rng = np.random.default_rng(seed)
    n2 = max(1, n // 2)
    pts = ((rng.random((n + n2, 3)) - 0.5) * [60, 40, 6]).astype(np.float32)
    lens = np.array([n, n2])
    p, l = OP.grid_subsample(pts, lens, voxel, "port")
    r, rl = OP.grid_subsample(pts, lens, voxel, "ref")
    assert np.array_equal(l, rl)
    assert np.array_equal(p.view(np.uint32), r.view(np.uint32))  # bit-exact incl. hashtable order


@needs_ref
def test_pyramid_port_vs_reference_bundled(scans):
    pts = np.concatenate([scans["s000000"], scans["s000004"]])
    lens = np.array([len(scans["s000000"]), len(scans["s000004"])])
    v = 0.6
    for _ in range(4):
        p, l = OP.grid_subsample(pts, lens, v, "port")
        r, rl = OP.grid_subsample(pts, lens, v, "ref")
        assert np.array_equal(l, rl) and np.array_equal(p.view(np.uint32), r.view(np.uint32))
        pts, lens, v = p, l, v * 2
    assert list(lens) == [431, 411]


@needs_ref
def test_radius_port_vs_reference(scans):
    a, b = scans["s000000"][::3], scans["s000007"][::3]
    s = np.concatenate([a, b])
    sl = np.array([len(a), len(b)])
    q, ql = OP.grid_subsample(s, sl, 0.9, "ref")
    for (qq, qql, ss, ssl, r) in [(s, sl, s, sl, 1.3), (q, ql, s, sl, 1.3), (s, sl, q, ql, 2.6)]:
        mine = OP.radius_neighbors(qq, ss, qql, ssl, r, "port")
        ref = OP.radius_neighbors(qq, ss, qql, ssl, r, "ref")
        assert mine.shape == ref.shape
        assert np.array_equal(np.sort(mine, 1), np.sort(ref, 1))  # same neighbour sets, same padding
        assert np.array_equal(mine, OP.canonicalize_ties(ref, qq, ss))  # same order outside exact-d2 ties


def test_radius_edge_cases():
    q = np.zeros((3, 3), np.float32)
    s = np.array([[0, 0, 0], [1, 0, 0], [0.5, 0, 0], [5, 5, 5]], np.float32)
    out = OP.radius_neighbors(q, s, np.array([2, 1]), np.array([3, 1]), 1.0, "port")
    # strict d2 < r2: the point at distance exactly 1.0 is excluded; second cloud has no neighbour -> all pad
    assert out.tolist() == [[0, 2], [0, 2], [4, 4]]


def test_backbone_blocks_vs_reference(golden_small):
    g = golden_small
    pts, p1 = torch.from_numpy(g["pyr_points0"]), torch.from_numpy(g["pyr_points1"])
    nb, sb = torch.from_numpy(g["pyr_nb0"]), torch.from_numpy(g["pyr_sub0"])
    x0 = torch.ones(pts.shape[0], 1)
    x1 = MO.conv_block(sub(g, "cb."), "", x0, pts, pts, nb, 0.7, 32)
    close(x1, g["cb_out"])
    x2 = MO.residual_block(sub(g, "rb."), "", torch.from_numpy(g["cb_out"]), pts, pts, nb, 0.7, 32, False)
    close(x2, g["rb_out"])
    x3 = MO.residual_block(sub(g, "rs."), "", torch.from_numpy(g["rb_out"]), p1, pts, sb, 0.7, 32, True)
    close(x3, g["rs_out"])
    close(MO.unary(sub(g, "ub."), "", torch.from_numpy(g["ub_in"]), 32), g["ub_out"])


def test_thdroformer_vs_reference(golden_small):
    g = golden_small
    ro, so = MO.thdroformer(sub(g, "tf."), "", *(torch.from_numpy(g[k]) for k in ("tf_rp", "tf_sp", "tf_rf", "tf_sf")))
    close(ro, g["tf_ro"])
    close(so, g["tf_so"])


def test_vote_nms_vs_reference(golden_small):
    g = golden_small
    x, f = MO.vote_layer(sub(g, "vl."), "", torch.from_numpy(g["vl_xyz"]), torch.from_numpy(g["vl_f"]))
    close(x, g["vl_oxyz"])
    close(f, g["vl_of"])
    nodes, ln = g["nms_nodes"], g["nms_len"]
    nb = OP.radius_search(nodes, nodes, ln, ln, 2.4, int(g["nms_limit"]))
    assert np.array_equal(MO.nms_greedy(nb).numpy(), g["nms_mask"])


def test_matching_vs_reference(golden_small):
    g = golden_small
    p2n, nm, knn, km = MO.point_to_node_partition(torch.from_numpy(g["part_pts"]), torch.from_numpy(g["part_nodes"]), 16)
    assert np.array_equal(p2n.numpy(), g["part_p2n"]) and np.array_equal(nm.numpy(), g["part_nm"])
    assert np.array_equal(knn.numpy(), g["part_knn"]) and np.array_equal(km.numpy(), g["part_km"])
    ri, si, sc = MO.superpoint_matching(*(torch.from_numpy(g[k]) for k in ("spm_rf", "spm_sf", "spm_rm", "spm_sm")),
                                        num_corr=50)
    assert np.array_equal(ri.numpy(), g["spm_ri"]) and np.array_equal(si.numpy(), g["spm_si"])
    close(sc, g["spm_sc"], 1e-6)
    o = MO.sinkhorn(torch.from_numpy(g["ot_in"]), torch.from_numpy(g["ot_rm"]), torch.from_numpy(g["ot_cm"]),
                    torch.tensor(1.6727))
    ref = g["ot_out"]
    live = ref > -1e11
    assert np.array_equal(o.numpy() > -1e11, live)
    close(o.numpy()[live], ref[live], 1e-5)


def test_procrustes_lgr_vs_reference(golden_small):
    g = golden_small
    T = MO.weighted_procrustes(*(torch.from_numpy(g[k]) for k in ("wp_src", "wp_ref", "wp_w")))
    close(T, g["wp_T"], 1e-5)
    rc, sc, cs, Te, _ = MO.lgr(*(torch.from_numpy(g[k]) for k in ("lgr_rk", "lgr_sk", "lgr_rm", "lgr_sm", "lgr_scores")))
    assert np.array_equal(rc.numpy(), g["lgr_rc"]) and np.array_equal(sc.numpy(), g["lgr_sc"])
    close(cs, g["lgr_cs"], 1e-6)
    close(Te, g["lgr_T"], 1e-5)


@needs_ref
def test_end_to_end_vs_reference_pair04(scans, golden_pairs, pretrained_state):
    """Whole forward of the oracle on bundled pair (0,4) with the pretrained checkpoint vs the reference's outputs."""
    g = golden_pairs
    a, b = scans["s000000"], scans["s000004"]
    pyr = OP.precompute_pyramid(np.concatenate([a, b]), [len(a), len(b)], 5, 0.3, 4.25 * 0.3, MO.DEFAULT_LIMITS, "ref")
    assert np.array_equal(np.stack(pyr["lengths"]), g["p04_lengths"])
    tp = MO.pyramid_to_torch(pyr)

    def nms_search(p, l):
        return OP.radius_search(p.numpy(), p.numpy(), l.numpy(), l.numpy(), 2.4, MO.DEFAULT_LIMITS[-1], "ref")

    torch.set_num_threads(8)
    with torch.no_grad():
        out = MO.forward(pretrained_state, tp, nms_search)
    close(out["feats_s5"][:64, :64], g["p04_feats_s5_head"], 2e-4)
    close(out["ref_feats_t1"][:64, :64], g["p04_t1_ref_head"], 2e-4)
    close(out["shifted_points_c"], g["p04_shifted"], 1e-4)
    assert np.array_equal(out["nms_masks"].numpy(), g["p04_nms"])
    close(out["ref_feats_c"], g["p04_ref_feats_c"], 2e-4)
    assert np.array_equal(out["ref_node_corr_indices"].numpy(), g["p04_ref_node_corr_indices"])
    assert np.array_equal(out["src_node_corr_indices"].numpy(), g["p04_src_node_corr_indices"])
    assert np.array_equal(out["ref_corr_points"].numpy(), g["p04_ref_corr_points"])
    assert np.array_equal(out["src_corr_points"].numpy(), g["p04_src_corr_points"])
    close(out["corr_scores"], g["p04_corr_scores"], 1e-3)
    close(out["estimated_transform"], g["p04_estimated_transform"], 1e-4)


def test_ingest_oracle_properties():
    """oracle/ingest_oracle.py (open3d voxel_down_sample semantics, parity unpinned against open3d itself: the package is absent):
    one output per occupied voxel, each the mean of exactly its points, voxels in first-occurrence order."""
    from oracle import ingest_oracle as IO
    rng = np.random.default_rng(5)
    p = np.concatenate([rng.normal(0, 3, (4000, 3)), rng.random((4000, 1))], 1).astype(np.float32)
    out = IO.voxel_downsample(p, 0.3)
    o = p[:, :3].min(0) - np.float32(0.15)
    idx = np.floor((p[:, :3] - o) / np.float32(0.3)).astype(np.int64)
    uniq, first, inv = np.unique(idx, axis=0, return_index=True, return_inverse=True)
    assert out.shape == (uniq.shape[0], 4)
    order = np.argsort(first)
    k = order[0]  # the first voxel of the output is the voxel of point 0
    assert np.array_equal(idx[0], uniq[k])
    assert np.allclose(out[0], p[inv.reshape(-1) == k].astype(np.float64).mean(0), atol=1e-6)
    assert abs(out[:, 3].mean() - 0.5) < 0.05  # intensities are averaged like coordinates
