"""GPU parity: grid_subsample / radius_search (CUDA, through the C ABI) vs the CPU oracle (oracle/pyramid_oracle.c)."""
import numpy as np
import pytest
import torch

from oracle import pyramid as OP

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_cloud(rng, n, extent=(60, 40, 6)):
    return ((rng.random((n, 3)) - 0.5) * np.array(extent)).astype(np.float32)


def test_bucket_table_matches_libstdcxx():
    from rdmnet_b200 import _lib as L
    L.call("rdm_selfcheck_bucket_table", 200000)


@pytest.mark.parametrize("seed,sizes,voxel", [
    (0, [5000, 2500], 0.5), (1, [20000, 17000], 0.3), (2, [300, 7, 900], 2.0), (3, [1, 1], 1.0),
    (4, [40000, 35000], 0.11), (5, [64, 64], 100.0), (6, [13, 14, 15, 29, 30, 31], 0.01)])
def test_grid_subsample_bit_exact(seed, sizes, voxel):
    from rdmnet_b200 import ops
    rng = np.random.default_rng(seed)
    pts = rand_cloud(rng, sum(sizes))
    if seed == 2:
        pts[100:200] = pts[100]  # many duplicates in one voxel
    lens = np.array(sizes, np.int64)
    ref_p, ref_l = OP.grid_subsample(pts, lens, voxel, "port")
    p, l = ops.grid_subsample(cu(pts), cu(lens), voxel)
    assert l.cpu().numpy().tolist() == ref_l.tolist()
    got = p.cpu().numpy()
    assert got.shape == ref_p.shape
    assert np.array_equal(got.view(np.uint32), ref_p.view(np.uint32))  # bit-exact, including hashtable order


def test_grid_subsample_bundled_pyramid(scans):
    from rdmnet_b200 import ops
    a, b = scans["s000000"], scans["s000004"]
    pts, lens = np.concatenate([a, b]), np.array([len(a), len(b)], np.int64)
    g, gl = cu(pts), cu(lens)
    v = 0.6
    for _ in range(4):
        pts, lens = OP.grid_subsample(pts, lens, v, "port")
        g, gl = ops.grid_subsample(g, gl, v)
        assert gl.cpu().numpy().tolist() == lens.tolist()
        assert np.array_equal(g.cpu().numpy().view(np.uint32), pts.view(np.uint32))
        v *= 2
    assert lens.tolist() == [431, 411]


def test_grid_subsample_cpu_tensors_and_errors():
    from rdmnet_b200 import ops
    rng = np.random.default_rng(9)
    pts = torch.from_numpy(rand_cloud(rng, 1000))
    lens = torch.tensor([600, 400])
    p, l = ops.grid_subsample(pts, lens, 1.0)  # CPU in -> CPU out, like the reference extension
    assert not p.is_cuda and not l.is_cuda
    rp, rl = OP.grid_subsample(pts.numpy(), lens.numpy(), 1.0, "port")
    assert np.array_equal(p.numpy().view(np.uint32), rp.view(np.uint32))
    with pytest.raises(RuntimeError):
        ops.grid_subsample(pts.double(), lens, 1.0)
    with pytest.raises(RuntimeError):
        ops.grid_subsample(pts, lens.int(), 1.0)
    with pytest.raises(RuntimeError):
        ops.grid_subsample(pts.t(), lens, 1.0)


@pytest.mark.parametrize("seed,nq,ns,radius,limit", [
    (0, [3000, 2000], [3000, 2000], 2.0, 30), (1, [800, 900], [4000, 5000], 3.0, 40),
    (2, [4000, 4100], [700, 650], 5.0, 64), (3, [500, 0, 300], [500, 10, 300], 4.0, 16),
    (4, [2000, 1], [2000, 1], 0.05, 8)])
def test_radius_search_vs_oracle_random(seed, nq, ns, radius, limit):
    from rdmnet_b200 import ops
    rng = np.random.default_rng(seed)
    same = nq == ns
    s = rand_cloud(rng, sum(ns))
    q = s if same else rand_cloud(rng, sum(nq))
    ql, sl = np.array(nq, np.int64), np.array(ns, np.int64)
    ref = OP.radius_search(q, s, ql, sl, radius, limit, "port")
    got = ops.radius_search(cu(q), cu(s), cu(ql), cu(sl), radius, limit)
    assert got.dtype == torch.int64
    assert tuple(got.shape) == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)


def test_radius_search_bundled_all_13(scans):
    """All 13 searches of precompute_data_stack_mode (utils/data.py:13-77) on bundled pair (0,4), exact equality."""
    from rdmnet_b200 import ops
    a, b = scans["s000000"], scans["s000004"]
    lim = [65, 63, 69, 70, 81]
    ref = OP.precompute_pyramid(np.concatenate([a, b]), [len(a), len(b)], 5, 0.3, 4.25 * 0.3, lim, "port")
    P = [cu(p) for p in ref["points"]]
    Ln = [cu(l) for l in ref["lengths"]]
    r = 4.25 * 0.3
    for i in range(5):
        got = ops.radius_search(P[i], P[i], Ln[i], Ln[i], r, lim[i]).cpu().numpy()
        assert np.array_equal(got, ref["neighbors"][i]), f"neighbors[{i}]"
        if i < 4:
            got = ops.radius_search(P[i + 1], P[i], Ln[i + 1], Ln[i], r, lim[i]).cpu().numpy()
            assert np.array_equal(got, ref["subsampling"][i]), f"subsampling[{i}]"
            got = ops.radius_search(P[i], P[i + 1], Ln[i], Ln[i + 1], 2 * r, lim[i + 1]).cpu().numpy()
            assert np.array_equal(got, ref["upsampling"][i]), f"upsampling[{i}]"
        r *= 2


def test_radius_neighbors_full_width_and_wide_limits(scans):
    from rdmnet_b200 import ops
    a = scans["s000007"][::4]
    q = a
    lens = np.array([len(a) // 2, len(a) - len(a) // 2], np.int64)
    ref = OP.radius_neighbors(q, q, lens, lens, 3.0, "port")
    got = ops.radius_neighbors(cu(q), cu(q), cu(lens), cu(lens), 3.0).cpu().numpy()
    assert got.shape == ref.shape and np.array_equal(got, ref)
    # limit above the number of neighbours -> width shrinks to max_count (radius_search.py:25-26)
    got = ops.radius_search(cu(q), cu(q), cu(lens), cu(lens), 3.0, 607).cpu().numpy()
    assert np.array_equal(got, ref)
    # int32 internal tables
    got32, _ = ops.radius_search_raw(cu(q), cu(q), cu(lens), cu(lens), 3.0, 20, index_dtype=torch.int32)
    assert np.array_equal(got32.cpu().numpy().astype(np.int64), ref[:, :20])


def test_radius_search_streaming_topk_overflow():
    """More than 2*KP hits per query: exercises the in-kernel sort-and-cut path."""
    from rdmnet_b200 import ops
    rng = np.random.default_rng(11)
    s = rand_cloud(rng, 3000, (4, 4, 4))
    lens = np.array([3000], np.int64)
    ref = OP.radius_search(s, s, lens, lens, 2.5, 12, "port")
    got = ops.radius_search(cu(s), cu(s), cu(lens), cu(lens), 2.5, 12).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("pair", [("s000000", "s000004"), ("s000000", "s000007")])
def test_build_pyramid_single_call_equals_oracle(scans, pair):
    """rdm_build_pyramid (one host call, one sync, int32 fixed-width tables) == the oracle's pyramid: points bit-exact,
    tables equal over the reference width and padding beyond it, upsampling = column 0, orders are permutations."""
    from rdmnet_b200.model import build_pyramid_gpu
    a, b = scans[pair[0]], scans[pair[1]]
    lim = [65, 63, 69, 70, 81]
    ref = OP.precompute_pyramid(np.concatenate([a, b]), [len(a), len(b)], 5, 0.3, 4.25 * 0.3, lim, "port")
    gp = build_pyramid_gpu(cu(np.concatenate([a, b])), torch.tensor([len(a), len(b)]).cuda(), 5, 0.3, 4.25 * 0.3, lim)
    d = gp.as_data_dict()
    for s in range(5):
        assert gp.lengths_host[s] == [int(x) for x in ref["lengths"][s]]
        assert np.array_equal(d["lengths"][s].cpu().numpy(), ref["lengths"][s])
        assert np.array_equal(d["points"][s].cpu().numpy().view(np.uint32), ref["points"][s].view(np.uint32)), f"points[{s}]"
        order = gp.table("order", s).cpu().numpy()
        assert np.array_equal(np.sort(order), np.arange(ref["points"][s].shape[0])), f"order[{s}]"

    def same(got, want, n_support, name, beyond=None):
        got = got.cpu().numpy()
        w = want.shape[1]
        assert got.shape[0] == want.shape[0] and got.shape[1] >= w, name
        assert np.array_equal(got[:, :w], want), name
        # columns beyond the reference's row width: plain padding (n_support) in the neighbour tables, the "column does not
        # exist" sentinel n_support + 1 in the subsampling tables (what the strided max-pool needs: rdm_mark_reference_width)
        assert (got[:, w:] == (n_support if beyond is None else beyond)).all(), name + " padding"

    for s in range(5):
        same(d["neighbors"][s], ref["neighbors"][s], ref["points"][s].shape[0], f"neighbors[{s}]")
        if s < 4:
            same(d["subsampling"][s], ref["subsampling"][s], ref["points"][s].shape[0], f"subsampling[{s}]",
                 beyond=ref["points"][s].shape[0] + 1)
            if s == 0:
                assert d["upsampling"][0] is None
            else:
                assert np.array_equal(d["upsampling"][s].cpu().numpy()[:, 0], ref["upsampling"][s][:, 0]), f"upsampling[{s}]"


def test_nms_fixpoint_equals_sequential_rule():
    """rdm_nms (parallel fixpoint) == the sequential greedy loop of rdmnet/vote/vote.py:33-40, incl. compaction."""
    from rdmnet_b200 import ops
    rng = np.random.default_rng(5)
    for n, split, r in ((842, 431, 2.4), (3000, 1400, 3.0), (50, 20, 100.0), (1, 1, 1.0)):
        pts = rand_cloud(rng, n, (80, 60, 4))
        lens = np.array([split, n - split], np.int64)
        nb = OP.radius_search(pts, pts, lens, lens, r, 81, "port")
        sel = np.zeros(n + 1, bool)
        for i in range(n):
            if sel[nb[i]].sum() == 0:
                sel[i] = True
        mask, idx, counts = ops.nms(cu(nb), split=split)
        assert np.array_equal(mask.cpu().numpy(), sel[:n])
        c = counts.tolist()
        want = np.nonzero(sel[:n])[0]
        assert c == [int((want < split).sum()), int((want >= split).sum())]
        assert np.array_equal(idx.cpu().numpy()[:len(want)], want)
        assert np.array_equal(ops.nms(cu(nb.astype(np.int32))).cpu().numpy(), sel[:n])
