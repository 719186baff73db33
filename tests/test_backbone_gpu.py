"""GPU parity: KPConv / GEMM / GroupNorm / LayerNorm / pooling kernels vs the torch-fp32 CPU oracle.
Tolerance (north_star): features within 1e-4 relative (scaled by the tensor's max magnitude)."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as MO
from oracle import pyramid as OP

pytestmark = pytest.mark.gpu
TOL = 1e-4


def close(got, ref, tol=TOL):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), f"max abs err {err} (ref max {ref.abs().max().item()})"


def make_neighbors(rng, m, n, h, fill=0.7):
    idx = rng.integers(0, n, size=(m, h))
    k = rng.integers(1, h + 1, size=m) if fill < 1 else np.full(m, h)
    idx[np.arange(h)[None, :] >= k[:, None]] = n
    return torch.from_numpy(idx.astype(np.int64))


@pytest.mark.parametrize("m,n,k,nk", [(300, 128, 64, True), (1000, 257, 768, True), (77, 1, 256, True),
                                       (842, 512, 7680, False), (5000, 64, 480, False), (2236, 1024, 1281, True),
                                       (130, 130, 37, False), (1, 259, 256, True)])
def test_linear(m, n, k, nk):
    from rdmnet_b200 import ops, _lib
    tc0 = _lib.lib().rdm_tc_gemm_count()
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5 if nk else torch.randn(k, n) / k ** 0.5
    b = torch.randn(n)
    ref = (x.double() @ (w.double().t() if nk else w.double()) + b.double()).float()
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), weight_is_kn=not nk)
    close(got, ref, 2e-5)
    got = ops.linear(x.cuda(), w.cuda(), None, weight_is_kn=not nk)
    close(got, ref - b, 2e-5)
    # weights with 16-byte row strides run on the tensor cores (tcgen05 tf32 x3) - [K,N]-layout ones through a transposed copy in
    # the workspace -, the rest on SIMT
    expect_tc = m >= 64 and n >= 8 and k % 4 == 0
    assert (_lib.lib().rdm_tc_gemm_count() - tc0 == 2) == expect_tc


@pytest.mark.parametrize("m,n,k", [(23319, 32, 64), (23319, 128, 32), (8841, 64, 960), (494, 512, 7680), (494, 2048, 512),
                                   (431, 128, 2048), (3078, 257, 768), (129, 72, 40)])
def test_linear_tensor_core_shapes(m, n, k):
    """The GEMM shapes of the backbone on the tcgen05 path (split-K included) vs an fp64 reference."""
    from rdmnet_b200 import ops, _lib
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5
    b = torch.randn(n)
    ref = (x.double() @ w.double().t() + b.double()).float()
    tc0 = _lib.lib().rdm_tc_gemm_count()
    for act in (0, 1):
        got = ops.linear(x.cuda(), w.cuda(), b.cuda(), act=act)
        close(got, torch.nn.functional.leaky_relu(ref, 0.1) if act else ref, 2e-5)
    assert _lib.lib().rdm_tc_gemm_count() - tc0 == 2


@pytest.mark.parametrize("cin,cout,m,n,h,idt", [(1, 64, 500, 500, 33, torch.int64), (32, 32, 700, 900, 65, torch.int64),
                                                 (64, 64, 300, 300, 63, torch.int32), (128, 128, 200, 250, 69, torch.int64),
                                                 (256, 256, 90, 120, 70, torch.int64), (512, 512, 50, 50, 81, torch.int32),
                                                 (48, 20, 100, 100, 17, torch.int64),
                                                 # large M: the non-split warp mappings of the main gather kernel
                                                 (32, 32, 10000, 10000, 40, torch.int32), (64, 64, 5000, 6000, 33, torch.int32),
                                                 (128, 128, 2500, 2500, 35, torch.int64), (256, 256, 1300, 1300, 37, torch.int32)])
def test_kpconv_vs_oracle(cin, cout, m, n, h, idt):
    from rdmnet_b200 import ops
    rng = np.random.default_rng(cin + m)
    torch.manual_seed(cin)
    s_pts = torch.from_numpy(((rng.random((n, 3)) - 0.5) * 3).astype(np.float32))
    q_pts = s_pts[:m].clone() if m <= n else torch.from_numpy(((rng.random((m, 3)) - 0.5) * 3).astype(np.float32))
    idx = make_neighbors(rng, m, n, h)
    feats = torch.ones(n, 1) if cin == 1 else torch.randn(n, cin)
    feats[::7] = -feats[::7].abs()  # rows with non-positive sums: excluded from the neighbour count (kpconv.py:113)
    w = torch.randn(15, cin, cout) / (15 * cin) ** 0.5
    kp = torch.randn(15, 3) * 0.6
    kp[0] = 0
    bias = torch.randn(cout)
    ref = MO.kpconv(feats, q_pts, s_pts, idx, w, kp, 0.9, bias)
    got = ops.kpconv(feats.cuda(), q_pts.cuda(), s_pts.cuda(), idx.to(idt).cuda(), w.cuda(), kp.cuda(), 0.9, bias.cuda())
    close(got, ref)


def test_groupnorm_layernorm_pooling():
    from rdmnet_b200 import ops
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    for n, c in [(1000, 64), (333, 256), (50, 2048), (7, 32)]:
        x = torch.randn(n, c) * 3 + 1
        g, b, r = torch.randn(c), torch.randn(c), torch.randn(n, c)
        ref = MO.group_norm(x, g, b, 32)
        close(ops.group_norm(x.cuda(), g.cuda(), b.cuda(), 32), ref)
        close(ops.group_norm(x.cuda(), g.cuda(), b.cuda(), 32, act=1), torch.nn.functional.leaky_relu(ref, 0.1))
        close(ops.group_norm(x.cuda(), g.cuda(), b.cuda(), 32, residual=r.cuda(), act=1),
              torch.nn.functional.leaky_relu(ref + r, 0.1))
        lref = torch.nn.functional.layer_norm(x + r, (c,), g, b)
        close(ops.layer_norm(x.cuda(), g.cuda(), b.cuda(), residual=r.cuda()), lref)
        close(ops.layer_norm(x.cuda(), g.cuda(), b.cuda(), relu=True), torch.relu(torch.nn.functional.layer_norm(x, (c,), g, b)))
    x = torch.randn(400, 128)
    idx = make_neighbors(rng, 150, 400, 40)
    close(ops.maxpool(x.cuda(), idx.cuda()), MO.maxpool(x, idx), 0)
    close(ops.maxpool(x.cuda(), idx.int().cuda()), MO.maxpool(x, idx), 0)
    skip = torch.randn(150, 36)
    x2 = torch.randn(400, 257)
    ref = torch.cat([MO.nearest_upsample(x2, idx), skip], 1)
    close(ops.nearest_upsample_concat(x2.cuda(), idx.cuda(), skip.cuda()), ref, 0)
    close(ops.activation(x.cuda(), 3), torch.sigmoid(x).clamp(0, 1), 1e-6)


def sub(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.ascontiguousarray(v)) for k, v in g.items() if k.startswith(prefix)}


def test_blocks_vs_reference_golden(golden_small):
    """ConvBlock / ResidualBlock (plain and strided) of the product host code vs the REFERENCE's outputs."""
    from rdmnet_b200 import modules as M
    g = golden_small
    pts, p1 = torch.from_numpy(g["pyr_points0"]).cuda(), torch.from_numpy(g["pyr_points1"]).cuda()
    nb, sb = torch.from_numpy(g["pyr_nb0"]).cuda(), torch.from_numpy(g["pyr_sub0"]).cuda()
    cb = M.ConvBlock(1, 64, 15, 1.5, 0.7, 32)
    cb.load_state_dict(sub(g, "cb."), strict=True)
    rb = M.ResidualBlock(64, 128, 15, 1.5, 0.7, 32)
    rb.load_state_dict(sub(g, "rb."), strict=True)
    rs = M.ResidualBlock(128, 128, 15, 1.5, 0.7, 32, strided=True)
    rs.load_state_dict(sub(g, "rs."), strict=True)
    cb, rb, rs = cb.cuda(), rb.cuda(), rs.cuda()
    with torch.no_grad():
        x1 = cb(torch.ones(pts.shape[0], 1, device="cuda"), pts, pts, nb)
        close(x1, torch.from_numpy(g["cb_out"]))
        x2 = rb(torch.from_numpy(np.ascontiguousarray(g["cb_out"])).cuda(), pts, pts, nb)
        close(x2, torch.from_numpy(g["rb_out"]))
        x3 = rs(torch.from_numpy(np.ascontiguousarray(g["rb_out"])).cuda(), p1, pts, sb)
        close(x3, torch.from_numpy(g["rs_out"]))
        ub = M.UnaryBlock(40, 64, 32)
        ub.load_state_dict(sub(g, "ub."), strict=True)
        close(ub.cuda()(torch.from_numpy(g["ub_in"]).cuda()), torch.from_numpy(g["ub_out"]))


def test_maxpool_reference_row_width():
    """A row that fills the whole reference width (max_count < limit) must NOT compete with the zero row: columns beyond the
    reference's tensor carry the sentinel N + 1 (rdm_mark_reference_width) and are skipped (kpconv/functional.py:54-67 on the
    narrowed table of ops/radius_search.py:25-26)."""
    import numpy as np
    from oracle import model_oracle as MO
    from rdmnet_b200 import ops
    rng = np.random.default_rng(1)
    n, m, w, limit, c = 50, 20, 6, 10, 8
    x = -torch.rand(n, c) - 0.1  # all negative: the zero row wins wherever it takes part
    narrow = torch.from_numpy(rng.integers(0, n, size=(m, w)).astype(np.int64))
    narrow[5:, 4:] = n  # rows 5.. have real padding; rows 0..4 fill the reference width
    wide = torch.full((m, limit), n + 1, dtype=torch.int64)
    wide[:, :w] = narrow
    ref = MO.maxpool(x, narrow)
    got = ops.maxpool(x.cuda(), wide.cuda()).cpu()
    assert torch.equal(got, ref)
    assert (ref[:5] < 0).all() and (ref[5:] == 0).all()
    got32 = ops.maxpool(x.cuda(), wide.to(torch.int32).cuda()).cpu()
    assert torch.equal(got32, ref)
