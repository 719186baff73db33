"""Training path: gradients of the operators with hand-written backward kernels against PyTorch
autograd through the CPU oracle (oracle/model_oracle.py is plain differentiable torch), at the north-star 1e-4 of the
tensor maximum; then one whole training step of the reference's UNMODIFIED experiments/model.py + loss.py on top of the
drop-in (forward in train mode, OverallLoss, backward, Adam step)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_oracle as MO
from oracle import pyramid as OP

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "RDMNet")


def rel(got, ref):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return ((got - ref).abs().max() / ref.abs().max().clamp(min=1e-30)).item()


def check(name, got, ref, tol=1e-4):
    e = rel(got, ref)
    print(f"[grad parity] {name}: {e:.2e} (bar {tol:.0e})")
    assert e <= tol, (name, e)


@pytest.fixture(scope="module")
def cloud():
    rng = np.random.default_rng(11)
    a = ((rng.random((900, 3)) - 0.5) * [14, 12, 3]).astype(np.float32)
    b = ((rng.random((800, 3)) - 0.5) * [14, 12, 3]).astype(np.float32)
    pts = np.concatenate([a, b])
    lens = np.array([900, 800], np.int64)
    p1, l1 = OP.grid_subsample(pts, lens, 0.9, "port")
    nb = OP.radius_search(pts, pts, lens, lens, 1.5, 24, "port")
    sub = OP.radius_search(p1, pts, l1, lens, 1.5, 24, "port")
    return {k: torch.from_numpy(v) for k, v in dict(pts=pts, p1=p1, nb=nb, sub=sub).items()}


def test_kpconv_gradients(cloud):
    from rdmnet_b200 import ops
    torch.manual_seed(0)
    pts, nb = cloud["pts"], cloud["nb"]
    n = pts.shape[0]
    for cin, cout in ((32, 32), (128, 64), (1, 16)):
        f = torch.randn(n, cin) if cin > 1 else torch.rand(n, 1) - 0.2
        w = torch.randn(15, cin, cout) * 0.2
        b = torch.randn(cout) * 0.1
        kp = torch.randn(15, 3) * 0.6
        gout = torch.randn(n, cout)
        fr, wr, br = f.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        yr = MO.kpconv(fr, pts, pts, nb, wr, kp, 0.9, br)
        yr.backward(gout)
        fg, wg, bg = f.cuda().requires_grad_(), w.cuda().requires_grad_(), b.cuda().requires_grad_()
        y = ops.kpconv(fg, pts.cuda(), pts.cuda(), nb.cuda(), wg, kp.cuda(), 0.9, bg)
        y.backward(gout.cuda())
        check(f"kpconv C{cin}->{cout} forward", y, yr)
        check(f"kpconv C{cin}->{cout} d feats", fg.grad, fr.grad)
        check(f"kpconv C{cin}->{cout} d weights", wg.grad, wr.grad)
        check(f"kpconv C{cin}->{cout} d bias", bg.grad, br.grad)


def test_linear_norm_pool_gradients(cloud):
    from rdmnet_b200 import ops
    torch.manual_seed(1)
    n, sub = cloud["pts"].shape[0], cloud["sub"]
    x = torch.randn(n, 48)
    w, b = torch.randn(64, 48) * 0.2, torch.randn(64) * 0.1
    g, be = torch.rand(64) + 0.5, torch.randn(64) * 0.1
    res = torch.randn(n, 64)
    gout = torch.randn(n, 64)
    # Linear -> GroupNorm(+residual) -> LeakyReLU  (UnaryBlock + the ResidualBlock tail, kpconv/modules.py:78-83, 222-224)
    xr, wr, br, gr, ber, rr = [t.clone().requires_grad_() for t in (x, w, b, g, be, res)]
    y = F.leaky_relu(MO.group_norm(F.linear(xr, wr, br), gr, ber, 8) + rr, 0.1)
    y.backward(gout)
    xg, wg, bg, gg, beg, rg = [t.cuda().requires_grad_() for t in (x, w, b, g, be, res)]
    yg = ops.group_norm(ops.linear(xg, wg, bg), gg, beg, 8, residual=rg, act=1)
    check("unary forward", yg, y)
    yg.backward(gout.cuda())
    for name, a_, b_ in (("d x", xg, xr), ("d W", wg, wr), ("d b", bg, br), ("d gamma", gg, gr), ("d beta", beg, ber), ("d residual", rg, rr)):
        check("linear+groupnorm " + name, a_.grad, b_.grad)
    # strided max-pool and nearest upsample + concat
    m = sub.shape[0]
    xr = x.clone().requires_grad_()
    MO.maxpool(xr, sub).backward(gout[:m, :48])
    xg = x.cuda().requires_grad_()
    ops.maxpool(xg, sub.cuda()).backward(gout[:m, :48].cuda())
    check("maxpool d x", xg.grad, xr.grad)
    up = torch.randint(0, m, (n, 3))
    coarse, skip = torch.randn(m, 40), torch.randn(n, 24)
    cr, sr = coarse.clone().requires_grad_(), skip.clone().requires_grad_()
    torch.cat([MO.nearest_upsample(cr, up), sr], 1).backward(gout)
    cg, sg = coarse.cuda().requires_grad_(), skip.cuda().requires_grad_()
    ops.nearest_upsample_concat(cg, up.cuda(), sg).backward(gout.cuda())
    check("upsample d coarse", cg.grad, cr.grad)
    check("upsample d skip", sg.grad, sr.grad)
    # LayerNorm(+residual)+ReLU, clamped sigmoid, index_select
    lw, lb = torch.rand(64) + 0.5, torch.randn(64) * 0.1
    hr, rr, lwr, lbr = [t.clone().requires_grad_() for t in (gout, res, lw, lb)]
    F.relu(F.layer_norm(hr + rr, (64,), lwr, lbr)).backward(x @ torch.randn(48, 64))
    torch.manual_seed(1)
    _ = torch.randn(n, 48), torch.randn(64, 48), torch.randn(64), torch.rand(64), torch.randn(64), torch.randn(n, 64), torch.randn(n, 64)
    hg, rg, lwg, lbg = [t.cuda().requires_grad_() for t in (gout, res, lw, lb)]
    go2 = hr.grad  # reuse shapes: recompute the same upstream gradient deterministically
    yl = ops.layer_norm(hg, lwg, lbg, residual=rg, relu=True)
    hr2, rr2, lwr2, lbr2 = [t.clone().requires_grad_() for t in (gout, res, lw, lb)]
    yr = F.relu(F.layer_norm(hr2 + rr2, (64,), lwr2, lbr2))
    up_g = torch.randn(n, 64)
    yr.backward(up_g)
    yl.backward(up_g.cuda())
    for name, a_, b_ in (("d x", hg, hr2), ("d residual", rg, rr2), ("d gamma", lwg, lwr2), ("d beta", lbg, lbr2)):
        check("layernorm " + name, a_.grad, b_.grad)
    s = torch.randn(500)
    sr_ = s.clone().requires_grad_()
    torch.clamp(torch.sigmoid(sr_), 0, 1).backward(torch.ones(500))
    sg_ = s.cuda().requires_grad_()
    ops.activation(sg_, 3).backward(torch.ones(500).cuda())
    check("sigmoid score d x", sg_.grad, sr_.grad)
    table, idx = torch.randn(300, 16), torch.randint(0, 300, (40, 7))
    tr = table.clone().requires_grad_()
    tr[idx.reshape(-1)].reshape(40, 7, 16).backward(torch.ones(40, 7, 16))
    tg = table.cuda().requires_grad_()
    ops.index_select(tg, idx.cuda(), 0).backward(torch.ones(40, 7, 16).cuda())
    check("index_select d data", tg.grad, tr.grad)


@pytest.mark.parametrize("strided", [False, True])
def test_residual_block_gradients(cloud, strided):
    """Every parameter gradient of a ResidualBlock (unary1 -> KPConv -> GN -> unary2 (+ shortcut [max-pool] [unary])) vs autograd
    through the oracle's residual_block."""
    from rdmnet_b200 import modules as M
    torch.manual_seed(3)
    pts, p1 = cloud["pts"], cloud["p1"]
    idx = cloud["sub"] if strided else cloud["nb"]
    q = p1 if strided else pts
    # 8 groups over the 32 mid channels: with one channel per group (32 groups) GroupNorm cancels the preceding bias exactly
    # and its true gradient is 0 - nothing to compare but rounding noise
    groups = 8
    blk = M.ResidualBlock(64, 128, 15, 1.5, 0.9, groups, strided=strided)
    x = torch.randn(pts.shape[0], 64)
    gout = torch.randn(q.shape[0], 128)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "kernel_points" not in k) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_()
    MO.residual_block(sd, "", xr, q, pts, idx, 0.9, groups, strided).backward(gout)
    blk = blk.cuda().train()
    xg = x.cuda().requires_grad_()
    y = blk(xg, q.cuda(), pts.cuda(), idx.cuda())
    y.backward(gout.cuda())
    check(f"ResidualBlock(strided={strided}) d input", xg.grad, xr.grad)
    worst = 0.0
    for name, p in blk.named_parameters():
        assert p.grad is not None, name
        e = rel(p.grad, sd[name].grad)
        print(f"[grad parity]   {name}: {e:.2e} (|ref| max {sd[name].grad.abs().max().item():.2e})")
        worst = max(worst, e)
    print(f"[grad parity] ResidualBlock(strided={strided}): worst parameter gradient error {worst:.2e} over {len(list(blk.parameters()))} tensors")
    assert worst <= 1e-4


@pytest.mark.parametrize("case", [0, 1])
def test_ground_truth_overlap_kernels_vs_reference_golden(case):
    """rdm_node_correspondences / rdm_node_distance_mask through rdmnet_b200.registration against the outputs of the reference's
    own get_node_correspondences / get_node_overlap / get_node_correspondences_disance (matching.py:252-503) on seeded patches
    (tests/golden/gt_small.npz, generated by tests/golden/make_golden_gt.py importing the reference): the sphere-test matrix and
    the correspondence index list exactly (row-major order), the overlap ratios to float rounding."""
    from rdmnet_b200 import registration as R
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "gt_small.npz")))
    p = f"c{case}_"
    c = lambda k, dt=None: torch.from_numpy(g[p + k]).cuda() if dt is None else torch.from_numpy(g[p + k]).to(dt).cuda()  # noqa: E731
    args = (c("ref_nodes"), c("src_nodes"), c("ref_knn"), c("src_knn"), c("T"), 0.6)
    kw = dict(ref_masks=c("ref_masks"), src_masks=c("src_masks"), ref_knn_masks=c("ref_knn_masks"), src_knn_masks=c("src_knn_masks"))
    mask = R.get_node_correspondences(*args, return_mask=True, **kw)
    assert torch.equal(mask.cpu(), torch.from_numpy(g[p + "sphere_mask"])), "sphere-test matrix"
    idx, ov = R.get_node_correspondences(*args, **kw)
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), g[p + "corr_indices"]), "correspondence indices"
    check(f"gt case {case} corr_overlaps", ov, torch.from_numpy(g[p + "corr_overlaps"]), 1e-6)
    b = min(g[p + "ref_nodes"].shape[0], g[p + "src_nodes"].shape[0])
    pov = R.get_node_overlap(args[0][:b], args[1][:b], args[2][:b], args[3][:b], args[4], 0.6, ref_masks=kw["ref_masks"][:b],
                             src_masks=kw["src_masks"][:b], ref_knn_masks=kw["ref_knn_masks"][:b], src_knn_masks=kw["src_knn_masks"][:b])
    check(f"gt case {case} pair overlaps", pov, torch.from_numpy(g[p + "pair_overlaps"]), 1e-6)
    dmask = R.get_node_correspondences_disance(args[0], args[1], args[4], 2.0, ref_masks=kw["ref_masks"], src_masks=kw["src_masks"])
    assert torch.equal(dmask.cpu(), torch.from_numpy(g[p + "distance_mask"])), "nearest-node distance mask"


def test_ground_truth_ball_query_equals_brute_force():
    """registration.get_correspondences (the GPU stand-in of the cKDTree ball query, geotransformer/utils/registration.py:203-217,
    called by experiments/loss.py:92,151) against an O(MN) double-precision distance matrix: the same pair set."""
    from rdmnet_b200 import registration as R
    rng = np.random.default_rng(4)
    ref = (rng.random((700, 3)) * [20, 20, 3]).astype(np.float32)
    src = (rng.random((650, 3)) * [20, 20, 3]).astype(np.float32)
    ang = 0.3
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [0.5, -1.0, 0.1]
    for radius in (0.45, 1.2):
        got = R.get_correspondences(ref, src, T, radius)
        moved = (src @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        d2 = ((ref[:, None, :].astype(np.float64) - moved[None].astype(np.float64)) ** 2).sum(-1)
        want = {tuple(x) for x in np.argwhere(d2 < radius * radius).tolist()}
        near = {tuple(x) for x in np.argwhere(np.abs(np.sqrt(d2) - radius) < 1e-5).tolist()}  # float-rounding band
        gs = {tuple(x) for x in got.tolist()}
        assert got.dtype == np.int64 and got.shape[1] == 2 and len(gs) == len(got)
        assert (gs ^ want) <= near, (len(gs), len(want), len(gs ^ want))
        print(f"[train] ball query r={radius}: {len(gs)} pairs")
    assert R.get_correspondences(ref[:0], src, None, 1.0).shape == (0, 2)
    assert R.get_correspondences(ref, src + 1000.0, None, 0.5).shape == (0, 2)


def test_transformer_layer_gradients():
    """RoPE + multi-head attention + output projection / LayerNorm / FFN (one self layer with rotary embedding, one cross layer)
    against autograd through the oracle's transformer_layer: rdm_rope_bwd, rdm_attention_bwd, Linear / LayerNorm backward."""
    from rdmnet_b200 import modules as M
    torch.manual_seed(7)
    n0, n1, c, heads = 217, 190, 128, 4
    for rotary in (True, False):
        layer = M.TransformerLayer(c, heads, rotary=rotary)
        with torch.no_grad():
            for prm in layer.parameters():
                prm.add_(torch.randn_like(prm) * 0.05)
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "div_term" not in k) for k, v in layer.state_dict().items()}
        x, mem = torch.randn(n0, c), (torch.randn(n0, c) if rotary else torch.randn(n1, c))
        ex, em = (torch.randn(n0, c // 2), None) if rotary else (None, None)
        gout = torch.randn(n0, c)
        xr, mr = x.clone().requires_grad_(), mem.clone().requires_grad_()
        exr = ex.clone().requires_grad_() if rotary else None
        if rotary:  # self attention: memory = input, one embedding for both sides (thdroformer.py:229-243)
            yr = MO.transformer_layer(sd, "", xr, xr, heads, exr, exr)
        else:
            yr = MO.transformer_layer(sd, "", xr, mr, heads)
        yr.backward(gout)
        layer = layer.cuda().train()
        xg, mg = x.cuda().requires_grad_(), mem.cuda().requires_grad_()
        exg = ex.cuda().requires_grad_() if rotary else None
        yg = layer(xg, xg, exg, exg) if rotary else layer(xg, mg)
        check(f"TransformerLayer(rotary={rotary}) forward", yg, yr)
        yg.backward(gout.cuda())
        check(f"TransformerLayer(rotary={rotary}) d input", xg.grad, xr.grad)
        if rotary:
            check("TransformerLayer d rotary embedding", exg.grad, exr.grad)
        else:
            check("TransformerLayer d memory", mg.grad, mr.grad)
        worst = 0.0
        scale = max(sd[name].grad.abs().max().item() for name, _ in layer.named_parameters())
        for name, prm in layer.named_parameters():
            assert prm.grad is not None, name
            rmax = sd[name].grad.abs().max().item()
            if rmax < 1e-5 * scale:
                # analytically zero (the key bias shifts every score of a row by the same q.b: softmax cancels it): both sides are
                # rounding noise, compared on the scale of the other gradients
                e = (prm.grad.detach().cpu() - sd[name].grad).abs().max().item() / scale
            else:
                e = rel(prm.grad, sd[name].grad)
            print(f"[grad parity]   {name}: {e:.2e} (|ref| max {rmax:.2e})")
            worst = max(worst, e)
        assert worst <= 1e-4, worst


def test_sinkhorn_gradients():
    """rdm_sinkhorn_bwd (forward re-run + reverse sweep over the 100 stored iterates) against autograd through the unrolled
    iterations of the oracle (learnable_sinkhorn.py:13-66): d scores and d alpha, with masked rows / columns and a full patch."""
    from rdmnet_b200 import ops
    torch.manual_seed(9)
    for (p, r, c, iters) in ((6, 128, 128, 100), (3, 40, 57, 30)):
        scores = torch.randn(p, r, c) * 1.5
        rm, cm = torch.rand(p, r) < 0.8, torch.rand(p, c) < 0.7
        rm[0], cm[0] = True, True  # one patch without masked points
        rm[:, 0], cm[:, 0] = True, True
        alpha = torch.tensor(0.7)
        # upstream gradient: zero on masked entries (the reference's losses read those as the constant 1e12, loss.py:263-271)
        live = torch.ones(p, r + 1, c + 1, dtype=torch.bool)
        live[:, :r] &= rm[:, :, None]
        live[:, :, :c] &= cm[:, None, :]
        gout = torch.randn(p, r + 1, c + 1) * live
        sr, ar = scores.clone().requires_grad_(), alpha.clone().requires_grad_()
        out_r = MO.sinkhorn(sr, rm, cm, ar, iters)
        out_r.backward(gout)
        sg, ag = scores.cuda().requires_grad_(), alpha.cuda().requires_grad_()
        out_g = ops.sinkhorn(sg, rm.cuda(), cm.cuda(), ag, iters)
        assert torch.equal((out_g.cpu() < -1e11), ~live)
        fe = ((out_g.cpu() - out_r.detach()).abs() * live).max().item()
        print(f"[grad parity] sinkhorn {p}x{r}x{c} forward (live entries): {fe:.2e} abs")
        assert fe < 2e-3
        out_g.backward(gout.cuda())
        check(f"sinkhorn {p}x{r}x{c} d scores", sg.grad, sr.grad)
        check(f"sinkhorn {p}x{r}x{c} d alpha", ag.grad.reshape(()), ar.grad.reshape(()))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "experiments")), reason="reference tree not staged (baseline/_ref/RDMNet)")
def test_reference_training_step_runs_on_the_dropin(pretrained_state):
    """experiments/model.py (train-mode forward with ground truth) + experiments/loss.py OverallLoss, UNMODIFIED, on top of
    rdmnet_b200.dropin: loss finite, every trained parameter receives a finite gradient, an Adam step changes the loss."""
    from rdmnet_b200 import dropin as D, synthetic
    names = ("geotransformer", "rdmnet", "config", "backbone", "model_infer", "model", "loss", "dataset")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in names}
    for k in saved:
        del sys.modules[k]
    path = list(sys.path)
    try:
        D.install(reference_root=REF)
        config, model_mod, loss_mod = (importlib.import_module(m) for m in ("config", "model", "loss"))
        data = importlib.import_module("geotransformer.utils.data")
        cfg = config.make_cfg()
        cfg.test.vis = False
        cfg.neighbor_limits = [65, 63, 69, 70, 81]
        model = model_mod.create_model(cfg)
        model.load_state_dict(pretrained_state, strict=True)
        model = model.cuda().train()
        loss_fn = loss_mod.OverallLoss(cfg).cuda()
        ne, na = synthetic.SIZE_CLASSES["8k"]
        p = synthetic.make_pair(pair_id=21, n_elev=ne, n_azim=na)
        item = dict(ref_points=p["ref_points"], src_points=p["src_points"], ref_feats=np.ones((len(p["ref_points"]), 1), np.float32),
                    src_feats=np.ones((len(p["src_points"]), 1), np.float32), transform=p["transform"])
        dd = data.registration_collate_fn_stack_mode([item], 5, 0.3, 4.25 * 0.3, cfg.neighbor_limits)
        dd = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in dd.items()}
        dd["testing"] = False
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-6)  # experiments/trainval.py:34
        np.random.seed(0)
        out = model(dd)
        # loss.py:240-243 builds its index helpers with bare torch.arange (CPU) and masks them with CUDA tensors - torch 1.8
        # accepted that, torch 2.x does not. The default-device context is the environment shim (like np.int in dropin):
        # the loss source stays unmodified.
        with torch.device("cuda"):
            losses = loss_fn(out, dd)
        loss0 = float(losses["loss"])
        assert np.isfinite(loss0)
        opt.zero_grad()
        losses["loss"].backward()
        missing = [n for n, q in model.named_parameters() if q.grad is None]
        bad = [n for n, q in model.named_parameters() if q.grad is not None and not torch.isfinite(q.grad).all()]
        print(f"[train] loss {loss0:.4f} {({k: round(float(v), 4) for k, v in losses.items()})}; parameters without gradient: {missing}")
        assert not bad, bad
        assert len(missing) <= 4, missing  # the reference prints these too (epoch_based_trainer.py:105-107)
        gn = float(torch.sqrt(sum((q.grad.double() ** 2).sum() for q in model.parameters() if q.grad is not None)))
        assert np.isfinite(gn) and gn > 0
        opt.step()
        np.random.seed(0)
        with torch.no_grad(), torch.device("cuda"):
            loss1 = float(loss_fn(model(dd), dd)["loss"])
        print(f"[train] gradient norm {gn:.4f}; loss after one Adam step {loss1:.4f}")
        assert np.isfinite(loss1) and loss1 != loss0
    finally:
        D.uninstall()
        for k in [k for k in sys.modules if k.split(".")[0] in names]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = path
