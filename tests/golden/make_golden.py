"""Generates the golden fixtures in this directory by IMPORTING the reference (read-only, /root/reference) in the
build container and running it on CPU. Not run on the GPU box (the reference does not exist there); the
resulting .npz files are committed. Usage:  python tests/golden/make_golden.py

Shims: stub open3d (PLY reader only), ipdb, IPython, matplotlib, coloredlogs, easydict,
rdmnet.utils.visualization; np.int alias; Tensor.cuda()/Module.cuda() -> identity; rdmnet.ext -> the reference's own
C++ core compiled in place (oracle/_ref/libref_ext.so).

Outputs:
  scans.npz          xyz of the three bundled scans (assets/pc/*.npy), float32
  modules_small.npz  per-module input/weight/output triples from the reference classes with small seeded weights
  pair_outputs.npz   end-to-end outputs of model_infer.RDMNet with the pretrained checkpoint on pairs (0,4), (0,7)
  _big/rdmnet_state.pt   (git-ignored) the checkpoint's 'model' state dict, for GPU-box parity runs
"""
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import pyramid as OP  # noqa: E402


def install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _PC:
        def __init__(self, pts):
            self.points = pts

    def read_point_cloud(path):
        raw = open(path, "rb").read()
        body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
        return _PC(np.frombuffer(body, dtype="<f8").reshape(-1, 3).copy())

    o3d = mod("open3d")
    o3d.io = mod("open3d.io", read_point_cloud=read_point_cloud)
    o3d.geometry = mod("open3d.geometry")
    o3d.utility = mod("open3d.utility")
    o3d.visualization = mod("open3d.visualization")
    o3d.pipelines = mod("open3d.pipelines")
    mod("ipdb", set_trace=lambda *a, **k: None)
    mod("IPython", embed=lambda *a, **k: None)
    mpl = mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")
    mpl.cm = mod("matplotlib.cm")
    mod("mpl_toolkits")
    mod("mpl_toolkits.mplot3d", Axes3D=object)
    mod("coloredlogs", ColoredFormatter=logging.Formatter)

    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    mod("easydict", EasyDict=EasyDict)
    mod("rdmnet.utils.visualization", vis_shifte_node=lambda *a, **k: None, visualization=lambda *a, **k: None,
        vis_node_grouping=lambda *a, **k: None)
    np.int = int
    np.float = float
    torch.Tensor.cuda = lambda s, *a, **k: s.contiguous()
    torch.nn.Module.cuda = lambda s, *a, **k: s

    # rdmnet.ext -> reference C++ core
    def grid_subsampling(points, lengths, voxel):
        p, l = OP.grid_subsample(points.numpy(), lengths.numpy(), float(voxel), impl="ref")
        return [torch.from_numpy(p), torch.from_numpy(l)]

    def radius_neighbors(q, s, ql, sl, r):
        return torch.from_numpy(OP.radius_neighbors(q.numpy(), s.numpy(), ql.numpy(), sl.numpy(), float(r), "ref"))

    pkg = types.ModuleType("rdmnet")
    pkg.__path__ = [os.path.join(REF, "rdmnet")]
    sys.modules["rdmnet"] = pkg
    mod("rdmnet.ext", grid_subsampling=grid_subsampling, radius_neighbors=radius_neighbors)
    utils = types.ModuleType("rdmnet.utils")
    utils.__path__ = []
    sys.modules["rdmnet.utils"] = utils
    sys.path[:0] = [REF, os.path.join(REF, "experiments")]


def t2n(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def sd_items(prefix, module):
    return {prefix + k: v.detach().numpy() for k, v in module.state_dict().items()}


def random_cloud(rng, n, extent):
    return (rng.random((n, 3)).astype(np.float32) - 0.5) * np.array(extent, np.float32)


def make_modules_small():
    from geotransformer.modules.kpconv import ConvBlock, ResidualBlock, UnaryBlock
    from geotransformer.modules.sinkhorn import LearnableLogOptimalTransport
    from geotransformer.modules.geotransformer import SuperPointMatching, LocalGlobalRegistration
    from geotransformer.modules.ops import point_to_node_partition
    from geotransformer.modules.registration.procrustes import weighted_procrustes
    from rdmnet.thdroformer import ThDRoFormer
    from rdmnet.vote import Vote_layer, NMS
    from easydict import EasyDict

    g = {}
    rng = np.random.default_rng(7351)
    torch.manual_seed(7351)

    # ---- pyramid on a small cloud pair (reference ext) ----
    a, b = random_cloud(rng, 700, (12, 10, 3)), random_cloud(rng, 600, (12, 10, 3))
    pts = np.concatenate([a, b])
    lens = np.array([700, 600], np.int64)
    p1, l1 = OP.grid_subsample(pts, lens, 0.9, impl="ref")
    nb0 = OP.radius_search(pts, pts, lens, lens, 1.5, 20, impl="ref")
    sub0 = OP.radius_search(p1, pts, l1, lens, 1.5, 20, impl="ref")
    nb0, sub0 = OP.canonicalize_ties(nb0, pts, pts), OP.canonicalize_ties(sub0, p1, pts)
    g.update(pyr_points0=pts, pyr_lengths0=lens, pyr_points1=p1, pyr_lengths1=l1, pyr_nb0=nb0, pyr_sub0=sub0)

    # ---- ConvBlock 1->64 and ResidualBlocks (plain, strided) ----
    tp, tl, tn, ts = map(torch.from_numpy, (pts, lens, nb0, sub0))
    tp1 = torch.from_numpy(p1)
    cb = ConvBlock(1, 64, 15, 1.5, 0.7, 32)
    rb = ResidualBlock(64, 128, 15, 1.5, 0.7, 32)
    rs = ResidualBlock(128, 128, 15, 1.5, 0.7, 32, strided=True)
    for m in (cb, rb, rs):
        for name, prm in m.named_parameters():
            if "norm" in name:  # non-trivial affine
                prm.data = torch.randn_like(prm) * 0.3 + (1.0 if name.endswith("weight") else 0.0)
    with torch.no_grad():
        x0 = torch.ones(pts.shape[0], 1)
        x1 = cb(x0, tp, tp, tn)
        x2 = rb(x1, tp, tp, tn)
        x3 = rs(x2, tp1, tp, ts)
    g.update(sd_items("cb.", cb)); g.update(sd_items("rb.", rb)); g.update(sd_items("rs.", rs))
    g.update(cb_out=x1.numpy(), rb_out=x2.numpy(), rs_out=x3.numpy())

    ub = UnaryBlock(40, 64, 32)
    with torch.no_grad():
        ux = torch.randn(90, 40)
        g.update(sd_items("ub.", ub)); g.update(ub_in=ux.numpy(), ub_out=ub(ux).numpy())

    # ---- ThDRoFormer (hidden 32, 4 heads, 2x(self,cross)) ----
    tf = ThDRoFormer(48, 40, 32, 4, 2)
    with torch.no_grad():
        rp, sp = torch.randn(1, 37, 3) * 5, torch.randn(1, 29, 3) * 5
        rf, sf = torch.randn(1, 37, 48), torch.randn(1, 29, 48)
        ro, so = tf(rp, sp, rf, sf)
    g.update(sd_items("tf.", tf))
    g.update(tf_rp=rp[0].numpy(), tf_sp=sp[0].numpy(), tf_rf=rf[0].numpy(), tf_sf=sf[0].numpy(),
             tf_ro=ro[0].numpy(), tf_so=so[0].numpy())

    # ---- Vote layer + NMS ----
    vcfg = EasyDict(MLPS=[64, 32], MAX_TRANSLATE_RANGE=[3.0, 3.0, 3.0], input_feats_dim=32, NMS_radius=2.4)
    vl = Vote_layer(vcfg, 1)
    with torch.no_grad():
        vl.ctr_reg.weight.mul_(20.0)  # make the clamp bite
        vx, vf = torch.randn(55, 3) * 4, torch.randn(55, 32)
        nx, nf_ = vl(vx, vf)
    g.update(sd_items("vl.", vl)); g.update(vl_xyz=vx.numpy(), vl_f=vf.numpy(), vl_oxyz=nx.numpy(), vl_of=nf_.numpy())
    nms = NMS(vcfg, [9, 9, 9, 9, 12])
    nodes = torch.from_numpy(random_cloud(rng, 160, (30, 20, 2)))
    nlen = torch.tensor([90, 70])
    g.update(nms_nodes=nodes.numpy(), nms_len=nlen.numpy(), nms_limit=np.int64(12), nms_mask=nms(nodes, nlen).numpy())

    # ---- partition / coarse matching ----
    pp, nn_ = torch.from_numpy(random_cloud(rng, 500, (20, 20, 2))), torch.from_numpy(random_cloud(rng, 24, (20, 20, 2)))
    p2n, nm, knn, km = point_to_node_partition(pp, nn_, 16)
    g.update(part_pts=pp.numpy(), part_nodes=nn_.numpy(), part_p2n=p2n.numpy(), part_nm=nm.numpy(),
             part_knn=knn.numpy(), part_km=km.numpy())
    spm = SuperPointMatching(50, True)
    rf = torch.nn.functional.normalize(torch.randn(24, 16), dim=1)
    sf = torch.nn.functional.normalize(torch.randn(19, 16), dim=1)
    rm, sm = torch.rand(24) > 0.2, torch.rand(19) > 0.2
    ri, si, sc = spm(rf, sf, rm, sm)
    g.update(spm_rf=rf.numpy(), spm_sf=sf.numpy(), spm_rm=rm.numpy(), spm_sm=sm.numpy(), spm_ri=ri.numpy(),
             spm_si=si.numpy(), spm_sc=sc.numpy())

    # ---- sinkhorn ----
    ot = LearnableLogOptimalTransport(100)
    with torch.no_grad():
        ot.alpha.fill_(1.6727)
        s = torch.randn(6, 20, 20) * 2
        rmk, cmk = torch.rand(6, 20) > 0.25, torch.rand(6, 20) > 0.25
        rmk[0], cmk[0] = True, True
        o = ot(s, rmk, cmk)
    g.update(ot_in=s.numpy(), ot_rm=rmk.numpy(), ot_cm=cmk.numpy(), ot_out=o.numpy())

    # ---- procrustes + LGR on a synthetic rigid problem ----
    src = torch.randn(7, 40, 3) * 3
    ang = 0.3
    R = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
    ref = src @ R.t() + torch.tensor([1.0, -2.0, 0.5]) + torch.randn(7, 40, 3) * 0.05
    w = torch.rand(7, 40)
    w[3, 20:] = 0
    T = weighted_procrustes(src, ref, w, return_transform=True)
    g.update(wp_src=src.numpy(), wp_ref=ref.numpy(), wp_w=w.numpy(), wp_T=T.numpy())

    lgr = LocalGlobalRegistration(1, 0.6, mutual=False, confidence_threshold=0, use_dustbin=True,
                                  use_global_score=False, correspondence_threshold=3, correspondence_limit=None,
                                  num_refinement_steps=5)
    B, K = 12, 16
    rk = torch.randn(B, K, 3) * 4
    perm = torch.stack([torch.randperm(K) for _ in range(B)])
    Rinv = R.t()
    skp = torch.stack([((rk[i] - torch.tensor([1.0, -2.0, 0.5])) @ Rinv.t())[perm[i]] for i in range(B)])
    skp = skp + torch.randn_like(skp) * 0.03
    logits = torch.full((B, K + 1, K + 1), -6.0)
    for i in range(B):
        for j in range(K):
            src_j = (perm[i] == j).nonzero()[0, 0]
            if (i * 7 + j) % 4 != 0:  # leave some unmatched
                logits[i, j, src_j] = -1.0 + 0.1 * torch.randn(())
    logits[:, -1, :] = -3.0
    logits[:, :, -1] = -3.0
    logits[5:7, :K, :K] = -6.0  # patches with < 3 correspondences
    rmk, smk = torch.rand(B, K) > 0.1, torch.rand(B, K) > 0.1
    with torch.no_grad():
        rc, sc_, cs, Te = lgr(rk, skp, rmk, smk, logits, torch.ones(B))
    g.update(lgr_rk=rk.numpy(), lgr_sk=skp.numpy(), lgr_rm=rmk.numpy(), lgr_sm=smk.numpy(), lgr_scores=logits.numpy(),
             lgr_rc=rc.numpy(), lgr_sc=sc_.numpy(), lgr_cs=cs.numpy(), lgr_T=Te.numpy())
    np.savez_compressed(os.path.join(HERE, "modules_small.npz"), **g)
    print("modules_small.npz:", len(g), "arrays")


def make_pair_outputs():
    from config import make_cfg
    import model_infer
    from geotransformer.utils.data import registration_collate_fn_stack_mode

    scans = {k: np.load(os.path.join(REF, "assets/pc", k + ".npy"))[:, :3].astype(np.float32)
             for k in ("000000", "000004", "000007")}
    np.savez_compressed(os.path.join(HERE, "scans.npz"), **{"s" + k: v for k, v in scans.items()})
    cfg = make_cfg()
    cfg.test.vis = False
    cfg.neighbor_limits = [65, 63, 69, 70, 81]  # what calibrate_neighbors_stack_mode returns on the two infer pairs
    model = model_infer.create_model(cfg)
    ck = torch.load(os.path.join(REF, "weights/rdmnet.pth.tar"), map_location="cpu", weights_only=False)
    model.load_state_dict(ck["model"], strict=True)
    os.makedirs(os.path.join(HERE, "_big"), exist_ok=True)
    torch.save({k: v.clone() for k, v in ck["model"].items()}, os.path.join(HERE, "_big", "rdmnet_state.pt"))
    model.eval()
    torch.set_grad_enabled(False)
    g = {}
    for tag, (a, b) in {"p04": ("000000", "000004"), "p07": ("000000", "000007")}.items():
        item = dict(ref_points=scans[a], src_points=scans[b], ref_feats=np.ones((len(scans[a]), 1), np.float32),
                    src_feats=np.ones((len(scans[b]), 1), np.float32))
        dd = registration_collate_fn_stack_mode([item], 5, 0.3, 4.25 * 0.3, cfg.neighbor_limits)
        for k in ("neighbors", "subsampling", "upsampling"):
            dd[k] = [t.contiguous() for t in dd[k]]
        dd["testing"] = True
        hooks = {}
        model.encoder.register_forward_hook(lambda m, i, o: hooks.__setitem__("enc", [t.clone() for t in o]))
        model.transformer.register_forward_hook(lambda m, i, o: hooks.__setitem__("t1", o))
        model.vote.register_forward_hook(lambda m, i, o: hooks.__setitem__("vote", o))
        model.nms.register_forward_hook(lambda m, i, o: hooks.__setitem__("nms", o))
        out = model(dd)
        keep = ("estimated_transform", "ref_node_corr_indices", "src_node_corr_indices", "ref_corr_points",
                "src_corr_points", "corr_scores", "ref_points_c", "src_points_c", "ref_feats_c", "src_feats_c")
        for k in keep:
            g[f"{tag}_{k}"] = out[k].numpy()
        g[f"{tag}_lengths"] = np.stack([l.numpy() for l in dd["lengths"]])
        g[f"{tag}_nms"] = hooks["nms"].numpy()
        g[f"{tag}_shifted"] = hooks["vote"][0].numpy()
        g[f"{tag}_feats_s5_head"] = hooks["enc"][-1][:64, :64].numpy()
        g[f"{tag}_feats_s5_absmean"] = np.float64(hooks["enc"][-1].abs().mean())
        g[f"{tag}_feats_s1_absmean"] = np.float64(hooks["enc"][0].abs().mean())
        g[f"{tag}_t1_ref_head"] = hooks["t1"][0][0, :64, :64].numpy()
        g[f"{tag}_feats_f_head"] = out["ref_feats_f"][:64, :64].numpy()
        g[f"{tag}_ms_sum"] = np.float64(out["matching_scores"].exp()[:, :-1, :-1].sum())
        print(tag, "ncorr", out["corr_scores"].shape[0], "\n", out["estimated_transform"].numpy())
    np.savez_compressed(os.path.join(HERE, "pair_outputs.npz"), **g)


if __name__ == "__main__":
    install_shims()
    make_modules_small()
    make_pair_outputs()
