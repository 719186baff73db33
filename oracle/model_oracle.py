"""TEST INFRASTRUCTURE ONLY (oracle): plain-PyTorch fp32 CPU restatement of the RDMNet inference forward.

Flat functional code over a checkpoint ``state_dict`` (key layout = the reference's); every
function cites the reference file:line it restates. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; rdmnet_b200 (the product) never does.

Parity status: PINNED against the reference's own Python executed in the build container
(tests/golden/make_golden.py imports /root/reference and writes tests/golden/*.npz; tests/test_oracle_cpu.py
replays them): module-level goldens on small seeded inputs and the end-to-end outputs for bundled pair (0,4).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

CFG = dict(  # experiments/config.py:86-161
    num_stages=5, voxel=0.3, radius=4.25 * 0.3, sigma=2.0 * 0.3, groups=32, kernel_size=15,
    hidden=128, heads=4, patch_k=128, sinkhorn_iters=100, num_corr=256, nms_radius=2.4, max_shift=3.0,
    acceptance_radius=0.6, corr_threshold=3, refine_steps=5,
)
DEFAULT_LIMITS = [65, 63, 69, 70, 81]  # calibrated on the bundled pairs


# ----------------------------------------------------------------------------- backbone
def gather_rows(x, idx):
    """geotransformer/modules/ops/index_select.py:4-30 with dim=0."""
    return x[idx.reshape(-1)].reshape(*idx.shape, *x.shape[1:])


def kpconv(s_feats, q_pts, s_pts, idx, weights, kernel_points, sigma, bias=None):
    """geotransformer/modules/kpconv/kpconv.py:79-122."""
    s_pad = torch.cat([s_pts, torch.full((1, 3), 1e6, dtype=s_pts.dtype)], 0)  # :91 shadow point
    nb = gather_rows(s_pad, idx) - q_pts[:, None, :]  # (M,H,3)
    diff = nb[:, :, None, :] - kernel_points[None, None]  # (M,H,K,3)
    sq = (diff ** 2).sum(3)
    w = torch.clamp(1 - torch.sqrt(sq) / sigma, min=0.0).transpose(1, 2)  # (M,K,H) :98-100
    f_pad = torch.cat([s_feats, torch.zeros(1, s_feats.shape[1])], 0)
    nf = gather_rows(f_pad, idx)  # (M,H,C)
    wf = torch.matmul(w, nf).permute(1, 0, 2)  # (K,M,C)
    out = torch.matmul(wf, weights).sum(0)  # (M,Cout)
    cnt = (nf.sum(-1) > 0).sum(-1).clamp(min=1)  # :113-115
    out = out / cnt[:, None]
    if bias is not None:
        out = out + bias
    return out


def group_norm(x, w, b, groups):
    """kpconv/modules.py:46-50: nn.GroupNorm over (1,C,N) i.e. statistics over (C/groups x N) jointly."""
    return F.group_norm(x.t()[None], groups, w, b, 1e-5)[0].t()


def unary(sd, p, x, groups, relu=True):
    """kpconv/modules.py:78-83 UnaryBlock."""
    x = F.linear(x, sd[p + "mlp.weight"], sd[p + "mlp.bias"])
    x = group_norm(x, sd[p + "norm.norm.weight"], sd[p + "norm.norm.bias"], groups)
    return F.leaky_relu(x, 0.1) if relu else x


def maxpool(x, idx):
    """kpconv/functional.py:54-67."""
    return gather_rows(torch.cat([x, torch.zeros(1, x.shape[1])], 0), idx).max(1)[0]


def nearest_upsample(x, idx):
    """kpconv/functional.py:6-22."""
    return torch.cat([x, torch.zeros(1, x.shape[1])], 0)[idx[:, 0]]


def conv_block(sd, p, feats, q, s, idx, sigma, groups):
    """kpconv/modules.py:143-147 ConvBlock."""
    x = kpconv(feats, q, s, idx, sd[p + "KPConv.weights"], sd[p + "KPConv.kernel_points"], sigma,
               sd.get(p + "KPConv.bias"))
    x = group_norm(x, sd[p + "norm.norm.weight"], sd[p + "norm.norm.bias"], groups)
    return F.leaky_relu(x, 0.1)


def residual_block(sd, p, feats, q, s, idx, sigma, groups, strided):
    """kpconv/modules.py:205-225 ResidualBlock."""
    x = unary(sd, p + "unary1.", feats, groups) if (p + "unary1.mlp.weight") in sd else feats
    x = kpconv(x, q, s, idx, sd[p + "KPConv.weights"], sd[p + "KPConv.kernel_points"], sigma,
               sd.get(p + "KPConv.bias"))
    x = F.leaky_relu(group_norm(x, sd[p + "norm_conv.norm.weight"], sd[p + "norm_conv.norm.bias"], groups), 0.1)
    x = unary(sd, p + "unary2.", x, groups, relu=False)
    sc = maxpool(feats, idx) if strided else feats
    if (p + "unary_shortcut.mlp.weight") in sd:
        sc = unary(sd, p + "unary_shortcut.", sc, groups, relu=False)
    return F.leaky_relu(x + sc, 0.1)


def encoder(sd, feats, pyr, sigma0=CFG["sigma"], groups=CFG["groups"], p="encoder."):
    """experiments/backbone.py:72-107 (sigma per block: :11-70)."""
    P, NB, SUB = pyr["points"], pyr["neighbors"], pyr["subsampling"]
    out = []
    x = conv_block(sd, p + "encoder1_1.", feats, P[0], P[0], NB[0], sigma0, groups)
    x = residual_block(sd, p + "encoder1_2.", x, P[0], P[0], NB[0], sigma0, groups, False)
    out.append(x)
    for s in range(1, 5):
        sig_prev, sig = sigma0 * 2 ** (s - 1), sigma0 * 2 ** s
        x = residual_block(sd, p + f"encoder{s + 1}_1.", x, P[s], P[s - 1], SUB[s - 1], sig_prev, groups, True)
        x = residual_block(sd, p + f"encoder{s + 1}_2.", x, P[s], P[s], NB[s], sig, groups, False)
        x = residual_block(sd, p + f"encoder{s + 1}_3.", x, P[s], P[s], NB[s], sig, groups, False)
        out.append(x)
    return out


def decoder(sd, feats, pyr, groups=CFG["groups"], p="decoder."):
    """experiments/backbone.py:118-151."""
    UP = pyr["upsampling"]
    l4 = unary(sd, p + "decoder4.", torch.cat([nearest_upsample(feats[4], UP[3]), feats[3]], 1), groups)
    l3 = unary(sd, p + "decoder3.", torch.cat([nearest_upsample(l4, UP[2]), feats[2]], 1), groups)
    l2 = torch.cat([nearest_upsample(l3, UP[1]), feats[1]], 1)
    return F.linear(l2, sd[p + "decoder2.mlp.weight"], sd[p + "decoder2.mlp.bias"])


# ----------------------------------------------------------------------------- ThDRoFormer
def rope(x, emb):
    """rdmnet/thdroformer/thdroformer.py:56-85. x (H,N,D); emb (H,N,D/2)."""
    rot = torch.stack([-x[..., 1::2], x[..., 0::2]], -1).reshape(x.shape)  # :71-73
    theta = torch.sigmoid(emb.repeat_interleave(2, dim=-1)) * 3.14159265359 * 2  # :76-78
    return x * torch.cos(theta) + rot * torch.sin(theta)


def mha(sd, p, xq, xk, heads, emb_q=None, emb_k=None):
    """thdroformer.py:108-139 (self, RoPE) / transformer/vanilla_transformer.py:31-70 (cross)."""
    n, c = xq.shape
    d = c // heads
    q = F.linear(xq, sd[p + "proj_q.weight"], sd[p + "proj_q.bias"]).reshape(n, heads, d).transpose(0, 1)
    k = F.linear(xk, sd[p + "proj_k.weight"], sd[p + "proj_k.bias"]).reshape(-1, heads, d).transpose(0, 1)
    v = F.linear(xk, sd[p + "proj_v.weight"], sd[p + "proj_v.bias"]).reshape(-1, heads, d).transpose(0, 1)
    if emb_q is not None:
        q = rope(q, emb_q.reshape(n, heads, -1).transpose(0, 1))
        k = rope(k, emb_k.reshape(-1, heads, emb_k.shape[1] // heads).transpose(0, 1))
    s = torch.softmax(torch.matmul(q, k.transpose(1, 2)) / d ** 0.5, -1)  # thdroformer.py:20-25
    return torch.matmul(s, v).transpose(0, 1).reshape(n, c)


def transformer_layer(sd, p, x, mem, heads, emb_x=None, emb_mem=None):
    """thdroformer.py:141-202 / vanilla_transformer.py:73-129 + transformer/output_layer.py:15-21."""
    h = mha(sd, p + "attention.attention.", x, mem, heads, emb_x, emb_mem)
    h = F.linear(h, sd[p + "attention.linear.weight"], sd[p + "attention.linear.bias"])
    h = F.layer_norm(h + x, (x.shape[1],), sd[p + "attention.norm.weight"], sd[p + "attention.norm.bias"])
    e = F.relu(F.linear(h, sd[p + "output.expand.weight"], sd[p + "output.expand.bias"]))
    e = F.linear(e, sd[p + "output.squeeze.weight"], sd[p + "output.squeeze.bias"])
    return F.layer_norm(h + e, (x.shape[1],), sd[p + "output.norm.weight"], sd[p + "output.norm.bias"])


def thdroformer(sd, p, ref_pts, src_pts, ref_f, src_f, heads=CFG["heads"]):
    """thdroformer.py:304-347, layer schedule :229-251 (sequential cross attention)."""
    e0 = F.linear(ref_pts, sd[p + "embedding.proj.weight"], sd[p + "embedding.proj.bias"])
    e1 = F.linear(src_pts, sd[p + "embedding.proj.weight"], sd[p + "embedding.proj.bias"])
    f0 = F.linear(ref_f, sd[p + "in_proj.weight"], sd[p + "in_proj.bias"])
    f1 = F.linear(src_f, sd[p + "in_proj.weight"], sd[p + "in_proj.bias"])
    i = 0
    while (p + f"transformer.layers.{i}.attention.linear.weight") in sd:
        lp = p + f"transformer.layers.{i}."
        if i % 2 == 0:
            f0 = transformer_layer(sd, lp, f0, f0, heads, e0, e0)
            f1 = transformer_layer(sd, lp, f1, f1, heads, e1, e1)
        else:
            f0 = transformer_layer(sd, lp, f0, f1, heads)
            f1 = transformer_layer(sd, lp, f1, f0, heads)  # sees the updated f0 (:244-245)
        i += 1
    return (F.linear(f0, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"]),
            F.linear(f1, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"]))


# ----------------------------------------------------------------------------- vote / NMS
def vote_layer(sd, p, xyz, feats, max_shift=CFG["max_shift"]):
    """rdmnet/vote/vote.py:78-117."""
    h = feats
    for lin, ln in ((0, 1), (3, 4)):
        h = F.linear(h, sd[p + f"mlp_modules.{lin}.weight"], sd[p + f"mlp_modules.{lin}.bias"])
        h = F.relu(F.layer_norm(h, (h.shape[1],), sd[p + f"mlp_modules.{ln}.weight"], sd[p + f"mlp_modules.{ln}.bias"]))
    off = F.linear(h, sd[p + "ctr_reg.weight"], sd[p + "ctr_reg.bias"])
    xyz_new = xyz + off[:, :3].clamp(-max_shift, max_shift)
    f = feats + off[:, 3:]
    return xyz_new, F.layer_norm(f, (f.shape[1],), sd[p + "out_proj.0.weight"], sd[p + "out_proj.0.bias"])


def nms_greedy(nbr_idx):
    """rdmnet/vote/vote.py:33-40. nbr_idx (N,H) padded with N."""
    n = nbr_idx.shape[0]
    sel = np.zeros(n + 1, dtype=bool)
    nb = np.asarray(nbr_idx)
    for i in range(n):
        if not sel[nb[i]].any():
            sel[i] = True
    return torch.from_numpy(sel[:-1])


# ----------------------------------------------------------------------------- matching
def pairwise_distance(x, y, normalized=False):
    """geotransformer/modules/ops/pairwise_distance.py:4-31."""
    xy = torch.matmul(x, y.t())
    if normalized:
        d = 2.0 - 2.0 * xy
    else:
        d = (x ** 2).sum(-1)[:, None] - 2 * xy + (y ** 2).sum(-1)[None, :]
    return d.clamp(min=1e-12)


def point_to_node_partition(points, nodes, k):
    """geotransformer/modules/ops/pointcloud_partition.py:60-107."""
    d = pairwise_distance(nodes, points)  # (M,N)
    p2n = d.min(0)[1]
    node_masks = torch.zeros(nodes.shape[0], dtype=torch.bool)
    node_masks[p2n] = True
    own = torch.zeros_like(d, dtype=torch.bool)
    own[p2n, torch.arange(points.shape[0])] = True
    d = d.masked_fill(~own, 1e12)
    # the reference uses d.topk(k, largest=False) (:95); exact-distance ties come out of torch.topk in an
    # implementation-defined order (CPU and CUDA differ), so the oracle fixes the canonical (d, index) order
    knn = torch.sort(d, dim=1, stable=True)[1][:, :k]
    knn_masks = p2n[knn] == torch.arange(nodes.shape[0])[:, None]
    knn = knn.masked_fill(~knn_masks, points.shape[0])
    return p2n, node_masks, knn, knn_masks


def superpoint_matching(ref_f, src_f, ref_masks, src_masks, num_corr=CFG["num_corr"]):
    """geotransformer/modules/geotransformer/superpoint_matching.py:14-83 (dual normalisation on)."""
    ri = torch.nonzero(ref_masks, as_tuple=True)[0]
    si = torch.nonzero(src_masks, as_tuple=True)[0]
    s = torch.exp(-pairwise_distance(ref_f[ri], src_f[si], normalized=True))
    s = (s / s.sum(1, keepdim=True)) * (s / s.sum(0, keepdim=True))
    kk = min(num_corr, s.numel())
    sc, ci = s.reshape(-1).topk(kk, largest=True)
    return ri[ci // s.shape[1]], si[ci % s.shape[1]], sc


def sinkhorn(scores, row_masks, col_masks, alpha, iters=CFG["sinkhorn_iters"], inf=1e12):
    """geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66."""
    b, m, n = scores.shape
    prm = torch.zeros(b, m + 1, dtype=torch.bool)
    prm[:, :m] = ~row_masks
    pcm = torch.zeros(b, n + 1, dtype=torch.bool)
    pcm[:, :n] = ~col_masks
    ps = torch.cat([torch.cat([scores, alpha.expand(b, m, 1)], -1), alpha.expand(b, 1, n + 1)], 1)
    ps = ps.masked_fill(prm[:, :, None] | pcm[:, None, :], -inf)
    nr, nc = row_masks.float().sum(1), col_masks.float().sum(1)
    norm = -torch.log(nr + nc)
    log_mu = torch.empty(b, m + 1)
    log_mu[:, :m] = norm[:, None]
    log_mu[:, m] = torch.log(nc) + norm
    log_mu[prm] = -inf
    log_nu = torch.empty(b, n + 1)
    log_nu[:, :n] = norm[:, None]
    log_nu[:, n] = torch.log(nr) + norm
    log_nu[pcm] = -inf
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(ps + v[:, None, :], 2)
        v = log_nu - torch.logsumexp(ps + u[:, :, None], 1)
    return ps + u[:, :, None] + v[:, None, :] - norm[:, None, None]


def weighted_procrustes(src, ref, w, eps=1e-5):
    """geotransformer/modules/registration/procrustes.py:6-73. (B,N,3),(B,N,3),(B,N) -> (B,4,4)."""
    w = torch.where(w < 0.0, torch.zeros_like(w), w)
    w = (w / (w.sum(1, keepdim=True) + eps))[:, :, None]
    sc = (src * w).sum(1, keepdim=True)
    rc = (ref * w).sum(1, keepdim=True)
    H = (src - sc).transpose(1, 2) @ (w * (ref - rc))
    U, _, Vh = torch.linalg.svd(H)
    V = Vh.transpose(1, 2)
    Ut = U.transpose(1, 2)
    eye = torch.eye(3).repeat(src.shape[0], 1, 1)
    eye[:, 2, 2] = torch.sign(torch.det(V @ Ut))
    R = V @ eye @ Ut
    t = rc.transpose(1, 2) - R @ sc.transpose(1, 2)
    T = torch.eye(4).repeat(src.shape[0], 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = t[:, :, 0]
    return T


def apply_transform(pts, T):
    """geotransformer/modules/ops/transformation.py:7-60."""
    return pts @ T[..., :3, :3].transpose(-1, -2) + T[..., None, :3, 3]


def lgr(ref_knn_pts, src_knn_pts, ref_knn_masks, src_knn_masks, score_mat, cfg=CFG):
    """geotransformer/modules/geotransformer/local_global_registration.py:49-91,138-243
    (k=1, mutual=False, use_dustbin=True, use_global_score=False, correspondence_limit=None)."""
    s = torch.exp(score_mat)
    b, m, n = s.shape
    mask = ref_knn_masks[:, :, None] & src_knn_masks[:, None, :]
    rv, ri = s.max(2)  # topk(k=1, dim=2)
    ref_mat = torch.zeros_like(s)
    ref_mat.scatter_(2, ri[:, :, None], rv[:, :, None])
    ref_corr = ref_mat > s[:, :, -1:]
    sv, si = s.max(1)
    src_mat = torch.zeros_like(s)
    src_mat.scatter_(1, si[:, None, :], sv[:, None, :])
    src_corr = src_mat > s[:, -1:, :]
    corr = (ref_corr | src_corr)[:, :-1, :-1] & mask
    s = s[:, :-1, :-1] * corr.float()
    bi, ri, si = torch.nonzero(corr, as_tuple=True)
    ref_c, src_c, sc = ref_knn_pts[bi, ri], src_knn_pts[bi, si], s[bi, ri, si]
    # chunks :163-171
    counts = torch.bincount(bi, minlength=b)
    keep = torch.nonzero(counts >= cfg["corr_threshold"], as_tuple=True)[0]
    thr = cfg["acceptance_radius"]
    if keep.numel() > 0:
        kmax = int(counts[keep].max())
        starts = torch.cumsum(counts, 0) - counts
        br = torch.zeros(keep.numel(), kmax, 3)
        bs = torch.zeros(keep.numel(), kmax, 3)
        bw = torch.zeros(keep.numel(), kmax)
        for j, pb in enumerate(keep.tolist()):
            a, c = int(starts[pb]), int(counts[pb])
            br[j, :c], bs[j, :c], bw[j, :c] = ref_c[a:a + c], src_c[a:a + c], sc[a:a + c]
        Ts = weighted_procrustes(bs, br, bw)
        res = torch.linalg.norm(ref_c[None] - apply_transform(src_c[None], Ts), dim=2)
        inl = res < thr
        best = inl.sum(1).argmax()
        cur = sc * inl[best].float()
    else:
        T = weighted_procrustes(src_c[None], ref_c[None], sc[None])[0]
        cur = sc * (torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1) < thr).float()
    T = weighted_procrustes(src_c[None], ref_c[None], cur[None])[0]
    for _ in range(cfg["refine_steps"] - 1):
        cur = sc * (torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1) < thr).float()
        T = weighted_procrustes(src_c[None], ref_c[None], cur[None])[0]
    return ref_c, src_c, sc, T, (bi, ri, si)


# ----------------------------------------------------------------------------- whole forward
def forward(sd, pyr, nms_search, cfg=CFG, use_vote=True):
    """experiments/model_infer.py:109-354. pyr = tensors of oracle.pyramid.precompute_pyramid;
    nms_search(points (N,3) tensor, lengths) -> (N,H) int64 neighbour table (vote.py:24-31).
    use_vote=False: the Mulran configuration (experiments/infer.py:119-120, model_infer.py:59,180): vote layer, NMS and the
    second transformer are skipped, the superpoints are the coarsest pyramid level."""
    out = {}
    P, L = pyr["points"], pyr["lengths"]
    nc, nf = int(L[-1][0]), int(L[1][0])
    pts_c, pts_f = P[-1], P[1]
    feats = torch.ones(P[0].shape[0], 1)
    fl = encoder(sd, feats, pyr)
    out["feats_s5"] = fl[-1]
    rf, sf = thdroformer(sd, "transformer.", pts_c[:nc], pts_c[nc:], fl[-1][:nc], fl[-1][nc:])
    out["ref_feats_t1"], out["src_feats_t1"] = rf, sf
    wn, bn = sd["proj_n2p_score.weight"], sd["proj_n2p_score.bias"]
    rn, sn = F.linear(rf, wn, bn), F.linear(sf, wn, bn)
    fl[-1] = torch.cat([torch.cat([rf, rn], 1), torch.cat([sf, sn], 1)], 0)
    dec = decoder(sd, fl, pyr)
    feats_f = dec[:, :-1]
    out["feats_f"] = feats_f
    out["p2p_logit"] = dec[:, -1]
    if use_vote:
        shifted, fc = vote_layer(sd, "vote.", pts_c, torch.cat([rf, sf], 0))
        out["shifted_points_c"], out["vote_feats_c"] = shifted, fc
        masks = nms_greedy(nms_search(shifted, L[-1]))
        out["nms_masks"] = masks
        rm, sm = masks[:nc], masks[nc:]
        ref_pc, src_pc = shifted[:nc][rm], shifted[nc:][sm]
        rf2, sf2 = thdroformer(sd, "transformer2.", ref_pc, src_pc, fc[:nc][rm], fc[nc:][sm])
    else:
        ref_pc, src_pc, rf2, sf2 = pts_c[:nc], pts_c[nc:], rf, sf
    rfn, sfn = F.normalize(rf2, p=2, dim=1), F.normalize(sf2, p=2, dim=1)
    out["ref_points_c"], out["src_points_c"], out["ref_feats_c"], out["src_feats_c"] = ref_pc, src_pc, rfn, sfn
    ref_pf, src_pf = pts_f[:nf], pts_f[nf:]
    _, rnm, rknn, rkm = point_to_node_partition(ref_pf, ref_pc, cfg["patch_k"])
    _, snm, sknn, skm = point_to_node_partition(src_pf, src_pc, cfg["patch_k"])
    out["ref_node_knn_indices"], out["src_node_knn_indices"] = rknn, sknn
    rci, sci, ncs = superpoint_matching(rfn, sfn, rnm, snm)
    out["ref_node_corr_indices"], out["src_node_corr_indices"], out["node_corr_scores"] = rci, sci, ncs
    rpp = torch.cat([ref_pf, torch.zeros(1, 3)], 0)
    spp = torch.cat([src_pf, torch.zeros(1, 3)], 0)
    rpf = torch.cat([feats_f[:nf], torch.zeros(1, feats_f.shape[1])], 0)
    spf = torch.cat([feats_f[nf:], torch.zeros(1, feats_f.shape[1])], 0)
    rk, sk = rknn[rci], sknn[sci]
    rkp, skp = rpp[rk], spp[sk]
    ms = torch.einsum("bnd,bmd->bnm", rpf[rk], spf[sk]) / feats_f.shape[1] ** 0.5
    ms = sinkhorn(ms, rkm[rci], skm[sci], sd["optimal_transport.alpha"])
    out["matching_scores"] = ms
    rc, sc_, cs, T, _ = lgr(rkp, skp, rkm[rci], skm[sci], ms)
    out["ref_corr_points"], out["src_corr_points"], out["corr_scores"], out["estimated_transform"] = rc, sc_, cs, T
    return out


def pyramid_to_torch(pyr):
    return {k: [torch.from_numpy(np.ascontiguousarray(a)) for a in v] for k, v in pyr.items()}
