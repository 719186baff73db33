"""Achieved float error of every stage of the GPU path against the CPU oracle, on bundled pair (0,4) with the pretrained
checkpoint: (a) end to end (errors accumulate through the network) and (b) stage by stage with the ORACLE's input fed to
each GPU stage (isolates which kernel loses precision). Prints `max|got-ref| / max|ref|` per tensor. GPU box only.

    python scripts/diag_parity.py [--pair p04|p07]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import model_oracle as MO  # noqa: E402
from oracle import pyramid as OP  # noqa: E402


def rel(got, ref):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pair", default="p04")
    ap.add_argument("--synthetic", default=None, help="size class (4k/8k/16k/32k) of a synthetic pair instead of a bundled one")
    ap.add_argument("--pair-id", type=int, default=11)
    args = ap.parse_args()
    from rdmnet_b200 import ops
    from rdmnet_b200.model import create_model
    scans = dict(np.load(os.path.join(ROOT, "tests", "golden", "scans.npz")))
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "_big", "rdmnet_state.pt"), map_location="cpu", weights_only=True)
    a, b = scans["s000000"], scans["s000004" if args.pair == "p04" else "s000007"]
    if args.synthetic:
        from rdmnet_b200 import synthetic
        ne, na = synthetic.SIZE_CLASSES[args.synthetic]
        pr = synthetic.make_pair(pair_id=args.pair_id, n_elev=ne, n_azim=na)
        a, b = pr["ref_points"], pr["src_points"]
    pyr = OP.precompute_pyramid(np.concatenate([a, b]), [len(a), len(b)], 5, 0.3, 4.25 * 0.3, MO.DEFAULT_LIMITS,
                                "ref" if OP.ref_available() else "port")
    tp = MO.pyramid_to_torch(pyr)
    torch.set_num_threads(os.cpu_count() or 8)
    model = create_model()
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    dd = {k: [t.cuda() for t in v] for k, v in tp.items()}
    P, NB, SUB = tp["points"], tp["neighbors"], tp["subsampling"]
    gP, gNB, gSUB = dd["points"], dd["neighbors"], dd["subsampling"]
    g = MO.CFG["groups"]
    sig0 = MO.CFG["sigma"]
    names = ["encoder1_1", "encoder1_2"] + [f"encoder{s}_{j}" for s in range(2, 6) for j in (1, 2, 3)]
    print("== encoder, block by block: e2e = GPU chain vs oracle chain; iso = GPU block fed the oracle's input")
    x_ref = torch.ones(P[0].shape[0], 1)
    x_gpu = x_ref.cuda()
    with torch.no_grad():
        for name in names:
            st = int(name[7]) - 1
            j = int(name[9])
            strided = j == 1 and st > 0
            if strided:
                q, s, idx, sig = P[st], P[st - 1], SUB[st - 1], sig0 * 2 ** (st - 1)
                gq, gs, gidx = gP[st], gP[st - 1], gSUB[st - 1]
            else:
                q, s, idx, sig = P[st], P[st], NB[st], sig0 * 2 ** st
                gq, gs, gidx = gP[st], gP[st], gNB[st]
            p = "encoder." + name + "."
            blk = getattr(model.encoder, name)
            if name == "encoder1_1":
                y_ref = MO.conv_block(sd, p, x_ref, q, s, idx, sig, g)
            else:
                y_ref = MO.residual_block(sd, p, x_ref, q, s, idx, sig, g, strided)
            y_iso = blk(x_ref.contiguous().cuda(), gq, gs, gidx)
            y_gpu = blk(x_gpu, gq, gs, gidx)
            # inside the block: the KPConv alone (gather + weight GEMM) on the oracle's input
            kp = blk.KPConv
            xin = x_ref
            if name != "encoder1_1" and (p + "unary1.mlp.weight") in sd:
                xin = MO.unary(sd, p + "unary1.", x_ref, g)
            k_ref = MO.kpconv(xin, q, s, idx, sd[p + "KPConv.weights"], sd[p + "KPConv.kernel_points"], sig, sd.get(p + "KPConv.bias"))
            k_gpu = kp(xin.contiguous().cuda(), gq, gs, gidx)
            # gather alone vs fp64 recomputation is not available from the oracle; report the fused op
            print(f"{name:12s} C={tuple(y_ref.shape)} e2e {rel(y_gpu, y_ref):.2e}  iso {rel(y_iso, y_ref):.2e}  kpconv-only {rel(k_gpu, k_ref):.2e}")
            x_ref, x_gpu = y_ref.contiguous(), y_gpu
    # unary GEMM precision alone: one big Linear vs fp64
    with torch.no_grad():
        w = sd["encoder.encoder3_2.unary2.mlp.weight"]
        xin = torch.randn(6255, w.shape[1])
        ref64 = (xin.double() @ w.double().t())
        got = ops.linear(xin.cuda(), w.cuda(), None)
        cpu32 = F.linear(xin, w)
        print(f"linear 6255x{w.shape[1]}x{w.shape[0]}: gpu vs fp64 {rel(got, ref64):.2e}   torch-cpu fp32 vs fp64 {rel(cpu32, ref64):.2e}")
    print("== whole forward vs oracle")
    with torch.no_grad():
        ref = MO.forward(sd, tp, lambda p_, l_: OP.radius_search(p_.numpy(), p_.numpy(), l_.numpy(), l_.numpy(), 2.4, 81, "port"))
        dd["features"] = torch.ones(tp["points"][0].shape[0], 1).cuda()
        out = model(dd)
        feats = model.encoder(dd["features"], dd)
    nc = int(tp["lengths"][-1][0])
    nf = out["ref_feats_f"].shape[0]
    print(f"feats_s5 (runner)      {rel(feats[-1], ref['feats_s5']):.2e}")
    print(f"shifted_points_c       {rel(out['shifted_ref_points_c'], ref['shifted_points_c'][:nc]):.2e}")
    print(f"feats_f                {rel(out['ref_feats_f'], ref['feats_f'][:nf]):.2e}")
    print(f"nms mask equal         {bool(np.array_equal(out['mask'].cpu().numpy(), ref['nms_masks'].numpy()))}")
    if out["ref_feats_c"].shape == ref["ref_feats_c"].shape:
        print(f"ref_feats_c            {rel(out['ref_feats_c'], ref['ref_feats_c']):.2e}")
        print(f"src_feats_c            {rel(out['src_feats_c'], ref['src_feats_c']):.2e}")
    if out["corr_scores"].shape == ref["corr_scores"].shape:
        print(f"corr_scores            {rel(out['corr_scores'], ref['corr_scores']):.2e}")
    print(f"estimated_transform    {rel(out['estimated_transform'], ref['estimated_transform']):.2e}")
    # isolated stages on the oracle's inputs
    print("== isolated stages on the oracle's inputs")
    with torch.no_grad():
        pc = tp["points"][-1]
        r1, s1 = model.transformer(pc[:nc].cuda(), pc[nc:].cuda(), ref["feats_s5"][:nc].contiguous().cuda(), ref["feats_s5"][nc:].contiguous().cuda())
        print(f"transformer1 ref/src   {rel(r1, ref['ref_feats_t1']):.2e} {rel(s1, ref['src_feats_t1']):.2e}")
        tfc = torch.cat([ref["ref_feats_t1"], ref["src_feats_t1"]], 0)
        sh, vf = model.vote(pc.cuda(), tfc.cuda())
        print(f"vote xyz / feats       {rel(sh, ref['shifted_points_c']):.2e} {rel(vf, ref['vote_feats_c']):.2e}")
        rm, sm = ref["nms_masks"][:nc], ref["nms_masks"][nc:]
        r2, s2 = model.transformer2(ref["ref_points_c"].cuda(), ref["src_points_c"].cuda(), ref["vote_feats_c"][:nc][rm].cuda(),
                                    ref["vote_feats_c"][nc:][sm].cuda())
        r2n = F.normalize(r2, p=2, dim=1)
        print(f"transformer2+norm ref  {rel(r2n, ref['ref_feats_c']):.2e}")
        # sinkhorn on the oracle's patch scores
        rknn, sknn = ref["ref_node_knn_indices"], ref["src_node_knn_indices"]
        rci, sci = ref["ref_node_corr_indices"], ref["src_node_corr_indices"]
        ff = ref["feats_f"]
        rpf = torch.cat([ff[:nf], torch.zeros(1, ff.shape[1])], 0)
        spf = torch.cat([ff[nf:], torch.zeros(1, ff.shape[1])], 0)
        ms_in = torch.einsum("bnd,bmd->bnm", rpf[rknn[rci]], spf[sknn[sci]]) / ff.shape[1] ** 0.5
        rkm = rknn[rci] < nf
        skm = sknn[sci] < (ff.shape[0] - nf)
        got = ops.sinkhorn(ms_in.cuda(), rkm.cuda(), skm.cuda(), sd["optimal_transport.alpha"].cuda(), 100)
        want = MO.sinkhorn(ms_in, rkm, skm, sd["optimal_transport.alpha"])
        live = want > -1e11
        err = (got.cpu()[live] - want[live]).abs().max().item()
        print(f"sinkhorn abs err on live entries {err:.2e} (|ref| max {want[live].abs().max().item():.2f}); exp-domain rel "
              f"{rel(got.cpu()[live].exp(), want[live].exp()):.2e}")
        ps = ops.patch_scores(ff[:nf].cuda().contiguous(), ff[nf:].cuda().contiguous(), rknn.cuda(), sknn.cuda(), rci.cuda(), sci.cuda())
        print(f"patch_scores           {rel(ps, ms_in):.2e}")


if __name__ == "__main__":
    main()
