"""Micro-benchmark of the KPConv neighbour-gather kernel on the real pyramid of one synthetic pair (GPU box).
Prints per-layer time and G-roofline GB/s. Knobs come from the environment (RDM_GATHER_VEC=2|4)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdmnet_b200 import ops, synthetic, _lib as L
from rdmnet_b200.model import build_pyramid_gpu

src = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
if src == "bundled":
    sc = dict(np.load(os.path.join(ROOT, "tests/golden/scans.npz")))
    a, b = sc["s000000"], sc["s000004"]
else:
    p = synthetic.make_pair(0)
    a, b = p["ref_points"], p["src_points"]
pts = torch.from_numpy(np.concatenate([a, b])).cuda()
lens = torch.tensor([len(a), len(b)]).cuda()
gp = build_pyramid_gpu(pts, lens, 5, 0.3, 4.25 * 0.3, [65, 63, 69, 70, 81])
pyr = gp.as_data_dict()
P, NB, SUB = pyr["points"], pyr["neighbors"], pyr["subsampling"]
ORDER = [gp.table("order", s) for s in range(5)] if os.environ.get("BG_ORDER", "1") == "1" else [None] * 5
# (C_in, C_out, query stage, support stage, table) of the 14 KPConv calls (experiments/backbone.py:11-70)
layers = [(1, 64, 0, 0, NB[0]), (32, 32, 0, 0, NB[0])]
for s_ in range(1, 5):
    c = 32 * 2 ** (s_ - 1)
    layers += [(c, c, s_, s_ - 1, SUB[s_ - 1]), (2 * c, 2 * c, s_, s_, NB[s_]), (2 * c, 2 * c, s_, s_, NB[s_])]
torch.manual_seed(0)
# kernel points of the pretrained checkpoint, layer by layer (their spread scales with the stage: the sparsity of the
# influences - ~1.7 non-zero of 15 per neighbour - depends on it; a fixed random set made the deep stages dense)
CK = os.path.join(ROOT, "tests", "golden", "_big", "rdmnet_state.pt")
SD = torch.load(CK, map_location="cpu", weights_only=True) if os.path.exists(CK) else None
NAMES = ["encoder1_1", "encoder1_2"] + [f"encoder{s}_{j}" for s in range(2, 6) for j in (1, 2, 3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot_b = tot_t = 0.0
print(f"RDM_GATHER_MODE={os.environ.get('RDM_GATHER_MODE', 'auto')} BG_ORDER={os.environ.get('BG_ORDER', '1')} source={src}")
for li, (cin, cout, qs, ss, tab) in enumerate(layers):
    m, h = tab.shape
    n = P[ss].shape[0]
    feats = torch.randn(n, cin, device="cuda") if cin > 1 else torch.ones(n, 1, device="cuda")
    w = torch.zeros(15, cin, 1, device="cuda")
    sigma = 0.6 * 2 ** ss
    kp = (SD[f'encoder.{NAMES[li]}.KPConv.kernel_points'] if SD is not None else torch.randn(15, 3) * 0.66 * sigma).cuda().contiguous()
    out = torch.empty((m, 15 * cin), device="cuda")
    rowpos = torch.empty(n, dtype=torch.uint8, device="cuda")
    hk = ops._host_copy(kp)
    def run():
        L.call("rdm_kpconv_gather", L.ptr(feats), L.ptr(P[qs]), L.ptr(P[ss]), L.ptr(tab), 4, L.ptr(kp), hk.data_ptr(),
               float(sigma), m, n, h, cin, L.ptr(ORDER[qs]), L.ptr(out), L.ptr(rowpos), L.stream())
    iters = int(os.environ.get('BG_ITERS', '10'))
    for _ in range(3 if iters > 1 else 0):
        run()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    g = ops.kpconv_gather_bytes(m, h, cin, cout, 4)
    valid = float((tab < n).float().mean())
    tot_b += g; tot_t += t
    print(f"C{cin:4d} M{m:6d} N{n:6d} H{h:3d} fill {valid:.2f}  {t*1e3:7.1f} us  {g/t/1e6:7.0f} GB/s")
print(f"TOTAL {tot_t*1e3:.1f} us  {tot_b/tot_t/1e6:.0f} GB/s  frac {tot_b/tot_t/1e6/6539.9:.3f}")
