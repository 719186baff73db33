"""Which knob changes the src_feats_c error on the 4k synthetic pair (GPU box)? Prints the error of the full GPU path
(GPU-built pyramid) and of the tables path (oracle pyramid fed as a data_dict) against the CPU oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as MO, pyramid as OP
from rdmnet_b200 import synthetic
from rdmnet_b200.model import create_model

def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max()).item()

sd = torch.load(os.path.join(ROOT, "tests/golden/_big/rdmnet_state.pt"), map_location="cpu", weights_only=True)
ne, na = synthetic.SIZE_CLASSES[sys.argv[1] if len(sys.argv) > 1 else "4k"]
p = synthetic.make_pair(pair_id=11, n_elev=ne, n_azim=na)
pts = np.concatenate([p["ref_points"], p["src_points"]]); lens = [len(p["ref_points"]), len(p["src_points"])]
pyr = OP.precompute_pyramid(pts, lens, 5, 0.3, 4.25 * 0.3, MO.DEFAULT_LIMITS, "port")
tp = MO.pyramid_to_torch(pyr)
torch.set_num_threads(16)
with torch.no_grad():
    ref = MO.forward(sd, tp, lambda q, l: OP.radius_search(q.numpy(), q.numpy(), l.numpy(), l.numpy(), 2.4, 81, "port"))
m = create_model(); m.load_state_dict(sd, strict=True); m = m.cuda().eval()
with torch.no_grad():
    out = m({"points": torch.from_numpy(pts).cuda(), "lengths": torch.tensor(lens, dtype=torch.int64).cuda()})
    dd = {k: [t.cuda() for t in v] for k, v in tp.items()}
    out2 = m(dd)
knobs = {k: os.environ[k] for k in os.environ if k.startswith("RDM_")}
nc = int(tp["lengths"][-1][0])
for name, o in (("gpu-pyramid", out), ("tables", out2)):
    print(knobs, name, "src_feats_c %.2e ref_feats_c %.2e feats_f %.2e shifted %.2e" % (
        rel(o["src_feats_c"], ref["src_feats_c"]), rel(o["ref_feats_c"], ref["ref_feats_c"]),
        rel(o["ref_feats_f"], ref["feats_f"][:o["ref_feats_f"].shape[0]]), rel(o["shifted_ref_points_c"], ref["shifted_points_c"][:nc])))
