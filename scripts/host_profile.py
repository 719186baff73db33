"""cProfile of the single host thread that drives the pair pipeline (GPU box): where do the ~3.5 ms of host time per pair go?"""
import cProfile, os, pstats, sys, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdmnet_b200 import synthetic
from rdmnet_b200.model import PairPipeline, create_model
m = create_model()
ck = os.path.join(ROOT, "tests/golden/_big/rdmnet_state.pt")
if os.path.exists(ck):
    m.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True), strict=True)
m = m.cuda().eval()
items = []
for i in range(4):
    p = synthetic.make_pair(pair_id=i)
    items.append((torch.from_numpy(np.concatenate([p["ref_points"], p["src_points"]])).cuda(),
                  torch.tensor([len(p["ref_points"]), len(p["src_points"])], dtype=torch.int64).cuda()))
pipe = PairPipeline(m)
for _ in pipe.run([items[i % 4] for i in range(12)]):
    pass
torch.cuda.synchronize()
N = 60
pr = cProfile.Profile()
pr.enable()
for _ in pipe.run([items[i % 4] for i in range(N)]):
    pass
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
ps = pstats.Stats(pr, stream=s).sort_stats("tottime")
ps.print_stats(32)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:5000])
