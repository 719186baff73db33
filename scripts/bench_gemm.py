"""Micro-benchmark of rdm_linear on the GEMM shapes of one synthetic pair (GPU box): back-to-back launches, CUDA events."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdmnet_b200 import ops, _lib as L

shapes = [(23319, 32, 64), (23319, 32, 480), (23319, 128, 32), (23319, 128, 64), (8841, 64, 960), (8841, 256, 64),
          (8841, 257, 768), (3078, 128, 1920), (3078, 512, 1536), (3078, 512, 128), (1106, 256, 3840), (1106, 1024, 256),
          (494, 512, 7680), (494, 2048, 512), (494, 1024, 1284), (431, 128, 2048), (128, 64, 32), (128, 64, 128)]
iters = int(os.environ.get("BG_ITERS", "20"))
print("RDM_GEMM_TC=", os.environ.get("RDM_GEMM_TC", "1"), "presplit registry:", os.environ.get("RDM_LINEAR_USE_REGISTRY", "0"))
tot = 0.0
for m, n, k in shapes:
    x = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    b = torch.randn(n, device="cuda")
    if os.environ.get("RDM_LINEAR_USE_REGISTRY", "0") == "1" and k % 4 == 0:  # pre-split weights (what the runners use)
        split = torch.empty((2, n, k), device="cuda")
        L.call("rdm_presplit_weight", w.data_ptr(), n, k, split.data_ptr(), L.stream())
        L.lib().rdm_presplit_register(w.data_ptr(), split.data_ptr())
    for _ in range(2 if iters > 1 else 0):
        ops.linear(x, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get("BG_GRAPH", "1") == "1" and iters > 1:
        # replay a CUDA graph of `iters` back-to-back calls: device time per launch without the Python launch cost
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                y = ops.linear(x, w, b)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(iters):
            y = ops.linear(x, w, b)
        e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters
    ref = (x.double() @ w.double().t() + b.double())
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    tot += t
    print(f"M{m:6d} N{n:5d} K{k:5d}  {t*1e3:8.1f} us  {2*m*n*k/t/1e9:8.1f} TFLOP/s  rel.err {err:.2e}")
print(f"TOTAL {tot*1e3:.1f} us")
