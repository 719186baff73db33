import ctypes, sys
sys.path.insert(0, '/root/repo')
import torch
from rdmnet_b200 import _lib as L
lib = ctypes.CDLL(L.LIB_PATH)
lib.rdm_debug_gemm_timeline.argtypes = [ctypes.c_int] * 3
torch.cuda.init(); torch.zeros(1).cuda()
for m, n, k in ((23319, 32, 480), (8841, 64, 960), (494, 512, 7680), (3078, 512, 1536)):
    lib.rdm_debug_gemm_timeline(m, n, k)
