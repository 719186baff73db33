"""Both tcgen05 GEMM kernels against an fp64 reference on the shapes of the backbone: variant 1 = A operand in tensor
memory (gemm_tc_atmem.cu, the default), variant 0 = all operands in shared memory (gemm_tc.cu, RDM_GEMM_ATMEM=0)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(300, 128, 64), (23319, 32, 64), (8841, 64, 960), (494, 512, 7680), (494, 2048, 512), (3078, 257, 768),
          (129, 72, 40), (2236, 1024, 1284)]


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_linear_tcgen05_variants_vs_fp64(m, n, k, variant):
    from rdmnet_b200 import ops, _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.rdm_debug_gemm_variant.argtypes = [ctypes.c_int]
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5
    b = torch.randn(n)
    ref = (x.double() @ w.double().t() + b.double())
    lib.rdm_debug_gemm_variant(variant)
    try:
        n0 = _lib.lib().rdm_tc_gemm_count()
        got = ops.linear(x.cuda(), w.cuda(), b.cuda())
        torch.cuda.synchronize()
        used_tc = _lib.lib().rdm_tc_gemm_count() > n0
    finally:
        lib.rdm_debug_gemm_variant(1)
    err = (got.cpu().double() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    print(f"variant {variant} {m}x{n}x{k}: rel err vs fp64 {err:.2e} (tensor cores: {used_tc})")
    assert err <= 2e-5, err
