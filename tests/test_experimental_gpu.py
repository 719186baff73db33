"""EXPERIMENTAL kernels that are compiled but off by default. Not part of the default GPU suite: a protocol bug in a
tcgen05 kernel traps the context and would take the following tests with it. Run explicitly on a GPU box:

    RDM_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu -q
"""
import ctypes
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RDM_TEST_EXPERIMENTAL", "0") != "1",
                                 reason="experimental kernels run only with RDM_TEST_EXPERIMENTAL=1")]


@pytest.mark.parametrize("m,n,k", [(300, 128, 64), (23319, 32, 64), (8841, 64, 960), (494, 512, 7680), (494, 2048, 512),
                                   (3078, 257, 768), (129, 72, 40), (2236, 1024, 1284)])
def test_linear_a_in_tmem_variant(m, n, k):
    """gemm_tc_atmem.cu (A operand staged in TMEM by tcgen05.st, 'ts' MMAs) vs an fp64 reference, same bar as the
    default tcgen05 GEMM (tests/test_backbone_gpu.py::test_linear_tensor_core_shapes)."""
    from rdmnet_b200 import ops, _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.rdm_debug_gemm_variant.argtypes = [ctypes.c_int]
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k)
    w = torch.randn(n, k) / k ** 0.5
    b = torch.randn(n)
    ref = (x.double() @ w.double().t() + b.double())
    lib.rdm_debug_gemm_variant(1)
    try:
        got = ops.linear(x.cuda(), w.cuda(), b.cuda())
        torch.cuda.synchronize()
    finally:
        lib.rdm_debug_gemm_variant(0)
    err = (got.cpu().double() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), err
    base = ops.linear(x.cuda(), w.cuda(), b.cuda())
    assert (got - base).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
