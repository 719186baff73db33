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


@pytest.mark.parametrize("m,n,h,c", [(23000, 23000, 65, 32), (9000, 23000, 65, 32), (9000, 9000, 63, 64), (3000, 3000, 69, 128),
                                     (1100, 3000, 69, 128), (500, 500, 81, 512)])
def test_kpconv_gather_persistent_variant_bit_equal(m, n, h, c):
    """kpconv_gather_v4p_kernel (persistent grid pulling work from a counter) == the default v4 kernel, bit for bit
    (same per-item code), including back-to-back launches that rely on the counters re-arming themselves."""
    import numpy as np
    from rdmnet_b200 import ops, _lib as L
    lib = ctypes.CDLL(L.LIB_PATH)
    lib.rdm_debug_gather_persist.argtypes = [ctypes.c_int]
    rng = np.random.default_rng(m + c)
    pts_s = torch.from_numpy(((rng.random((n, 3)) - 0.5) * [60, 40, 4]).astype(np.float32)).cuda()
    pts_q = pts_s[:m].contiguous() if m <= n else torch.from_numpy(((rng.random((m, 3)) - 0.5) * [60, 40, 4]).astype(np.float32)).cuda()
    idx = torch.from_numpy(rng.integers(0, n, size=(m, h)).astype(np.int32))
    k = rng.integers(1, h + 1, size=m)
    idx[torch.from_numpy(np.arange(h)[None, :] >= k[:, None])] = n
    idx = idx.cuda()
    feats = torch.randn(n, c, device="cuda")
    kp = (torch.randn(15, 3) * 0.5).cuda()
    hk = ops._host_copy(kp)
    rowpos = torch.empty(n, dtype=torch.uint8, device="cuda")

    def run():
        out = torch.empty((m, 15 * c), device="cuda")
        L.call("rdm_kpconv_gather", L.ptr(feats), L.ptr(pts_q), L.ptr(pts_s), L.ptr(idx), 4, L.ptr(kp), hk.data_ptr(), 1.2, m, n, h, c,
               None, L.ptr(out), L.ptr(rowpos), L.stream())
        return out

    base = run()
    lib.rdm_debug_gather_persist(1)
    try:
        got = [run() for _ in range(3)]
        torch.cuda.synchronize()
    finally:
        lib.rdm_debug_gather_persist(0)
    for g in got:
        assert torch.equal(g, base)
