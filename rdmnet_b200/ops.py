"""Functional operators of the hot path (host side). Mirrors ``geotransformer.modules.ops`` and the functional
parts of ``geotransformer.modules.kpconv`` (same names, argument meaning and error behaviour), executing on the GPU
through librdm_sm100.so. Reference citations are relative to /root/reference.
"""
import torch

from . import _lib as L


def _chk(t, dtype, name, ndim=None):
    # same preconditions as the reference extension (geotransformer/extensions/common/torch_helper.h:6-35)
    if not torch.is_tensor(t):
        raise RuntimeError(f"{name} must be a tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {dtype} tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if ndim is not None and t.ndim != ndim:
        raise RuntimeError(f"{name} must be {ndim}-d")


def _dev(t, device):
    return t if t.is_cuda else t.to(device, non_blocking=True)


def _cuda_device(*ts):
    for t in ts:
        if t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("rdmnet_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------------- pyramid
def grid_subsample(points, lengths, voxel_size):
    """geotransformer/modules/ops/grid_subsample.py:7-22 / rdmnet.ext.grid_subsampling.

    points (N,3) f32, lengths (B,) i64 -> (s_points (M,3), s_lengths (B,)), bit-exact incl. row order.
    Outputs live on the device of `points` (CPU inputs are staged through the GPU, results copied back).
    """
    _chk(points, torch.float32, "points", 2)
    _chk(lengths, torch.int64, "lengths", 1)
    dev = _cuda_device(points, lengths)
    p, l = _dev(points, dev), _dev(lengths, dev)
    n, b = p.shape[0], l.shape[0]
    out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
    out_len = torch.empty(b, dtype=torch.int64, device=dev)
    wsb = L.lib().rdm_grid_subsample_workspace(n, b)
    ws = _ws(wsb, dev)
    with torch.cuda.device(dev):
        L.call("rdm_grid_subsample", L.ptr(p), L.ptr(l), b, n, float(voxel_size), L.ptr(out), L.ptr(out_len), L.ptr(ws),
               wsb, L.stream())
    m = int(out_len.sum().item())  # the one host sync: the output shape is data dependent
    s_points, s_lengths = out[:m], out_len
    if not points.is_cuda:
        s_points = s_points.cpu()
    if not lengths.is_cuda:
        s_lengths = s_lengths.cpu()
    return s_points, s_lengths


def radius_search_raw(q_points, s_points, q_lengths, s_lengths, radius, limit, index_dtype=torch.int64, counts=False):
    """Fixed-width search: returns (indices (Nq, limit), max_count device int32 tensor[, counts (Nq,)]). No host sync."""
    _chk(q_points, torch.float32, "q_points", 2)
    _chk(s_points, torch.float32, "s_points", 2)
    _chk(q_lengths, torch.int64, "q_lengths", 1)
    _chk(s_lengths, torch.int64, "s_lengths", 1)
    if q_lengths.shape[0] != s_lengths.shape[0]:
        raise RuntimeError("q_lengths and s_lengths must have the same batch size")
    dev = _cuda_device(q_points, s_points)
    q, s, ql, sl = (_dev(t, dev) for t in (q_points, s_points, q_lengths, s_lengths))
    nq, ns, b = q.shape[0], s.shape[0], ql.shape[0]
    out = torch.empty((nq, max(limit, 0)), dtype=index_dtype, device=dev)
    maxc = torch.empty(1, dtype=torch.int32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev) if counts else None
    wsb = L.lib().rdm_radius_search_workspace(ns, b)
    ws = _ws(wsb, dev)
    with torch.cuda.device(dev):
        L.call("rdm_radius_search", L.ptr(q), L.ptr(s), L.ptr(ql), L.ptr(sl), b, nq, ns, ns, float(radius), int(limit),
               L.ptr(out), out.element_size(), L.ptr(cnt), L.ptr(maxc), L.ptr(ws), wsb, L.stream())
    return (out, maxc, cnt) if counts else (out, maxc)


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius):
    """rdmnet.ext.radius_neighbors (radius_neighbors.cpp:5-67): full-width (Nq, max_count) table."""
    _, maxc = radius_search_raw(q_points, s_points, q_lengths, s_lengths, radius, 0)
    width = int(maxc.item())
    out, _ = radius_search_raw(q_points, s_points, q_lengths, s_lengths, radius, max(width, 1))
    out = out[:, :width]
    return out if q_points.is_cuda else out.cpu()


def radius_search(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit):
    """geotransformer/modules/ops/radius_search.py:7-27: (Nq, min(max_count, limit)) int64, padded with Ns."""
    if neighbor_limit <= 0:
        return radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius)
    out, maxc = radius_search_raw(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit)
    width = min(int(maxc.item()), neighbor_limit)
    if width < neighbor_limit:
        out = out[:, :width].contiguous()
    return out if q_points.is_cuda else out.cpu()


# ----------------------------------------------------------------------------------------------- backbone
def _idx_bytes(idx):
    if idx.dtype == torch.int64:
        return 8
    if idx.dtype == torch.int32:
        return 4
    raise RuntimeError("neighbor indices must be int64 or int32")


def linear(x, weight, bias=None, weight_is_kn=False):
    """torch.nn.functional.linear on the fp32 SIMT/tensor path. weight (out,in) or, if weight_is_kn, (in,out)."""
    _chk(x, torch.float32, "x", 2)
    m, k = x.shape
    n = weight.shape[1] if weight_is_kn else weight.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    wsb = L.lib().rdm_linear_workspace(m, n, k) if m * n <= (1 << 20) else 0
    ws = _ws(wsb, x.device) if wsb else None
    L.call("rdm_linear", L.ptr(x), k, L.ptr(weight), weight.shape[1], 0 if weight_is_kn else 1, L.ptr(bias), L.ptr(out),
           n, m, n, k, L.ptr(ws), wsb, L.stream())
    return out


def kpconv(s_feats, q_points, s_points, neighbor_indices, weights, kernel_points, sigma, bias=None):
    """KPConv.forward (geotransformer/modules/kpconv/kpconv.py:79-122)."""
    _chk(s_feats, torch.float32, "s_feats", 2)
    _chk(neighbor_indices, neighbor_indices.dtype, "neighbor_indices", 2)
    m, h = neighbor_indices.shape
    n, c = s_feats.shape
    kk, cin, cout = weights.shape
    if kk != 15 or cin != c:
        raise RuntimeError("kpconv: weights must be (15, C_in, C_out)")
    dev = s_feats.device
    gathered = torch.empty((m, kk * c), dtype=torch.float32, device=dev)
    rowpos = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    L.call("rdm_kpconv_gather", L.ptr(s_feats), L.ptr(q_points), L.ptr(s_points), L.ptr(neighbor_indices),
           _idx_bytes(neighbor_indices), L.ptr(kernel_points), float(sigma), m, n, h, c, L.ptr(gathered), L.ptr(rowpos),
           L.stream())
    return linear(gathered, weights.reshape(kk * c, cout), bias, weight_is_kn=True)


def maxpool(x, neighbor_indices):
    """geotransformer/modules/kpconv/functional.py:54-67."""
    m, h = neighbor_indices.shape
    n, c = x.shape
    out = torch.empty((m, c), dtype=torch.float32, device=x.device)
    L.call("rdm_maxpool", L.ptr(x), L.ptr(neighbor_indices), _idx_bytes(neighbor_indices), m, n, h, c, L.ptr(out), L.stream())
    return out


def nearest_upsample_concat(x, upsample_indices, skip):
    """nearest_upsample (functional.py:6-22) fused with torch.cat([up, skip], dim=1) (experiments/backbone.py:129-141)."""
    m = upsample_indices.shape[0]
    n, c1 = x.shape
    c2 = skip.shape[1] if skip is not None else 0
    out = torch.empty((m, c1 + c2), dtype=torch.float32, device=x.device)
    if upsample_indices.stride(1) != 1:
        raise RuntimeError("upsample_indices must be row-major")
    L.lib()  # ensure loaded
    L.call("rdm_upsample_concat", L.ptr(x), upsample_indices.data_ptr(), _idx_bytes(upsample_indices),
           upsample_indices.stride(0), L.ptr(skip), m, n, c1, c2, L.ptr(out), L.stream())
    return out


def nearest_upsample(x, upsample_indices):
    return nearest_upsample_concat(x, upsample_indices, None)


def group_norm(x, weight, bias, groups, residual=None, act=0, slope=0.1, eps=1e-5):
    """GroupNorm over stacked (N,C) features (kpconv/modules.py:33-50) + optional residual add + LeakyReLU."""
    _chk(x, torch.float32, "x", 2)
    n, c = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(2 * groups, dtype=torch.float64, device=x.device)
    L.call("rdm_groupnorm", L.ptr(x), L.ptr(weight), L.ptr(bias), L.ptr(residual), L.ptr(y), n, c, groups, eps, act, slope,
           L.ptr(stats), L.stream())
    return y


def layer_norm(x, weight, bias, residual=None, relu=False, eps=1e-5):
    _chk(x, torch.float32, "x", 2)
    n, c = x.shape
    y = torch.empty_like(x)
    L.call("rdm_layernorm", L.ptr(x), L.ptr(residual), L.ptr(weight), L.ptr(bias), L.ptr(y), n, c, eps, 2 if relu else 0,
           L.stream())
    return y


def activation(x, act, slope=0.1):
    """act: 1 LeakyReLU(slope), 2 ReLU, 3 clamp(sigmoid(x), 0, 1)."""
    x = x.contiguous()
    y = torch.empty_like(x)
    L.call("rdm_activation", L.ptr(x), L.ptr(y), x.numel(), act, slope, L.stream())
    return y
