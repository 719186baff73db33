"""Host-side mirror of the registration helpers the unmodified experiments/{model,model_infer,loss,infer}.py import:

  geotransformer.modules.registration.matching   get_node_correspondences (:252-366), get_node_overlap (:368-436),
                                                  get_node_correspondences_disance (:441-503)
  geotransformer.modules.registration.metrics    relative_rotation_error, relative_translation_error,
                                                  isotropic_transform_error (:46-111)
  geotransformer.modules.registration.procrustes weighted_procrustes, WeightedProcrustes (:6-91)
  geotransformer.utils.open3d                    registration_with_ransac_from_correspondences (:173-203)
  geotransformer.utils.registration              get_correspondences (:203-217; the ground-truth ball query of experiments/loss.py:92,151)

The ground-truth overlap / mask computations and RANSAC run as CUDA kernels of librdm_sm100.so (csrc/gt.cu, lgr.cu); the
metrics are a handful of scalar operations on one 4x4 pose and stay in torch, like in the reference.
"""
import numpy as np
import torch

from . import _lib as L
from . import ops
from .modules import WeightedProcrustes  # noqa: F401  (re-export)
from .ops import weighted_procrustes  # noqa: F401  (re-export)


def _u8(t):
    return None if t is None else t.to(torch.uint8).contiguous()


def _f32(t):
    if t.dtype != torch.float32:
        raise RuntimeError("float32 tensors expected")
    return t.contiguous()


def _node_overlaps(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, ref_masks, src_masks,
                   ref_knn_masks, src_knn_masks, sphere_filter, compact):
    m, n, k = ref_nodes.shape[0], src_nodes.shape[0], ref_knn_points.shape[1]
    if src_knn_points.shape[1] != k:
        raise RuntimeError("ref and src patches must hold the same number of points")
    dev = ref_nodes.device
    overlaps = torch.empty((m, n), dtype=torch.float32, device=dev)
    intersect = torch.empty((m, n), dtype=torch.uint8, device=dev)
    idx = torch.empty((max(m * n, 1), 2), dtype=torch.int64, device=dev) if compact else None
    val = torch.empty(max(m * n, 1), dtype=torch.float32, device=dev) if compact else None
    cnt = torch.zeros(1, dtype=torch.int32, device=dev) if compact else None
    wsb = int(L.lib().rdm_node_correspondences_workspace(m, n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = [_u8(ref_masks), _u8(src_masks), _u8(ref_knn_masks), _u8(src_knn_masks)]
    L.call("rdm_node_correspondences", L.ptr(_f32(ref_nodes)), L.ptr(_f32(src_nodes)), L.ptr(_f32(ref_knn_points)),
           L.ptr(_f32(src_knn_points)), L.ptr(_f32(transform)) if transform is not None else None, float(pos_radius), m, n, k,
           L.ptr(keep[0]), L.ptr(keep[1]), L.ptr(keep[2]), L.ptr(keep[3]), 1 if sphere_filter else 0, L.ptr(overlaps),
           L.ptr(intersect), L.ptr(idx), L.ptr(val), L.ptr(cnt), L.ptr(ws), wsb, L.stream())
    return overlaps, intersect, idx, val, cnt


@torch.no_grad()
def get_node_correspondences(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, ref_masks=None,
                             src_masks=None, ref_knn_masks=None, src_knn_masks=None, return_mask=False):
    """matching.py:252-366 -> (corr_indices (C,2) int64, corr_overlaps (C,)); with return_mask the (M,N) sphere-test
    matrix. One host sync for C (the output shape is data dependent)."""
    overlaps, intersect, idx, val, cnt = _node_overlaps(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform,
                                                        pos_radius, ref_masks, src_masks, ref_knn_masks, src_knn_masks, True,
                                                        not return_mask)
    if return_mask:
        return intersect.view(torch.bool)
    c = int(cnt.item())
    return idx[:c], val[:c]


@torch.no_grad()
def get_node_overlap(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, ref_masks=None,
                     src_masks=None, ref_knn_masks=None, src_knn_masks=None, return_mask=False):
    """matching.py:368-436: patch-overlap ratio of patch i of the ref side against patch i of the src side -> (B,), no
    sphere pre-filter. (The kernel evaluates the whole (B,B) pair matrix; the reference's result is its diagonal.)"""
    if ref_knn_points.shape[0] != src_knn_points.shape[0]:
        raise RuntimeError("get_node_overlap pairs patch i with patch i: both sides need the same number of patches")
    overlaps, _, _, _, _ = _node_overlaps(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, None,
                                          None, ref_knn_masks, src_knn_masks, False, False)
    return overlaps.diagonal().contiguous()


@torch.no_grad()
def get_node_correspondences_disance(ref_nodes, src_nodes, transform, pos_radius, ref_masks=None, src_masks=None):
    """matching.py:441-503 (sic): (M,N) bool mask of mutually-nearest node pairs closer than the threshold."""
    m, n = ref_nodes.shape[0], src_nodes.shape[0]
    mask = torch.empty((m, n), dtype=torch.uint8, device=ref_nodes.device)
    keep = [_u8(ref_masks), _u8(src_masks)]
    L.call("rdm_node_distance_mask", L.ptr(_f32(ref_nodes)), L.ptr(_f32(src_nodes)), L.ptr(_f32(transform)), float(pos_radius), m,
           n, L.ptr(keep[0]), L.ptr(keep[1]), L.ptr(mask), L.stream())
    return mask.view(torch.bool)


def compact_nonzero(mat):
    """torch.nonzero(mat > 0) in row-major order (+ values) without leaving the device until the count is needed."""
    rows, cols = mat.shape
    dev = mat.device
    idx = torch.empty((max(rows * cols, 1), 2), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    if mat.dtype == torch.float32:
        val = torch.empty(max(rows * cols, 1), dtype=torch.float32, device=dev)
        L.call("rdm_compact_nonzero", L.ptr(mat.contiguous()), None, rows, cols, L.ptr(idx), L.ptr(val), L.ptr(cnt), L.stream())
        c = int(cnt.item())
        return idx[:c], val[:c]
    b = mat.to(torch.uint8).contiguous()
    L.call("rdm_compact_nonzero", None, L.ptr(b), rows, cols, L.ptr(idx), None, L.ptr(cnt), L.stream())
    return idx[:int(cnt.item())], None


def get_correspondences(ref_points, src_points, transform=None, matching_radius=None):
    """geotransformer/utils/registration.py:203-217: all (ref index, src index) pairs closer than `matching_radius` after moving
    src by `transform`. The reference builds a scipy cKDTree on the host per call (experiments/loss.py:92 and :151 hand it
    `.detach().cpu().numpy()` arrays every iteration); here the ball query is the uniform-grid radius search of librdm_sm100
    at its full width and a device-side compaction. Numpy (or tensors) in, (K,2) int64 numpy out like the reference, rows in
    (ref, ascending distance) order - the callers only scatter 1s with them. `d < r` strict (nanoflann's rule, as everywhere
    in this library) against cKDTree's `d <= r`: differs only for a pair at exactly the radius."""
    if not torch.cuda.is_available():
        raise RuntimeError("rdmnet_b200 needs a CUDA device: there is no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device())
    as_t = lambda a: (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).to(dev, torch.float32).contiguous()  # noqa: E731
    ref, src = as_t(ref_points), as_t(src_points)
    if transform is not None:
        src = ops.apply_transform(src, as_t(transform))
    m, n = ref.shape[0], src.shape[0]
    if m == 0 or n == 0:
        return np.zeros((0, 2), dtype=np.int64)
    ql = torch.tensor([m], dtype=torch.int64, device=dev)
    sl = torch.tensor([n], dtype=torch.int64, device=dev)
    _, maxc = ops.radius_search_raw(ref, src, ql, sl, float(matching_radius), 0, index_dtype=torch.int32)
    width = int(maxc.item())  # the widest ball: the table below is complete (no neighbour limit, like the KD-tree query)
    if width == 0:
        return np.zeros((0, 2), dtype=np.int64)
    table, _ = ops.radius_search_raw(ref, src, ql, sl, float(matching_radius), width, index_dtype=torch.int32)  # padded with n
    pos, _ = compact_nonzero(table < n)  # (K, 2) = (row, column) of the live slots, row-major
    if pos.shape[0] == 0:
        return np.zeros((0, 2), dtype=np.int64)
    flat = (pos[:, 0] * table.shape[1] + pos[:, 1]).contiguous()
    j = ops.index_select(table.reshape(-1).contiguous(), flat, 0)
    return torch.stack([pos[:, 0], j.to(torch.int64)], 1).cpu().numpy().astype(np.int64)


# ---------------------------------------------------------------------------------------------------------- metrics
def relative_rotation_error(gt_rotations, rotations):
    """metrics.py:46-65: acos((trace(R^T R_gt) - 1) / 2) in degrees."""
    mat = torch.matmul(rotations.transpose(-1, -2), gt_rotations)
    trace = mat[..., 0, 0] + mat[..., 1, 1] + mat[..., 2, 2]
    x = (0.5 * (trace - 1.0)).clamp(min=-1.0, max=1.0)
    return 180.0 * torch.arccos(x) / np.pi


def relative_translation_error(gt_translations, translations):
    """metrics.py:68-81."""
    return torch.linalg.norm(gt_translations - translations, dim=-1)


def isotropic_transform_error(gt_transforms, transforms, reduction="mean"):
    """metrics.py:84-111."""
    assert reduction in ["mean", "sum", "none"]
    gt_r, gt_t = ops.get_rotation_translation_from_transform(gt_transforms)
    r, t = ops.get_rotation_translation_from_transform(transforms)
    rre, rte = relative_rotation_error(gt_r, r), relative_translation_error(gt_t, t)
    if reduction == "mean":
        return rre.mean(), rte.mean()
    if reduction == "sum":
        return rre.sum(), rte.sum()
    return rre, rte


# ------------------------------------------------------------------------------------------------------------ RANSAC
def registration_with_ransac_from_correspondences(src_points, ref_points, correspondences=None, distance_threshold=0.05,
                                                  ransac_n=3, num_iterations=10000, seed=7351, return_info=False):
    """geotransformer/utils/open3d.py:173-203 with the hypotheses evaluated on the GPU (rdm_ransac_correspondences):
    numpy (or tensor) points in, (4,4) float64 numpy transform from src to ref out, like the open3d call it replaces.
    Sampling is a counter-based generator, so results are reproducible for a given `seed` (open3d's are not)."""
    if not torch.cuda.is_available():
        raise RuntimeError("rdmnet_b200 needs a CUDA device: there is no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device())
    src = torch.as_tensor(np.asarray(src_points) if not torch.is_tensor(src_points) else src_points, dtype=torch.float32).to(dev)
    ref = torch.as_tensor(np.asarray(ref_points) if not torch.is_tensor(ref_points) else ref_points, dtype=torch.float32).to(dev)
    if correspondences is not None:
        corr = torch.as_tensor(np.asarray(correspondences), dtype=torch.int64).to(dev)
        src, ref = ops.index_select(src, corr[:, 0].contiguous(), 0), ops.index_select(ref, corr[:, 1].contiguous(), 0)
    src, ref = src.contiguous(), ref.contiguous()
    c = src.shape[0]
    if c < ransac_n:
        return (np.eye(4), {"inliers": 0, "iteration": -1}) if return_info else np.eye(4)
    T = torch.empty((4, 4), dtype=torch.float32, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    wsb = int(L.lib().rdm_ransac_workspace(c))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    L.call("rdm_ransac_correspondences", L.ptr(src), L.ptr(ref), c, float(distance_threshold), int(ransac_n), int(num_iterations),
           int(seed), L.ptr(T), L.ptr(meta), L.ptr(ws), wsb, L.stream())
    Th = T.cpu().numpy().astype(np.float64)
    if return_info:
        inl, it = meta.tolist()
        return Th, {"inliers": int(inl), "iteration": int(it)}
    return Th
