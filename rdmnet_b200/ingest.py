"""Ingest side of the path: raw KITTI velodyne scan -> 0.3 m voxel-barycentre downsample on the GPU, the
step the reference performs offline with open3d (preporcess/downsample_pcd_kitti.py:20-36) before its dataset loader
reads the resulting .npy files (rdmnet/datasets/registration/kitti/dataset.py:154-160)."""
import numpy as np
import torch

from . import _lib as L


def load_kitti_bin(path):
    """KITTI velodyne .bin: float32 (x, y, z, intensity) rows (downsample_pcd_kitti.py:20)."""
    return np.fromfile(path, dtype=np.float32).reshape(-1, 4)


def voxel_downsample(points, voxel_size=0.3):
    """open3d `voxel_down_sample` semantics on the GPU (rdm_voxel_downsample): points (N,3) xyz or (N,4) xyzi, numpy or
    tensor -> tensor of voxel means on the device, voxels in first-occurrence order. One host sync (the output shape)."""
    if not torch.cuda.is_available():
        raise RuntimeError("rdmnet_b200 needs a CUDA device: there is no CPU path")
    dev = points.device if torch.is_tensor(points) and points.is_cuda else torch.device("cuda", torch.cuda.current_device())
    p = torch.as_tensor(points, dtype=torch.float32).to(dev).contiguous()
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise RuntimeError("points must be (N,3) or (N,4)")
    n, stride = p.shape
    out = torch.empty_like(p)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = int(L.lib().rdm_voxel_downsample_workspace(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    L.call("rdm_voxel_downsample", L.ptr(p), stride, n, float(voxel_size), L.ptr(out), L.ptr(cnt), L.ptr(ws), wsb, L.stream())
    return out[:int(cnt.item())]


def downsample_kitti_scan(path, voxel_size=0.3):
    """.bin file -> (M,4) float32 numpy array, the content of the reference's downsampled_xyzi/<seq>/<frame>.npy."""
    return voxel_downsample(load_kitti_bin(path), voxel_size).cpu().numpy()
