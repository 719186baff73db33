"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the voxel downsample the reference runs offline through open3d
(preporcess/downsample_pcd_kitti.py:28, `pcd.voxel_down_sample(0.3)` with the intensity as colour).

open3d is a third-party dependency that is absent from /root/reference and from this image (requirements.txt pins no
version); its published algorithm (open3d/geometry/PointCloud.cpp, VoxelDownSample): voxel_min_bound = min_bound -
voxel_size * 0.5, index = floor((p - voxel_min_bound) / voxel_size), one output point per occupied voxel = the mean of its
points / colours, emitted in the iteration order of a std::unordered_map (unspecified). PARITY UNPINNED against open3d
itself: the values follow the published definition; the order here is the canonical first-occurrence order."""
import numpy as np


def voxel_downsample(points, voxel):
    p = np.asarray(points, np.float32)
    o = p[:, :3].min(0) - np.float32(0.5) * np.float32(voxel)
    idx = np.floor((p[:, :3] - o) / np.float32(voxel)).astype(np.int64)
    _, first, inv = np.unique(idx, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")  # voxels by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    g = rank[inv]
    cnt = np.bincount(g, minlength=order.shape[0]).astype(np.float64)
    out = np.stack([np.bincount(g, weights=p[:, a].astype(np.float64), minlength=order.shape[0]) / cnt for a in range(p.shape[1])], 1)
    return out.astype(np.float32)
