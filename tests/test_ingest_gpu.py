"""Ingest: GPU voxel-barycentre downsample of a raw scan vs the numpy restatement of open3d's
voxel_down_sample semantics (oracle/ingest_oracle.py) - same voxel set, same first-occurrence order, means to 1e-5 m."""
import numpy as np
import pytest
import torch

from oracle import ingest_oracle as IO

pytestmark = pytest.mark.gpu


def raw_scan(n, seed, with_intensity=True):
    rng = np.random.default_rng(seed)
    xyz = np.concatenate([rng.normal(0, [25, 18, 1.2], (n, 3)), rng.uniform(-60, 60, (n // 4, 3)) * [1, 1, 0.05]]).astype(np.float32)
    if not with_intensity:
        return xyz
    return np.concatenate([xyz, rng.random((xyz.shape[0], 1)).astype(np.float32)], 1)


@pytest.mark.parametrize("n,stride4,voxel", [(120000, True, 0.3), (20000, False, 0.3), (5000, True, 1.0), (1, True, 0.3), (64, False, 0.05)])
def test_voxel_downsample_vs_oracle(n, stride4, voxel):
    from rdmnet_b200 import ingest
    pts = raw_scan(n, n + 3, stride4)
    got = ingest.voxel_downsample(pts, voxel).cpu().numpy()
    ref = IO.voxel_downsample(pts, voxel)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref).max()
    print(f"[parity] voxel_downsample n={pts.shape[0]} -> {got.shape[0]} voxels, max |err| {err:.2e} m")
    assert err <= 1e-5
    again = ingest.voxel_downsample(torch.from_numpy(pts).cuda(), voxel).cpu().numpy()
    assert np.array_equal(got, again), "deterministic (fixed-point accumulation)"


def test_voxel_downsample_feeds_the_path(tmp_path):
    """raw .bin file -> GPU downsample -> the hot path accepts it (the chain preporcess -> dataset -> model of the reference)."""
    from rdmnet_b200 import ingest
    pts = raw_scan(60000, 9)
    f = tmp_path / "000000.bin"
    pts.tofile(f)
    out = ingest.downsample_kitti_scan(str(f))
    assert out.dtype == np.float32 and out.shape[1] == 4 and 1000 < out.shape[0] < pts.shape[0]
    # idempotence property: every output point is alone in its voxel of the same grid
    o = pts[:, :3].min(0) - np.float32(0.15)
    idx = np.floor((out[:, :3] - o) / np.float32(0.3)).astype(np.int64)
    assert np.unique(idx, axis=0).shape[0] >= out.shape[0] - 8  # means on a voxel face may round into the neighbour cell
