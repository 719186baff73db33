"""Host-side mirror of geotransformer/utils/data.py (the data-preparation half of the hot path): same function names,
arguments and returned dict layout, with the voxel pyramid built on the GPU by librdm_sm100.so instead of inside
forked CPU DataLoader workers.

  precompute_data_stack_mode          data.py:13-77    4x grid_subsample + 13x radius_search -> dict of lists
  single/registration_collate_fn_stack_mode  data.py:80-192   stack the clouds of a batch (+ the pyramid)
  calibrate_neighbors_stack_mode      data.py:195-220  neighbour limits = keep_ratio percentile of the neighbourhood sizes
  build_dataloader_stack_mode         data.py:223-253  DataLoader factory

CUDA cannot run in forked worker processes, so the loader returned by build_dataloader_stack_mode lets its workers
stack the points only (the reference's own `precompute_data=False` branch, data.py:184-189) and builds the pyramid in
the consuming process, on the GPU, as each batch comes out (SURVEY 8(b) "Threading").
"""
import os
from functools import partial

import numpy as np
import torch

from . import ops


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits):
    """data.py:13-77 on the GPU: returns {'points','lengths','neighbors','subsampling','upsampling'} with int64 tables of
    the reference's widths (min(max_count, limit)), on the device the pyramid was built on."""
    from .model import precompute_data_stack_mode as gpu_precompute
    dev = ops._cuda_device(points, lengths)
    d = gpu_precompute(points.to(dev).contiguous(), lengths.to(dev), num_stages, voxel_size, radius,
                       [int(x) for x in neighbor_limits])
    d.pop("lengths_host", None)
    return d


def _merge(data_dicts):
    collated = {}
    for dd in data_dicts:
        for key, value in dd.items():
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(value)
            collated.setdefault(key, []).append(value)
    return collated


def _finish(collated, feats, points_list, batch_size, num_stages, voxel_size, search_radius, neighbor_limits, precompute_data):
    lengths = torch.LongTensor([p.shape[0] for p in points_list])
    points = torch.cat(points_list, dim=0)
    if batch_size == 1:  # data.py:120-123 / :177-180: unwrap single-sample lists
        for key, value in collated.items():
            collated[key] = value[0]
    collated["features"] = feats
    if precompute_data:
        collated.update(precompute_data_stack_mode(points, lengths, num_stages, voxel_size, search_radius, neighbor_limits))
    else:
        collated["points"] = points
        collated["lengths"] = lengths
    collated["batch_size"] = batch_size
    return collated


def single_collate_fn_stack_mode(data_dicts, num_stages, voxel_size, search_radius, neighbor_limits, precompute_data=True):
    """data.py:80-136: clouds stacked as [P_1, ..., P_B]."""
    collated = _merge(data_dicts)
    normals = torch.cat(collated.pop("normals"), dim=0) if "normals" in collated else None
    feats = torch.cat(collated.pop("feats"), dim=0)
    points_list = collated.pop("points")
    out = _finish(collated, feats, points_list, len(data_dicts), num_stages, voxel_size, search_radius, neighbor_limits,
                  precompute_data)
    if normals is not None:
        out["normals"] = normals
    return out


def registration_collate_fn_stack_mode(data_dicts, num_stages, voxel_size, search_radius, neighbor_limits,
                                       precompute_data=True):
    """data.py:139-192: clouds stacked as [ref_1..ref_B, src_1..src_B]."""
    collated = _merge(data_dicts)
    feats = torch.cat(collated.pop("ref_feats") + collated.pop("src_feats"), dim=0)
    points_list = collated.pop("ref_points") + collated.pop("src_points")
    return _finish(collated, feats, points_list, len(data_dicts), num_stages, voxel_size, search_radius, neighbor_limits,
                   precompute_data)


def calibrate_neighbors_stack_mode(dataset, collate_fn, num_stages, voxel_size, search_radius, keep_ratio=0.8,
                                   sample_threshold=2000):
    """data.py:195-220 with the neighbourhood sizes counted and histogrammed on the GPU: per sample the pyramid's points
    come from rdm_grid_subsample, the per-query neighbour counts of the 5 self searches from rdm_radius_search (count
    mode: no table is written) and the histogram from rdm_neighbor_histogram; one small readback per sample for the
    early-exit test of :213-214."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (search_radius / voxel_size + 1) ** 3))
    if not torch.cuda.is_available():
        raise RuntimeError("rdmnet_b200 needs a CUDA device: there is no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device())
    hists = torch.zeros((num_stages, hist_n), dtype=torch.int32, device=dev)
    for i in range(len(dataset)):
        dd = collate_fn([dataset[i]], num_stages, voxel_size, search_radius, [hist_n] * num_stages, precompute_data=False)
        points, lengths = dd["points"].to(dev).contiguous(), dd["lengths"].to(dev)
        voxel, radius = voxel_size, search_radius
        for s in range(num_stages):
            if s > 0:
                points, lengths = ops.grid_subsample(points, lengths, voxel_size=voxel)
            _, _, counts = ops.radius_search_raw(points, points, lengths, lengths, radius, 0, index_dtype=torch.int32,
                                                 counts=True)
            ops.neighbor_histogram(counts, hist_n, hists[s])
            voxel *= 2
            radius *= 2
        if int(hists.sum(dim=1).min().item()) > sample_threshold:
            break
    neighbor_hists = hists.cpu().numpy()
    cum_sum = np.cumsum(neighbor_hists.T, axis=0)
    return np.sum(cum_sum < (keep_ratio * cum_sum[hist_n - 1, :]), axis=0)


def reset_seed_worker_init_fn(worker_id):
    """geotransformer/utils/torch.py:40-45."""
    import random
    seed = torch.initial_seed() % (2 ** 32)
    np.random.seed(seed)
    random.seed(seed)


class GpuPyramidLoader:
    """Iterable over a torch DataLoader whose workers only stack points; every batch gets its pyramid built on the GPU
    here, in the consuming process. Everything else (len, dataset, sampler, ...) is the wrapped loader's."""

    def __init__(self, loader, num_stages, voxel_size, search_radius, neighbor_limits, precompute_data=True):
        self.loader = loader
        self.args = (num_stages, voxel_size, search_radius, [int(x) for x in neighbor_limits])
        self.precompute_data = precompute_data

    def __len__(self):
        return len(self.loader)

    def __getattr__(self, name):
        return getattr(self.loader, name)

    def __iter__(self):
        for dd in self.loader:
            if self.precompute_data:
                points, lengths = dd.pop("points"), dd.pop("lengths")
                dd.update(precompute_data_stack_mode(points, lengths, *self.args))
            yield dd


def build_dataloader_stack_mode(dataset, collate_fn, num_stages, voxel_size, search_radius, neighbor_limits, batch_size=1,
                                num_workers=1, shuffle=False, drop_last=False, distributed=False, precompute_data=True):
    """data.py:223-253 (+ build_dataloader, geotransformer/utils/torch.py:48-77)."""
    sampler = torch.utils.data.DistributedSampler(dataset) if distributed else None
    loader = torch.utils.data.DataLoader(
        dataset, batch_size=batch_size, num_workers=num_workers, shuffle=False if distributed else shuffle, sampler=sampler,
        collate_fn=partial(collate_fn, num_stages=num_stages, voxel_size=voxel_size, search_radius=search_radius,
                           neighbor_limits=neighbor_limits, precompute_data=False),
        worker_init_fn=reset_seed_worker_init_fn, pin_memory=os.environ.get("RDM_LOADER_PIN", "1") == "1", drop_last=drop_last)
    return GpuPyramidLoader(loader, num_stages, voxel_size, search_radius, neighbor_limits, precompute_data)
