"""Host-buffer entry point of the hot path: what experiments/infer.py does per pair between the DataLoader and
``after_test_step`` (geotransformer/engine/single_tester.py:104-117: to_cuda -> model(data_dict) -> release_cuda),
with the stacked raw points as the only input. Host tensors are staged through pinned memory; the voxel pyramid is
built on the GPU inside the model (rdmnet_b200.model.RDMNet.build_pyramid)."""
import numpy as np
import torch

RESULT_KEYS = ("estimated_transform", "ref_corr_points", "src_corr_points", "corr_scores")


class PairRegistrar:
    """Reusable pinned staging buffers + one model. ``register(ref_points, src_points)`` takes host arrays (N,3) f32
    and returns host numpy results; every call does one H2D of the points and one D2H of the results."""

    def __init__(self, model, max_points=1 << 18, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("rdmnet_b200 needs a CUDA device: there is no CPU path")
        self.model = model
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.h_points = torch.empty((max_points, 3), dtype=torch.float32).pin_memory()
        self.h_lengths = torch.empty(2, dtype=torch.int64).pin_memory()
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def register(self, ref_points, src_points, keys=RESULT_KEYS):
        nr, ns = ref_points.shape[0], src_points.shape[0]
        if nr + ns > self.h_points.shape[0]:
            raise RuntimeError("PairRegistrar: pair larger than the staging buffer")
        self.h_points[:nr] = torch.as_tensor(ref_points)
        self.h_points[nr:nr + ns] = torch.as_tensor(src_points)
        self.h_lengths[0], self.h_lengths[1] = nr, ns
        pts = self.h_points[:nr + ns].to(self.device, non_blocking=True)
        lens = self.h_lengths.to(self.device, non_blocking=True)
        self.h2d_bytes = pts.numel() * 4 + lens.numel() * 8
        out = self.model({"points": pts, "lengths": lens})
        res = {}
        for k in keys:  # D2H (synchronises); the pose already came back through pinned memory inside rdm_match_forward
            if k == "estimated_transform" and "estimated_transform_host" in out:
                res[k] = out["estimated_transform_host"].numpy()
            else:
                res[k] = out[k].cpu().numpy()
        self.d2h_bytes = int(sum(v.nbytes for v in res.values()))
        return res


class PairStreamRegistrar(PairRegistrar):
    """Throughput form of PairRegistrar: ``register_stream(pairs)`` takes an iterable of (ref_points, src_points) host
    arrays and yields one result dict per pair, in order. Pair i+1 is staged (pinned copy + H2D) and its voxel pyramid
    built on a side stream while pair i is in the network (rdmnet_b200.model.PairPipeline) - the role the reference gives
    to its DataLoader workers. Same results as register() pair by pair."""

    def __init__(self, model, max_points=1 << 18, device=None):
        super().__init__(model, max_points, device)
        from .model import PairPipeline
        self.pipe = PairPipeline(model, self.device)
        self.h_pts = [self.h_points, torch.empty_like(self.h_points).pin_memory()]
        self.h_len = [self.h_lengths, torch.empty_like(self.h_lengths).pin_memory()]

    def _stager(self, ref_points, src_points, slot):
        def stage():  # runs under the pipeline's side stream
            nr, ns = ref_points.shape[0], src_points.shape[0]
            if nr + ns > self.h_pts[slot].shape[0]:
                raise RuntimeError("PairStreamRegistrar: pair larger than the staging buffer")
            hp, hl = self.h_pts[slot], self.h_len[slot]
            hp[:nr] = torch.as_tensor(ref_points)
            hp[nr:nr + ns] = torch.as_tensor(src_points)
            hl[0], hl[1] = nr, ns
            pts = hp[:nr + ns].to(self.device, non_blocking=True)
            lens = hl.to(self.device, non_blocking=True)
            self.h2d_bytes = pts.numel() * 4 + lens.numel() * 8
            return pts, lens
        return stage

    def register_stream(self, pairs, keys=RESULT_KEYS):
        items = (self._stager(r, s, i & 1) for i, (r, s) in enumerate(pairs))
        for out in self.pipe.run(items):
            res = {}
            for k in keys:
                if k == "estimated_transform" and "estimated_transform_host" in out:
                    res[k] = out["estimated_transform_host"].numpy()
                else:
                    res[k] = out[k].cpu().numpy()
            self.d2h_bytes = int(sum(v.nbytes for v in res.values()))
            yield res


def register_pair(model, ref_points, src_points):
    return PairRegistrar(model, max_points=ref_points.shape[0] + src_points.shape[0]).register(
        np.ascontiguousarray(ref_points, np.float32), np.ascontiguousarray(src_points, np.float32))
