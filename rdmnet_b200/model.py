"""RDMNet forward orchestration on the GPU: host-side mirror of experiments/backbone.py (Encoder/Decoder),
experiments/model_infer.py (RDMNet, create_model), experiments/config.py (the constants the model reads) and
geotransformer/utils/data.py:13-77 (precompute_data_stack_mode), with every operator executed by librdm_sm100.so.

Differences from the reference that are part of the design (none changes results):
  * the voxel pyramid (4x grid_subsample + radius searches) is built on the GPU inside forward() when the
    data_dict does not carry 'neighbors' (the reference builds it in CPU DataLoader workers and ships ~75 MB of
    int64 tables over PCIe per pair); upsampling[0], which the reference computes and never reads
    (experiments/backbone.py:144), is skipped on that path;
  * NMS, partition, Sinkhorn and the pose solver run without host round trips; the only host syncs left are the
    data-dependent output shapes (pyramid lengths, NMS survivors, number of correspondences).
"""
import ctypes
import os
import types

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .modules import _Module, cache_key
from .modules import (ConvBlock, LastUnaryBlock, LearnableLogOptimalTransport, LocalGlobalRegistration, NMS,
                      ResidualBlock, SuperPointMatching, ThDRoFormer, UnaryBlock, Vote_layer)

DEFAULT_NEIGHBOR_LIMITS = [65, 63, 69, 70, 81]  # calibrate_neighbors_stack_mode on the bundled pairs


class _NS(types.SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)


def make_cfg():
    """The subset of experiments/config.py:10-188 that the model reads (same attribute paths)."""
    c = _NS()
    c.seed = 7351
    c.backbone = _NS(num_stages=5, init_voxel_size=0.3, kernel_size=15, base_radius=4.25, base_sigma=2.0,
                     init_radius=4.25 * 0.3, init_sigma=2.0 * 0.3, group_norm=32, input_dim=1, init_dim=64, output_dim=256)
    c.model = _NS(ground_truth_matching_radius=0.6, num_points_in_patch=128, num_sinkhorn_iterations=100,
                  ground_truth_corres_radius=2.4, n2p_score_threshold=0.1, p2p_score_threshold=0.1)
    c.coarse_matching = _NS(num_targets=128, overlap_threshold=0.1, num_correspondences=256, dual_normalization=True)
    c.thdroformer = _NS(input_dim=2048, hidden_dim=128, output_dim=256, num_heads=4, num_layers=4, input_dim2=256,
                        num_layers2=4, k2=None)
    c.Vote = _NS(model_use_vote=True, inference_use_vote=True, MAX_TRANSLATE_RANGE=[3.0, 3.0, 3.0], MLPS=[512, 256],
                 NMS_radius=2.4, n2n_overlap_threshold=1.2, n2p_overlap_threshold=0.6, p2p_overlap_threshold=0.6)
    c.fine_matching = _NS(acceptance_radius=0.6, mutual=False, topk=1, confidence_threshold=0, use_dustbin=True,
                          use_global_score=False, correspondence_threshold=3, correspondence_limit=None,
                          num_refinement_steps=5)
    c.test = _NS(vis=False)
    c.neighbor_limits = list(DEFAULT_NEIGHBOR_LIMITS)
    return c


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits,
                               index_dtype=torch.int64, skip_unused=False):
    """geotransformer/utils/data.py:13-77 on the GPU. Returns the same dict (+ 'lengths_host')."""
    points_list, lengths_list, host = [], [], []
    for i in range(num_stages):
        if i > 0:
            points, lengths = ops.grid_subsample(points, lengths, voxel_size=voxel_size)
        points_list.append(points)
        lengths_list.append(lengths)
        voxel_size *= 2
    neighbors, subsampling, upsampling, widths = [], [], [], []

    def search(q, s, ql, sl, r, lim):
        out, maxc = ops.radius_search_raw(q, s, ql, sl, r, lim, index_dtype=index_dtype)
        widths.append((out, maxc, lim))
        return out

    for i in range(num_stages):
        P, Ln = points_list[i], lengths_list[i]
        neighbors.append(search(P, P, Ln, Ln, radius, neighbor_limits[i]))
        if i < num_stages - 1:
            S, Sl = points_list[i + 1], lengths_list[i + 1]
            subsampling.append(search(S, P, Sl, Ln, radius, neighbor_limits[i]))
            if skip_unused and i == 0:
                upsampling.append(None)
            else:
                upsampling.append(search(P, S, Ln, Sl, radius * 2, neighbor_limits[i + 1]))
        radius *= 2
    # one D2H for all data-dependent sizes: per-stage lengths and the max neighbour counts (row widths)
    meta = torch.cat([torch.stack(lengths_list).reshape(-1).to(torch.int64)] + [w[1].to(torch.int64) for w in widths]).tolist()
    nb = lengths_list[0].shape[0]
    host = [meta[i * nb:(i + 1) * nb] for i in range(num_stages)]
    maxcs = meta[num_stages * nb:]

    def narrow(t):
        if t is None:
            return None
        for (out, _, lim), mc in zip(widths, maxcs):
            if out is t:
                return t if mc >= lim else t[:, :mc].contiguous()  # radius_search.py:25-26 width semantics
        return t

    return {
        "points": points_list, "lengths": lengths_list, "lengths_host": host,
        "neighbors": [narrow(t) for t in neighbors], "subsampling": [narrow(t) for t in subsampling],
        "upsampling": [narrow(t) for t in upsampling],
    }


class GpuPyramid:
    """Result of rdm_build_pyramid: one device buffer + the host descriptor the C runners consume. Tensors are views
    into the buffer, created on demand (`points(s)`, `lengths(s)`, `table(kind, s)`, `as_data_dict()`)."""

    def __init__(self, buf, desc, lengths_host, d_lengths, points0, batch):
        self.buf, self.desc, self.lengths_host, self._d_lengths, self._points0, self.batch = (
            buf, desc, lengths_host, d_lengths, points0, batch)
        self._base = buf.data_ptr()

    def _view(self, ptr, nbytes, dtype):
        off = ptr - self._base
        return self.buf[off:off + nbytes].view(dtype)

    def points(self, s):
        if s == 0:
            return self._points0
        return self._view(self.desc.points[s], self.desc.n[s] * 12, torch.float32).view(-1, 3)

    def lengths(self, s):
        return self._view(self._d_lengths[s], 8 * self.batch, torch.int64)

    def table(self, kind, s):
        d = self.desc
        ptr, rows, w = {"neighbors": (d.neighbors[s], d.n[s], d.nb_width[s]),
                        "subsampling": (d.subsampling[s], d.n[s + 1], d.sub_width[s]),
                        "upsampling": (d.upsampling[s], d.n[s], d.up_width[s]),
                        "order": (d.order[s], d.n[s], 1)}[kind]
        if not ptr:
            return None
        t = self._view(ptr, rows * w * 4, torch.int32)
        return t if kind == "order" else t.view(rows, w)

    def as_data_dict(self):
        S = self.desc.num_stages
        return {"points": [self.points(s) for s in range(S)], "lengths": [self.lengths(s) for s in range(S)],
                "lengths_host": self.lengths_host,
                "neighbors": [self.table("neighbors", s) for s in range(S)],
                "subsampling": [self.table("subsampling", s) for s in range(S - 1)],
                "upsampling": [self.table("upsampling", s) for s in range(S - 1)]}


class PyramidJob:
    """One pyramid in flight through rdm_build_pyramid_begin / _finish (the two-phase form of rdm_build_pyramid)."""

    def __init__(self):
        self.handle = L.lib().rdm_pyramid_job_create()
        if not self.handle:
            raise RuntimeError("rdm_pyramid_job_create failed: " + L.lib().rdm_last_error().decode())
        self._live = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                L.lib().rdm_pyramid_job_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def begin(self, points, lengths, num_stages, voxel_size, radius, neighbor_limits, skip_unused=True, up_nearest_only=True):
        """Queues the chained subsamplings + the size readback on the current stream; returns immediately."""
        points = points.contiguous()
        n0, batch = points.shape[0], lengths.shape[0]
        cfg = L.PyramidCfg()
        cfg.num_stages, cfg.batch, cfg.first_voxel, cfg.first_radius = num_stages, batch, voxel_size * 2, radius
        for i, v in enumerate(neighbor_limits):
            cfg.limits[i] = int(v)
        cfg.skip_up0, cfg.up_nearest_only = int(skip_unused), int(up_nearest_only)
        lib = L.lib()
        ob, wb = lib.rdm_build_pyramid_bytes(n0, ctypes.byref(cfg)), lib.rdm_build_pyramid_workspace(n0, ctypes.byref(cfg))
        buf = torch.empty(int(ob), dtype=torch.uint8, device=points.device)
        ws = torch.empty(int(wb), dtype=torch.uint8, device=points.device)
        with torch.cuda.device(points.device):
            L.call("rdm_build_pyramid_begin", self.handle, L.ptr(points), L.ptr(lengths), n0, ctypes.byref(cfg), L.ptr(buf), int(ob),
                   L.ptr(ws), int(wb), L.stream())
        self._live = (points, lengths, buf, ws, num_stages, batch)

    def finish(self):
        """Waits (host) for the stage sizes, queues every radius search on the current stream -> GpuPyramid."""
        points, lengths, buf, ws, num_stages, batch = self._live
        self._live = None
        desc = L.PyramidDesc()
        h_len = (ctypes.c_int64 * (num_stages * batch))()
        h_dl = (ctypes.c_void_p * 8)()
        with torch.cuda.device(points.device):
            L.call("rdm_build_pyramid_finish", self.handle, ctypes.byref(desc), ctypes.cast(h_len, ctypes.c_void_p),
                   ctypes.cast(h_dl, ctypes.c_void_p), L.stream())
        host = [[int(h_len[s * batch + b]) for b in range(batch)] for s in range(num_stages)]
        gp = GpuPyramid(buf, desc, host, [int(h_dl[s]) for s in range(num_stages)], points, batch)
        gp._keep = (lengths, ws)
        return gp


_DEFAULT_JOB = None


def build_pyramid_gpu(points, lengths, num_stages, voxel_size, radius, neighbor_limits, skip_unused=True,
                      up_nearest_only=True):
    """geotransformer/utils/data.py:13-77 through rdm_build_pyramid: one synchronisation, int32 tables of fixed
    width = the neighbour limit, plus a cell-sorted query order per stage for the KPConv gather."""
    global _DEFAULT_JOB
    if _DEFAULT_JOB is None:
        _DEFAULT_JOB = PyramidJob()
    _DEFAULT_JOB.begin(points, lengths, num_stages, voxel_size, radius, neighbor_limits, skip_unused, up_nearest_only)
    return _DEFAULT_JOB.finish()


_PIPE_STREAMS = {}


class PairPipeline:
    """Software pipeline over a stream of scan pairs, single host thread, results in input order and identical to
    model(data_dict) pair by pair. Default: two pairs in flight - while pair i runs through the network on the main stream, the
    voxel pyramid of pair i+1 is built on a side stream (what the reference's DataLoader workers do on CPU cores,
    geotransformer/utils/data.py:223-253 with num_workers=8).

    RDM_PIPE_OVERLAP=1 (opt-in) keeps three in flight:
        side stream      pyramid of pair i+2
        net stream A/B   encoder -> transformer 1 -> decoder of pair i+1   (forward_head, asynchronous)
        net stream B/A   vote / NMS / transformer 2 / matching / pose of pair i   (forward_tail, two host syncs)
    with the matching tail of pair i and the radius searches of pair i+2 held behind the event that rdm_backbone_forward records
    after the encoder of pair i+1, so that the bandwidth-bound KPConv gathers keep the machine to themselves. Measured
    (profiles/README.md, r2u): median step 4.10-4.16 ms against 4.18-4.22 ms, end to end 232-240 against 234-236 pairs/s - within
    noise, because the cycle stays encoder(i+1) -> tail(i) -> [host] -> encoder(i+2): the host thread is blocked inside the tail's two
    stream synchronisations and cannot queue the next head earlier. It stays off until rdm_match_forward loses its host syncs."""

    def __init__(self, model, device=None):
        self.model = model
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # high priority: the side stream's kernels are few and small, and the host waits on the first of them (the
        # subsampling chain) before it can queue the rest of the pair in flight - they must not starve behind the
        # network's large grids
        # the streams are shared by every pipeline of a device: the caching allocator keeps one block pool per stream, and a
        # fresh stream would start with an empty pool (cudaMalloc inside somebody's timed region)
        res = _PIPE_STREAMS.get(self.device)
        if res is None:
            side = torch.cuda.Stream(self.device, priority=int(os.environ.get("RDM_PIPE_PRIORITY", "-1")))
            nets = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            enc_done = [torch.cuda.Event(), torch.cuda.Event()]
            for e, st in zip(enc_done, nets):
                e.record(st)  # instantiates the CUDA event (torch creates it lazily)
            res = _PIPE_STREAMS[self.device] = (side, nets, enc_done)
        self.side, self.nets, self.enc_done = res
        self.jobs = [PyramidJob(), PyramidJob()]
        self.match_jobs = None
        # the next pair's radius searches are held back until the pair in the network has left its encoder, so that they share
        # the SMs with the (latency-bound) rest instead of with the KPConv gathers: measured +10 % gather
        # bandwidth and +2 % pairs/s inside the pipelined bench region (profiles/README.md, r2a). RDM_PIPE_DEFER_SEARCH=0
        # restores the eager order.
        self.defer_searches = os.environ.get("RDM_PIPE_DEFER_SEARCH", "1") == "1"
        self.overlap = os.environ.get("RDM_PIPE_OVERLAP", "1") == "1"
        # RDM_PIPE_GATHER_ALONE=0 lets the matching tail of pair i start at once, on top of the encoder of pair i+1 (A/B knob)
        self.gather_alone = os.environ.get("RDM_PIPE_GATHER_ALONE", "1") == "1"
        # RDM_PIPE_TAIL_EARLY=1: the small-grid front of the tail (transformer 2, partitions, coarse matching) may overlap the next
        # pair's encoder; only the patch stage (patch scores, Sinkhorn, pose) waits for it
        self.tail_early = os.environ.get("RDM_PIPE_TAIL_EARLY", "0") == "1"

    def _begin(self, item, slot):
        points, lengths = item() if callable(item) else item  # a callable may stage host data (runs on the side stream)
        b, cfg = self.model.cfg.backbone, self.model.cfg
        self.jobs[slot].begin(points, lengths, b.num_stages, b.init_voxel_size, b.init_radius, cfg.neighbor_limits)

    def _finish(self, slot):
        gp = self.jobs[slot].finish()
        ev = torch.cuda.Event()
        ev.record(self.side)
        return gp, ev

    def run(self, items, before_step=None, after_step=None):
        """items: iterable of (points (N,3) f32 cuda, lengths (2,) i64 cuda) or of callables returning such a tuple
        (called under the side stream). Yields the model's output dict per pair. before_step(i), if given, runs on pair i's
        network stream right before pair i enters the network (bench.py: L2 flush); after_step(i) right after its last
        kernel (bench.py: timing event)."""
        if self.overlap:
            yield from self._run_overlapped(items, before_step, after_step)
            return
        main = torch.cuda.current_stream(self.device)
        it = iter(items)
        first = next(it, None)
        if first is None:
            return
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self._begin(first, 0)
            cur = self._finish(0)
        i = 0
        while cur is not None:
            gp, ready = cur
            nxt_item = next(it, None)
            if before_step is not None:
                before_step(i)
            main.wait_event(ready)
            # the pyramid buffer, its workspace and the staged points were allocated under the side stream but are read by
            # main-stream kernels (and by the caller through views in the output dict): tell the caching allocator, or
            # the blocks could be handed to the next side-stream pyramid while the main stream still reads them
            for t in (gp.buf, gp._points0) + tuple(gp._keep):
                t.record_stream(main)
            state = self.model.forward_head(None, gp=gp)  # asynchronous launches on the main stream
            cur = None
            if nxt_item is not None:
                backbone_done = None
                if self.defer_searches:
                    backbone_done = torch.cuda.Event()
                    backbone_done.record(main)
                with torch.cuda.stream(self.side):
                    self._begin(nxt_item, (i + 1) & 1)
                    if backbone_done is not None:
                        self.side.wait_event(backbone_done)  # ordered before the searches that _finish queues
                    cur = self._finish((i + 1) & 1)  # host waits for the (short) subsampling chain only
            out = self.model.forward_tail(state)
            if after_step is not None:
                after_step(i)
            yield out
            i += 1

    # ---- three pairs in flight
    def _head(self, i, pyr, before_step):
        """Queues encoder + transformer 1 + decoder of pair i on net stream i & 1 (asynchronous)."""
        gp, ready = pyr
        s = self.nets[i & 1]
        with torch.cuda.stream(s):
            if before_step is not None:
                before_step(i)
            s.wait_event(ready)
            for t in (gp.buf, gp._points0) + tuple(gp._keep):  # allocated under the side stream, read here (see run())
                t.record_stream(s)
            L.call("rdm_backbone_set_encoder_event", self.enc_done[i & 1].cuda_event)
            try:
                state = self.model.forward_head(None, gp=gp)
            finally:
                L.call("rdm_backbone_set_encoder_event", None)
        return state

    def _run_overlapped(self, items, before_step, after_step):
        """Host order per iteration i (nothing here blocks on more than it needs):
            begin(i)      phase 1 of the tail of pair i, queued right behind its head on net stream i & 1
            head(i+1)     on the other net stream (after launching the subsampling chain of pair i+2 on the side stream)
            pyramid(i+2)  finish: its radius searches wait for the encoder of pair i+1
            continue(i)   host waits for the NMS counts of pair i; the rest of its tail waits for the encoder of pair i+1 (stream)
            finish(i-1)   host waits for the result of pair i-1 -> yield
        GPU order: net stream i & 1 carries head(i), tail(i), head(i+2), ... so encoder(i+2) starts the moment tail(i) ends, with no
        host round trip in between; the cycle is encoder + max(transformer 1 + decoder, tail)."""
        caller = torch.cuda.current_stream(self.device)
        if self.match_jobs is None:
            self.match_jobs = [L.lib().rdm_match_job_create() for _ in range(4)]
            if any(j is None for j in self.match_jobs):
                raise RuntimeError("rdm_match_job_create failed: " + L.lib().rdm_last_error().decode())
        it = iter(items)
        first = next(it, None)
        if first is None:
            return
        for s in [self.side] + self.nets:
            s.wait_stream(caller)
        with torch.cuda.stream(self.side):
            self._begin(first, 0)
            pyr = self._finish(0)
        nxt = next(it, None)
        if nxt is not None:
            with torch.cuda.stream(self.side):
                self._begin(nxt, 1)
        state = self._head(0, pyr, before_step)
        pyr_next = None
        if nxt is not None:
            with torch.cuda.stream(self.side):
                if self.defer_searches:
                    self.side.wait_event(self.enc_done[0])
                pyr_next = self._finish(1)
        def emit(p):  # p = (index, ctx or None, finished output or None) of the pair whose tail is in flight
            idx, ctx, sync_out = p
            out = self.model._match_finish(ctx) if ctx is not None else sync_out
            for v in out.values():  # produced on a net stream, consumed by the caller on its own stream
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(caller)
            return out

        try:
            yield from self._overlapped_loop(it, state, pyr_next, before_step, after_step, emit)
        finally:  # a consumer that stops early (or an error) must not leave jobs "in flight" for the next run
            for j in self.match_jobs:
                L.call("rdm_match_job_reset", j)

    def _overlapped_loop(self, it, state, pyr_next, before_step, after_step, emit):
        pending, i = None, 0
        while state is not None:
            s = self.nets[i & 1]
            with torch.cuda.stream(s):
                ctx = self.model.tail_begin(state, self.match_jobs[i & 3])
            state_next = None
            if pyr_next is not None:
                nxt = next(it, None)
                if nxt is not None:
                    with torch.cuda.stream(self.side):
                        self._begin(nxt, i & 1)  # job slot of pair i, whose pyramid was handed over long ago
                state_next = self._head(i + 1, pyr_next, before_step)
                pyr_next = None
                if nxt is not None:
                    with torch.cuda.stream(self.side):
                        if self.defer_searches:
                            self.side.wait_event(self.enc_done[(i + 1) & 1])
                        pyr_next = self._finish(i & 1)
            sync_out = None
            with torch.cuda.stream(s):
                early = self.tail_early and ctx is not None and state_next is not None and self.gather_alone
                if state_next is not None and self.gather_alone and not early:
                    s.wait_event(self.enc_done[(i + 1) & 1])  # the gathers of pair i+1 run alone
                if ctx is not None:
                    if early:  # only the patch stage waits for the encoder of pair i+1 (inside rdm_match_continue)
                        L.call("rdm_match_set_patch_wait_event", self.enc_done[(i + 1) & 1].cuda_event)
                    try:
                        self.model._match_continue(ctx)
                    finally:
                        if early:
                            L.call("rdm_match_set_patch_wait_event", None)
                else:  # configurations without the runner tail (no vote branch / stepwise): synchronous
                    sync_out = self.model.forward_tail(state)
                if after_step is not None:
                    after_step(i)  # right behind the pair's last kernel, before the head of pair i+2 lands on this stream
            if pending is not None:
                yield emit(pending)
            pending = (i, ctx, sync_out)
            state, i = state_next, i + 1
        if pending is not None:
            yield emit(pending)


def _state_key(module):
    return cache_key(module)


def _p(t):
    return None if t is None else t.data_ptr()


def _presplit_reset(owner):
    """Unregisters the pre-split weights `owner` registered earlier (their buffers are about to be dropped)."""
    for key in getattr(owner, "_ps_keys", []):
        L.lib().rdm_presplit_register(key, None)
    owner._ps_keys = []


def _presplit(owner, w, keep):
    """Registers the tf32 hi / lo split of a constant nn.Linear-layout weight (rows, ld) with the library, so that the
    runners' tcgen05 GEMMs stream it through two TMA descriptors instead of splitting the tile in every CTA (include/
    rdm_sm100.h: rdm_presplit_weight). Called only from the fingerprint-guarded descriptor builders."""
    if w is None or not w.is_cuda or w.dtype != torch.float32 or w.ndim != 2 or not w.is_contiguous():
        return
    if w.shape[1] % 4 != 0 or w.data_ptr() % 16 != 0 or w.shape[0] < 8 or w.shape[1] < 8:
        return  # not TMA-eligible: the GEMM dispatcher takes the SIMT kernel for it anyway
    split = torch.empty((2,) + tuple(w.shape), dtype=torch.float32, device=w.device)
    L.call("rdm_presplit_weight", w.data_ptr(), w.shape[0], w.shape[1], split.data_ptr(), L.stream())
    L.lib().rdm_presplit_register(w.data_ptr(), split.data_ptr())
    keep.append(split)
    owner._ps_keys.append(w.data_ptr())


def _unary_desc(u, keep=None, owner=None):
    """rdm_unary_desc of a UnaryBlock / LastUnaryBlock (NULL weight for nn.Identity / None). With `keep` (a list that
    outlives the descriptor) a weight whose in_features is not a multiple of 4 is passed as a zero-padded copy with a
    16-byte row stride, which qualifies the layer for the TMA-fed tensor-core GEMM."""
    if u is None or isinstance(u, nn.Identity):
        return L.UnaryDesc()
    norm = getattr(u, "norm", None)
    w, ldw = u.mlp.weight, 0
    if keep is not None and u.mlp.in_features % 4 != 0:
        ldw = (u.mlp.in_features + 3) // 4 * 4
        wp = torch.zeros((u.mlp.out_features, ldw), dtype=torch.float32, device=w.device)
        wp[:, :u.mlp.in_features] = w.detach()
        keep.append(wp)
        w = wp
    if owner is not None and keep is not None:
        _presplit(owner, w.detach(), keep)
    return L.UnaryDesc(_p(w), _p(u.mlp.bias), _p(norm.norm.weight) if norm is not None else None,
                       _p(norm.norm.bias) if norm is not None else None, u.mlp.in_features, u.mlp.out_features, ldw)


def pyramid_desc(data_dict, lengths_host):
    """rdm_pyramid_desc over a data_dict (ours or the reference's). Returns (desc, keepalive tensors)."""
    d = L.PyramidDesc()
    pts, nb, sub, up = data_dict["points"], data_dict["neighbors"], data_dict["subsampling"], data_dict["upsampling"]
    keep, dtypes = [], set()

    def table(t):
        t = t.contiguous()
        keep.append(t)
        dtypes.add(t.dtype)
        return t

    d.num_stages = len(pts)
    for s in range(len(pts)):
        p = pts[s].contiguous()
        keep.append(p)
        d.points[s], d.n[s] = p.data_ptr(), int(sum(lengths_host[s]))
        t = table(nb[s])
        d.neighbors[s], d.nb_width[s] = t.data_ptr(), t.shape[1]
        if s < len(pts) - 1:
            t = table(sub[s])
            d.subsampling[s], d.sub_width[s] = t.data_ptr(), t.shape[1]
            if up[s] is not None:
                t = table(up[s])
                d.upsampling[s], d.up_width[s] = t.data_ptr(), t.shape[1]
    if len(dtypes) != 1 or next(iter(dtypes)) not in (torch.int32, torch.int64):
        raise RuntimeError("neighbour tables must all be int64 or all be int32")
    d.index_bytes = 8 if next(iter(dtypes)) == torch.int64 else 4
    return d, keep


class Encoder(_Module):
    """experiments/backbone.py:7-107."""

    def __init__(self, input_dim, init_dim, kernel_size, init_radius, init_sigma, group_norm):
        super().__init__()
        d, r, s, g, k = init_dim, init_radius, init_sigma, group_norm, kernel_size
        self.encoder1_1 = ConvBlock(input_dim, d, k, r, s, g)
        self.encoder1_2 = ResidualBlock(d, d * 2, k, r, s, g)
        self.encoder2_1 = ResidualBlock(d * 2, d * 2, k, r, s, g, strided=True)
        self.encoder2_2 = ResidualBlock(d * 2, d * 4, k, r * 2, s * 2, g)
        self.encoder2_3 = ResidualBlock(d * 4, d * 4, k, r * 2, s * 2, g)
        self.encoder3_1 = ResidualBlock(d * 4, d * 4, k, r * 2, s * 2, g, strided=True)
        self.encoder3_2 = ResidualBlock(d * 4, d * 8, k, r * 4, s * 4, g)
        self.encoder3_3 = ResidualBlock(d * 8, d * 8, k, r * 4, s * 4, g)
        self.encoder4_1 = ResidualBlock(d * 8, d * 8, k, r * 4, s * 4, g, strided=True)
        self.encoder4_2 = ResidualBlock(d * 8, d * 16, k, r * 8, s * 8, g)
        self.encoder4_3 = ResidualBlock(d * 16, d * 16, k, r * 8, s * 8, g)
        self.encoder5_1 = ResidualBlock(d * 16, d * 16, k, r * 8, s * 8, g, strided=True)
        self.encoder5_2 = ResidualBlock(d * 16, d * 32, k, r * 16, s * 16, g)
        self.encoder5_3 = ResidualBlock(d * 32, d * 32, k, r * 16, s * 16, g)

    def _block_descs(self):
        """rdm_block_desc array of the 14 blocks (cached until a parameter changes)."""
        key = _state_key(self)
        if getattr(self, "_desc_key", None) != key:
            names = ["encoder1_1", "encoder1_2"] + [f"encoder{s}_{j}" for s in range(2, 6) for j in (1, 2, 3)]
            arr = (L.BlockDesc * len(names))()
            keep = []
            _presplit_reset(self)
            for i, name in enumerate(names):
                b, stage = getattr(self, name), int(name[7]) - 1
                kp = b.KPConv
                hk = ops._host_copy(kp.kernel_points)
                keep.append(hk)
                d = arr[i]
                if isinstance(b, ConvBlock):
                    d.norm_conv_w, d.norm_conv_b = _p(b.norm.norm.weight), _p(b.norm.norm.bias)
                    d.c_in, d.c_out, d.strided = b.in_channels, b.out_channels, 0
                else:
                    d.unary1, d.unary2, d.shortcut = (_unary_desc(b.unary1, keep, self), _unary_desc(b.unary2, keep, self),
                                                      _unary_desc(b.unary_shortcut, keep, self))
                    d.norm_conv_w, d.norm_conv_b = _p(b.norm_conv.norm.weight), _p(b.norm_conv.norm.bias)
                    d.c_in, d.c_out, d.strided = b.in_channels, b.out_channels, 1 if b.strided else 0
                wt = kp.weights.detach().reshape(-1, kp.out_channels).t().contiguous()  # [C_out, 15*C_in]
                keep.append(wt)
                _presplit(self, wt, keep)
                d.kpconv_w, d.kpconv_wt, d.kpconv_b = _p(kp.weights), _p(wt), _p(kp.bias)
                d.kernel_points, d.h_kernel_points = _p(kp.kernel_points), hk.data_ptr()
                d.c_mid_in, d.c_mid_out, d.stage, d.sigma = kp.in_channels, kp.out_channels, stage, float(kp.sigma)
            self._descs, self._desc_keep, self._desc_key = arr, keep, key
        return self._descs

    def forward(self, feats, data_dict, pyr=None):
        """One host call for the 14 blocks (rdm_encoder_forward). Returns the 5 stage outputs."""
        blocks = self._block_descs()
        if pyr is None:
            lh = data_dict.get("lengths_host") or [l.tolist() for l in data_dict["lengths"]]
            pyr = pyramid_desc(data_dict, lh)
        desc, _keep = pyr
        dev = feats.device
        last = {}
        for d in blocks:
            last[d.stage] = d.c_out
        outs = [torch.empty((desc.n[s], last[s]), dtype=torch.float32, device=dev) for s in range(desc.num_stages)]
        out_ptrs = (ctypes.c_void_p * 8)(*[o.data_ptr() for o in outs])
        lib = L.lib()
        wsb = lib.rdm_encoder_workspace(ctypes.cast(blocks, ctypes.c_void_p), len(blocks), ctypes.byref(desc), 32)
        ws = torch.empty(max(int(wsb), 1), dtype=torch.uint8, device=dev)
        feats = feats.contiguous()
        L.call("rdm_encoder_forward", ctypes.cast(blocks, ctypes.c_void_p), len(blocks), ctypes.byref(desc),
               self.encoder1_1.norm.num_groups, L.ptr(feats), ctypes.cast(out_ptrs, ctypes.c_void_p), L.ptr(ws), int(wsb),
               L.stream())
        return outs

    def forward_modules(self, feats, data_dict):
        """The same wiring through the per-block Python modules (used by the module-level parity tests)."""
        P, NB, SUB = data_dict["points"], data_dict["neighbors"], data_dict["subsampling"]
        out = []
        x = self.encoder1_1(feats, P[0], P[0], NB[0])
        x = self.encoder1_2(x, P[0], P[0], NB[0])
        out.append(x)
        for s in range(1, 5):
            x = getattr(self, f"encoder{s + 1}_1")(x, P[s], P[s - 1], SUB[s - 1])
            x = getattr(self, f"encoder{s + 1}_2")(x, P[s], P[s], NB[s])
            x = getattr(self, f"encoder{s + 1}_3")(x, P[s], P[s], NB[s])
            out.append(x)
        return out


class Decoder(_Module):
    """experiments/backbone.py:110-151."""

    def __init__(self, output_dim, init_dim, group_norm):
        super().__init__()
        self.decoder4 = UnaryBlock(init_dim * 20 + 1, init_dim * 16, group_norm)
        self.decoder3 = UnaryBlock(init_dim * 24, init_dim * 8, group_norm)
        self.decoder2 = LastUnaryBlock(init_dim * 12, output_dim + 1)

    def unary_descs(self):
        """rdm_unary_desc[3] = decoder4, decoder3, decoder2 (cached; see _Module)."""
        key = _state_key(self)
        if getattr(self, "_desc_key", None) != key:
            keep = []
            _presplit_reset(self)
            arr = (L.UnaryDesc * 3)(_unary_desc(self.decoder4, keep, self), _unary_desc(self.decoder3, keep, self),
                                    _unary_desc(self.decoder2, keep, self))
            self._descs, self._desc_keep, self._desc_key = arr, keep, key
        return self._descs

    def forward(self, feats, data_dict, pyr=None):
        """rdm_decoder_forward: three (nearest upsample || skip -> unary) levels in one host call. Returns [l2]."""
        self.unary_descs()
        if pyr is None:
            lh = data_dict.get("lengths_host") or [l.tolist() for l in data_dict["lengths"]]
            pyr = pyramid_desc(data_dict, lh)
        desc, _keep = pyr
        coarse = feats[4].contiguous()
        skips = [feats[3].contiguous(), feats[2].contiguous(), feats[1].contiguous()]
        skip_ptrs = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in skips])
        # rows padded to a multiple of 4 floats (257 -> 260): 16-byte rows for the GEMM epilogue and for the consumers of
        # the feature columns (rdm_patch_scores); the returned tensor is the [:, :out_channels] view
        c_out = self.decoder2.out_channels
        ld_out = (c_out + 3) // 4 * 4
        out_full = torch.empty((desc.n[1], ld_out), dtype=torch.float32, device=coarse.device)
        out = out_full[:, :c_out]
        lib = L.lib()
        wsb = lib.rdm_decoder_workspace(ctypes.cast(self._descs, ctypes.c_void_p), 3, ctypes.byref(desc), 4, 32)
        ws = torch.empty(max(int(wsb), 1), dtype=torch.uint8, device=coarse.device)
        L.call("rdm_decoder_forward", ctypes.cast(self._descs, ctypes.c_void_p), 3, ctypes.byref(desc), 4,
               self.decoder4.norm.num_groups, L.ptr(coarse), coarse.shape[1], ctypes.cast(skip_ptrs, ctypes.c_void_p),
               L.ptr(out_full), ld_out, L.ptr(ws), int(wsb), L.stream())
        return [out]

    def forward_modules(self, feats, data_dict):
        UP = data_dict["upsampling"]
        l4 = self.decoder4(ops.nearest_upsample_concat(feats[4], UP[3], feats[3]))
        l3 = self.decoder3(ops.nearest_upsample_concat(l4, UP[2], feats[2]))
        l2 = self.decoder2(ops.nearest_upsample_concat(l3, UP[1], feats[1]))
        return [l2, l3, l4]


class RDMNet(_Module):
    """experiments/model_infer.py:26-354 (inference forward). Same submodule names => same checkpoint keys."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.num_points_in_patch = cfg.model.num_points_in_patch
        b, t = cfg.backbone, cfg.thdroformer
        self.encoder = Encoder(b.input_dim, b.init_dim, b.kernel_size, b.init_radius, b.init_sigma, b.group_norm)
        self.decoder = Decoder(b.output_dim, b.init_dim, b.group_norm)
        self.transformer = ThDRoFormer(t.input_dim, t.output_dim, t.hidden_dim, t.num_heads, t.num_layers)
        self.use_vote = cfg.Vote.inference_use_vote and cfg.Vote.model_use_vote
        if cfg.Vote.model_use_vote:
            cfg.Vote.input_feats_dim = t.output_dim
            self.vote = Vote_layer(cfg.Vote, 1)
            self.nms = NMS(cfg.Vote, cfg.neighbor_limits)
            self.proj_n2n_score = nn.Linear(t.output_dim, 1)
            self.transformer2 = ThDRoFormer(t.input_dim2, t.output_dim, t.hidden_dim, t.num_heads, t.num_layers2, t.k2)
        self.proj_n2p_score = nn.Linear(t.output_dim, 1)
        self.coarse_matching = SuperPointMatching(cfg.coarse_matching.num_correspondences,
                                                  cfg.coarse_matching.dual_normalization, cfg.model.n2p_score_threshold)
        f = cfg.fine_matching
        self.fine_matching = LocalGlobalRegistration(
            f.topk, f.acceptance_radius, mutual=f.mutual, confidence_threshold=f.confidence_threshold,
            use_dustbin=f.use_dustbin, use_global_score=f.use_global_score,
            correspondence_threshold=f.correspondence_threshold, correspondence_limit=f.correspondence_limit,
            num_refinement_steps=f.num_refinement_steps)
        self.optimal_transport = LearnableLogOptimalTransport(cfg.model.num_sinkhorn_iterations)

    def _ones(self, n, device):
        """Input features of the KITTI configuration: ones (N,1) (kitti/dataset.py:188-189), cached by size."""
        c = getattr(self, "_ones_cache", None)
        if c is None or c.shape[0] < n or c.device != device:
            c = torch.ones((max(n, 1 << 16), 1), dtype=torch.float32, device=device)
            self._ones_cache = c
        return c[:n]

    def build_pyramid(self, points, lengths):
        b = self.cfg.backbone
        return precompute_data_stack_mode(points.contiguous(), lengths, b.num_stages, b.init_voxel_size, b.init_radius,
                                          self.cfg.neighbor_limits, index_dtype=torch.int32, skip_unused=True)

    def _backbone(self, desc, nc_ref, feats):
        """rdm_backbone_forward: encoder + transformer 1 + n2p head + decoder + p2p head, one asynchronous host call.
        Returns (tf (nc, 256), n2p (nc,), dec (nf, 257) view with row stride 260, p2p (nf,))."""
        key = self._fwd_key if getattr(self, "_fwd_key", None) is not None else cache_key(self)
        if getattr(self, "_bdesc_key", None) != key:
            blocks, dec, t1 = self.encoder._block_descs(), self.decoder.unary_descs(), self.transformer.runner_desc()
            d = L.BackboneDesc()
            d.h_blocks, d.num_blocks, d.groups = ctypes.addressof(blocks), len(blocks), self.encoder.encoder1_1.norm.num_groups
            d.h_transformer1 = ctypes.addressof(t1)
            d.n2p_w, d.n2p_b = _p(self.proj_n2p_score.weight), _p(self.proj_n2p_score.bias)
            d.h_dec, d.num_dec = ctypes.addressof(dec), 3
            self._bdesc, self._bdesc_keep, self._bdesc_key = d, (blocks, dec, t1), key
        d = self._bdesc
        dev = feats.device
        S = desc.num_stages
        nc, nf = desc.n[S - 1], desc.n[S - 1 - d.num_dec]
        c, c_out = self.transformer.out_proj.out_features, self.decoder.decoder2.out_channels
        ld = (c_out + 3) // 4 * 4
        tf = torch.empty((nc, c), dtype=torch.float32, device=dev)
        n2p = torch.empty(nc, dtype=torch.float32, device=dev)
        dec_full = torch.empty((nf, ld), dtype=torch.float32, device=dev)
        p2p = torch.empty(nf, dtype=torch.float32, device=dev)
        o = L.BackboneOut(tf.data_ptr(), n2p.data_ptr(), dec_full.data_ptr(), ld, p2p.data_ptr())
        lib = L.lib()
        wsb = int(lib.rdm_backbone_workspace(ctypes.byref(d), ctypes.byref(desc), nc_ref))
        if wsb == 0:
            raise RuntimeError("rdm_backbone_workspace failed: " + lib.rdm_last_error().decode())
        # the backbone's scratch (~330 MB of transient activations for a 30k-point pair) is kept per stream and only ever grows:
        # the next use on the same stream is ordered behind this one, and its size varies from pair to pair by a few per cent,
        # which made the caching allocator fall back to cudaMalloc (a 10-20 ms stall) every now and then
        key = ("backbone", torch.cuda.current_stream(dev).cuda_stream)
        cache = self.__dict__.setdefault("_ws_cache", {})
        ws = cache.get(key)
        if ws is None or ws.numel() < wsb:
            ws = cache[key] = torch.empty(int(wsb * 1.25) + (1 << 20), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.call("rdm_backbone_forward", ctypes.byref(d), ctypes.byref(desc), nc_ref, L.ptr(feats.contiguous()), ctypes.byref(o),
                   L.ptr(ws), wsb, L.stream())
        return tf, n2p, dec_full[:, :c_out], p2p

    def _match_desc(self):
        """rdm_match_desc over this model's vote / score / transformer2 / matching parameters (cached; see _Module)."""
        key = self._fwd_key if getattr(self, "_fwd_key", None) is not None else cache_key(self)
        if getattr(self, "_mdesc_key", None) != key:
            v, cfg = self.vote, self.cfg
            mods = list(v.mlp_modules)
            if len(mods) != 6 or v.max_offset_limit is None:
                raise RuntimeError("rdm_match_forward expects the RDMNet vote head: two (Linear, LayerNorm, ReLU) stages")
            d = L.MatchDesc()
            ps_keep = []
            _presplit_reset(self)
            for w_ in (mods[0].weight, mods[3].weight, v.ctr_reg.weight):
                _presplit(self, w_.detach(), ps_keep)
            d.v_w0, d.v_b0, d.v_g0, d.v_e0 = _p(mods[0].weight), _p(mods[0].bias), _p(mods[1].weight), _p(mods[1].bias)
            d.v_w1, d.v_b1, d.v_g1, d.v_e1 = _p(mods[3].weight), _p(mods[3].bias), _p(mods[4].weight), _p(mods[4].bias)
            d.v_wr, d.v_br = _p(v.ctr_reg.weight), _p(v.ctr_reg.bias)
            d.v_go, d.v_eo = _p(v.out_proj[0].weight), _p(v.out_proj[0].bias)
            for i, x in enumerate(v.max_offset_limit.tolist()):
                d.max_offset[i] = float(x)
            d.c, d.h0, d.h1 = mods[0].in_features, mods[0].out_features, mods[3].out_features
            d.n2n_w, d.n2n_b = _p(self.proj_n2n_score.weight), _p(self.proj_n2n_score.bias)
            t2 = self.transformer2.runner_desc()
            d.h_transformer2 = ctypes.addressof(t2)
            alpha = self.optimal_transport.alpha.detach().reshape(1).float().contiguous()
            d.ot_alpha = alpha.data_ptr()
            fm, ot = self.fine_matching, self.optimal_transport
            d.nms_radius, d.acceptance_radius, d.sinkhorn_inf = float(self.nms.NMS_radius), float(fm.acceptance_radius), float(ot.inf)
            d.nms_limit, d.point_limit = int(self.nms.neighbor_limits), int(self.num_points_in_patch)
            d.num_correspondences = int(self.coarse_matching.num_correspondences)
            d.dual_normalization = 1 if self.coarse_matching.dual_normalization else 0
            d.sinkhorn_iterations, d.correspondence_threshold = int(ot.num_iterations), int(fm.correspondence_threshold)
            d.refinement_steps = int(fm.num_refinement_steps)
            self._mdesc, self._mdesc_keep, self._mdesc_key = d, (t2, alpha, ps_keep), key
        return self._mdesc

    def _match_tail(self, out, points_c, lengths_c, nc_ref, tf, n2p, points_f, nf_ref, feats_f, job=None):
        """model_infer.py:180-354 through rdm_match_forward (one host call, two internal synchronisations) or, with a match
        `job`, only its asynchronous first phase (rdm_match_begin)."""
        d = self._match_desc()
        if feats_f.shape[1] != d.c:
            raise RuntimeError(f"rdm_match_forward gathers {d.c}-channel fine features (the vote / transformer width) and scales "
                               f"the patch scores by 1/sqrt({d.c}); got backbone.output_dim = {feats_f.shape[1]}")
        dev = tf.device
        nc, nf, c, K, P = points_c.shape[0], points_f.shape[0], d.c, d.point_limit, d.num_correspondences
        f32, i64, u8 = torch.float32, torch.int64, torch.uint8
        E = lambda shape, dt=f32: torch.empty(shape, dtype=dt, device=dev)
        cap = P * 2 * K
        B = dict(shifted=E((nc, 3)), vote_feats=E((nc, c)), n2n=E(nc), mask=E(nc, u8), sel=E(nc, i64), sel_points=E((nc, 3)),
                 sel_feats=E((nc, c)), sel_n2p=E(nc), sel_n2n=E(nc), node_masks=E(nc, u8), knn=E((nc, K), i64),
                 knn_masks=E((nc, K), u8), corr_ref=E(P, i64), corr_src=E(P, i64), corr_sc=E(P), ms=E((P, K + 1, K + 1)),
                 rcp=E((cap, 3)), scp=E((cap, 3)), cs=E(cap), bij=E((cap, 3), torch.int32), T=E((4, 4)))
        io = L.MatchIO()
        io.points_c, io.lengths_c, io.nc, io.nc_ref = L.ptr(points_c), L.ptr(lengths_c), nc, nc_ref
        io.feats_c, io.n2p_scores = L.ptr(tf), L.ptr(n2p)
        io.points_f, io.nf, io.nf_ref = L.ptr(points_f), nf, nf_ref
        if feats_f.stride(1) != 1 or feats_f.stride(0) % 4 != 0:
            feats_f = feats_f.contiguous()
        io.feats_f, io.ld_feats_f = feats_f.data_ptr(), feats_f.stride(0)
        (io.shifted_points, io.vote_feats, io.n2n_scores, io.nms_mask, io.selected, io.sel_points, io.sel_feats_norm, io.sel_n2p,
         io.sel_n2n, io.node_masks, io.knn_indices, io.knn_masks, io.corr_ref, io.corr_src, io.corr_node_scores,
         io.matching_scores, io.ref_corr_points, io.src_corr_points, io.corr_scores, io.corr_bij, io.transform) = [
            B[k].data_ptr() for k in ("shifted", "vote_feats", "n2n", "mask", "sel", "sel_points", "sel_feats", "sel_n2p", "sel_n2n",
                                      "node_masks", "knn", "knn_masks", "corr_ref", "corr_src", "corr_sc", "ms", "rcp", "scp", "cs",
                                      "bij", "T")]
        lib = L.lib()
        wsb = int(lib.rdm_match_workspace(ctypes.byref(d), nc, nc_ref, nf, nf_ref))
        ws = torch.empty(max(wsb, 1), dtype=u8, device=dev)
        res = L.MatchResult()
        if job is not None:  # asynchronous form: phase 1 only; _match_continue / _match_finish complete it
            with torch.cuda.device(dev):
                L.call("rdm_match_begin", job, ctypes.byref(d), ctypes.byref(io), L.ptr(ws), wsb, L.stream())
            # everything the job still points at until rdm_match_finish: the buffers, and the descriptor structs (the job keeps a
            # COPY of `d`, but `d` itself points at the transformer-2 descriptor and the weight tensors of this weights epoch)
            return dict(out=out, B=B, io=io, ws=ws, res=res, job=job, nc_ref=nc_ref,
                        keep=(points_c, lengths_c, tf, n2p, points_f, feats_f, d, self._mdesc_keep))
        with torch.cuda.device(dev):
            L.call("rdm_match_forward", ctypes.byref(d), ctypes.byref(io), ctypes.byref(res), L.ptr(ws), wsb, L.stream())
        return self._match_outputs(out, B, res, nc_ref)

    def _match_continue(self, ctx):
        """Waits for the NMS survivor counts of the pair (host), queues the rest of the tail (rdm_match_continue)."""
        L.call("rdm_match_continue", ctx["job"], ctypes.byref(ctx["res"]))

    def _match_finish(self, ctx):
        """Waits for the pair's result counts + pose (host) and assembles the output dict."""
        L.call("rdm_match_finish", ctx["job"], ctypes.byref(ctx["res"]))
        return self._match_outputs(ctx["out"], ctx["B"], ctx["res"], ctx["nc_ref"])

    @staticmethod
    def _match_outputs(out, B, res, nc_ref):
        f32 = torch.float32
        n0, n1, k, ncorr = res.n_ref_sel, res.n_src_sel, res.num_patches, res.num_corr
        out["shifted_ref_points_c"], out["shifted_src_points_c"] = B["shifted"][:nc_ref], B["shifted"][nc_ref:]
        out["mask"] = B["mask"].view(torch.bool)  # 0/1 bytes reinterpreted: no conversion kernel
        out["ref_points_c"], out["src_points_c"] = B["sel_points"][:n0], B["sel_points"][n0:n0 + n1]
        out["ref_feats_c"], out["src_feats_c"] = B["sel_feats"][:n0], B["sel_feats"][n0:n0 + n1]
        out["ref_n2p_scores_c"], out["src_n2p_scores_c"] = B["sel_n2p"][:n0], B["sel_n2p"][n0:n0 + n1]
        out["ref_n2n_scores_c"], out["src_n2n_scores_c"] = B["sel_n2n"][:n0], B["sel_n2n"][n0:n0 + n1]
        out["ref_node_corr_indices"], out["src_node_corr_indices"] = B["corr_ref"][:k], B["corr_src"][:k]
        out["node_corr_scores"] = B["corr_sc"][:k]
        out["ref_node_knn_indices"], out["src_node_knn_indices"] = B["knn"][:n0], B["knn"][n0:n0 + n1]
        km = B["knn_masks"].view(torch.bool)
        out["ref_node_knn_masks"], out["src_node_knn_masks"] = km[:n0], km[n0:n0 + n1]
        out["matching_scores"] = B["ms"][:k]
        out["ref_corr_points"], out["src_corr_points"], out["corr_scores"] = B["rcp"][:ncorr], B["scp"][:ncorr], B["cs"][:ncorr]
        out["corr_patch_ij"] = B["bij"][:ncorr]
        out["estimated_transform"] = B["T"]
        out["estimated_transform_host"] = torch.tensor(list(res.transform), dtype=f32).view(4, 4)
        return out

    @torch.no_grad()
    def forward_head(self, data_dict, gp=None):
        """Pyramid (unless `gp`, a ready GpuPyramid, or a reference-style data_dict is given) + encoder + first
        transformer + decoder: asynchronous launches only once the pyramid exists. Returns the state forward_tail needs."""
        out = {}
        self._fwd_key = cache_key(self)  # one fingerprint walk per forward, shared by the backbone / match descriptors
        data_dict = data_dict if data_dict is not None else {}
        if gp is not None or "neighbors" not in data_dict:  # raw stacked points in: build the pyramid here, on the GPU
            if gp is None:
                b = self.cfg.backbone
                gp = build_pyramid_gpu(data_dict["points"], data_dict["lengths"], b.num_stages, b.init_voxel_size,
                                       b.init_radius, self.cfg.neighbor_limits)
            S = gp.desc.num_stages
            L_host = gp.lengths_host
            points_c, points_f, points = gp.points(S - 1), gp.points(1), gp.points(0)
            lengths_c = gp.lengths(S - 1)
            pyr = (gp.desc, [gp])
            out["pyramid"] = gp
        else:
            L_host = data_dict.get("lengths_host")
            if L_host is None:
                L_host = [l.tolist() for l in data_dict["lengths"]]
            points_c, points_f, points = data_dict["points"][-1], data_dict["points"][1], data_dict["points"][0]
            lengths_c = data_dict["lengths"][-1]
            pyr = pyramid_desc(data_dict, L_host)
        nc, nf, n0 = int(L_host[-1][0]), int(L_host[1][0]), int(L_host[0][0])
        feats = data_dict.get("features")
        if feats is None:
            feats = self._ones(points.shape[0], points.device)
        out["ori_ref_points_c"], out["ori_src_points_c"] = points_c[:nc], points_c[nc:]
        ref_points_f, src_points_f = points_f[:nf].contiguous(), points_f[nf:].contiguous()
        out["ref_points_f"], out["src_points_f"] = ref_points_f, src_points_f
        out["ref_points"], out["src_points"] = points[:n0], points[n0:]

        if data_dict.get("stepwise", False):  # per-module host calls (kept for the module-level parity tests)
            feats_list = self.encoder(feats, data_dict, pyr)
            feats_c = feats_list[-1]
            ref_feats_c, src_feats_c = self.transformer(points_c[:nc].contiguous(), points_c[nc:].contiguous(),
                                                        feats_c[:nc], feats_c[nc:])
            tf = torch.cat([ref_feats_c, src_feats_c], 0)
            n2p_logit = ops.linear(tf, self.proj_n2p_score.weight, self.proj_n2p_score.bias)  # (Nc,1)
            n2p = ops.activation(n2p_logit.view(-1), 3)
            feats_list[-1] = torch.cat([tf, n2p_logit], 1)
            dec = self.decoder(feats_list, data_dict, pyr)[0]
            p2p = ops.activation(dec[:, -1].contiguous(), 3)
        else:
            tf, n2p, dec, p2p = self._backbone(pyr[0], nc, feats)
        feats_f = dec[:, :-1]  # strided view (row stride 260): no copy of the 16 MB feature table
        out["ref_p2p_scores_c"], out["src_p2p_scores_c"] = p2p[:nf], p2p[nf:]
        ref_n2p, src_n2p = n2p[:nc], n2p[nc:]
        out["ref_feats_f"], out["src_feats_f"] = feats_f[:nf], feats_f[nf:]

        return (out, data_dict, points_c, lengths_c, nc, tf, n2p, points_f, nf, feats_f, ref_points_f, src_points_f, ref_n2p,
                src_n2p)

    @torch.no_grad()
    def tail_begin(self, state, job):
        """Asynchronous first phase of forward_tail (vote + NMS) on the current stream -> context for tail_continue / tail_finish.
        Only the runner path (vote branch on, not stepwise) has it; returns None otherwise."""
        (out, data_dict, points_c, lengths_c, nc, tf, n2p, points_f, nf, feats_f, ref_points_f, src_points_f, ref_n2p,
         src_n2p) = state
        if not (self.use_vote and not data_dict.get("stepwise", False)):
            return None
        return self._match_tail(out, points_c.contiguous(), lengths_c, nc, tf, n2p, points_f.contiguous(), nf, feats_f, job=job)

    @torch.no_grad()
    def forward_tail(self, state):
        """Everything after the decoder (model_infer.py:180-354)."""
        (out, data_dict, points_c, lengths_c, nc, tf, n2p, points_f, nf, feats_f, ref_points_f, src_points_f, ref_n2p,
         src_n2p) = state
        if self.use_vote and not data_dict.get("stepwise", False):
            return self._match_tail(out, points_c.contiguous(), lengths_c, nc, tf, n2p, points_f.contiguous(), nf, feats_f)
        if self.use_vote:
            shifted, vf = self.vote(points_c, tf)
            out["shifted_ref_points_c"], out["shifted_src_points_c"] = shifted[:nc], shifted[nc:]
            n2n = ops.activation(ops.linear(vf, self.proj_n2n_score.weight, self.proj_n2n_score.bias).view(-1), 3)
            masks, sel, sel_counts = self.nms(shifted.contiguous(), lengths_c, split=nc)
            out["mask"] = masks
            n_ref_sel, n_src_sel = sel_counts.tolist()  # data-dependent counts: the one host sync of this section
            ref_sel, src_sel = sel[:n_ref_sel], sel[n_ref_sel:n_ref_sel + n_src_sel] - nc
            ref_points_c, src_points_c = shifted[:nc][ref_sel].contiguous(), shifted[nc:][src_sel].contiguous()
            ref_feats_c, src_feats_c = vf[:nc][ref_sel], vf[nc:][src_sel]
            out["ref_n2p_scores_c"], out["src_n2p_scores_c"] = ref_n2p[ref_sel], src_n2p[src_sel]
            out["ref_n2n_scores_c"], out["src_n2n_scores_c"] = n2n[:nc][ref_sel], n2n[nc:][src_sel]
            ref_feats_c, src_feats_c = self.transformer2(ref_points_c, src_points_c, ref_feats_c, src_feats_c)
        else:  # cfg.Vote.inference_use_vote = False (the Mulran configuration, experiments/infer.py:119-120)
            ref_points_c, src_points_c = points_c[:nc].contiguous(), points_c[nc:].contiguous()
            ref_feats_c, src_feats_c = tf[:nc], tf[nc:]
            out["ref_n2p_scores_c"], out["src_n2p_scores_c"] = ref_n2p, src_n2p
        out["ref_points_c"], out["src_points_c"] = ref_points_c, src_points_c
        ref_feats_c_norm = torch.nn.functional.normalize(ref_feats_c, p=2, dim=1)
        src_feats_c_norm = torch.nn.functional.normalize(src_feats_c, p=2, dim=1)
        out["ref_feats_c"], out["src_feats_c"] = ref_feats_c_norm, src_feats_c_norm

        k = self.num_points_in_patch
        _, ref_node_masks, ref_knn, ref_knn_masks = ops.point_to_node_partition(ref_points_f, ref_points_c, k)
        _, src_node_masks, src_knn, src_knn_masks = ops.point_to_node_partition(src_points_f, src_points_c, k)
        ref_feats_f, src_feats_f = feats_f[:nf], feats_f[nf:]
        out["ref_feats_f"], out["src_feats_f"] = ref_feats_f, src_feats_f

        ref_ci, src_ci, node_corr_scores = self.coarse_matching(ref_feats_c_norm, src_feats_c_norm, ref_node_masks,
                                                                src_node_masks)
        out["ref_node_corr_indices"], out["src_node_corr_indices"] = ref_ci, src_ci
        out["node_corr_scores"] = node_corr_scores
        out["ref_node_knn_indices"], out["src_node_knn_indices"] = ref_knn, src_knn
        out["ref_node_knn_masks"], out["src_node_knn_masks"] = ref_knn_masks, src_knn_masks

        scores = ops.patch_scores(ref_feats_f, src_feats_f, ref_knn, src_knn, ref_ci, src_ci)
        rm8, sm8 = ref_knn_masks.to(torch.uint8), src_knn_masks.to(torch.uint8)
        ot = self.optimal_transport
        matching_scores = ops.sinkhorn(scores, rm8, sm8, ot.alpha, ot.num_iterations, ot.inf, row_gather=ref_ci,
                                       col_gather=src_ci)
        out["matching_scores"] = matching_scores
        fm = self.fine_matching
        ref_corr, src_corr, corr_scores, T, bij = ops.local_global_registration(
            matching_scores, ref_points_f, src_points_f, ref_knn, src_knn, rm8, sm8, ref_ci, src_ci, fm.acceptance_radius,
            fm.correspondence_threshold, fm.num_refinement_steps)
        out["ref_corr_points"], out["src_corr_points"], out["corr_scores"] = ref_corr, src_corr, corr_scores
        out["estimated_transform"] = T
        out["corr_patch_ij"] = bij
        return out


    def forward(self, data_dict):
        return self.forward_tail(self.forward_head(data_dict))


def create_model(cfg=None):
    return RDMNet(cfg if cfg is not None else make_cfg())
