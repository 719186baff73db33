"""Reference arm of bench.py: runs the UNMODIFIED reference (nubot-nudt/RDMNet, staged under baseline/_ref/RDMNet by
__graft_entry__.build()) through its own public code path -

    geotransformer.utils.data.registration_collate_fn_stack_mode   (its C++ extension: 4x grid_subsampling + 13x radius_neighbors)
    experiments/model_infer.RDMNet.forward                          (its PyTorch modules)

- on the host cores (`device='cpu'`: the arm the driver's ratio is taken against) and, for BASELINE.md 3.4, eagerly on
one GPU fed by CPU collate workers (`device='cuda'`: the denominator of the north star's ">= 10x the reference's 1-GPU
pairs/s"). None of rdmnet_b200's kernels, modules or engine are on this path. Test/bench infrastructure: only
bench.py imports it.

What has to be supplied around the reference for it to import at all (none of it is arithmetic of the
path): stubs for open3d (PLY reader only), ipdb, IPython, matplotlib, coloredlogs, easydict, the no-op
rdmnet.utils.visualization, the removed numpy aliases, and `rdmnet.ext` bound to the reference's own C++ core compiled in
place (oracle/_ref/libref_ext.so; its pybind wrapper needs the torch extension toolchain, the core does not).
"""
import contextlib
import logging
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "RDMNet")
_NAMES = ("geotransformer", "rdmnet", "config", "backbone", "model_infer", "model", "loss", "dataset", "open3d")


def available():
    if not os.path.isdir(os.path.join(REF, "experiments")):
        return False, "baseline/_ref/RDMNet not staged"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import pyramid as OP
    if not OP.ref_available():
        return False, "oracle/_ref/libref_ext.so (the reference's C++ core) not built"
    return True, ""


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Import shims + rdmnet.ext over the reference's C++ core. Idempotent."""
    if getattr(install, "done", False):
        return
    from oracle import pyramid as OP

    class _PC:
        def __init__(self, pts):
            self.points = pts

    def read_point_cloud(path):
        raw = open(path, "rb").read()
        body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
        return _PC(np.frombuffer(body, dtype="<f8").reshape(-1, 3).copy())

    def missing(name):
        try:
            __import__(name)
            return False
        except Exception:
            return True

    if missing("open3d"):
        o3d = _mod("open3d")
        o3d.io = _mod("open3d.io", read_point_cloud=read_point_cloud)
        for sub in ("geometry", "utility", "visualization", "pipelines"):
            setattr(o3d, sub, _mod("open3d." + sub))
    if missing("ipdb"):
        _mod("ipdb", set_trace=lambda *a, **k: None)
    if missing("IPython"):
        _mod("IPython", embed=lambda *a, **k: None)
    if missing("matplotlib"):
        mpl = _mod("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _mod("matplotlib.pyplot")
        mpl.cm = _mod("matplotlib.cm")
        _mod("mpl_toolkits")
        _mod("mpl_toolkits.mplot3d", Axes3D=object)
    if missing("coloredlogs"):
        _mod("coloredlogs", ColoredFormatter=logging.Formatter)
    if missing("easydict"):
        class EasyDict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        _mod("easydict", EasyDict=EasyDict)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float

    def grid_subsampling(points, lengths, voxel):
        p, l = OP.grid_subsample(points.numpy(), lengths.numpy(), float(voxel), impl="ref")
        return [torch.from_numpy(p), torch.from_numpy(l)]

    def radius_neighbors(q, s, ql, sl, r):
        return torch.from_numpy(OP.radius_neighbors(q.numpy(), s.numpy(), ql.numpy(), sl.numpy(), float(r), "ref"))

    pkg = types.ModuleType("rdmnet")
    pkg.__path__ = [os.path.join(REF, "rdmnet")]
    sys.modules["rdmnet"] = pkg
    _mod("rdmnet.ext", grid_subsampling=grid_subsampling, radius_neighbors=radius_neighbors)
    utils = types.ModuleType("rdmnet.utils")
    utils.__path__ = []
    sys.modules["rdmnet.utils"] = utils
    noop = lambda *a, **k: None  # noqa: E731
    _mod("rdmnet.utils.visualization", vis_shifte_node=noop, visualization=noop, vis_node_grouping=noop)
    for p in (os.path.join(REF, "experiments"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    install.done = True


@contextlib.contextmanager
def cpu_mode():
    """The reference hard-codes `.cuda()` in its forward (pointcloud_partition.py:87, learnable_sinkhorn.py:34-58,
    local_global_registration.py:54-59, procrustes.py:54-63, vote.py:34); on the host cores those become no-ops."""
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda s, *a, **k: s.contiguous()
    torch.nn.Module.cuda = lambda s, *a, **k: s
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def build_model(state, limits, device):
    install()
    import config
    import model_infer
    cfg = config.make_cfg()
    cfg.test.vis = False
    cfg.neighbor_limits = list(limits)
    model = model_infer.create_model(cfg)
    model.load_state_dict(state, strict=True)
    model.eval()
    return (model.cuda() if device == "cuda" else model), cfg


class _Pairs(torch.utils.data.Dataset):
    def __init__(self, pairs, n):
        self.pairs, self.n = pairs, n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        p = self.pairs[i % len(self.pairs)]
        return dict(ref_points=p["ref_points"], src_points=p["src_points"], ref_feats=np.ones((len(p["ref_points"]), 1), np.float32),
                    src_feats=np.ones((len(p["src_points"]), 1), np.float32))


def _collate(items, cfg):
    from geotransformer.utils.data import registration_collate_fn_stack_mode
    dd = registration_collate_fn_stack_mode(items, cfg.backbone.num_stages, cfg.backbone.init_voxel_size, cfg.backbone.init_radius,
                                            cfg.neighbor_limits)
    for k in ("neighbors", "subsampling", "upsampling"):  # radius_search returns non-contiguous views
        dd[k] = [t.contiguous() for t in dd[k]]
    return dd


class _Collate:
    def __init__(self, cfg):
        self.cfg = cfg

    def __call__(self, items):
        install()
        return _collate(items, self.cfg)


def run_cpu(model, cfg, pairs, n_steps, n_warm, budget_s):
    """The reference's CPU path, one pair at a time on the calling process (collate 1 thread as in a DataLoader worker, the
    forward on all torch threads). -> (pairs done, seconds, last estimated_transform)."""
    ds = _Pairs(pairs, n_warm + n_steps)
    T = None
    with cpu_mode(), torch.no_grad():
        for i in range(n_warm):
            dd = _collate([ds[i]], cfg)
            dd["testing"] = True
            model(dd)
        t0, done = time.perf_counter(), 0
        for i in range(n_warm, n_warm + n_steps):
            dd = _collate([ds[i]], cfg)
            dd["testing"] = True
            T = model(dd)["estimated_transform"]
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return done, dt, T.numpy(), int(dd["points"][0].shape[0])


def run_gpu_eager(model, cfg, pairs, n_steps, n_warm, workers=8):
    """the reference's PyTorch-eager model on one GPU fed by `workers` CPU collate processes (its own
    DataLoader arrangement, geotransformer/utils/torch.py:48-77), RANSAC / .npz writing off. Wall-clock pairs/s over the
    timed pairs + CUDA-event time around the forward (single_tester.py:113-117)."""
    from geotransformer.utils.torch import to_cuda
    loader = torch.utils.data.DataLoader(_Pairs(pairs, n_warm + n_steps), batch_size=1, num_workers=workers, shuffle=False,
                                         collate_fn=_Collate(cfg), persistent_workers=False)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fwd_ms, t0, done, T = 0.0, None, 0, None
    with torch.no_grad():
        for i, dd in enumerate(loader):
            if i == n_warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            dd = to_cuda(dd)
            dd["testing"] = True
            ev0.record()
            out = model(dd)
            ev1.record()
            torch.cuda.synchronize()
            if i >= n_warm:
                fwd_ms += ev0.elapsed_time(ev1)
                done += 1
                T = out["estimated_transform"]
    dt = time.perf_counter() - t0
    return done, dt, fwd_ms / max(done, 1), T.cpu().numpy()


class _TrainPairs(_Pairs):
    def __getitem__(self, i):
        d = super().__getitem__(i)
        d["transform"] = self.pairs[i % len(self.pairs)]["transform"].astype(np.float32)
        return d


def run_gpu_train(state, limits, pairs, n_steps, n_warm, workers=8):
    """Config 4 denominator: the reference's UNMODIFIED training step on one GPU - experiments/model.py (train-mode forward with
    ground truth) + experiments/loss.py OverallLoss + loss.backward() + Adam (experiments/trainval.py:34, epoch_based_trainer.py:104),
    PyTorch eager over its own modules, CPU collate (its C++ extension core) in `workers` DataLoader processes, cKDTree ground
    truth on the host as in loss.py:92,151. The only shim beyond install(): loss.py:240-243 builds index helpers with bare
    torch.arange and masks them with CUDA tensors, which torch 2.x rejects - the loss runs under `with torch.device('cuda')`.
    -> (steps done, seconds, mean CUDA-event ms of forward+loss+backward+step, first and last loss)."""
    install()
    import config
    import loss as loss_mod
    import model as model_mod
    from geotransformer.utils.torch import to_cuda
    cfg = config.make_cfg()
    cfg.test.vis = False
    cfg.neighbor_limits = list(limits)
    net = model_mod.create_model(cfg)
    net.load_state_dict(state, strict=True)
    net = net.cuda().train()
    loss_fn = loss_mod.OverallLoss(cfg).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=1e-6)
    loader = torch.utils.data.DataLoader(_TrainPairs(pairs, n_warm + n_steps), batch_size=1, num_workers=workers, shuffle=False,
                                         collate_fn=_Collate(cfg), persistent_workers=False)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms, t0, done, losses = 0.0, None, 0, []
    for i, dd in enumerate(loader):
        if i == n_warm:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        dd = to_cuda(dd)
        dd["testing"] = False
        ev0.record()
        np.random.seed(1000 + i)
        out = net(dd)
        with torch.device("cuda"):
            ls = loss_fn(out, dd)
        opt.zero_grad()
        ls["loss"].backward()
        opt.step()
        ev1.record()
        val = float(ls["loss"].detach())
        if i >= n_warm:
            dev_ms += ev0.elapsed_time(ev1)
            done += 1
            losses.append(val)
    dt = time.perf_counter() - t0
    return done, dt, dev_ms / max(done, 1), (losses[0], losses[-1])
