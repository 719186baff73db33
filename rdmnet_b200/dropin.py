"""Drop-in installer: makes the UNMODIFIED reference entry code - experiments/{backbone,model,model_infer,loss,dataset,
infer,trainval}.py - bind to rdmnet_b200 instead of the reference's own `geotransformer.*` / `rdmnet.*` packages.

    import rdmnet_b200.dropin as dropin
    dropin.install(reference_root="/path/to/RDMNet")     # before the first `import geotransformer` / `import rdmnet`
    sys.path.insert(0, "/path/to/RDMNet/experiments");  import model_infer, config

What it registers in sys.modules (every name the reference's experiments/*.py import from the hot path):

    rdmnet.ext                                   -> rdmnet_b200.ext_shim       (geotransformer/extensions/pybind.cpp:8-17)
    geotransformer.modules.ops (+ submodules)    -> rdmnet_b200.ops            (modules/ops/__init__.py:1-21)
    geotransformer.modules.kpconv                -> rdmnet_b200.modules        (modules/kpconv/__init__.py)
    geotransformer.modules.sinkhorn              -> LearnableLogOptimalTransport
    geotransformer.modules.geotransformer        -> SuperPointMatching, SuperPointTargetGenerator, LocalGlobalRegistration
    geotransformer.modules.registration (+ .matching/.metrics/.procrustes) -> rdmnet_b200.registration
    geotransformer.modules.transformer           -> TransformerLayer, AttentionLayer, MultiHeadAttention, AttentionOutput
    geotransformer.utils.data                    -> rdmnet_b200.data           (GPU pyramid instead of CPU collate workers)
    geotransformer.utils.open3d                  -> GPU RANSAC                 (utils/open3d.py:173-203)
    geotransformer.utils.registration            -> GPU ground-truth ball query get_correspondences (utils/registration.py:203-217;
                                                    other names fall through to the reference's file)
    rdmnet.thdroformer, rdmnet.vote              -> ThDRoFormer / Vote_layer, NMS
    rdmnet.utils.visualization                   -> headless no-ops (the reference module cannot be imported: SURVEY 3.1)

With `reference_root` the remaining, non-hot-path modules (geotransformer.engine, geotransformer.utils.{common,torch,
timer,...}, rdmnet.datasets.*) resolve to the reference's own files, which is how infer.py / trainval.py run unchanged.
Third-party modules the reference imports but this path never calls (easydict, IPython, ipdb, coloredlogs) get minimal
stand-ins only if they are not installed.
"""
import importlib
import logging
import os
import sys
import types

from . import data as _data
from . import ext_shim as _ext
from . import modules as _modules
from . import ops as _ops
from . import registration as _registration

_INSTALLED = {}


def _module(name, attrs=None, path=None, doc=None):
    m = types.ModuleType(name, doc)
    if path is not None:
        m.__path__ = list(path)
    if attrs:
        m.__dict__.update(attrs)
    sys.modules[name] = m
    _INSTALLED[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules and not hasattr(sys.modules[parent], child):
        setattr(sys.modules[parent], child, m)  # (a function of the same name, e.g. ops.grid_subsample, wins: as in the reference)
    return m


def _pick(src, names):
    return {n: getattr(src, n) for n in names}


def _pairwise_distance_np(x, y, normalized=False, channel_first=False):
    """geotransformer/modules/ops/pairwise_distance.py:34-60 (numpy helper used by the dataset-side utilities)."""
    import numpy as np
    if channel_first:
        xy = np.matmul(x.swapaxes(-1, -2), y)
        x2 = np.expand_dims(np.sum(x ** 2, axis=-2), -1)
        y2 = np.expand_dims(np.sum(y ** 2, axis=-2), -2)
    else:
        xy = np.matmul(x, y.swapaxes(-1, -2))
        x2 = np.expand_dims(np.sum(x ** 2, axis=-1), -1)
        y2 = np.expand_dims(np.sum(y ** 2, axis=-1), -2)
    d = 2.0 - 2.0 * xy if normalized else x2 - 2 * xy + y2
    return np.maximum(d, 1e-12)


def install_third_party_stubs():
    """Stand-ins for packages the reference imports at module level but the hot path never calls. Real installations win."""
    def have(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not have("easydict"):
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    setattr(self, k, v)

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                self[k] = v

        _module("easydict", {"EasyDict": EasyDict})
    if not have("IPython"):
        _module("IPython", {"embed": lambda *a, **k: None})
    if not have("ipdb"):
        _module("ipdb", {"set_trace": lambda *a, **k: None})
    if not have("coloredlogs"):
        _module("coloredlogs", {"ColoredFormatter": logging.Formatter})


def install(reference_root=None, third_party_stubs=True):
    """Registers the alias modules (idempotent). Returns the dict {module name: module} it installed."""
    if _INSTALLED.get("__done__") == (reference_root,):
        return _INSTALLED
    import numpy as np
    if not hasattr(np, "int"):  # geotransformer/utils/open3d.py:88 and friends use the removed numpy aliases
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if third_party_stubs:
        install_third_party_stubs()
    gt_path = [os.path.join(reference_root, "geotransformer")] if reference_root else []
    rd_path = [os.path.join(reference_root, "rdmnet")] if reference_root else []
    sub = lambda base, *p: [os.path.join(b, *p) for b in base]  # noqa: E731

    # package skeletons: __path__ lets everything that is NOT aliased below come from the reference tree
    _module("geotransformer", path=gt_path)
    _module("geotransformer.modules", path=sub(gt_path, "modules"))
    _module("geotransformer.utils", path=sub(gt_path, "utils"))
    _module("rdmnet", path=rd_path)
    _module("rdmnet.utils", path=[])  # the reference's rdmnet/utils/__init__ chain imports a missing module

    _module("rdmnet.ext", _pick(_ext, ["grid_subsampling", "radius_neighbors"]), doc=_ext.__doc__)

    op_names = ["grid_subsample", "radius_search", "index_select", "pairwise_distance", "point_to_node_partition",
                "apply_transform", "apply_rotation", "inverse_transform", "get_rotation_translation_from_transform",
                "get_transform_from_rotation_translation"]
    op_attrs = _pick(_ops, op_names)
    op_attrs["pairwise_distance_np"] = _pairwise_distance_np
    op_attrs["get_point_to_node_indices"] = lambda points, nodes, return_counts=False: (
        _ops.point_to_node_partition(points, nodes, 1, return_count=True)[:2] if return_counts
        else _ops.point_to_node_partition(points, nodes, 1)[0])
    _module("geotransformer.modules.ops", op_attrs, path=[], doc=_ops.__doc__)
    for subname, names in (("grid_subsample", ["grid_subsample"]), ("radius_search", ["radius_search"]),
                           ("index_select", ["index_select"]), ("pairwise_distance", ["pairwise_distance", "pairwise_distance_np"]),
                           ("pointcloud_partition", ["point_to_node_partition", "get_point_to_node_indices"]),
                           ("transformation", ["apply_transform", "apply_rotation", "inverse_transform",
                                               "get_rotation_translation_from_transform",
                                               "get_transform_from_rotation_translation"])):
        _module("geotransformer.modules.ops." + subname, {n: op_attrs[n] for n in names})

    kp = _pick(_modules, ["KPConv", "ConvBlock", "ResidualBlock", "UnaryBlock", "LastUnaryBlock", "GroupNorm"])
    kp.update(nearest_upsample=_ops.nearest_upsample, maxpool=_ops.maxpool)
    _module("geotransformer.modules.kpconv", kp, path=[])
    _module("geotransformer.modules.kpconv.kpconv", {"KPConv": _modules.KPConv})
    _module("geotransformer.modules.kpconv.modules", {k: v for k, v in kp.items() if k[0].isupper()})
    _module("geotransformer.modules.kpconv.functional", {"nearest_upsample": _ops.nearest_upsample, "maxpool": _ops.maxpool})

    _module("geotransformer.modules.sinkhorn", {"LearnableLogOptimalTransport": _modules.LearnableLogOptimalTransport}, path=[])
    _module("geotransformer.modules.sinkhorn.learnable_sinkhorn",
            {"LearnableLogOptimalTransport": _modules.LearnableLogOptimalTransport})

    geo = _pick(_modules, ["SuperPointMatching", "SuperPointTargetGenerator", "LocalGlobalRegistration"])
    _module("geotransformer.modules.geotransformer", geo, path=[])
    _module("geotransformer.modules.geotransformer.superpoint_matching", {"SuperPointMatching": _modules.SuperPointMatching})
    _module("geotransformer.modules.geotransformer.superpoint_target",
            {"SuperPointTargetGenerator": _modules.SuperPointTargetGenerator})
    _module("geotransformer.modules.geotransformer.local_global_registration",
            {"LocalGlobalRegistration": _modules.LocalGlobalRegistration})

    reg_names = ["get_node_correspondences", "get_node_correspondences_disance", "get_node_overlap", "weighted_procrustes",
                 "WeightedProcrustes", "relative_rotation_error", "relative_translation_error", "isotropic_transform_error"]
    reg = _pick(_registration, reg_names)
    _module("geotransformer.modules.registration", reg, path=[])
    _module("geotransformer.modules.registration.matching", {n: reg[n] for n in reg_names[:3]})
    _module("geotransformer.modules.registration.metrics", {n: reg[n] for n in reg_names[5:]})
    _module("geotransformer.modules.registration.procrustes", {n: reg[n] for n in reg_names[3:5]})

    tr = _pick(_modules, ["TransformerLayer", "AttentionLayer", "MultiHeadAttention", "AttentionOutput"])
    _module("geotransformer.modules.transformer", tr, path=[])
    _module("geotransformer.modules.transformer.vanilla_transformer", {k: v for k, v in tr.items() if k != "AttentionOutput"})
    _module("geotransformer.modules.transformer.output_layer", {"AttentionOutput": _modules.AttentionOutput})

    _module("geotransformer.utils.data", _pick(_data, ["precompute_data_stack_mode", "single_collate_fn_stack_mode",
                                                       "registration_collate_fn_stack_mode", "calibrate_neighbors_stack_mode",
                                                       "build_dataloader_stack_mode"]), doc=_data.__doc__)
    _module("geotransformer.utils.open3d", {"registration_with_ransac_from_correspondences":
                                            _registration.registration_with_ransac_from_correspondences})

    reg_mod = _module("geotransformer.utils.registration", {"get_correspondences": _registration.get_correspondences})
    if reference_root:  # every other name of that module (offline evaluation helpers) comes from the reference's own file
        ref_file = os.path.join(reference_root, "geotransformer", "utils", "registration.py")

        def _fallback(name, _cache={}):
            if "m" not in _cache:
                import importlib.util
                spec = importlib.util.spec_from_file_location("geotransformer.utils._reference_registration", ref_file)
                _cache["m"] = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(_cache["m"])
            try:
                return getattr(_cache["m"], name)
            except AttributeError:
                raise AttributeError("module 'geotransformer.utils.registration' has no attribute %r" % name) from None

        reg_mod.__getattr__ = _fallback
    _module("rdmnet.thdroformer", {"ThDRoFormer": _modules.ThDRoFormer}, path=[])
    _module("rdmnet.thdroformer.thdroformer", {"ThDRoFormer": _modules.ThDRoFormer})
    _module("rdmnet.vote", _pick(_modules, ["Vote_layer", "NMS"]), path=[])
    _module("rdmnet.vote.vote", _pick(_modules, ["Vote_layer", "NMS"]))
    noop = lambda *a, **k: None  # noqa: E731
    _module("rdmnet.utils.visualization", {"vis_shifte_node": noop, "visualization": noop, "vis_node_grouping": noop})

    if reference_root:
        exp = os.path.join(reference_root, "experiments")
        for p in (reference_root, exp):
            if p not in sys.path:
                sys.path.insert(0, p)
    _INSTALLED["__done__"] = (reference_root,)
    return _INSTALLED


def uninstall():
    """Removes every alias module again (tests)."""
    for name in [n for n in _INSTALLED if n != "__done__"]:
        if sys.modules.get(name) is _INSTALLED[name]:
            del sys.modules[name]
    _INSTALLED.clear()
