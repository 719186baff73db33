"""Drop-in boundary, CPU part (no GPU): with rdmnet_b200.dropin installed, the UNMODIFIED reference files
experiments/{config,backbone,model_infer,model,loss,dataset}.py import, `create_model(cfg)` builds the model out of
rdmnet_b200's module classes and the pretrained checkpoint loads with strict=True. The reference tree is the staged copy
under baseline/_ref/RDMNet (written by __graft_entry__.build() when /root/reference is mounted)."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "RDMNet")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "experiments")),
                                reason="reference tree not staged (baseline/_ref/RDMNet)")


@pytest.fixture()
def dropin():
    from rdmnet_b200 import dropin as D
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("geotransformer", "rdmnet", "config", "backbone",
                                                                        "model_infer", "model", "loss", "dataset")}
    for k in saved:
        del sys.modules[k]
    path = list(sys.path)
    D.install(reference_root=REF)
    yield D
    D.uninstall()
    for k in [k for k in sys.modules if k.split(".")[0] in ("geotransformer", "rdmnet", "config", "backbone", "model_infer",
                                                            "model", "loss", "dataset")]:
        del sys.modules[k]
    sys.modules.update(saved)
    sys.path[:] = path


def test_alias_modules_expose_the_reference_names(dropin):
    # every hot-path name the reference's experiments/*.py import
    want = {
        "rdmnet.ext": ["grid_subsampling", "radius_neighbors"],
        "geotransformer.modules.ops": ["point_to_node_partition", "index_select", "radius_search", "apply_transform",
                                       "pairwise_distance", "grid_subsample"],
        "geotransformer.modules.kpconv": ["ConvBlock", "ResidualBlock", "UnaryBlock", "LastUnaryBlock", "nearest_upsample"],
        "geotransformer.modules.registration": ["get_node_correspondences", "get_node_correspondences_disance", "get_node_overlap"],
        "geotransformer.modules.registration.metrics": ["isotropic_transform_error"],
        "geotransformer.modules.sinkhorn": ["LearnableLogOptimalTransport"],
        "geotransformer.modules.geotransformer": ["SuperPointMatching", "SuperPointTargetGenerator", "LocalGlobalRegistration"],
        "rdmnet.thdroformer": ["ThDRoFormer"], "rdmnet.vote": ["Vote_layer", "NMS"],
        "rdmnet.utils.visualization": ["vis_shifte_node", "visualization", "vis_node_grouping"],
        "geotransformer.utils.data": ["registration_collate_fn_stack_mode", "calibrate_neighbors_stack_mode",
                                      "build_dataloader_stack_mode"],
        "geotransformer.utils.open3d": ["registration_with_ransac_from_correspondences"],
    }
    for mod, names in want.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"
    # geotransformer/modules/ops/grid_subsample.py:4 binds through importlib
    import rdmnet_b200.ext_shim as shim
    assert importlib.import_module("rdmnet.ext").grid_subsampling is shim.grid_subsampling


def test_unmodified_reference_model_builds_and_loads_checkpoint(dropin, pretrained_state):
    import rdmnet_b200.modules as M
    config = importlib.import_module("config")          # baseline/_ref/RDMNet/experiments/config.py, unmodified
    model_infer = importlib.import_module("model_infer")  # ... model_infer.py, unmodified
    assert os.path.realpath(model_infer.__file__).startswith(os.path.realpath(REF))
    cfg = config.make_cfg()
    cfg.test.vis = False
    cfg.neighbor_limits = [65, 63, 69, 70, 81]
    model = model_infer.create_model(cfg)
    # the reference's wiring, our operators
    assert isinstance(model.encoder.encoder1_1, M.ConvBlock) and isinstance(model.encoder.encoder3_2, M.ResidualBlock)
    assert isinstance(model.transformer, M.ThDRoFormer) and isinstance(model.vote, M.Vote_layer)
    assert isinstance(model.optimal_transport, M.LearnableLogOptimalTransport)
    assert isinstance(model.fine_matching, M.LocalGlobalRegistration) and isinstance(model.coarse_target, M.SuperPointTargetGenerator)
    missing, unexpected = model.load_state_dict(pretrained_state, strict=True)  # base_tester.py:97-107
    assert not missing and not unexpected
    assert len(model.state_dict()) == 497


def test_training_side_reference_files_import(dropin):
    # experiments/model.py (train/val forward) and dataset.py import on top of the aliases as well
    model = importlib.import_module("model")
    assert hasattr(model, "create_model")
    dataset = importlib.import_module("dataset")
    assert hasattr(dataset, "infer_data_loader")


def test_no_cpu_path_behind_the_aliases(dropin):
    ops = importlib.import_module("geotransformer.modules.ops")
    if torch.cuda.is_available():
        pytest.skip("CPU-container check")
    with pytest.raises(RuntimeError):
        ops.grid_subsample(torch.zeros(10, 3), torch.tensor([10]), 0.3)
