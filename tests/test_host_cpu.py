"""Host-side logic that needs no GPU: cache invalidation keys of the module mirror, the collate mirror of
geotransformer/utils/data.py (stacking only), the synthetic workload generator's conventions."""
import numpy as np
import torch


def test_cache_key_sees_in_place_weight_edits_in_eval_mode():
    from rdmnet_b200 import modules as M
    blk = M.UnaryBlock(8, 16, 4).eval()
    k0 = M.cache_key(blk)
    assert M.cache_key(blk) == k0
    with torch.no_grad():
        blk.mlp.weight.copy_(torch.randn_like(blk.mlp.weight))  # in-place edit: no epoch bump
    k1 = M.cache_key(blk)
    assert k1 != k0
    blk.mlp.weight.data = torch.randn_like(blk.mlp.weight)  # storage swap
    k2 = M.cache_key(blk)
    assert k2 != k1
    blk.mlp.load_state_dict({"weight": torch.zeros(16, 8), "bias": torch.zeros(16)})  # plain nn.Linear child, loaded directly
    assert M.cache_key(blk) != k2
    M.invalidate_caches()
    assert M.cache_key(blk)[1] == M._WEIGHTS_EPOCH[0]


def test_collate_mirror_stacks_like_the_reference():
    from rdmnet_b200 import data
    rng = np.random.default_rng(0)
    a, b = rng.random((50, 3)).astype(np.float32), rng.random((40, 3)).astype(np.float32)
    item = dict(ref_points=a, src_points=b, ref_feats=np.ones((50, 1), np.float32), src_feats=np.ones((40, 1), np.float32),
                seq_id=0, ref_frame=0, src_frame=4)
    dd = data.registration_collate_fn_stack_mode([item], 5, 0.3, 1.275, [1] * 5, precompute_data=False)
    assert dd["lengths"].tolist() == [50, 40] and dd["points"].shape == (90, 3) and dd["features"].shape == (90, 1)
    assert dd["seq_id"] == 0 and dd["src_frame"] == 4 and dd["batch_size"] == 1  # single-sample lists are unwrapped (data.py:177-180)
    assert torch.equal(dd["points"][:50], torch.from_numpy(a))
    two = data.registration_collate_fn_stack_mode([item, item], 5, 0.3, 1.275, [1] * 5, precompute_data=False)
    assert two["lengths"].tolist() == [50, 50, 40, 40]  # [ref_1..ref_B, src_1..src_B] (data.py:142)


def test_synthetic_pair_direction_and_determinism():
    from rdmnet_b200 import synthetic
    p = synthetic.make_pair(pair_id=3, n_elev=32, n_azim=500)
    q = synthetic.make_pair(pair_id=3, n_elev=32, n_azim=500)
    assert np.array_equal(p["ref_points"], q["ref_points"]) and np.array_equal(p["transform"], q["transform"])
    T = p["transform"].astype(np.float64)
    assert -12.5 < T[0, 3] < -7.5, "src -> ref translation points backwards (the reference's KITTI pair convention)"
    R = T[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    # ground truth consistency: a large part of the transformed source lies on the reference's surfaces
    from scipy.spatial import cKDTree
    d, _ = cKDTree(p["ref_points"]).query(p["src_points"] @ R.T.astype(np.float32) + T[:3, 3].astype(np.float32))
    assert np.quantile(d, 0.3) < 0.6, np.quantile(d, [0.1, 0.3, 0.5])
