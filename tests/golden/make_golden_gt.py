"""Golden vectors of the training-side ground-truth generators, produced by IMPORTING the reference
(geotransformer/modules/registration/matching.py:252-503) in the build container and running it on CPU; same shims as
make_golden.py. Usage: python tests/golden/make_golden_gt.py  ->  tests/golden/gt_small.npz (committed)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG.install_shims()
    from geotransformer.modules.registration.matching import (get_node_correspondences, get_node_correspondences_disance,
                                                              get_node_overlap)
    rng = np.random.default_rng(2024)
    g = {}
    for case, (m, n, k) in enumerate(((37, 41, 16), (90, 75, 64))):
        ang = 0.4 + 0.1 * case
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
        T[:3, 3] = [1.5, -0.7, 0.2]
        Tinv = np.linalg.inv(T).astype(np.float32)
        ref_nodes = (rng.random((m, 3)).astype(np.float32) - 0.5) * np.array([30, 30, 4], np.float32)
        # src nodes: some are ref nodes moved into the src frame (+ noise), the rest random
        src_world = np.concatenate([ref_nodes[: n // 2] + rng.normal(0, 0.4, (n // 2, 3)).astype(np.float32),
                                    (rng.random((n - n // 2, 3)).astype(np.float32) - 0.5) * np.array([30, 30, 4], np.float32)])
        src_nodes = (src_world @ Tinv[:3, :3].T + Tinv[:3, 3]).astype(np.float32)
        ref_knn = ref_nodes[:, None, :] + rng.normal(0, 1.2, (m, k, 3)).astype(np.float32)
        src_knn_world = src_world[:, None, :] + rng.normal(0, 1.2, (n, k, 3)).astype(np.float32)
        src_knn = (src_knn_world @ Tinv[:3, :3].T + Tinv[:3, 3]).astype(np.float32)
        ref_masks, src_masks = rng.random(m) < 0.9, rng.random(n) < 0.9
        ref_knn_masks, src_knn_masks = rng.random((m, k)) < 0.8, rng.random((n, k)) < 0.8
        ref_knn_masks[:, 0], src_knn_masks[:, 0] = True, True
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
        args = (t(ref_nodes), t(src_nodes), t(ref_knn), t(src_knn), t(T), 0.6)
        kw = dict(ref_masks=t(ref_masks), src_masks=t(src_masks), ref_knn_masks=t(ref_knn_masks), src_knn_masks=t(src_knn_masks))
        idx, ov = get_node_correspondences(*args, **kw)
        mask = get_node_correspondences(*args, return_mask=True, **kw)
        b = min(m, n)
        pair_ov = get_node_overlap(t(ref_nodes[:b]), t(src_nodes[:b]), t(ref_knn[:b]), t(src_knn[:b]), t(T), 0.6,
                                   ref_masks=t(ref_masks[:b]), src_masks=t(src_masks[:b]), ref_knn_masks=t(ref_knn_masks[:b]),
                                   src_knn_masks=t(src_knn_masks[:b]))
        dmask = get_node_correspondences_disance(t(ref_nodes), t(src_nodes), t(T), 2.0, ref_masks=t(ref_masks), src_masks=t(src_masks))
        p = f"c{case}_"
        g.update({p + "ref_nodes": ref_nodes, p + "src_nodes": src_nodes, p + "ref_knn": ref_knn, p + "src_knn": src_knn, p + "T": T,
                  p + "ref_masks": ref_masks, p + "src_masks": src_masks, p + "ref_knn_masks": ref_knn_masks,
                  p + "src_knn_masks": src_knn_masks, p + "corr_indices": idx.numpy(), p + "corr_overlaps": ov.numpy(),
                  p + "sphere_mask": mask.numpy(), p + "pair_overlaps": pair_ov.numpy(), p + "distance_mask": dmask.numpy()})
        print(case, "pairs", idx.shape[0], "sphere", int(mask.sum()), "pair overlaps > 0:", int((pair_ov > 0).sum()), "distance mask", int(dmask.sum()))
    np.savez_compressed(os.path.join(HERE, "gt_small.npz"), **g)


if __name__ == "__main__":
    main()
