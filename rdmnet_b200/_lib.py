"""ctypes loader for librdm_sm100.so (the C ABI of include/rdm_sm100.h). Fails loudly: no fallback."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librdm_sm100.so")
_lib = None

c_void_p, c_int, c_i64, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/rdm_sm100.h (tests/test_abi_cpu.py checks it)
SIGNATURES = {
    "rdm_last_error": (ctypes.c_char_p, []),
    "rdm_version": (c_int, []),
    "rdm_launch_count": (ctypes.c_ulonglong, []),
    "rdm_tc_gemm_count": (ctypes.c_ulonglong, []),
    "rdm_abi_layout": (c_int, [c_void_p, c_int]),
    "rdm_prof_enable": (None, [c_int]),
    "rdm_prof_read": (c_int, [c_void_p, c_int]),
    "rdm_grid_subsample_workspace": (c_size_t, [c_i64, c_int]),
    "rdm_grid_subsample": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_selfcheck_bucket_table": (c_int, [c_i64]),
    "rdm_radius_search_workspace": (c_size_t, [c_i64, c_int]),
    "rdm_radius_search": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_float, c_int,
                                  c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_build_pyramid_bytes": (c_size_t, [c_i64, c_void_p]),
    "rdm_build_pyramid_workspace": (c_size_t, [c_i64, c_void_p]),
    "rdm_build_pyramid": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "rdm_pyramid_job_create": (c_void_p, []),
    "rdm_pyramid_job_destroy": (None, [c_void_p]),
    "rdm_build_pyramid_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t,
                                        c_void_p]),
    "rdm_build_pyramid_finish": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_kpconv_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_maxpool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rdm_upsample_concat": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rdm_linear_workspace": (c_size_t, [c_int, c_int, c_int]),
    "rdm_linear": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                           c_void_p, c_size_t, c_void_p]),
    "rdm_groupnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int,
                              c_float, c_void_p, c_void_p]),
    "rdm_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p]),
    "rdm_activation": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_float, c_void_p]),
    "rdm_rope": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rdm_attention": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_void_p]),
    "rdm_tf_layer_blob_floats": (c_size_t, []),
    "rdm_tf_project": (c_int, [c_void_p, c_int, c_void_p]),
    "rdm_tf_attend": (c_int, [c_void_p, c_int, c_void_p]),
    "rdm_encoder_workspace": (c_size_t, [c_void_p, c_int, c_void_p, c_int]),
    "rdm_encoder_forward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_decoder_workspace": (c_size_t, [c_void_p, c_int, c_void_p, c_int, c_int]),
    "rdm_decoder_forward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                    c_void_p, c_size_t, c_void_p]),
    "rdm_thdroformer_workspace": (c_size_t, [c_int, c_int, c_int]),
    "rdm_thdroformer_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_point_to_node_workspace": (c_size_t, [c_int, c_int]),
    "rdm_point_to_node": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "rdm_coarse_matching": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "rdm_patch_scores": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_int, c_float, c_void_p, c_void_p]),
    "rdm_sinkhorn": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_float, c_void_p, c_void_p]),
    "rdm_backbone_workspace": (c_size_t, [c_void_p, c_void_p, c_int]),
    "rdm_backbone_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_match_workspace": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "rdm_match_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_weighted_procrustes": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "rdm_lgr_workspace": (c_size_t, [c_int, c_int]),
    "rdm_lgr": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_size_t, c_void_p]),
    "rdm_index_select": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_int, c_i64, c_void_p, c_void_p, c_void_p]),
    "rdm_apply_transform": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_neighbor_histogram": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rdm_kpconv_gather_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p, c_void_p]),
    "rdm_transpose": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rdm_colsum": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rdm_groupnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "rdm_maxpool_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rdm_upsample_concat_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rdm_linear_bwd_workspace": (c_size_t, [c_int, c_int, c_int]),
    "rdm_linear_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "rdm_linear_kn_workspace": (c_size_t, [c_int, c_int, c_int]),
    "rdm_scatter_add_rows": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_int, c_i64, c_void_p, c_void_p]),
    "rdm_match_job_create": (c_void_p, []),
    "rdm_match_job_destroy": (None, [c_void_p]),
    "rdm_match_job_reset": (c_int, [c_void_p]),
    "rdm_match_set_patch_wait_event": (c_int, [c_void_p]),
    "rdm_match_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_match_continue": (c_int, [c_void_p, c_void_p]),
    "rdm_match_finish": (c_int, [c_void_p, c_void_p]),
    "rdm_backbone_set_encoder_event": (c_int, [c_void_p]),
    "rdm_set_precision": (c_int, [c_int]),
    "rdm_get_precision": (c_int, []),
    "rdm_rope_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rdm_attention_bwd_workspace": (c_size_t, [c_int, c_int]),
    "rdm_attention_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rdm_sinkhorn_bwd_workspace": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rdm_sinkhorn_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p,
                                 c_size_t, c_void_p, c_void_p, c_void_p]),
    "rdm_activation_bwd": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_float, c_void_p, c_void_p]),
    "rdm_voxel_downsample_workspace": (c_size_t, [c_int]),
    "rdm_voxel_downsample": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_presplit_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rdm_presplit_register": (c_int, [c_void_p, c_void_p]),
    "rdm_presplit_clear": (None, []),
    "rdm_ransac_workspace": (c_size_t, [c_int]),
    "rdm_ransac_correspondences": (c_int, [c_void_p, c_void_p, c_int, c_float, c_int, c_int, ctypes.c_ulonglong, c_void_p,
                                           c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_node_correspondences_workspace": (c_size_t, [c_int, c_int]),
    "rdm_node_correspondences": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rdm_compact_nonzero": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_node_distance_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p]),
}


class TfProjJob(ctypes.Structure):
    _fields_ = [("x", c_void_p), ("wt", c_void_p), ("bias", c_void_p), ("emb", c_void_p), ("y", c_void_p),
                ("n", c_int), ("ldx", c_int), ("lde", c_int), ("ldy_t", c_int)]


class TfAttnJob(ctypes.Structure):
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("x", c_void_p), ("blob", c_void_p), ("out", c_void_p),
                ("nq", c_int), ("nk", c_int), ("ldx", c_int), ("ldk_t", c_int)]


class UnaryDesc(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("b", c_void_p), ("gn_w", c_void_p), ("gn_b", c_void_p), ("c_in", c_int), ("c_out", c_int),
                ("ldw", c_int)]


class BlockDesc(ctypes.Structure):
    _fields_ = [("unary1", UnaryDesc), ("unary2", UnaryDesc), ("shortcut", UnaryDesc), ("kpconv_w", c_void_p),
                ("kpconv_wt", c_void_p), ("kpconv_b", c_void_p), ("kernel_points", c_void_p), ("h_kernel_points", c_void_p),
                ("norm_conv_w", c_void_p), ("norm_conv_b", c_void_p), ("c_in", c_int), ("c_mid_in", c_int),
                ("c_mid_out", c_int), ("c_out", c_int), ("strided", c_int), ("stage", c_int), ("sigma", c_float)]


class PyramidDesc(ctypes.Structure):
    _fields_ = [("points", c_void_p * 8), ("neighbors", c_void_p * 8), ("subsampling", c_void_p * 8),
                ("upsampling", c_void_p * 8), ("n", c_int * 8), ("nb_width", c_int * 8), ("sub_width", c_int * 8),
                ("up_width", c_int * 8), ("num_stages", c_int), ("index_bytes", c_int), ("order", c_void_p * 8)]


class PyramidCfg(ctypes.Structure):
    _fields_ = [("num_stages", c_int), ("batch", c_int), ("first_voxel", c_float), ("first_radius", c_float),
                ("limits", c_int * 8), ("skip_up0", c_int), ("up_nearest_only", c_int)]


class ThdroformerDesc(ctypes.Structure):
    _fields_ = [("emb_w", c_void_p), ("emb_b", c_void_p), ("in_w", c_void_p), ("in_b", c_void_p), ("out_w", c_void_p),
                ("out_b", c_void_p), ("layer_blobs", c_void_p * 32), ("is_self", c_int * 32), ("num_layers", c_int),
                ("c_in", c_int), ("c_out", c_int)]


class BackboneDesc(ctypes.Structure):
    _fields_ = [("h_blocks", c_void_p), ("num_blocks", c_int), ("groups", c_int), ("h_transformer1", c_void_p),
                ("n2p_w", c_void_p), ("n2p_b", c_void_p), ("h_dec", c_void_p), ("num_dec", c_int)]


class BackboneOut(ctypes.Structure):
    _fields_ = [("feats_c", c_void_p), ("n2p_scores", c_void_p), ("feats_f", c_void_p), ("ld_feats_f", c_int),
                ("p2p_scores", c_void_p)]


class MatchDesc(ctypes.Structure):
    _fields_ = [("v_w0", c_void_p), ("v_b0", c_void_p), ("v_g0", c_void_p), ("v_e0", c_void_p),
                ("v_w1", c_void_p), ("v_b1", c_void_p), ("v_g1", c_void_p), ("v_e1", c_void_p),
                ("v_wr", c_void_p), ("v_br", c_void_p), ("v_go", c_void_p), ("v_eo", c_void_p),
                ("max_offset", c_float * 3), ("c", c_int), ("h0", c_int), ("h1", c_int),
                ("n2n_w", c_void_p), ("n2n_b", c_void_p), ("h_transformer2", c_void_p), ("ot_alpha", c_void_p),
                ("nms_radius", c_float), ("acceptance_radius", c_float), ("sinkhorn_inf", c_float),
                ("nms_limit", c_int), ("point_limit", c_int), ("num_correspondences", c_int), ("dual_normalization", c_int),
                ("sinkhorn_iterations", c_int), ("correspondence_threshold", c_int), ("refinement_steps", c_int)]


class MatchIO(ctypes.Structure):
    _fields_ = [("points_c", c_void_p), ("lengths_c", c_void_p), ("nc", c_int), ("nc_ref", c_int), ("feats_c", c_void_p),
                ("n2p_scores", c_void_p), ("points_f", c_void_p), ("nf", c_int), ("nf_ref", c_int), ("feats_f", c_void_p),
                ("ld_feats_f", c_int),
                ("shifted_points", c_void_p), ("vote_feats", c_void_p), ("n2n_scores", c_void_p), ("nms_mask", c_void_p),
                ("selected", c_void_p), ("sel_points", c_void_p), ("sel_feats_norm", c_void_p), ("sel_n2p", c_void_p),
                ("sel_n2n", c_void_p), ("node_masks", c_void_p), ("knn_indices", c_void_p), ("knn_masks", c_void_p),
                ("corr_ref", c_void_p), ("corr_src", c_void_p), ("corr_node_scores", c_void_p), ("matching_scores", c_void_p),
                ("ref_corr_points", c_void_p), ("src_corr_points", c_void_p), ("corr_scores", c_void_p), ("corr_bij", c_void_p),
                ("transform", c_void_p)]


class MatchResult(ctypes.Structure):
    _fields_ = [("n_ref_sel", c_int), ("n_src_sel", c_int), ("num_patches", c_int), ("num_corr", c_int),
                ("transform", c_float * 16)]


def lib():
    """Returns the loaded library, raising RuntimeError if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C rdmnet_b200/csrc). rdmnet_b200 has no CPU or PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("rdmnet_b200: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("rdmnet_b200: expected a contiguous tensor")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(code, what):
    if code != 0:
        raise RuntimeError(f"{what} failed (code {code}): {lib().rdm_last_error().decode()}")


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


class ProfRecord(ctypes.Structure):
    _fields_ = [("tag", c_int), ("ms", c_float), ("m", c_int), ("n", c_int), ("h", c_int), ("c", c_int)]


def prof_enable(on):
    """In-library CUDA-event timing of the KPConv gather / weight-GEMM launches (rdm_prof_enable)."""
    lib().rdm_prof_enable(int(on))  # bit mask: 1 = gather brackets, 2 = weight-GEMM brackets


def prof_read(max_records=1 << 16):
    """-> list of (tag, ms, m, n, h, c); synchronises the recorded events."""
    buf = (ProfRecord * max_records)()
    n = lib().rdm_prof_read(ctypes.cast(buf, c_void_p), max_records)
    return [(buf[i].tag, buf[i].ms, buf[i].m, buf[i].n, buf[i].h, buf[i].c) for i in range(n)]


def launch_count():
    return int(lib().rdm_launch_count())
