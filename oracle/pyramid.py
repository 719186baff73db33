"""TEST INFRASTRUCTURE ONLY (oracle): ctypes bindings for

* ``liboracle.so``  - the plain-C restatement (oracle/pyramid_oracle.c) of the reference's
  ``grid_subsampling`` / ``radius_neighbors`` (geotransformer/extensions/cpu/*), and
* ``_ref/libref_ext.so`` - the reference's own C++ core compiled in place (oracle/Makefile, oracle/ref_shim.cpp).

plus ``precompute_pyramid`` = the call pattern of geotransformer/utils/data.py:13-77.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def _load(path):
    if not os.path.exists(path):
        build()
    return ctypes.CDLL(path)


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        _port = _load(os.path.join(_HERE, "liboracle.so"))
        _port.oracle_grid_subsample.restype = ctypes.c_int64
        _port.oracle_grid_subsample.argtypes = [_f32p, _i64p, ctypes.c_int64, ctypes.c_float, _f32p, _i64p]
        _port.oracle_radius_neighbors.restype = ctypes.c_int64
        _port.oracle_radius_neighbors.argtypes = [_f32p, _f32p, _i64p, _i64p, ctypes.c_int64, ctypes.c_float,
                                                  ctypes.c_int64, _i64p]
    return _port


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_ext.so"))


def ref_lib():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_ext.so"))
        _ref.ref_grid_subsampling.restype = ctypes.c_int64
        _ref.ref_grid_subsampling.argtypes = [_f32p, ctypes.c_int64, _i64p, ctypes.c_int64, ctypes.c_float, _f32p,
                                              _i64p]
        _ref.ref_radius_neighbors.restype = ctypes.c_int64
        _ref.ref_radius_neighbors.argtypes = [_f32p, ctypes.c_int64, _f32p, ctypes.c_int64, _i64p, _i64p,
                                              ctypes.c_int64, ctypes.c_float, _i64p]
    return _ref


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _l(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def grid_subsample(points, lengths, voxel, impl="port"):
    """(N,3) f32, (B,) i64 -> (M,3) f32, (B,) i64.  impl: 'port' (C restatement) | 'ref' (reference C++)."""
    pts, pp = _f(points)
    lens, lp = _l(lengths)
    out = np.empty((max(pts.shape[0], 1), 3), np.float32)
    olen = np.zeros(lens.shape[0], np.int64)
    if impl == "port":
        m = port_lib().oracle_grid_subsample(pp, lp, lens.shape[0], voxel, out.ctypes.data_as(_f32p),
                                             olen.ctypes.data_as(_i64p))
    else:
        m = ref_lib().ref_grid_subsampling(pp, pts.shape[0], lp, lens.shape[0], voxel, out.ctypes.data_as(_f32p),
                                           olen.ctypes.data_as(_i64p))
    return out[:m].copy(), olen


def radius_neighbors(q, s, q_len, s_len, radius, impl="port"):
    """Full-width result (Nq, max_count) i64, padded with Ns (radius_neighbors_cpu.cpp:85)."""
    qa, qp = _f(q)
    sa, sp = _f(s)
    ql, qlp = _l(q_len)
    sl, slp = _l(s_len)
    if impl == "port":
        lib = port_lib()
        w = lib.oracle_radius_neighbors(qp, sp, qlp, slp, ql.shape[0], radius, 0, None)
        out = np.empty((qa.shape[0], w), np.int64)
        lib.oracle_radius_neighbors(qp, sp, qlp, slp, ql.shape[0], radius, w, out.ctypes.data_as(_i64p))
    else:
        lib = ref_lib()
        w = lib.ref_radius_neighbors(qp, qa.shape[0], sp, sa.shape[0], qlp, slp, ql.shape[0], radius, None)
        out = np.empty((qa.shape[0], w), np.int64)
        lib.ref_radius_neighbors(qp, qa.shape[0], sp, sa.shape[0], qlp, slp, ql.shape[0], radius,
                                 out.ctypes.data_as(_i64p))
    return out


def radius_search(q, s, q_len, s_len, radius, limit, impl="port"):
    """geotransformer/modules/ops/radius_search.py:7-27 (made contiguous)."""
    nb = radius_neighbors(q, s, q_len, s_len, radius, impl)
    if limit > 0:
        nb = nb[:, :limit]
    return np.ascontiguousarray(nb)


def canonicalize_ties(nbrs, q, s):
    """Re-sort every run of exactly-equal d2 inside each row by ascending index (pads stay last).

    nanoflann sorts with std::sort on d2 only (nanoflann.hpp:1287) so the order inside such runs is unspecified;
    this makes a reference result comparable with the (d2, idx) order used by the port and the CUDA path.
    """
    nbrs = nbrs.copy()
    ns = s.shape[0]
    sp = np.concatenate([s, np.full((1, 3), 1e18, np.float32)], 0).astype(np.float32)
    d = q[:, None, :] - sp[nbrs]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    d2 = np.where(nbrs == ns, np.float32(np.inf), d2)
    order = np.lexsort((nbrs, d2), axis=1)
    return np.take_along_axis(nbrs, order, axis=1)


def precompute_pyramid(points, lengths, num_stages, voxel, radius, limits, impl="port"):
    """geotransformer/utils/data.py:13-77."""
    pts_l, len_l, nb_l, sub_l, up_l = [], [], [], [], []
    pts, lens = np.ascontiguousarray(points, np.float32), np.asarray(lengths, np.int64)
    for i in range(num_stages):
        if i > 0:
            pts, lens = grid_subsample(pts, lens, voxel, impl)
        pts_l.append(pts)
        len_l.append(lens)
        voxel *= 2
    for i in range(num_stages):
        nb_l.append(radius_search(pts_l[i], pts_l[i], len_l[i], len_l[i], radius, limits[i], impl))
        if i < num_stages - 1:
            sub_l.append(radius_search(pts_l[i + 1], pts_l[i], len_l[i + 1], len_l[i], radius, limits[i], impl))
            up_l.append(radius_search(pts_l[i], pts_l[i + 1], len_l[i], len_l[i + 1], radius * 2, limits[i + 1],
                                      impl))
        radius *= 2
    return {"points": pts_l, "lengths": len_l, "neighbors": nb_l, "subsampling": sub_l, "upsampling": up_l}
