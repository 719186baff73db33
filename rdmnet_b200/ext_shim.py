"""Drop-in for the reference's native module ``rdmnet.ext`` (geotransformer/extensions/pybind.cpp:8-17): the same two
functions with the same signatures and return conventions, executed by librdm_sm100.so.

    import rdmnet_b200.ext_shim as shim; shim.install()     # sys.modules['rdmnet.ext'] = this module

after which geotransformer/modules/ops/{grid_subsample,radius_search}.py:4 (`importlib.import_module('rdmnet.ext')`)
bind to it unchanged."""
import sys

from . import ops


def grid_subsampling(points, lengths, voxel_size):
    """grid_subsampling.h:6-10: -> [s_points (M,3) f32, s_lengths (B,) i64] on the devices of the inputs."""
    s_points, s_lengths = ops.grid_subsample(points, lengths, voxel_size)
    return [s_points, s_lengths]


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius):
    """radius_neighbors.h:5-11: -> (Nq, max_count) i64, rows sorted by distance, padded with Ns."""
    return ops.radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius)


def install(name="rdmnet.ext"):
    sys.modules[name] = sys.modules[__name__]
    return sys.modules[__name__]
