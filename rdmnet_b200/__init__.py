"""rdmnet_b200: B200-native (sm_100a) implementation of the RDMNet dense-matching hot path.

Host side = Python/PyTorch (tensors, streams, torch.distributed); compute = hand-written CUDA in
librdm_sm100.so behind the C ABI declared in include/rdm_sm100.h. There is no CPU fallback: every op raises
if the CUDA library is missing or the tensors are not on a CUDA device.
"""
__version__ = "0.1.0"


def set_precision(mode):
    """Precision of the dense contractions (BASELINE config 3): "fp32" (default; 3-term tf32 split, the mode every parity test
    runs in) or "tf32" (one tensor-core product per k-step, 10-bit mantissa like fp16; geometry, normalisation statistics,
    Sinkhorn and the pose solver stay fp32). Process-wide (rdm_set_precision)."""
    from . import _lib
    code = {"fp32": 0, "f32": 0, 0: 0, "tf32": 1, "fast": 1, 1: 1}.get(mode)
    if code is None:
        raise ValueError("precision must be 'fp32' or 'tf32'")
    _lib.call("rdm_set_precision", code)


def get_precision():
    from . import _lib
    return "tf32" if _lib.lib().rdm_get_precision() == 1 else "fp32"
