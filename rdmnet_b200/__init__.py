"""rdmnet_b200: B200-native (sm_100a) implementation of the RDMNet dense-matching hot path.

Host side = Python/PyTorch (tensors, streams, torch.distributed); compute = hand-written CUDA in
librdm_sm100.so behind the C ABI declared in include/rdm_sm100.h. There is no CPU fallback: every op raises
if the CUDA library is missing or the tensors are not on a CUDA device.
"""
__version__ = "0.1.0"
