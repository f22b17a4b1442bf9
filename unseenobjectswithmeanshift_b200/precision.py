"""Numerical policy of the PyTorch-side (cuDNN / cuBLAS) layers around the CUDA kernels.

The reference computes this path in fp32 and the parity bar is 1e-3 relative on mask logits
(BASELINE.json). cuDNN convolutions default to TF32 (10-bit mantissa, ~5e-4 relative error per
layer), which eats that budget, so the pixel decoders run their convolutions with TF32 off
unless ``set_strict_fp32(False)`` is called. cuBLAS matmuls are fp32 by PyTorch's default.
"""
import contextlib

import torch

_strict = True


def set_strict_fp32(flag=True):
    global _strict
    _strict = bool(flag)


def strict_fp32():
    return _strict


@contextlib.contextmanager
def conv_precision():
    old = torch.backends.cudnn.allow_tf32
    if _strict:
        torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32 = old
