"""Drop-in for the reference's ONE native interface: the pybind module ``MultiScaleDeformableAttention`` built from
``MSMFormer/meanshiftformer/modeling/pixel_decoder/ops/src`` (``vision.cpp:18-21``, dispatcher ``ms_deform_attn.h:25-67``)
and imported by ``ops/functions/ms_deform_attn_func.py:21-22`` as ``import MultiScaleDeformableAttention as MSDA``.

Put this directory on the reference's ``PYTHONPATH`` *instead of* compiling ``ops/`` (``sh make.sh``):

    PYTHONPATH=/path/to/unseenobjectswithmeanshift_b200/dropin:$PYTHONPATH python tools/test_image_with_ms_transformer.py ...

The reference's ``MSDeformAttnFunction`` then calls the sm_100a kernels of ``libmsmformer_b200.so`` through the C ABI
(``msm_ms_deform_attn_fwd`` / ``msm_ms_deform_attn_bwd``, ``include/msmformer_b200.h``). Same argument lists, dtypes
(fp32 tensors, int64 ``spatial_shapes [L,2]`` / ``level_start_index [L]``), return values and error behaviour
(``RuntimeError`` on non-CUDA / non-contiguous inputs, as ``ms_deform_attn_cuda.cu:33-43``) as the pybind module;
``im2col_step`` is accepted and checked like the reference does (``batch % min(batch, im2col_step) == 0``, ``:57``) but
the kernel does not chunk the batch.
"""
import os
import sys

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from unseenobjectswithmeanshift_b200 import ops as _ops  # noqa: E402


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> output [N, Lq, M*D]   (ms_deform_attn.h:25-45, ms_deform_attn_cuda.cu:25-85)"""
    return _ops.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                       im2col_step)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]   (ms_deform_attn.h:47-67, ms_deform_attn_cuda.cu:88-158)"""
    return _ops.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                        grad_output, im2col_step)
