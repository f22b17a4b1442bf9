"""Autograd wrapper of the deformable-attention op - mirror of the reference's
pixel_decoder/ops/functions/ms_deform_attn_func.py:32-49, calling libmsmformer_b200 instead of the
``MultiScaleDeformableAttention`` pybind module.

The reference's pure-PyTorch ``ms_deform_attn_core_pytorch`` ("for debug and test only", :52-72)
is deliberately NOT mirrored here: this package has no non-CUDA path. Its arithmetic lives in
``oracle/pixel_decoder.py`` for the tests.
"""
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ...... import ops


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        value, sampling_locations, attention_weights = (value.detach(), sampling_locations.detach(),
                                                        attention_weights.detach())
        output = ops.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                            sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, aw = ctx.saved_tensors
        grad_value, grad_loc, grad_aw = ops.ms_deform_attn_backward(value, shapes, lsi, loc, aw,
                                                                    grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_aw, None
