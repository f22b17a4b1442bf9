"""Multi-scale deformable attention module - mirror of the reference's
pixel_decoder/ops/modules/ms_deform_attn.py:34-125 (same parameters: sampling_offsets,
attention_weights, value_proj, output_proj; same special initialisation :66-80).

Unlike the reference (:116-121) there is no ``try/except`` that silently reroutes every failure of
the CUDA op to a PyTorch ``grid_sample`` path: errors propagate.
"""
import math
import os
import warnings

import torch
from torch import nn
from torch.nn import functional as F
from torch.nn.init import constant_, xavier_uniform_

from ...... import ops
from ..functions import MSDeformAttnFunction


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("MSDeformAttn: a power-of-2 head dimension lets the kernel use vector loads.")
        self.im2col_step = 128
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # offsets start as a ring of directions per head, scaled by the point index (reference :66-80)
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2)
        grid = grid.repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, project=True, query_pos=None):
        """query [N,Lq,C]; reference_points [N,Lq,L,2] (or 4); input_flatten [N,S,C];
        input_spatial_shapes int64 [L,2]; input_level_start_index int64 [L] -> [N,Lq,C]."""
        N, Lq, _ = query.shape
        S = input_flatten.shape[1]
        pos_table = None
        if query_pos is not None:
            # extension: the caller passes query and its positional embedding separately; on the tensor-core path
            # the embedding becomes a cached [Lq, N_out] row bias of the offsets / logits GEMM (identical for every
            # image: sine embeddings + level embedding), otherwise it is added here
            # measured on B200 (R50 config, M = 50400 rows): the per-row bias reads in the GEMM epilogue cost more
            # than the separate add kernel saves, so the fold is opt-in (MSM_FOLD_POS=1)
            fold = (os.environ.get("MSM_FOLD_POS", "0") == "1" and not torch.is_grad_enabled()
                    and input_padding_mask is None and reference_points.shape[-1] == 2
                    and ops.tc_linear_enabled() and query.is_contiguous())
            if fold:
                pos_table = query_pos
            else:
                query = query + query_pos
        if not torch.cuda.is_current_stream_capturing():  # the check reads the device (not capturable)
            assert int((input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum()) == S
        M, L, P = self.n_heads, self.n_levels, self.n_points
        value = ops.dense(input_flatten, self.value_proj.weight, self.value_proj.bias)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, S, M, self.d_model // M)
        # offsets and attention logits share their input: one GEMM over the concatenated weights
        n_off = M * L * P * 2
        w_ow = ops.cached_cat(self, "w_ow", [self.sampling_offsets.weight, self.attention_weights.weight])
        b_ow = ops.cached_cat(self, "b_ow", [self.sampling_offsets.bias, self.attention_weights.bias])
        # autograd route whenever ANY tensor that reaches the op needs a gradient - not only when the offset weights
        # train: a frozen pixel decoder under a trainable backbone / value_proj must still differentiate through it
        needs_grad = torch.is_grad_enabled() and (
            query.requires_grad or input_flatten.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            ow = torch.cat([ops.dense(query, self.sampling_offsets.weight, self.sampling_offsets.bias),
                            ops.dense(query, self.attention_weights.weight, self.attention_weights.bias)], -1)
        else:
            if pos_table is not None and ops.linear_supported(query, w_ow):
                tab = ops.cached_value(self, "ow_pos", [pos_table, w_ow],
                                       lambda: F.linear(pos_table[0], w_ow).contiguous())
                ow = ops.linear_fused(query, w_ow, b_ow, rowbias=tab)
            else:
                ow = ops.dense(query if pos_table is None else query + pos_table, w_ow, b_ow)
            if reference_points.shape[-1] == 2 and input_padding_mask is None:
                # inference: softmax + sampling locations + gather in one kernel
                output = ops.ms_deform_attn_fused_forward(value.contiguous(), input_spatial_shapes,
                                                          input_level_start_index, ow, reference_points, L, P)
                if not project:  # extension: the caller fuses output_proj with its residual + LayerNorm
                    return output
                return ops.dense(output, self.output_proj.weight, self.output_proj.bias)
        offsets = ow[..., :n_off].reshape(N, Lq, M, L, P, 2)
        weights = F.softmax(ow[..., n_off:].reshape(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
        if reference_points.shape[-1] == 2:
            wh = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            locations = reference_points[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            locations = reference_points[:, :, None, :, None, :2] \
                + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(
                "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        output = MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                            locations.contiguous(), weights.contiguous(), self.im2col_step)
        if not project:
            return output
        return ops.dense(output, self.output_proj.weight, self.output_proj.bias)
