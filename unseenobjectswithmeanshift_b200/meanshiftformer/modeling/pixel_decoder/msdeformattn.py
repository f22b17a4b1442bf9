"""MSDeformAttn pixel decoder - mirror of the reference's modeling/pixel_decoder/msdeformattn.py
(encoder :23-161, MSDeformAttnPixelDecoder :164-358): same class names, constructor keywords,
``from_config`` keys and parameter names. Dense layers stay PyTorch/cuBLAS/cuDNN; the sampling op
is csrc/ms_deform_attn.cu.
"""
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn.init import normal_

from .... import ops
from ....d2compat import SEM_SEG_HEADS_REGISTRY, Conv2d, ShapeSpec, c2_xavier_fill, configurable, get_norm
from ....precision import conv_precision
from ..transformer_decoder.position_encoding import PositionEmbeddingSine
from ..transformer_decoder.transformer import _get_activation_fn, _get_clones
from .ops.modules import MSDeformAttn


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        relu = self.activation is F.relu   # ops.dense: tensor-core GEMM, also under autograd (fp32 training)
        h = ops.dense(src, self.linear1.weight, self.linear1.bias, relu=relu)
        if not relu:
            h = self.activation(h)
        return self.norm2(src + self.dropout3(ops.dense(self.dropout2(h), self.linear2.weight, self.linear2.bias)))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        attn = self.self_attn
        if not self.training and self.activation is F.relu and src.is_contiguous() \
                and self.linear1.weight.shape[0] % 32 == 0 and padding_mask is None \
                and reference_points.shape[-1] == 2 \
                and ops.linear_ln_supported(src, attn.output_proj.weight, src, self.norm1) \
                and ops.linear_ln_supported(src, attn.output_proj.weight, src, self.norm2):
            # inference: both post-norm residual blocks end in a GEMM whose epilogue does the add + LayerNorm
            # query = src + pos enters one GEMM only: (src + pos) W^T = src W^T + (pos W^T), a cached row-bias table
            sampled = attn(src, reference_points, src, spatial_shapes, level_start_index, padding_mask, project=False,
                           query_pos=pos)
            src = ops.linear_ln(sampled, attn.output_proj.weight, attn.output_proj.bias, src, self.norm1)
            if ops.ffn_ln_supported(src, self.linear1.weight, self.linear2.weight, self.norm2):
                return ops.ffn_ln(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                                  self.norm2)
            h = ops.dense(src, self.linear1.weight, self.linear1.bias, relu=True)
            return ops.linear_ln(h, self.linear2.weight, self.linear2.bias, src, self.norm2)
        a = attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes, level_start_index, padding_mask)
        return self.forward_ffn(self.norm1(src + self.dropout1(a)))


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Pixel centres of every level in normalised coordinates, [B, sum(HW), L, 2] (x, y)."""
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes.tolist()):
            ys = torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device)
            xs = torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device)
            ry, rx = torch.meshgrid(ys, xs, indexing="ij")
            ry = ry.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            rx = rx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((rx, ry), -1))
        ref = torch.cat(pts, 1)
        return ref[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                reference_points=None):
        out = src
        ref = reference_points
        if ref is None:
            ref = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        for layer in self.layers:
            out = layer(out, pos, ref, spatial_shapes, level_start_index, padding_mask)
        return out


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1,
                 activation="relu", num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                    num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.level_embed)

    def _geometry(self, shapes, batch, device):
        """Device-side constants of a feature pyramid (level shapes, level starts, pixel-centre reference
        points), built once per (shapes, batch, device): no host<->device traffic on later calls, which also
        keeps the forward capturable in a CUDA graph."""
        key = (shapes, batch, str(device))
        cache = self.__dict__.setdefault("_geom_cache", {})
        if key not in cache:
            spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=device)
            starts = [0]
            for h, w in shapes[:-1]:
                starts.append(starts[-1] + h * w)
            level_start_index = torch.as_tensor(starts, dtype=torch.long, device=device)
            valid_ratios = torch.ones(batch, len(shapes), 2, dtype=torch.float32, device=device)
            ref = self.encoder.get_reference_points(spatial_shapes, valid_ratios, device)
            cache[key] = (spatial_shapes, level_start_index, valid_ratios, ref)
        return cache[key]

    def forward(self, srcs, pos_embeds):
        """srcs / pos_embeds: per-level [B,C,H,W], coarse to fine. No padding: valid ratios are 1."""
        B = srcs[0].shape[0]
        dev = srcs[0].device
        shapes = tuple((int(s.shape[-2]), int(s.shape[-1])) for s in srcs)
        src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)

        def build_pos():
            return torch.cat([p.flatten(2).transpose(1, 2) + self.level_embed[l].view(1, 1, -1)
                              for l, p in enumerate(pos_embeds)], 1)
        if torch.is_grad_enabled():
            pos = build_pos()
        else:  # sine embeddings depend on the shapes only: one tensor per pyramid until level_embed changes
            pos = ops.cached_value(self, "pos%s_%d_%s" % (shapes, B, dev), [self.level_embed], build_pos)
        spatial_shapes, level_start_index, valid_ratios, ref = self._geometry(shapes, B, dev)
        memory = self.encoder(src, spatial_shapes, level_start_index, valid_ratios, pos, None, reference_points=ref)
        return memory, shapes, level_start_index


@SEM_SEG_HEADS_REGISTRY.register()
class MSDeformAttnPixelDecoder(nn.Module):
    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, transformer_dropout: float, transformer_nheads: int,
                 transformer_dim_feedforward: int, transformer_enc_layers: int, conv_dim: int, mask_dim: int,
                 norm: Optional[Union[str, Callable]] = None, transformer_in_features: List[str],
                 common_stride: int):
        super().__init__()
        by_stride = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in by_stride]  # "res2" .. "res5"
        self.feature_strides = [v.stride for _, v in by_stride]
        self.feature_channels = [v.channels for _, v in by_stride]
        enc_in = sorted(((k, v) for k, v in input_shape.items() if k in transformer_in_features),
                        key=lambda kv: kv[1].stride)
        self.transformer_in_features = [k for k, _ in enc_in]
        enc_channels = [v.channels for _, v in enc_in]
        self.transformer_feature_strides = [v.stride for _, v in enc_in]
        self.transformer_num_feature_levels = len(self.transformer_in_features)
        chans = enc_channels[::-1] if self.transformer_num_feature_levels > 1 else enc_channels[-1:]
        self.input_proj = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)) for c in chans])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)
        self.mask_dim = mask_dim
        self.mask_features = Conv2d(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        c2_xavier_fill(self.mask_features)
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        stride = min(self.transformer_feature_strides)
        self.num_fpn_levels = int(np.log2(stride) - np.log2(self.common_stride))
        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        for idx, in_channels in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lateral = Conv2d(in_channels, conv_dim, kernel_size=1, bias=use_bias, norm=get_norm(norm, conv_dim))
            output = Conv2d(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias,
                            norm=get_norm(norm, conv_dim), activation=F.relu)
            c2_xavier_fill(lateral)
            c2_xavier_fill(output)
            self.add_module("adapter_{}".format(idx + 1), lateral)
            self.add_module("layer_{}".format(idx + 1), output)
            lateral_convs.append(lateral)
            output_convs.append(output)
        self.lateral_convs = lateral_convs[::-1]  # top-down order
        self.output_convs = output_convs[::-1]

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        head = cfg.MODEL.SEM_SEG_HEAD
        return {
            "input_shape": {k: v for k, v in input_shape.items() if k in head.IN_FEATURES},
            "conv_dim": head.CONVS_DIM,
            "mask_dim": head.MASK_DIM,
            "norm": head.NORM,
            "transformer_dropout": cfg.MODEL.MASK_FORMER.DROPOUT,
            "transformer_nheads": cfg.MODEL.MASK_FORMER.NHEADS,
            "transformer_dim_feedforward": 1024,  # hard-wired in the reference (:305-306)
            "transformer_enc_layers": head.TRANSFORMER_ENC_LAYERS,
            "transformer_in_features": head.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES,
            "common_stride": head.COMMON_STRIDE,
        }

    def forward_features(self, features):
        """-> (mask_features [B,mask_dim,H/4,W/4], coarsest encoder map, [3 maps coarse to fine]).
        Always fp32, like the reference's @autocast(enabled=False) + .float() (:314,320)."""
        with torch.autocast("cuda", enabled=False), conv_precision():
            srcs, pos = [], []
            for idx, f in enumerate(self.transformer_in_features[::-1]):
                x = features[f].float()
                proj = self.input_proj[idx]
                srcs.append(proj[1](ops.conv1x1_layer(proj[0], x)))
                pos.append(self.pe_layer(x))
            y, shapes, level_start_index = self.transformer(srcs, pos)
            B = y.shape[0]
            sizes = [h * w for h, w in shapes]
            out = [z.transpose(1, 2).reshape(B, -1, *shapes[i]) for i, z in enumerate(torch.split(y, sizes, dim=1))]
            for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
                cur = ops.conv1x1_layer(self.lateral_convs[idx], features[f].float())
                if cur.is_cuda and not torch.is_grad_enabled():   # fused top-down step (one pass over the map)
                    merged = ops.upsample_add(out[-1].float(), cur)
                else:
                    merged = cur + F.interpolate(out[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
                out.append(ops.conv_layer(self.output_convs[idx], merged))
            multi_scale = out[:self.maskformer_num_feature_levels]
            return ops.conv1x1_layer(self.mask_features, out[-1]), out[0], multi_scale
