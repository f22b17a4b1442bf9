"""``build_pixel_decoder`` and ``SimpleBasePixelDecoder`` - mirror of the reference's
modeling/pixel_decoder/fpn.py:21-33, 161-290. (``BasePixelDecoder`` / ``TransformerEncoderPixelDecoder``
are not selected by any UOIS config and are out of scope)
"""
import logging
from typing import Callable, Dict, Optional, Union

from torch import nn

from .... import ops
from ....d2compat import SEM_SEG_HEADS_REGISTRY, Conv2d, ShapeSpec, c2_xavier_fill, configurable
from ....precision import conv_precision


def build_pixel_decoder(cfg, input_shape):
    name = cfg.MODEL.SEM_SEG_HEAD.PIXEL_DECODER_NAME
    model = SEM_SEG_HEADS_REGISTRY.get(name)(cfg, input_shape)
    if not callable(getattr(model, "forward_features", None)):
        raise ValueError("Only SEM_SEG_HEADS with forward_features method can be used as pixel decoder. "
                         f"Please implement forward_features for {name} to only return mask features.")
    return model


@SEM_SEG_HEADS_REGISTRY.register()
class SimpleBasePixelDecoder(nn.Module):
    """UCN RGB-D configs: the (already unit-normalised) 64-d embedding map is the single
    multi-scale feature; mask features = 3x3 conv to ``mask_dim`` at full resolution."""

    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, conv_dim: int, mask_dim: int,
                 norm: Optional[Union[str, Callable]] = None):
        super().__init__()
        by_stride = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in by_stride]
        self.mask_dim = mask_dim
        if self.mask_dim != 64:  # reference :238-246
            self.mask_features = Conv2d(conv_dim, mask_dim, kernel_size=3, stride=1, padding=1)
            c2_xavier_fill(self.mask_features)
        self.maskformer_num_feature_levels = 1

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        head = cfg.MODEL.SEM_SEG_HEAD
        return {"input_shape": {k: v for k, v in input_shape.items() if k in head.IN_FEATURES},
                "conv_dim": head.CONVS_DIM, "mask_dim": head.MASK_DIM, "norm": head.NORM}

    def forward_features(self, features):
        multi_scale = []
        y = None
        for f in self.in_features[::-1]:
            y = features[f]
            if len(multi_scale) < self.maskformer_num_feature_levels:
                multi_scale.append(y)
        if self.mask_dim == 64:
            return y, None, multi_scale
        with conv_precision():
            return ops.conv_layer(self.mask_features, y.float()), None, multi_scale

    def forward(self, features, targets=None):
        logging.getLogger(__name__).warning("Calling forward() may cause unpredicted behavior of PixelDecoder module.")
        return self.forward_features(features)
