"""Import the reference's hot-path modules by file path, in THIS container only.

This is fixture-generation tooling, not product code: ``make_golden.py`` uses it to run the
reference's own PyTorch path on CPU and dump seeded input/output vectors into
``tests/golden/*.npz``. Nothing in ``tests/`` (at run time), ``bench.py`` or the package imports
this file; ``/root/reference`` does not exist on the GPU box.

The reference needs detectron2 + fvcore, which are not installed. Only a handful of symbols
are touched by the hot-path files, so they are stubbed here with their documented behaviour:

* ``detectron2.config.configurable``   - decorator; we always call ``__init__`` with explicit
  kwargs, so the stub is the identity (the ``from_config`` path is not exercised).
* ``detectron2.layers.Conv2d``         - ``nn.Conv2d`` + optional ``norm`` / ``activation``.
* ``detectron2.layers.get_norm``       - ``"GN" -> GroupNorm(32, C)``, ``"" -> None``.
* ``detectron2.layers.ShapeSpec``      - namedtuple(channels, height, width, stride).
* ``detectron2.utils.registry.Registry`` / ``detectron2.modeling.SEM_SEG_HEADS_REGISTRY``.
* ``fvcore.nn.weight_init.c2_xavier_fill`` - ``kaiming_uniform_(a=1)`` + zero bias.
* ``MultiScaleDeformableAttention``    - empty module, so ``MSDeformAttn.forward`` falls into the
  reference's own pure-PyTorch ``ms_deform_attn_core_pytorch``
  (ops/modules/ms_deform_attn.py:116-121).
"""
import collections
import importlib
import os
import sys
import types

import torch
from torch import nn
from torch.nn import functional as F

REF_ROOT = os.environ.get("MSM_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "MSMFormer", "meanshiftformer")


class _Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


class _Conv2d(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    if norm == "GN":
        return nn.GroupNorm(32, out_channels)
    raise ValueError(norm)


def _c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def _configurable(init_func=None, *, from_config=None):
    if init_func is not None:
        return init_func
    return lambda f: f


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if "refmsm" in sys.modules:
        return
    if not os.path.isdir(REF_PKG):
        raise FileNotFoundError(f"reference not found at {REF_PKG}")
    ShapeSpec = collections.namedtuple("ShapeSpec", ["channels", "height", "width", "stride"],
                                       defaults=[None, None, None, None])
    _mod("detectron2")
    _mod("detectron2.config", configurable=_configurable)
    _mod("detectron2.layers", Conv2d=_Conv2d, ShapeSpec=ShapeSpec, get_norm=_get_norm, DeformConv=None)
    _mod("detectron2.modeling", SEM_SEG_HEADS_REGISTRY=_Registry("SEM_SEG_HEADS"))
    _mod("detectron2.utils")
    _mod("detectron2.utils.registry", Registry=_Registry)
    _mod("fvcore")
    _mod("fvcore.nn")
    wi = _mod("fvcore.nn.weight_init", c2_xavier_fill=_c2_xavier_fill)
    sys.modules["fvcore.nn"].weight_init = wi
    _mod("MultiScaleDeformableAttention")

    # namespace-style parents whose __init__.py files are NOT executed (they pull in datasets,
    # swin, evaluators ...); relative imports inside the hot-path files resolve through __path__.
    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    # training-side imports of criterion.py / matcher.py: detectron2's PointRend helpers are third-party code absent
    # from /root/reference; their published algorithm (detectron2 0.6, point_rend/point_features.py) is restated here.
    # Random draws go through ``RAND`` so that the golden generator can record / replay them.
    def point_sample(input, point_coords, **kwargs):
        add_dim = point_coords.dim() == 3
        if add_dim:
            point_coords = point_coords.unsqueeze(2)
        out = F.grid_sample(input, 2.0 * point_coords - 1.0, **kwargs)
        return out.squeeze(3) if add_dim else out

    def get_uncertain_point_coords_with_randomness(coarse_logits, uncertainty_func, num_points, oversample_ratio,
                                                   importance_sample_ratio):
        num_boxes = coarse_logits.shape[0]
        num_sampled = int(num_points * oversample_ratio)
        point_coords = RAND[0](num_boxes, num_sampled, 2, device=coarse_logits.device)
        point_logits = point_sample(coarse_logits, point_coords, align_corners=False)
        point_uncertainties = uncertainty_func(point_logits)
        num_uncertain_points = int(importance_sample_ratio * num_points)
        num_random_points = num_points - num_uncertain_points
        idx = torch.topk(point_uncertainties[:, 0, :], k=num_uncertain_points, dim=1)[1]
        shift = num_sampled * torch.arange(num_boxes, dtype=torch.long, device=coarse_logits.device)
        idx += shift[:, None]
        point_coords = point_coords.view(-1, 2)[idx.view(-1), :].view(num_boxes, num_uncertain_points, 2)
        if num_random_points > 0:
            point_coords = torch.cat(
                [point_coords, RAND[0](num_boxes, num_random_points, 2, device=coarse_logits.device)], dim=1)
        return point_coords

    _mod("detectron2.utils.comm", get_world_size=lambda: 1)
    _mod("detectron2.projects")
    _mod("detectron2.projects.point_rend")
    _mod("detectron2.projects.point_rend.point_features", point_sample=point_sample,
         get_uncertain_point_coords_with_randomness=get_uncertain_point_coords_with_randomness)

    pkg("refmsm", REF_PKG)
    pkg("refmsm.utils", os.path.join(REF_PKG, "utils"))
    pkg("refmsm.modeling", os.path.join(REF_PKG, "modeling"))
    pkg("refmsm.modeling.transformer_decoder", os.path.join(REF_PKG, "modeling", "transformer_decoder"))
    pkg("refmsm.modeling.pixel_decoder", os.path.join(REF_PKG, "modeling", "pixel_decoder"))
    pkg("refmsm.modeling.pixel_decoder.ops", os.path.join(REF_PKG, "modeling", "pixel_decoder", "ops"))
    # ops/modules and ops/functions have harmless __init__.py files: let them import normally


RAND = [torch.rand]  # the random source of the restated PointRend helpers (replaced by the golden generator)


class TorchProxy:
    """``torch`` with ``rand`` replaced: assigned to a reference module's global ``torch`` to replay recorded draws."""

    def __init__(self, rand):
        self.rand = rand

    def __getattr__(self, name):
        return getattr(torch, name)


def ref(name):
    """ref('modeling.transformer_decoder.attention_util') -> reference module object."""
    install()
    return importlib.import_module("refmsm." + name)
