"""The few detectron2 / fvcore symbols the hot-path modules touch.

When detectron2 is importable its own objects are used, so the classes below register into the
SAME registries the reference's ``build_model`` / ``build_sem_seg_head`` consult and the package
is a drop-in. When it is not (this image), bundled equivalents with the same behaviour are used
so the modules, ``from_config`` included, still work stand-alone.
"""
import collections
import functools
import inspect

import torch
from torch import nn
from torch.nn import functional as F

try:  # pragma: no cover - detectron2 is not installed in the build image
    from detectron2.config import configurable
    from detectron2.layers import Conv2d, ShapeSpec, get_norm
    from detectron2.modeling import META_ARCH_REGISTRY, SEM_SEG_HEADS_REGISTRY
    from detectron2.utils.registry import Registry
    HAVE_DETECTRON2 = True
except ImportError:
    HAVE_DETECTRON2 = False

    class Registry(dict):
        """name -> class, with detectron2's ``register`` / ``get`` surface."""

        def __init__(self, name):
            super().__init__()
            self._name = name

        def register(self, obj=None):
            if obj is None:
                return lambda o: self.register(o)
            if obj.__name__ in self:
                raise KeyError(f"{obj.__name__} already registered in {self._name}")
            self[obj.__name__] = obj
            return obj

        def get(self, name):
            if name not in self:
                raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
            return self[name]

    META_ARCH_REGISTRY = Registry("META_ARCH")
    SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")

    ShapeSpec = collections.namedtuple("ShapeSpec", ["channels", "height", "width", "stride"],
                                       defaults=[None, None, None, None])

    def _is_cfg(x):
        return hasattr(x, "MODEL") and not isinstance(x, (torch.Tensor, nn.Module))

    def configurable(init_func=None, *, from_config=None):
        """``@configurable`` for ``__init__``: ``Cls(cfg, ...)`` is routed through ``Cls.from_config``."""
        assert init_func is not None and inspect.isfunction(init_func)

        @functools.wraps(init_func)
        def wrapped(self, *args, **kwargs):
            first = args[0] if args else kwargs.get("cfg")
            if _is_cfg(first):
                explicit = type(self).from_config(*args, **kwargs)
                init_func(self, **explicit)
            else:
                init_func(self, *args, **kwargs)

        return wrapped

    class Conv2d(nn.Conv2d):
        """nn.Conv2d with optional ``norm`` and ``activation`` (detectron2.layers.Conv2d)."""

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

    def get_norm(norm, out_channels):
        if norm is None or norm == "":
            return None
        if callable(norm):
            return norm(out_channels)
        if norm == "GN":
            return nn.GroupNorm(32, out_channels)
        if norm == "BN":
            return nn.BatchNorm2d(out_channels)
        raise ValueError(f"norm '{norm}' needs detectron2")


def c2_xavier_fill(module):
    """fvcore.nn.weight_init.c2_xavier_fill."""
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)
