"""Helpers the pixel decoder imports from the reference's modeling/transformer_decoder/transformer.py:357-369."""
import copy

from torch import nn
from torch.nn import functional as F


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


def _get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {activation}.")
