"""Hypersphere (vMF) attention - mirror of the reference's
modeling/transformer_decoder/attention_util.py with the math in CUDA (csrc/vmf_attention*.cu).

Same public names and signatures: ``KAPPA``, ``hypersphere_attention``,
``hypersphere_attention_forward``, ``MeanShiftAttention`` (an ``nn.MultiheadAttention`` subclass
with identical parameter names/shapes: in_proj_weight [3E,E], in_proj_bias [3E], out_proj.*).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor, nn
from torch.nn import functional as F

from .... import ops

# Hyperparameter of the vMF kernel (reference attention_util.py:26)
KAPPA = 30


def hypersphere_attention(q: Tensor, k: Tensor, v: Tensor, attn_mask: Optional[Tensor] = None,
                          dropout_p: float = 0.0, kappa: float = KAPPA,
                          need_weights: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """Reference attention_util.py:30-82. q [G,Nt,E], k/v [G,Ns,E], attn_mask additive float
    [G,Nt,Ns] or [Nt,Ns]; returns (out [G,Nt,E] unit rows, attn [G,Nt,Ns]).

    ``need_weights=False`` (extension) skips the [G,Nt,Ns] weight matrix, which the decoder
    layers discard; the default keeps the reference's return value."""
    if dropout_p > 0.0:
        raise NotImplementedError("dropout on attention weights: every MSMFormer config uses 0.0")
    G, Nt, E = q.shape
    Ns = k.shape[1]
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            attn_mask = torch.zeros(attn_mask.shape, dtype=torch.float32, device=q.device).masked_fill_(
                attn_mask, float("-inf"))
        if attn_mask.dim() == 2:
            attn_mask = attn_mask.unsqueeze(0).expand(G, Nt, Ns)
        attn_mask = attn_mask.contiguous()
    q4, k4, v4 = q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)  # [G,1,L,E] views: batch=G, heads=1
    out = torch.empty(G, Nt, E, device=q.device, dtype=torch.float32)
    res = ops.vmf_attention(q4, k4, v4, add_mask=attn_mask, kappa=kappa, out=out.unsqueeze(1),
                            return_den=need_weights)
    if not need_weights:
        return out, None
    _, den = res
    attn = ops.vmf_attention_weights(q4, k4, den, add_mask=attn_mask, kappa=kappa)
    return out, attn


def ms_in_projection_packed(q: Tensor, k: Tensor, v: Tensor, w: Tensor, b: Optional[Tensor] = None):
    """Reference attention_util.py:84-140: packed [3E,E] projection, q|k|v order."""
    E = q.size(-1)
    if k is v:
        if q is k:
            return ops.dense(q, w, b).chunk(3, dim=-1)
        w_q, w_kv = w.split([E, E * 2])
        b_q, b_kv = (None, None) if b is None else b.split([E, E * 2])
        return (ops.dense(q, w_q, b_q),) + ops.dense(k, w_kv, b_kv).chunk(2, dim=-1)
    w_q, w_k, w_v = w.chunk(3)
    b_q, b_k, b_v = (None, None, None) if b is None else b.chunk(3)
    return ops.dense(q, w_q, b_q), ops.dense(k, w_k, b_k), ops.dense(v, w_v, b_v)


def hypersphere_attention_forward(query: Tensor, key: Tensor, value: Tensor, embed_dim_to_check: int, num_heads: int,
                                  in_proj_weight: Tensor, in_proj_bias: Optional[Tensor],
                                  bias_k: Optional[Tensor], bias_v: Optional[Tensor], add_zero_attn: bool,
                                  dropout_p: float, out_proj_weight: Tensor, out_proj_bias: Optional[Tensor],
                                  training: bool = True, key_padding_mask: Optional[Tensor] = None,
                                  need_weights: bool = True, attn_mask: Optional[Tensor] = None,
                                  **unsupported) -> Tuple[Tensor, Optional[Tensor]]:
    """Reference attention_util.py:198-432 for the arguments MeanShiftAttention passes.

    query [L,N,E], key/value [S,N,E] (seq-first); attn_mask bool/float [N*h,L,S] or [L,S].
    The projected q/k/v stay in their [len, N, E] buffers: the kernel addresses head h of batch n
    through strides, so the reference's head-split transposes (:364-375) are not materialised."""
    if bias_k is not None or bias_v is not None or add_zero_attn or key_padding_mask is not None:
        raise NotImplementedError("bias_k/bias_v/add_zero_attn/key_padding_mask are never used by MSMFormer")
    if any(v is not None and v is not False for v in unsupported.values()):
        raise NotImplementedError(f"unsupported arguments: {sorted(unsupported)}")
    L, N, E = query.shape
    S = key.shape[0]
    assert E == embed_dim_to_check, f"was expecting embedding dimension of {embed_dim_to_check}, but got {E}"
    hd = E // num_heads
    assert hd * num_heads == E, f"embed_dim {E} not divisible by num_heads {num_heads}"
    assert key.shape == value.shape, f"key shape {key.shape} does not match value shape {value.shape}"
    q, k, v = ms_in_projection_packed(query, key, value, in_proj_weight, in_proj_bias)
    add_mask = None
    if attn_mask is not None:
        if attn_mask.dim() == 2:
            if attn_mask.shape != (L, S):
                raise RuntimeError(f"The shape of the 2D attn_mask is {attn_mask.shape}, but should be {(L, S)}.")
            attn_mask = attn_mask.unsqueeze(0).expand(N * num_heads, L, S)
        elif attn_mask.shape != (N * num_heads, L, S):
            raise RuntimeError(
                f"The shape of the 3D attn_mask is {attn_mask.shape}, but should be {(N * num_heads, L, S)}.")
        if attn_mask.dtype == torch.bool:
            add_mask = torch.zeros(attn_mask.shape, dtype=torch.float32, device=q.device).masked_fill_(
                attn_mask, float("-inf"))
        else:
            add_mask = attn_mask.float().contiguous()

    def heads_view(t, length):  # [len, N, E] -> [N, h, len, hd] view
        return t.contiguous().view(length, N, num_heads, hd).permute(1, 2, 0, 3)

    q4, k4, v4 = heads_view(q, L), heads_view(k, S), heads_view(v, S)
    o = torch.empty(L, N, E, device=q.device, dtype=torch.float32)
    res = ops.vmf_attention(q4, k4, v4, add_mask=add_mask, kappa=KAPPA, out=heads_view(o, L), return_den=need_weights)
    attn_output = ops.dense(o, out_proj_weight, out_proj_bias)
    if not need_weights:
        return attn_output, None
    _, den = res
    w = ops.vmf_attention_weights(q4, k4, den, add_mask=add_mask, kappa=KAPPA)
    return attn_output, w.view(N, num_heads, L, S).sum(dim=1) / num_heads


class MeanShiftAttention(nn.MultiheadAttention):
    """Reference attention_util.py:434-540. The constructor, like the reference's (:469-472),
    forwards only (embed_dim, num_heads): dropout is 0, bias True, no kdim/vdim, seq-first."""

    def __init__(self, embed_dim, num_heads=1, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False,
                 kdim=None, vdim=None, batch_first=False, device=None, dtype=None) -> None:
        super().__init__(embed_dim, num_heads)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, key_padding_mask: Optional[Tensor] = None,
                need_weights: bool = True, attn_mask: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
        return hypersphere_attention_forward(
            query, key, value, self.embed_dim, self.num_heads, self.in_proj_weight, self.in_proj_bias,
            self.bias_k, self.bias_v, self.add_zero_attn, self.dropout, self.out_proj.weight, self.out_proj.bias,
            training=self.training, key_padding_mask=key_padding_mask, need_weights=need_weights,
            attn_mask=attn_mask)
