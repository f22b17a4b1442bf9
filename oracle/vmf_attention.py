"""Oracle: von-Mises-Fisher ("hypersphere") attention. Test infrastructure only (see oracle/__init__.py).

Follows MSMFormer/meanshiftformer/modeling/transformer_decoder/attention_util.py.
"""
import torch
import torch.nn.functional as F

KAPPA = 30.0  # attention_util.py:26


def hypersphere_attention(q, k, v, attn_mask=None, kappa=KAPPA):
    """attention_util.py:64-82.

    q [G,Nt,E], k/v [G,Ns,E], attn_mask additive float [G,Nt,Ns] (0 / -inf) or None.
    Returns (out [G,Nt,E], attn [G,Nt,Ns]); every step is the reference's op, in its order:
    unit-normalise q and k (eps 1e-12), cosine scores, times kappa, plus mask, softmax over keys,
    weighted sum of v, unit-normalise the result.
    """
    qn = F.normalize(q, p=2.0, dim=-1)
    kn = F.normalize(k, p=2.0, dim=-1)
    scores = torch.bmm(qn, kn.transpose(-2, -1))
    scores = kappa * scores
    if attn_mask is not None:
        scores = scores + attn_mask
    attn = F.softmax(scores, dim=-1)
    out = F.normalize(torch.bmm(attn, v), p=2.0, dim=-1)
    return out, attn


def in_projection(query, key, value, w, b):
    """attention_util.py:121-140 (ms_in_projection_packed): packed [3E,E] weight, q|k|v order.

    The branch is chosen by object identity exactly as the reference does, because the three
    branches call differently-shaped GEMMs (same math, possibly different rounding).
    """
    E = query.size(-1)
    if key is value:
        if query is key:
            return F.linear(query, w, b).chunk(3, dim=-1)
        wq, wkv = w.split([E, 2 * E])
        bq, bkv = b.split([E, 2 * E])
        return (F.linear(query, wq, bq),) + tuple(F.linear(key, wkv, bkv).chunk(2, dim=-1))
    wq, wk, wv = w.chunk(3)
    bq, bk, bv = b.chunk(3)
    return F.linear(query, wq, bq), F.linear(key, wk, bk), F.linear(value, wv, bv)


def meanshift_attention(query, key, value, in_w, in_b, out_w, out_b, num_heads, blocked=None, kappa=KAPPA):
    """attention_util.py:198-432 restricted to what MeanShiftAttention.forward (474-540) reaches:
    no bias_k/v, no zero-attn, no key_padding_mask, dropout 0, batch_first False.

    query [L,N,E], key/value [S,N,E]; ``blocked`` bool [N*h, L, S] (True = may not attend) or None.
    Returns (out [L,N,E], head-averaged attention [N,L,S]).
    """
    L, N, E = query.shape
    S = key.shape[0]
    hd = E // num_heads
    q, k, v = in_projection(query, key, value, in_w, in_b)
    fmask = None
    if blocked is not None:  # :411-414 bool -> additive -inf
        fmask = torch.zeros(blocked.shape, dtype=torch.float32)
        fmask.masked_fill_(blocked, float("-inf"))
    # :364-375 split heads, batch-first
    q = q.contiguous().view(L, N * num_heads, hd).transpose(0, 1)
    k = k.contiguous().view(S, N * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(S, N * num_heads, hd).transpose(0, 1)
    o, attn = hypersphere_attention(q, k, v, fmask, kappa)
    o = o.transpose(0, 1).contiguous().view(L, N, E)
    o = F.linear(o, out_w, out_b)  # :425
    return o, attn.view(N, num_heads, L, S).sum(dim=1) / num_heads  # :427-430


def hypersphere_attention_backward(q, k, v, attn_mask, kappa, grad_out):
    """Gradient of hypersphere_attention (attention_util.py:64-82) wrt q, k, v, written out by hand (no autograd):
    the statement the CUDA backward (csrc/vmf_attention_bwd.cu) follows. Pinned on torch.autograd run through the
    reference's own function (tests/golden/hypersphere_attention_bwd.npz).

    With qn = unit(q), kn = unit(k), s = kappa qn kn^T + mask, p = softmax(s), o = p v, out = unit(o):
      g_o  = (grad_out - out <out, grad_out>) / |o|            (F.normalize, norm above its eps)
      g_v  = p^T g_o
      g_s  = p * (g_o v^T - <g_o, o>)                           (softmax; <g_o, o> = 0 up to rounding as g_o _|_ out)
      g_qn = kappa g_s kn,   g_kn = kappa g_s^T qn
      g_q  = (g_qn - qn <qn, g_qn>) / |q|,   g_k likewise.
    Returns (g_q, g_k, g_v)."""
    qnorm = q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    knorm = k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    qn, kn = q / qnorm, k / knorm
    s = kappa * torch.bmm(qn, kn.transpose(-2, -1))
    if attn_mask is not None:
        s = s + attn_mask
    p = F.softmax(s, dim=-1)
    o = torch.bmm(p, v)
    onorm = o.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    out = o / onorm
    g_o = (grad_out - out * (out * grad_out).sum(-1, keepdim=True)) / onorm
    g_v = torch.bmm(p.transpose(-2, -1), g_o)
    delta = (g_o * o).sum(-1, keepdim=True)
    g_s = p * (torch.bmm(g_o, v.transpose(-2, -1)) - delta)
    g_qn = kappa * torch.bmm(g_s, kn)
    g_kn = kappa * torch.bmm(g_s.transpose(-2, -1), qn)
    g_q = (g_qn - qn * (qn * g_qn).sum(-1, keepdim=True)) / qnorm
    g_k = (g_kn - kn * (kn * g_kn).sum(-1, keepdim=True)) / knorm
    return g_q, g_k, g_v
