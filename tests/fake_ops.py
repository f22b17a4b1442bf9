"""CPU stand-ins for the device kernels of csrc/two_stage.cu, used ONLY to exercise the host-side decision logic of
unseenobjectswithmeanshift_b200/fcn/test_dataset.py without a GPU (the kernels themselves are checked by the
``-m gpu`` tests). Each function follows the C ABI contract in include/msmformer_b200.h, not the reference."""
import numpy as np
import torch
import torch.nn.functional as F


def label_stats(labels, depth=None, num_ids=None):
    N, H, W = labels.shape
    L = int(num_ids) if num_ids is not None else int(labels.max().item()) + 1
    stats = torch.zeros(N, L, 6, dtype=torch.int32)
    for n in range(N):
        for l in torch.unique(labels[n]).long().tolist():
            if not 0 <= l < L:
                continue
            ys, xs = torch.nonzero(labels[n] == l, as_tuple=True)
            valid = int((depth[n, 2][labels[n] == l] > 0).sum()) if depth is not None else 0
            stats[n, l] = torch.tensor([len(ys), valid, W - int(xs.min()), H - int(ys.min()), int(xs.max()) + 1,
                                        int(ys.max()) + 1], dtype=torch.int32)
    return stats


def relabel_lut(labels, lut, lo=0):
    out = labels.clone()
    L = lut.shape[1]
    for n in range(labels.shape[0]):
        idx = labels[n].long() - lo
        ok = (idx >= 0) & (idx < L)
        out[n][ok] = lut[n][idx[ok]]
    return out


def crop_resize(rgb, depth, labels, rois, ids, crop_size):
    S = int(crop_size)
    num = rois.shape[0]
    rgb_crops, mask_crops = torch.zeros(num, 3, S, S), torch.zeros(num, S, S)
    depth_crops = torch.zeros(num, 3, S, S) if depth is not None else None
    for c in range(num):
        x0, y0, x1, y1 = (int(v) for v in rois[c])
        win = (slice(y0, y1 + 1), slice(x0, x1 + 1))
        rgb_crops[c] = F.interpolate(rgb[(slice(None),) + win][None], size=(S, S), mode="bilinear", align_corners=True)[0]
        if depth is not None:
            depth_crops[c] = F.interpolate(depth[(slice(None),) + win][None], size=(S, S), mode="bilinear",
                                           align_corners=True)[0]
        mask_crops[c] = F.interpolate((labels[win] == ids[c]).float()[None, None], size=(S, S), mode="nearest")[0, 0]
    return rgb_crops, depth_crops, mask_crops


def crop_label_stats(labels_crop, init_crop, depth_crop=None, num_ids=None):
    num = labels_crop.shape[0]
    L = int(num_ids) if num_ids is not None else int(labels_crop.max().item()) + 1
    stats = torch.zeros(num, L, 4, dtype=torch.int32)
    dsum = torch.zeros(num, L, dtype=torch.float64)
    for c in range(num):
        for l in range(L):
            where = labels_crop[c] == l
            stats[c, l, 0] = int(where.sum())
            stats[c, l, 1] = int((where & (init_crop[c] != 0)).sum())
            if depth_crop is not None:
                z = depth_crop[c, 2][where]
                stats[c, l, 2] = int((z > 0).sum())
                dsum[c, l] = z[z > 0].double().sum()
    return stats, dsum


def paste_crops(labels_crop, new_label, order, rois, height, width):
    refined = torch.zeros(int(height), int(width))
    S = labels_crop.shape[-1]
    for c in order.tolist():
        x0, y0, x1, y1 = (int(v) for v in rois[c])
        lut = new_label[c]
        idx = labels_crop[c].long()
        ok = (idx >= 0) & (idx < lut.shape[0])
        mapped = torch.zeros(S, S)
        mapped[ok] = lut[idx[ok]]
        back = F.interpolate(mapped[None, None], size=(y1 - y0 + 1, x1 - x0 + 1), mode="nearest")[0, 0]
        view = refined[y0:y1 + 1, x0:x1 + 1]
        view[back != 0] = back[back != 0]
    return refined


# ---------------------------------------------------------------------------------------------------------------
# Stand-ins for the attention / mask-head entry points (contract of msm_vmf_attention_fwd / _bwd, msm_mask_logits,
# msm_mask_to_attn_bits in include/msmformer_b200.h), used to exercise the AUTOGRAD WIRING of ops.py and of the
# decoder's training path on CPU. The backward stand-in is written from the saved (den, |o|) planes like the kernel,
# not with torch.autograd.
def _unpack_bits(bits, row_open, Ns):
    B, Q, words = bits.shape
    sh = torch.arange(32, dtype=torch.int64)
    blocked = ((bits.long().unsqueeze(-1) >> sh) & 1).bool().reshape(B, Q, words * 32)[..., :Ns]
    if row_open is not None:
        blocked = blocked & (row_open != 0).unsqueeze(-1)
    return blocked


def _weights(q, k, blocked_bits, row_open, add_mask, kappa, normalize_q, normalize_k):
    B, H, Nq, _ = q.shape
    Ns = k.shape[2]
    qn = F.normalize(q, dim=-1, eps=1e-12) if normalize_q else q
    kn = F.normalize(k, dim=-1, eps=1e-12) if normalize_k else k
    e = kappa * (qn @ kn.transpose(-1, -2)) - kappa
    if add_mask is not None:
        e = e + add_mask.view(B, H, Nq, Ns)
    w = torch.exp(e)
    if blocked_bits is not None:
        w = w.masked_fill(_unpack_bits(blocked_bits, row_open, Ns).unsqueeze(1), 0.0)
    return qn, kn, w


def _bhld_buffer(B, H, L, hd, dtype):
    return torch.empty(B, L, H, hd, dtype=dtype).permute(0, 2, 1, 3)


def vmf_attention(q, k, v, *, blocked_bits=None, row_open=None, add_mask=None, kappa=30.0, normalize_q=True,
                  normalize_k=True, out=None, return_den=False, save_norm=False):
    B, H, Nq, hd = q.shape
    _, _, w = _weights(q, k, blocked_bits, row_open, add_mask, kappa, normalize_q, normalize_k)
    den = w.sum(-1)
    o = (w @ v) / den.unsqueeze(-1)
    norm = o.norm(dim=-1)
    if out is None:
        out = _bhld_buffer(B, H, Nq, hd, q.dtype)
    out.copy_(o / norm.clamp_min(1e-12).unsqueeze(-1))
    if not return_den:
        return out
    den = den.reshape(B * H, Nq)
    return out, (torch.stack([den, norm.reshape(B * H, Nq)]) if save_norm else den)


def vmf_attention_bwd(q, k, v, out, grad_out, den, *, blocked_bits=None, row_open=None, add_mask=None, kappa=30.0,
                      normalize_q=True, normalize_k=True):
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    assert tuple(den.shape) == (2, B * H, Nq)
    qn, kn, w = _weights(q, k, blocked_bits, row_open, add_mask, kappa, normalize_q, normalize_k)
    p = w / den[0].view(B, H, Nq, 1)
    onorm = den[1].view(B, H, Nq, 1).clamp_min(1e-12)
    g_o = (grad_out - out * (out * grad_out).sum(-1, keepdim=True)) / onorm
    delta = (g_o * out).sum(-1, keepdim=True) * onorm
    g_v = p.transpose(-1, -2) @ g_o
    g_s = kappa * p * (g_o @ v.transpose(-1, -2) - delta)
    g_q, g_k = g_s @ kn, g_s.transpose(-1, -2) @ qn
    if normalize_q:
        g_q = (g_q - qn * (qn * g_q).sum(-1, keepdim=True)) / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    if normalize_k:
        g_k = (g_k - kn * (kn * g_k).sum(-1, keepdim=True)) / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    res = []
    for g, L in ((g_q, Nq), (g_k, Ns), (g_v, Ns)):
        buf = _bhld_buffer(B, H, L, hd, q.dtype)
        buf.copy_(g)
        res.append(buf)
    return tuple(res)


def mask_logits(mask_embed, mask_features, out=None):
    res = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)
    if out is not None:
        out.copy_(res)
        return out
    return res


def mask_to_attn_bits(masks, target_size):
    """msm_mask_to_attn_bits contract: bit set = key blocked (sigmoid < 0.5 after the bilinear resize), row_open = 0
    for rows that block every key."""
    B, Q = masks.shape[:2]
    m = F.interpolate(masks, size=tuple(int(v) for v in target_size), mode="bilinear", align_corners=False)
    blocked = m.sigmoid().flatten(2) < 0.5
    S = blocked.shape[-1]
    words = (S + 31) // 32
    pad = torch.zeros(B, Q, words * 32, dtype=torch.bool)
    pad[..., :S] = blocked
    v = (pad.view(B, Q, words, 32).long() << torch.arange(32)).sum(-1)
    bits = torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()
    return bits, (~blocked).any(-1).to(torch.int32).contiguous()


def dense(x, weight, bias=None, relu=False):
    y = F.linear(x, weight, bias)
    return torch.relu(y) if relu else y


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=128):
    """msm_ms_deform_attn_fwd contract through the oracle's pure-torch core: [N, Lq, M*D]."""
    from oracle import pixel_decoder as opd
    return opd.ms_deform_attn_core(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=128):
    """msm_ms_deform_attn_bwd contract: (grad_value, grad_sampling_loc, grad_attn_weight)."""
    from oracle import pixel_decoder as opd
    with torch.enable_grad():
        v, l, a = (t.detach().clone().requires_grad_() for t in (value, sampling_loc, attn_weight))
        out = opd.ms_deform_attn_core(v, spatial_shapes, level_start_index, l, a)
        return torch.autograd.grad(out, (v, l, a), grad_output)
