"""Oracle: mean-shift transformer decoder + prediction heads. Test infrastructure only.

Follows MSMFormer/meanshiftformer/modeling/transformer_decoder/meanshiftformer_transformer_decoder.py
(MeanShiftTransformerDecoder :343-695, PretrainedMeanShiftTransformerDecoder :697-1048 - the two
differ only in the number of feature levels, 3 vs 1) and position_encoding.py:29-52.

Weights come in as a ``state_dict``-style mapping with the reference's own key names, so the same
fixture drives the reference, this oracle and the CUDA-backed modules.
Only the configuration every UOIS YAML selects is restated: post-norm, mean-shift cross- and
self-attention, dropout 0 (configs/mixture_ResNet50.yaml:46-64, mixture_UCN.yaml:46-66).
"""
import math

import torch
import torch.nn.functional as F

from .vmf_attention import meanshift_attention


def position_embedding_sine(batch, height, width, num_pos_feats, temperature=10000.0, scale=2 * math.pi):
    """position_encoding.py:29-52 with mask=None, normalize=True. Returns [B, 2*npf, H, W]."""
    ones = torch.ones(batch, height, width, dtype=torch.float32)
    y_embed = ones.cumsum(1)
    x_embed = ones.cumsum(2)
    eps = 1e-6
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    px = x_embed[..., None] / dim_t
    py = y_embed[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def prediction_heads(sd, output, mask_features, target_size, num_heads, need_mask=True):
    """forward_prediction_heads, decoder.py:660-682 / :1012-1035.

    output [Q,B,C]; mask_features [B,Cm,h,w]. Returns (class logits [B,Q,K+1], mask logits
    [B,Q,h,w], blocked bool [B*heads,Q,Ht*Wt] or None).
    """
    dec = _ln(output, sd, "decoder_norm").transpose(0, 1)
    logits = F.linear(dec, sd["class_embed.weight"], sd["class_embed.bias"])
    e = dec
    for i in range(3):  # MLP :329-341
        e = F.linear(e, sd[f"mask_embed.layers.{i}.weight"], sd[f"mask_embed.layers.{i}.bias"])
        if i < 2:
            e = F.relu(e)
    masks = torch.einsum("bqc,bchw->bqhw", e, mask_features)
    blocked = None
    if need_mask:
        m = F.interpolate(masks, size=target_size, mode="bilinear", align_corners=False)
        blocked = (m.sigmoid().flatten(2).unsqueeze(1).repeat(1, num_heads, 1, 1).flatten(0, 1) < 0.5).bool()
    return logits, masks, blocked


def decoder_forward(sd, x, mask_features, *, num_heads, num_layers, decoder_block_norm=True,
                    disable_attention_mask=False, trace=None):
    """forward, decoder.py:540-658 / :894-1010.

    x: list of [B,Cin,hl,wl] (1 or 3 levels; their count selects the decoder variant);
    returns {'pred_logits','pred_masks','aux_outputs'} like the reference.
    ``trace`` (optional list) receives per-layer intermediates for teacher-forced kernel tests.
    """
    L = len(x)
    C = sd["query_feat.weight"].shape[1]
    has_proj = "input_proj.0.weight" in sd
    src, pos, sizes = [], [], []
    for i in range(L):
        B, _, h, w = x[i].shape
        sizes.append((h, w))
        pos.append(position_embedding_sine(B, h, w, C // 2).flatten(2).permute(2, 0, 1))
        s = F.conv2d(x[i], sd[f"input_proj.{i}.weight"], sd[f"input_proj.{i}.bias"]) if has_proj else x[i]
        s = s.flatten(2) + sd["level_embed.weight"][i][None, :, None]
        src.append(s.permute(2, 0, 1))
    B = src[0].shape[1]
    query_pos = sd["query_embed.weight"].unsqueeze(1).repeat(1, B, 1)
    out = sd["query_feat.weight"].unsqueeze(1).repeat(1, B, 1)

    all_logits, all_masks = [], []
    logits, masks, blocked = prediction_heads(sd, out, mask_features, sizes[0], num_heads, not disable_attention_mask)
    all_logits.append(logits)
    all_masks.append(masks)
    for i in range(num_layers):
        lvl = i % L
        if blocked is not None:  # :618 rows that block every key attend everywhere instead
            blocked[torch.where(blocked.sum(-1) == blocked.shape[-1])] = False
        if trace is not None:
            trace.append({"layer": i, "level": lvl, "tgt_in": out.clone(),
                          "blocked": None if blocked is None else blocked.clone()})
        p = f"transformer_cross_attention_layers.{i}."
        # MeanShiftCrossAttentionLayer.forward_post :245-260
        a, _ = meanshift_attention(out + query_pos, src[lvl] + pos[lvl], src[lvl],
                                   sd[p + "meanshift_attn.in_proj_weight"], sd[p + "meanshift_attn.in_proj_bias"],
                                   sd[p + "meanshift_attn.out_proj.weight"], sd[p + "meanshift_attn.out_proj.bias"],
                                   num_heads, blocked)
        out = _ln(out + a, sd, p + "norm")
        if trace is not None:
            trace[-1]["after_cross"] = out.clone()
        p = f"transformer_self_attention_layers.{i}."
        # MeanShiftSelfAttentionLayer.forward_post :171-181  (q = k = tgt + query_pos, value = tgt)
        qk = out + query_pos
        a, _ = meanshift_attention(qk, qk, out,
                                   sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                                   sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"],
                                   num_heads, None)
        out = _ln(out + a, sd, p + "norm")
        if trace is not None:
            trace[-1]["after_self"] = out.clone()
        p = f"transformer_ffn_layers.{i}."
        # FFNLayer.forward_post :300-304
        f = F.linear(F.relu(F.linear(out, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                     sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        out = _ln(out + f, sd, p + "norm")
        if decoder_block_norm:  # :637-638
            out = F.normalize(out, dim=-1)
        if trace is not None:
            trace[-1]["after_ffn"] = out.clone()
        logits, masks, blocked = prediction_heads(sd, out, mask_features, sizes[(i + 1) % L], num_heads,
                                                  not disable_attention_mask)
        all_logits.append(logits)
        all_masks.append(masks)
    return {"pred_logits": all_logits[-1], "pred_masks": all_masks[-1],
            "aux_outputs": [{"pred_logits": a, "pred_masks": b} for a, b in zip(all_logits[:-1], all_masks[:-1])]}
