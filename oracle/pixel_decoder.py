"""Oracle: multi-scale deformable attention and the two pixel decoders. Test infrastructure only.

Follows MSMFormer/meanshiftformer/modeling/pixel_decoder/{msdeformattn.py, fpn.py,
ops/modules/ms_deform_attn.py, ops/src/cuda/ms_deform_im2col_cuda.cuh}.
"""
import torch
import torch.nn.functional as F

from .decoder import position_embedding_sine


def ms_deform_attn_core(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """The op itself, restated from the CUDA forward kernel (ms_deform_im2col_cuda.cuh:242-304 and
    its bilinear helper :38-89) as dense gathers - NOT via grid_sample, so that it is an
    independent statement of the same arithmetic:

        out[b,q,m,:] = sum_{l,p} A[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p])
        pixel coords  h = loc_y*H - 0.5, w = loc_x*W - 0.5; a point contributes only if
        -1 < h < H and -1 < w < W; taps outside the map read as 0.

    value [N,S,M,D]; spatial_shapes int64 [L,2] (H,W); level_start_index int64 [L];
    sampling_locations [N,Lq,M,L,P,2] (x,y) in [0,1]; attention_weights [N,Lq,M,L,P].
    Returns [N,Lq,M*D].
    """
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = torch.zeros(N, Lq, M, D, dtype=value.dtype)
    for l in range(L):
        H, W = int(spatial_shapes[l, 0]), int(spatial_shapes[l, 1])
        start = int(level_start_index[l])
        v = value[:, start:start + H * W]  # [N,HW,M,D]
        loc = sampling_locations[:, :, :, l]  # [N,Lq,M,P,2]
        w_im = loc[..., 0] * W - 0.5
        h_im = loc[..., 1] * H - 0.5
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h0 = torch.floor(h_im)
        w0 = torch.floor(w_im)
        lh, lw = h_im - h0, w_im - w0
        h0, w0 = h0.long(), w0.long()
        acc = torch.zeros(N, Lq, M, P, D, dtype=value.dtype)
        for dh, dw, wt in ((0, 0, (1 - lh) * (1 - lw)), (0, 1, (1 - lh) * lw),
                           (1, 0, lh * (1 - lw)), (1, 1, lh * lw)):
            hh, ww = h0 + dh, w0 + dw
            ok = inside & (hh >= 0) & (hh <= H - 1) & (ww >= 0) & (ww <= W - 1)
            idx = (hh.clamp(0, H - 1) * W + ww.clamp(0, W - 1))  # [N,Lq,M,P]
            # gather value[n, idx, m, :]
            vm = v.permute(0, 2, 1, 3)  # [N,M,HW,D]
            g = torch.gather(vm, 2, idx.permute(0, 2, 1, 3).reshape(N, M, Lq * P, 1).expand(-1, -1, -1, D))
            g = g.view(N, M, Lq, P, D).permute(0, 2, 1, 3, 4)  # [N,Lq,M,P,D]
            acc = acc + g * (wt * ok.to(value.dtype))[..., None]
        out = out + (acc * attention_weights[:, :, :, l, :, None]).sum(3)
    return out.view(N, Lq, M * D)


def ms_deform_attn_module(sd, prefix, query, reference_points, input_flatten, spatial_shapes, level_start_index,
                          n_heads, n_levels, n_points):
    """MSDeformAttn.forward, ops/modules/ms_deform_attn.py:82-125 (2-d reference points, no padding mask)."""
    N, Lq, C = query.shape
    S = input_flatten.shape[1]
    value = F.linear(input_flatten, sd[prefix + "value_proj.weight"], sd[prefix + "value_proj.bias"])
    value = value.view(N, S, n_heads, C // n_heads)
    off = F.linear(query, sd[prefix + "sampling_offsets.weight"], sd[prefix + "sampling_offsets.bias"])
    off = off.view(N, Lq, n_heads, n_levels, n_points, 2)
    aw = F.linear(query, sd[prefix + "attention_weights.weight"], sd[prefix + "attention_weights.bias"])
    aw = F.softmax(aw.view(N, Lq, n_heads, n_levels * n_points), -1).view(N, Lq, n_heads, n_levels, n_points)
    normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)  # (W,H) :106
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = ms_deform_attn_core(value, spatial_shapes, level_start_index, loc, aw)
    return F.linear(out, sd[prefix + "output_proj.weight"], sd[prefix + "output_proj.bias"])


def encoder_reference_points(spatial_shapes, batch):
    """MSDeformAttnTransformerEncoder.get_reference_points, msdeformattn.py:140-153, with all
    valid ratios equal to 1 (the masks are all-False, :62). Returns [B, sum(HW), L, 2] (x,y)."""
    pts = []
    for H, W in spatial_shapes.tolist():
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
        pts.append(torch.stack((rx.reshape(-1) / W, ry.reshape(-1) / H), -1))
    ref = torch.cat(pts, 0)[None].expand(batch, -1, -1)
    return ref[:, :, None, :].expand(-1, -1, len(spatial_shapes), -1).contiguous()


def _gn_conv(x, sd, p, padding=0, relu=False):
    """detectron2 Conv2d wrapper with norm='GN' (GroupNorm(32,C)) and no conv bias."""
    y = F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), padding=padding)
    y = F.group_norm(y, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    return F.relu(y) if relu else y


def msdeform_pixel_decoder_forward(sd, features, *, n_heads, n_points=4, enc_layers, in_features=("res3", "res4", "res5"),
                                   fpn_features=("res2",)):
    """MSDeformAttnPixelDecoder.forward_features, msdeformattn.py:314-358 (+ the encoder :61-89,
    :122-131, :155-161). ``features``: dict res2..res5 of [B,C,H,W].
    Returns (mask_features, out[0], [3 multi-scale maps, coarse to fine])."""
    srcs, pos = [], []
    for idx, f in enumerate(in_features[::-1]):  # res5, res4, res3
        x = features[f].float()
        y = F.conv2d(x, sd[f"input_proj.{idx}.0.weight"], sd[f"input_proj.{idx}.0.bias"])
        y = F.group_norm(y, 32, sd[f"input_proj.{idx}.1.weight"], sd[f"input_proj.{idx}.1.bias"], 1e-5)
        srcs.append(y)
        pos.append(position_embedding_sine(x.shape[0], x.shape[2], x.shape[3], y.shape[1] // 2))
    B, C = srcs[0].shape[:2]
    L = len(srcs)
    shapes = torch.as_tensor([s.shape[-2:] for s in srcs], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    lvl_pos = torch.cat([p.flatten(2).transpose(1, 2) + sd["transformer.level_embed"][l].view(1, 1, -1)
                         for l, p in enumerate(pos)], 1)
    refp = encoder_reference_points(shapes, B)
    out = src
    for i in range(enc_layers):  # MSDeformAttnTransformerEncoderLayer.forward :122-131
        p = f"transformer.encoder.layers.{i}."
        a = ms_deform_attn_module(sd, p + "self_attn.", out + lvl_pos, refp, out, shapes, lsi, n_heads, L, n_points)
        out = F.layer_norm(out + a, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        f = F.linear(F.relu(F.linear(out, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                     sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        out = F.layer_norm(out + f, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    maps = []
    for l in range(L):
        H, W = shapes[l].tolist()
        z = out[:, int(lsi[l]):int(lsi[l]) + H * W]
        maps.append(z.transpose(1, 2).reshape(B, C, H, W))
    for idx, f in enumerate(fpn_features[::-1]):  # extra FPN levels :343-351
        n = len(fpn_features) - idx
        cur = _gn_conv(features[f].float(), sd, f"adapter_{n}")
        y = cur + F.interpolate(maps[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
        maps.append(_gn_conv(y, sd, f"layer_{n}", padding=1, relu=True))
    mask_features = F.conv2d(maps[-1], sd["mask_features.weight"], sd["mask_features.bias"])
    return mask_features, maps[0], maps[:3]


def simple_pixel_decoder_forward(sd, features, in_feature="res5"):
    """SimpleBasePixelDecoder.forward_features, fpn.py:261-284: the multi-scale feature is the
    input itself; mask features = 3x3 conv (only present when mask_dim != 64, :238-246)."""
    x = features[in_feature]
    if "mask_features.weight" in sd:
        return F.conv2d(x, sd["mask_features.weight"], sd["mask_features.bias"], padding=1), None, [x]
    return x, None, [x]
