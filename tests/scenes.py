"""Synthetic table-top scenes shared by the golden generator and the parity tests."""
import torch


def two_stage_scene(seed, H=96, W=128, objects=5, with_depth=True):
    """a synthetic table-top label map: ellipses of different sizes (one of them mostly without valid depth),
    an rgb image, a depth image [x, y, z] with holes."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    labels = torch.zeros(1, H, W)
    for k in range(objects):
        cy, cx = 12 + torch.rand(1, generator=g) * (H - 24), 12 + torch.rand(1, generator=g) * (W - 24)
        ry, rx = 5 + torch.rand(1, generator=g) * 12, 5 + torch.rand(1, generator=g) * 16
        labels[0][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = k + 2      # combine_masks numbers from 2
    rgb = torch.rand(1, 3, H, W, generator=g)
    depth = None
    if with_depth:
        z = 0.6 + 0.4 * torch.rand(H, W, generator=g) + 0.05 * labels[0]
        z[torch.rand(H, W, generator=g) < 0.15] = 0                               # sensor holes
        z[(labels[0] == 3) & (torch.rand(H, W, generator=g) < 0.7)] = 0           # object 3: mostly invalid depth
        depth = torch.stack([xx / W * z, yy / H * z, z])[None]
    return rgb, labels, depth


def two_stage_crop_labels(mask_crops, seed):
    """stand-in for the crop network's output: the initial mask split into two local objects plus a spurious
    blob outside of it (rejected by the overlap test), background 0."""
    g = torch.Generator().manual_seed(seed)
    num, S, _ = mask_crops.shape
    out = torch.zeros(num, S, S)
    cols = torch.arange(S)[None, :].expand(S, S)
    for i in range(num):
        split = int(S * (0.35 + 0.3 * torch.rand(1, generator=g)))
        out[i][(mask_crops[i] > 0) & (cols < split)] = 1
        out[i][(mask_crops[i] > 0) & (cols >= split)] = 2
        out[i][:S // 8, :S // 8] = 3
    return out


def probe_loss(out):
    """Deterministic scalar of every prediction of a decoder output (final + aux), with closed-form weights so that
    the generator and the tests build the same loss without sharing random state."""
    total = 0.0
    preds = out["aux_outputs"] + [{"pred_logits": out["pred_logits"], "pred_masks": out["pred_masks"]}]
    for i, p in enumerate(preds):
        for j, t in enumerate((p["pred_logits"], p["pred_masks"])):
            w = torch.cos(torch.arange(t.numel(), dtype=torch.float32, device=t.device) * 0.37 + 0.11 * i + 0.5 * j)
            total = total + (t.reshape(-1) * w.to(t.dtype)).sum() / t.numel() ** 0.5
    return total

