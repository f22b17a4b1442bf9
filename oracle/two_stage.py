"""Oracle: two-stage ("zoom-in") glue. Test infrastructure only.

CPU restatement of lib/fcn/test_dataset.py:62-198 (crop_rois, match_label_crop, filter_labels_depth) with
mask_to_tight_box from lib/utils/mask.py:180-187. cfg.TRAIN.SYN_CROP_SIZE (lib/fcn/config.py: 224) is the
``crop_size`` argument. ``F.upsample_bilinear`` = interpolate(bilinear, align_corners=True), ``F.upsample_nearest`` =
interpolate(nearest) (torch/nn/functional.py).
"""
import torch
import torch.nn.functional as F


def _object_ids(label_map, background):
    ids = torch.unique(label_map)
    return ids[1:] if len(ids) and ids[0] == background else ids


def tight_box(mask):
    """lib/utils/mask.py:180-187 -> x_min, y_min, x_max, y_max of the non-zero pixels."""
    ys, xs = torch.nonzero(mask, as_tuple=True)
    return int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())


def filter_labels_depth(labels, depth, threshold):
    """:186-199."""
    out = labels.clone()
    for i in range(labels.shape[0]):
        for obj in _object_ids(labels[i], 0):
            where = labels[i] == obj
            valid = (depth[i, 2][where] > 0).sum().float() / where.float().sum()
            if valid < threshold:
                out[i][where] = 0
    return out


def crop_rois(rgb, initial_masks, depth, crop_size=224, padding_percentage=0.25):
    """:62-114."""
    _, H, W = initial_masks.shape
    ids = _object_ids(initial_masks[0], 0)
    S = crop_size
    rgb_crops, mask_crops, rois = torch.zeros(len(ids), 3, S, S), torch.zeros(len(ids), S, S), torch.zeros(len(ids), 4)
    depth_crops = torch.zeros(len(ids), 3, S, S) if depth is not None else None
    for k, obj in enumerate(ids):
        mask = (initial_masks[0] == obj).float()
        x0, y0, x1, y1 = tight_box(mask)
        px = int(torch.round(torch.tensor(float(x1 - x0)) * padding_percentage))
        py = int(torch.round(torch.tensor(float(y1 - y0)) * padding_percentage))
        x0, x1, y0, y1 = max(x0 - px, 0), min(x1 + px, W - 1), max(y0 - py, 0), min(y1 + py, H - 1)
        rois[k] = torch.tensor([x0, y0, x1, y1], dtype=torch.float32)
        window = (slice(y0, y1 + 1), slice(x0, x1 + 1))
        rgb_crops[k] = F.interpolate(rgb[0][(slice(None),) + window][None], size=(S, S), mode="bilinear",
                                     align_corners=True)[0]
        mask_crops[k] = F.interpolate(mask[window][None, None], size=(S, S), mode="nearest")[0, 0]
        if depth is not None:
            depth_crops[k] = F.interpolate(depth[0][(slice(None),) + window][None], size=(S, S), mode="bilinear",
                                           align_corners=True)[0]
    return rgb_crops, mask_crops, rois, depth_crops


def match_label_crop(initial_masks, labels_crop, out_label_crop, rois, depth_crop):
    """:118-182. Returns (refined_masks, labels_crop with the rejected local objects set to -1)."""
    labels_crop = labels_crop.clone()
    num = labels_crop.shape[0]
    for i in range(num):
        for obj in torch.unique(labels_crop[i]):
            where = labels_crop[i] == obj
            if (where.float() * out_label_crop[i]).sum() / where.float().sum() < 0.5:
                labels_crop[i][where] = -1
    keys = []
    for i in range(num):
        if depth_crop is not None:
            alive = labels_crop[i] > -1
            z = depth_crop[i, 2][alive] if alive.sum() > 0 else depth_crop[i, 2]
            keys.append(torch.mean(z[z > 0]))
        else:
            keys.append((rois[i, 3] - rois[i, 1] + 1) * (rois[i, 2] - rois[i, 0] + 1))
    order = [i for i, _ in sorted(enumerate(keys), key=lambda t: t[1], reverse=True)]
    refined = torch.zeros_like(initial_masks).float()
    count = 0
    for i in order:
        renumbered = torch.zeros_like(labels_crop[i])
        for obj in _object_ids(labels_crop[i], -1):
            count += 1
            renumbered[labels_crop[i] == obj] = count
        x0, y0, x1, y1 = (int(v) for v in rois[i])
        back = F.interpolate(renumbered[None, None].float(), size=(y1 - y0 + 1, x1 - x0 + 1), mode="nearest")[0, 0]
        view = refined[0, y0:y1 + 1, x0:x1 + 1]
        view[back != 0] = back[back != 0]
    return refined, labels_crop
