"""Two-stage ("zoom-in") glue of the reference, lib/fcn/test_dataset.py:43-198, on the device.

Same function names, argument order and return values as the reference (``cfg.TRAIN.SYN_CROP_SIZE`` becomes the
``crop_size`` keyword, default 224 as in lib/fcn/config.py). The reference loops over objects in Python with
full-image ops and ``.item()`` round-trips per object; here each function is one statistics kernel over all objects,
one small device->host copy of those statistics (a few KB), the reference's integer logic on them, and one kernel
that writes the result (``ops.label_stats`` / ``crop_resize`` / ``crop_label_stats`` / ``paste_crops`` /
``relabel_lut``). All tensors are CUDA fp32; label maps hold small non-negative integer ids (< 1024).
"""
import numpy as np
import torch

from .. import ops
from ..meanshiftformer.modeling.transformer_decoder.mean_shift import clustering_features  # noqa: F401  (:43-59)


def _present(stats_row):
    """ids with at least one pixel, ascending (torch.unique order)."""
    return np.nonzero(stats_row[:, 0] > 0)[0]


def filter_labels_depth(labels, depth, threshold):
    """:186-199 - drop (set to 0) every object whose share of pixels with depth z > 0 is below ``threshold``.
    labels [N,H,W], depth [N,3,H,W]."""
    stats = ops.label_stats(labels, depth)
    host = stats.cpu().numpy()
    N, L = host.shape[:2]
    lut = np.tile(np.arange(L, dtype=np.float32), (N, 1))
    for i in range(N):
        for l in _present(host[i]):
            if l == 0:
                continue
            if np.float32(host[i, l, 1]) / np.float32(host[i, l, 0]) < np.float32(threshold):
                lut[i, l] = 0
    return ops.relabel_lut(labels, torch.from_numpy(lut).to(labels.device))


def crop_rois(rgb, initial_masks, depth, crop_size=224, padding_percentage=0.25):
    """:62-114 - one padded ROI per object of ``initial_masks[0]``, resized to crop_size x crop_size.
    rgb [1,3,H,W], initial_masks [1,H,W], depth [1,3,H,W] or None ->
    (rgb_crops [num,3,S,S], mask_crops [num,S,S], rois [num,4] fp32 x_min,y_min,x_max,y_max, depth_crops or None)."""
    _, H, W = initial_masks.shape
    host = ops.label_stats(initial_masks[:1]).cpu().numpy()[0]
    ids = _present(host)
    if len(ids) and ids[0] == 0:
        ids = ids[1:]
    rois = np.zeros((len(ids), 4), dtype=np.int32)
    for k, l in enumerate(ids):
        x_min, y_min, x_max, y_max = W - host[l, 2], H - host[l, 3], host[l, 4] - 1, host[l, 5] - 1
        x_pad = int(np.round(np.float32(x_max - x_min) * np.float32(padding_percentage)))
        y_pad = int(np.round(np.float32(y_max - y_min) * np.float32(padding_percentage)))
        rois[k] = (max(x_min - x_pad, 0), max(y_min - y_pad, 0), min(x_max + x_pad, W - 1), min(y_max + y_pad, H - 1))
    dev = initial_masks.device
    rois_dev = torch.from_numpy(rois).to(dev)
    rgb_crops, depth_crops, mask_crops = ops.crop_resize(
        rgb[0], depth[0] if depth is not None else None, initial_masks[0], rois_dev,
        torch.from_numpy(ids.astype(np.float32)).to(dev), crop_size)
    return rgb_crops, mask_crops, rois_dev.float(), depth_crops


def match_label_crop(initial_masks, labels_crop, out_label_crop, rois, depth_crop):
    """:118-182 - reject local objects that overlap the initial mask by less than half, order the crops (far to
    near by mean depth, or large to small without depth), renumber the surviving objects and paste them back.
    Returns (refined_masks [1,H,W], labels_crop with rejected ids set to -1)."""
    num = labels_crop.shape[0]
    _, H, W = initial_masks.shape
    dev = initial_masks.device
    if num == 0:
        return torch.zeros_like(initial_masks).float(), labels_crop
    stats, dsum = ops.crop_label_stats(labels_crop, out_label_crop, depth_crop)
    stats, dsum = stats.cpu().numpy(), dsum.cpu().numpy()
    rois_host = rois.cpu().numpy().astype(np.int64)
    L = stats.shape[1]
    keep = np.zeros((num, L), dtype=bool)
    for i in range(num):
        for l in _present(stats[i]):
            keep[i, l] = not (np.float32(stats[i, l, 1]) / np.float32(stats[i, l, 0]) < np.float32(0.5))
    keys = []
    for i in range(num):
        if depth_crop is not None:
            sel = keep[i] if stats[i, keep[i], 0].sum() > 0 else np.ones(L, dtype=bool)
            cnt = stats[i, sel, 2].sum()
            keys.append(np.float32(dsum[i, sel].sum() / cnt) if cnt else np.float32("nan"))
        else:
            keys.append((rois_host[i, 3] - rois_host[i, 1] + 1) * (rois_host[i, 2] - rois_host[i, 0] + 1))
    order = [i for i, _ in sorted(enumerate(keys), key=lambda x: x[1], reverse=True)]
    new_label = np.zeros((num, L), dtype=np.float32)
    mark = np.tile(np.arange(L, dtype=np.float32), (num, 1))
    count = 0
    for index in order:
        for l in np.nonzero(keep[index])[0]:
            count += 1
            new_label[index, l] = count
    mark[~keep] = -1
    refined = ops.paste_crops(labels_crop, torch.from_numpy(new_label).to(dev),
                              torch.tensor(order, dtype=torch.int32), torch.from_numpy(rois_host.astype(np.int32)), H, W)
    return refined.unsqueeze(0), ops.relabel_lut(labels_crop, torch.from_numpy(mark).to(dev))
