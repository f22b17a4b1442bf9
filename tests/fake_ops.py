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
