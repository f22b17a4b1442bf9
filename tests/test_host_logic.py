"""Host-side logic of the two-stage glue and of the instances -> label-map helpers, on CPU.

The device kernels are replaced by the CPU stand-ins of tests/fake_ops.py (same C-ABI contract), so what is checked
here is the part of unseenobjectswithmeanshift_b200/fcn/ that runs on the host: statistics decoding, ROI padding,
overlap rejection, crop ordering, renumbering - against the outputs of the reference's own functions
(tests/golden/two_stage.npz, instance_inference.npz). The kernels themselves are covered by the ``-m gpu`` tests."""
import pytest
import torch

import fake_ops
from oracle import instance_inference as oii


@pytest.fixture
def td(monkeypatch):
    from unseenobjectswithmeanshift_b200.fcn import test_dataset as mod
    for name in ("label_stats", "relabel_lut", "crop_resize", "crop_label_stats", "paste_crops"):
        monkeypatch.setattr(mod.ops, name, getattr(fake_ops, name))
    return mod


def test_two_stage_host_logic(td, golden):
    g, _ = golden("two_stage")
    S = int(g["crop_size"])
    assert torch.equal(td.filter_labels_depth(g["d_labels"], g["d_depth"], 0.5), g["d_filtered_05"])
    assert torch.equal(td.filter_labels_depth(g["d_labels"], g["d_depth"], 0.8), g["d_filtered_08"])
    for tag in "dn":
        depth = g["d_depth"] if tag == "d" else None
        labels = g["d_filtered_05"] if tag == "d" else g["n_labels"]
        rgb_crops, mask_crops, rois, depth_crops = td.crop_rois(g[tag + "_rgb"], labels, depth, crop_size=S)
        assert torch.equal(rois, g[tag + "_rois"]) and rois.dtype == torch.float32
        assert torch.equal(mask_crops, g[tag + "_mask_crops"]) and torch.equal(rgb_crops, g[tag + "_rgb_crops"])
        assert (depth_crops is None) == (depth is None)
        refined, marked = td.match_label_crop(labels, g[tag + "_labels_crop_in"], mask_crops, rois, depth_crops)
        assert torch.equal(refined, g[tag + "_refined"])
        assert torch.equal(marked, g[tag + "_labels_crop_out"])


def test_two_stage_host_logic_no_objects(td):
    labels = torch.zeros(1, 12, 16)
    rgb_crops, mask_crops, rois, depth_crops = td.crop_rois(torch.rand(1, 3, 12, 16), labels, None, crop_size=8)
    assert rgb_crops.shape == (0, 3, 8, 8) and mask_crops.shape == (0, 8, 8) and rois.shape == (0, 4)
    refined, marked = td.match_label_crop(labels, torch.zeros(0, 8, 8), mask_crops, rois, None)
    assert refined.shape == (1, 12, 16) and float(refined.abs().sum()) == 0


def test_confident_instances_and_combine_masks(golden):
    """fcn/test_utils.get_confident_instances / combine_masks are plain tensor code: run them on the oracle's
    instances (same CPU top-k order as the reference run) against the reference's label maps."""
    from unseenobjectswithmeanshift_b200.fcn import test_utils as tu
    g, _ = golden("instance_inference")
    T, H, W = int(g["topk"]), int(g["height"]), int(g["width"])
    K = g["pred_logits"].shape[-1] - 1
    for b, inst in enumerate(oii.inference_tail(g["pred_logits"], g["pred_masks"], (H, W), T)):
        for tag, kw in (("score", dict(topk=False, score=0.5)), ("topk", dict(topk=True, low_threshold=0.3))):
            conf = tu.get_confident_instances({"instances": inst}, num_class=K, **kw)
            assert torch.equal(tu.combine_masks(conf).double(), g[f"labelmap_{tag}_{b}"])
    empty = tu.get_confident_instances(inst, score=2.0)
    assert empty["scores"].numel() == 0 and float(tu.combine_masks(empty).abs().sum()) == 0


class _FakeStage:
    """stands in for a META_ARCH wrapper: ``label_maps`` returns a prepared label map."""

    def __init__(self, fn):
        self.fn, self.calls = fn, []

    def label_maps(self, batched_inputs, **kw):
        self.calls.append((batched_inputs, kw))
        return self.fn(batched_inputs), {}


def test_two_stage_orchestration(td, golden):
    """fcn/test_utils.two_stage_label_maps (inference part of test_sample_crop, lib/fcn/test_utils.py:245-420):
    with the two networks replaced by the label maps the golden run used, the refined label map must equal the
    reference's; stage 2 is called ONCE with every crop in the batch."""
    from scenes import two_stage_crop_labels
    from unseenobjectswithmeanshift_b200.fcn import test_utils as tu
    g, _ = golden("two_stage")
    S = int(g["crop_size"])
    for tag in "dn":
        depth = g["d_depth"] if tag == "d" else None
        stage1 = _FakeStage(lambda inputs: g[tag + "_labels"].clone())
        stage2 = _FakeStage(lambda inputs: two_stage_crop_labels(g[tag + "_mask_crops"], 7))
        out_label, refined = tu.two_stage_label_maps(stage1, stage2, g[tag + "_rgb"], depth, crop_size=S,
                                                     confident_score=0.6)
        assert torch.equal(out_label, g["d_filtered_05"] if tag == "d" else g["n_labels"])
        assert torch.equal(refined, g[tag + "_refined"])
        assert len(stage2.calls) == 1 and stage2.calls[0][0][0]["image"].shape == g[tag + "_rgb_crops"].shape
        assert ("depth" in stage2.calls[0][0][0]) == (depth is not None)
        assert stage1.calls[0][1]["score"] == 0.6 and stage1.calls[0][0][0]["image"].dim() == 3
    # no crop network: first-stage result only
    out_label, refined = tu.two_stage_label_maps(_FakeStage(lambda i: g["n_labels"].clone()), None, g["n_rgb"][0])
    assert refined is None and torch.equal(out_label, g["n_labels"])
    # nothing segmented: no second stage call
    stage2 = _FakeStage(lambda i: None)
    out_label, refined = tu.two_stage_label_maps(_FakeStage(lambda i: torch.zeros(1, 96, 128)), stage2, g["n_rgb"])
    assert refined is None and not stage2.calls
