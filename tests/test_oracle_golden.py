"""Pin the oracle (oracle/*.py) against vectors produced by the reference's own code
(tests/golden/*.npz, generator tests/golden/make_golden.py). CPU only."""
import numpy as np
import torch

from oracle import decoder as odec
from oracle import head as ohead
from oracle import mean_shift as oms
from oracle import pixel_decoder as opd
from oracle import vmf_attention as ovmf

TOL = dict(rtol=1e-5, atol=1e-6)


def _additive(blocked):
    return torch.zeros(blocked.shape).masked_fill_(blocked, float("-inf"))


def test_hypersphere_attention(golden):
    g, _ = golden("hypersphere_attention")
    assert float(g["kappa_default"]) == ovmf.KAPPA == 30.0
    out, attn = ovmf.hypersphere_attention(g["q"], g["k"], g["v"], _additive(g["blocked"]))
    torch.testing.assert_close(out, g["out_masked"], **TOL)
    torch.testing.assert_close(attn, g["attn_masked"], **TOL)
    out, attn = ovmf.hypersphere_attention(g["q"], g["k"], g["v"])
    torch.testing.assert_close(out, g["out_nomask"], **TOL)
    torch.testing.assert_close(attn, g["attn_nomask"], **TOL)
    out, attn = ovmf.hypersphere_attention(g["q"], g["k"], g["v"], None, 10.0)
    torch.testing.assert_close(out, g["out_kappa10"], **TOL)
    torch.testing.assert_close(attn, g["attn_kappa10"], **TOL)
    # outputs are unit vectors, attention rows sum to one, blocked keys get zero weight
    torch.testing.assert_close(g["out_masked"].norm(dim=-1), torch.ones(4, 10), **TOL)
    assert float(g["attn_masked"][g["blocked"]].abs().max()) == 0.0


def test_hypersphere_attention_backward(golden):
    """The hand-written backward of the oracle against torch.autograd through the reference's function."""
    g, _ = golden("hypersphere_attention_bwd")
    for tag, mask, kappa in (("masked", _additive(g["blocked"]), 30.0), ("nomask", None, 30.0), ("kappa10", None, 10.0)):
        out, _ = ovmf.hypersphere_attention(g["q"], g["k"], g["v"], mask, kappa)
        torch.testing.assert_close(out, g[f"out_{tag}"], **TOL)
        gq, gk, gv = ovmf.hypersphere_attention_backward(g["q"], g["k"], g["v"], mask, kappa, g["grad_out"])
        for got, name in ((gq, "gq"), (gk, "gk"), (gv, "gv")):
            want = g[f"{name}_{tag}"]
            torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-6 * max(1.0, float(want.abs().max())))
    # blocked keys receive no gradient through their value rows
    gv = g["gv_masked"]
    all_blocked = g["blocked"].all(dim=1)  # [BH, S]: keys no query of the problem may attend
    assert float(gv[all_blocked].abs().max() if all_blocked.any() else 0.0) == 0.0


def test_meanshift_attention(golden):
    g, sd = golden("meanshift_attention")
    H = int(g["num_heads"])
    out, w = ovmf.meanshift_attention(g["query"], g["key"], g["value"], sd["in_proj_weight"], sd["in_proj_bias"],
                                      sd["out_proj.weight"], sd["out_proj.bias"], H, g["blocked"])
    torch.testing.assert_close(out, g["out"], **TOL)
    torch.testing.assert_close(w, g["weights"], **TOL)
    q = g["query"]
    out, w = ovmf.meanshift_attention(q, q, q, sd["in_proj_weight"], sd["in_proj_bias"],
                                      sd["out_proj.weight"], sd["out_proj.bias"], H, None)
    torch.testing.assert_close(out, g["out_self"], **TOL)
    torch.testing.assert_close(w, g["weights_self"], **TOL)


def test_position_encoding(golden):
    g, _ = golden("position_encoding")
    torch.testing.assert_close(odec.position_embedding_sine(2, 5, 7, 16), g["pos16"], **TOL)
    torch.testing.assert_close(odec.position_embedding_sine(1, 15, 20, 128), g["pos128_15x20"], **TOL)


def _check_decoder(out, g, n_aux):
    torch.testing.assert_close(out["pred_logits"], g["pred_logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out["pred_masks"], g["pred_masks"], rtol=1e-4, atol=1e-5)
    assert len(out["aux_outputs"]) == n_aux
    for i, a in enumerate(out["aux_outputs"]):
        torch.testing.assert_close(a["pred_logits"], g[f"aux{i}_pred_logits"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(a["pred_masks"], g[f"aux{i}_pred_masks"], rtol=1e-4, atol=1e-5)


def test_decoder_multiscale(golden):
    g, sd = golden("decoder_multiscale")
    out = odec.decoder_forward(sd, [g["x0"], g["x1"], g["x2"]], g["mask_features"], num_heads=2, num_layers=4)
    _check_decoder(out, g, 4)


def test_decoder_pretrained(golden):
    g, sd = golden("decoder_pretrained")
    out = odec.decoder_forward(sd, [g["x0"]], g["mask_features"], num_heads=2, num_layers=3)
    _check_decoder(out, g, 3)


def test_msdeform_core_known_answer(golden):
    """The reference's own known-answer recipe, pixel_decoder/ops/test.py:24-63."""
    g, _ = golden("msdeform_core_testpy")
    args = (g["spatial_shapes"], g["level_start_index"])
    out64 = opd.ms_deform_attn_core(g["value"].double(), *args, g["sampling_locations"].double(),
                                    g["attention_weights"].double())
    assert torch.allclose(out64, g["out_fp64"])  # test.py:43 default tolerances
    out32 = opd.ms_deform_attn_core(g["value"], *args, g["sampling_locations"], g["attention_weights"])
    assert torch.allclose(out32, g["out_fp32"], rtol=1e-2, atol=1e-3)  # test.py:59
    torch.testing.assert_close(out32, g["out_fp32"], **TOL)


def test_msdeform_core_uois_geometry(golden):
    g, _ = golden("msdeform_core_uois")
    out = opd.ms_deform_attn_core(g["value"], g["spatial_shapes"], g["level_start_index"],
                                  g["sampling_locations"], g["attention_weights"])
    torch.testing.assert_close(out, g["out_fp32"], rtol=1e-4, atol=1e-5)
    out64 = opd.ms_deform_attn_core(g["value"].double(), g["spatial_shapes"], g["level_start_index"],
                                    g["sampling_locations"].double(), g["attention_weights"].double())
    torch.testing.assert_close(out64, g["out_fp64"], rtol=1e-9, atol=1e-10)


def test_msdeform_module(golden):
    g, sd = golden("msdeform_module")
    out = opd.ms_deform_attn_module(sd, "", g["query"], g["reference_points"], g["input_flatten"],
                                    g["spatial_shapes"], g["level_start_index"], 4, 3, 4)
    torch.testing.assert_close(out, g["out"], rtol=1e-4, atol=1e-5)


def _feats(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("in_")}


def test_pixel_decoder_msdeform(golden):
    g, sd = golden("pixel_decoder_msdeform")
    mf, first, ms = opd.msdeform_pixel_decoder_forward(sd, _feats(g), n_heads=4, enc_layers=2)
    torch.testing.assert_close(mf, g["mask_features"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(first, g["encoder_first"], rtol=1e-4, atol=1e-5)
    for i in range(3):
        torch.testing.assert_close(ms[i], g[f"ms{i}"], rtol=1e-4, atol=1e-5)


def test_pixel_decoder_simple(golden):
    g, sd = golden("pixel_decoder_simple")
    mf, none, ms = opd.simple_pixel_decoder_forward(sd, {"res5": g["x"]})
    assert none is None and ms[0] is g["x"]
    torch.testing.assert_close(mf, g["mask_features"], **TOL)


def test_head_r50style(golden):
    g, sd = golden("head_r50style")
    out, mf = ohead.head_forward(sd, _feats(g), pixel_decoder="MSDeformAttnPixelDecoder", num_heads=2,
                                 dec_layers=4, pd_heads=4, pd_enc_layers=2)
    torch.testing.assert_close(mf, g["last_feature_map"], rtol=1e-4, atol=1e-5)
    _check_decoder(out, g, 4)


def test_mean_shift(golden):
    g, _ = golden("mean_shift")
    X, Z0 = g["X"], g["Z0"]
    torch.testing.assert_close(oms.seed_hill_climbing_ball(X, Z0, 10, 10), g["Z_kappa10_it10"], **TOL)
    torch.testing.assert_close(oms.seed_hill_climbing_ball(X, Z0, 20, 4), g["Z_kappa20_it4"], **TOL)
    assert torch.equal(oms.connected_components(g["Z_kappa10_it10"], 0.04), g["cc_labels"])
    labels, Z = oms.mean_shift_with_seeds(X, Z0, 10, 10)
    assert torch.equal(labels, g["ws_labels"])
    torch.testing.assert_close(Z, g["ws_Z"], **TOL)
    first = int(g["first_seed_index"])
    seeds, idx = oms.select_smart_seeds(X, 12, first)
    assert torch.equal(idx, g["smart_indices"])
    torch.testing.assert_close(seeds, g["smart_seeds"], rtol=0, atol=0)
    labels, idx = oms.mean_shift_smart_init(X, 20, 12, 10, first)
    assert torch.equal(idx, g["smart_indices"])
    assert torch.equal(labels, g["smart_init_labels"])
    assert int(np.bincount(labels.numpy()).argmax()) == 0  # largest cluster carries label 0


def test_mean_shift_clusterer_d64(golden):
    """whole clusterer at the UOIS width (d = 64, 100 seeds, kappa 20): oracle == reference outputs."""
    g, _ = golden("mean_shift_d64")
    X, first = g["X"], int(g["first_seed_index"])
    seeds, idx = oms.select_smart_seeds(X, 100, first)
    assert torch.equal(idx, g["smart_indices"])
    assert torch.equal(seeds, g["smart_seeds"])
    seed_labels, Z = oms.mean_shift_with_seeds(X, seeds, 20, 10)
    torch.testing.assert_close(Z, g["Z"], **TOL)
    assert torch.equal(seed_labels, g["seed_labels"])
    assert torch.equal(oms.connected_components(g["Z"], 0.04), g["seed_labels"])
    labels, idx = oms.mean_shift_smart_init(X, 20, 100, 10, first)
    assert torch.equal(labels, g["smart_init_labels"])
    assert int(np.bincount(labels.numpy()).argmax()) == 0


def test_instance_inference_tail(golden):
    """mask upsample + instance_inference: oracle == the reference's own method (rows matched by score order;
    the three empty masks of image 1 tie at score 0 and are interchangeable)."""
    from oracle import instance_inference as oii
    g, _ = golden("instance_inference")
    T, H, W = int(g["topk"]), int(g["height"]), int(g["width"])
    res = oii.inference_tail(g["pred_logits"], g["pred_masks"], (H, W), T)
    for b, r in enumerate(res):
        mine = torch.argsort(r["scores"], descending=True, stable=True)
        ref = torch.argsort(g[f"scores_{b}"], descending=True, stable=True)
        torch.testing.assert_close(r["scores"][mine], g[f"scores_{b}"][ref], rtol=1e-6, atol=1e-7)
        assert torch.equal(r["pred_classes"][mine], g[f"classes_{b}"][ref])
        assert torch.equal(r["pred_boxes"][mine], g[f"boxes_{b}"][ref])
        assert torch.equal(r["pred_masks"][mine].to(torch.uint8), g[f"masks_{b}"][ref])
        # get_confident_instances + combine_masks (lib/fcn/test_utils.py:35-52, 93-112): same CPU topk order as the
        # reference run, so the label maps must be identical
        K = g["pred_logits"].shape[-1] - 1
        for tag, kw in (("score", dict(topk=False, score=0.5)), ("topk", dict(topk=True, low_threshold=0.3))):
            conf = oii.get_confident_instances(r, num_class=K, **kw)
            torch.testing.assert_close(conf["scores"], g[f"kept_{tag}_{b}"], rtol=1e-6, atol=1e-7)
            assert torch.equal(oii.combine_masks(conf), g[f"labelmap_{tag}_{b}"])


def test_two_stage_glue(golden):
    """crop_rois / match_label_crop / filter_labels_depth (lib/fcn/test_dataset.py:62-198): oracle == reference,
    with and without depth."""
    from oracle import two_stage as ots
    g, _ = golden("two_stage")
    S = int(g["crop_size"])
    filtered = ots.filter_labels_depth(g["d_labels"], g["d_depth"], 0.5)
    assert torch.equal(filtered, g["d_filtered_05"])
    assert torch.equal(ots.filter_labels_depth(g["d_labels"], g["d_depth"], 0.8), g["d_filtered_08"])
    for tag in "dn":
        depth = g["d_depth"] if tag == "d" else None
        labels = filtered if tag == "d" else g["n_labels"]
        rgb_crops, mask_crops, rois, depth_crops = ots.crop_rois(g[tag + "_rgb"], labels, depth, crop_size=S)
        assert torch.equal(rois, g[tag + "_rois"]) and torch.equal(mask_crops, g[tag + "_mask_crops"])
        assert torch.equal(rgb_crops, g[tag + "_rgb_crops"])
        assert depth_crops is None or torch.equal(depth_crops, g["d_depth_crops"])
        refined, marked = ots.match_label_crop(labels, g[tag + "_labels_crop_in"], mask_crops, rois, depth_crops)
        assert torch.equal(refined, g[tag + "_refined"]) and torch.equal(marked, g[tag + "_labels_crop_out"])
