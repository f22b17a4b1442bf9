"""Training side on CPU: the autograd wiring of ops.py with the kernels replaced by the
contract-level stand-ins of tests/fake_ops.py. What is checked here is the host code - which tensors are saved, how
the strided head views are passed, where each gradient is returned - against torch.autograd through the oracle
(itself pinned on the reference's function by tests/test_oracle_golden.py). The kernels are covered by the GPU tests."""
import pytest
import torch

import fake_ops
from oracle import vmf_attention as ovmf


@pytest.fixture
def ops(monkeypatch):
    from unseenobjectswithmeanshift_b200 import ops as mod
    for name in ("vmf_attention", "vmf_attention_bwd"):
        monkeypatch.setattr(mod, name, getattr(fake_ops, name))
    return mod


def _pack_bits(blocked):
    B, Q, S = blocked.shape
    words = (S + 31) // 32
    pad = torch.zeros(B, Q, words * 32, dtype=torch.bool)
    pad[..., :S] = blocked
    v = (pad.view(B, Q, words, 32).long() << torch.arange(32)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


@pytest.mark.parametrize("masked", [False, True])
def test_vmf_attention_function_gradients(ops, masked):
    torch.manual_seed(5)
    B, H, Q, S, hd = 2, 2, 7, 45, 8
    C = H * hd
    # seq-first projections as the attention module holds them: heads are addressed through strides
    q = torch.randn(Q, B, C, dtype=torch.float64, requires_grad=True)
    kv = torch.randn(S, B, 2 * C, dtype=torch.float64, requires_grad=True)
    k, v = kv[..., :C], kv[..., C:]
    heads = lambda t, L: t.reshape(L, B, H, hd).permute(1, 2, 0, 3)
    bits = ro = fmask = None
    if masked:
        blocked = torch.rand(B, Q, S) < 0.5
        blocked[:, 2] = True   # a row that blocks every key is treated as open (decoder.py:618)
        ro = (~blocked).any(-1).to(torch.int32)
        bits = _pack_bits(blocked)
        eff = blocked & (ro != 0).unsqueeze(-1)
        fmask = torch.zeros(B, 1, Q, S, dtype=torch.float64).masked_fill_(eff.unsqueeze(1), float("-inf"))
        fmask = fmask.expand(B, H, Q, S).reshape(B * H, Q, S)
    out = ops.vmf_attention_autograd(heads(q, Q), heads(k, S), heads(v, S), blocked_bits=bits, row_open=ro)
    assert out.shape == (B, H, Q, hd)
    gout = torch.randn(B, H, Q, hd, dtype=torch.float64)
    gq, gkv = torch.autograd.grad(out, (q, kv), gout)

    q2, kv2 = q.detach().clone().requires_grad_(), kv.detach().clone().requires_grad_()
    flat = lambda t: t.reshape(B * H, t.shape[2], hd)
    ref, _ = ovmf.hypersphere_attention(flat(heads(q2, Q)), flat(heads(kv2[..., :C], S)), flat(heads(kv2[..., C:], S)),
                                        fmask)
    torch.testing.assert_close(flat(out), ref, rtol=1e-9, atol=1e-12)
    rq, rkv = torch.autograd.grad(ref, (q2, kv2), flat(gout))
    torch.testing.assert_close(gq, rq, rtol=1e-7, atol=1e-10)
    torch.testing.assert_close(gkv, rkv, rtol=1e-7, atol=1e-10)


def test_vmf_attention_bwd_standin_matches_oracle_backward(golden):
    """The contract the kernel implements (gradients from den and |o|) equals the oracle's hand-written backward."""
    g, _ = golden("hypersphere_attention_bwd")
    q, k, v = (g[n].unsqueeze(1) for n in "qkv")   # [G,1,L,E]: batch = G, one head
    fmask = torch.zeros(g["blocked"].shape).masked_fill_(g["blocked"], float("-inf"))
    out, den = fake_ops.vmf_attention(q, k, v, add_mask=fmask, return_den=True, save_norm=True)
    torch.testing.assert_close(out.squeeze(1), g["out_masked"], rtol=1e-5, atol=1e-6)
    got = fake_ops.vmf_attention_bwd(q, k, v, out, g["grad_out"].unsqueeze(1), den, add_mask=fmask)
    for t, name in zip(got, ("gq", "gk", "gv")):
        want = g[f"{name}_masked"]
        torch.testing.assert_close(t.squeeze(1), want, rtol=1e-4, atol=1e-6 * max(1.0, float(want.abs().max())))


# --------------------------------------------------------------------------------------------------------------
# decoder training path: gradients of a probe loss wrt every parameter and input, against torch.autograd through
# the REFERENCE decoder (tests/golden/decoder_multiscale_bwd.npz)
DEC_KW = dict(num_classes=2, hidden_dim=32, num_queries=10, nheads=2, dim_feedforward=64, dec_layers=4,
              pre_norm=False, mask_dim=32, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)


def _check_grads(named_grads, want, rtol):
    seen = 0
    for name, g in named_grads:
        key = "grad::" + name
        if g is None:
            assert key not in want, f"{name}: no gradient, the reference has one"
            continue
        w = want[key]
        scale = max(float(w.abs().max()), 1e-6)
        err = float((g.float() - w).abs().max())
        assert err <= rtol * scale, f"{name}: max abs err {err:.3e} vs peak {scale:.3e}"
        seen += 1
    assert seen == sum(k.startswith("grad::") for k in want)


def test_oracle_decoder_training_gradients(golden):
    """The oracle decoder under torch.autograd reproduces the reference's gradients (it detaches the attention
    mask exactly where the reference does)."""
    from oracle import decoder as odec
    from scenes import probe_loss
    g, sd = golden("decoder_multiscale")
    want, _ = golden("decoder_multiscale_bwd")
    sd = {k: v.clone().requires_grad_() for k, v in sd.items()}
    x = [g[f"x{i}"].clone().requires_grad_() for i in range(3)]
    mf = g["mask_features"].clone().requires_grad_()
    loss = probe_loss(odec.decoder_forward(sd, x, mf, num_heads=2, num_layers=4))
    torch.testing.assert_close(loss.detach(), want["loss"], rtol=1e-4, atol=1e-5)
    names = list(sd) + ["x0", "x1", "x2", "mask_features"]
    grads = torch.autograd.grad(loss, list(sd.values()) + x + [mf], allow_unused=True)
    _check_grads(zip(names, grads), want, 2e-4)


def test_decoder_training_path_wiring(ops, monkeypatch, golden):
    """MeanShiftTransformerDecoder with grad enabled: autograd Functions around the (stand-in) kernels, torch for the
    other layers. Gradients reach every parameter the reference trains, with the reference's values."""
    from scenes import probe_loss
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder import (
        meanshiftformer_transformer_decoder as dec)
    for name in ("mask_logits", "mask_to_attn_bits", "dense"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    g, sd = golden("decoder_multiscale")
    want, _ = golden("decoder_multiscale_bwd")
    m = dec.MeanShiftTransformerDecoder(int(g["in_channels"]), True, **DEC_KW)
    m.load_state_dict(sd, strict=True)
    m.train()
    x = [g[f"x{i}"].clone().requires_grad_() for i in range(3)]
    mf = g["mask_features"].clone().requires_grad_()
    out = m(x, mf)
    torch.testing.assert_close(out["pred_masks"], g["pred_masks"], rtol=1e-4, atol=1e-5)
    loss = probe_loss(out)
    torch.testing.assert_close(loss.detach(), want["loss"], rtol=1e-4, atol=1e-5)
    params = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [p for _, p in params] + x + [mf], allow_unused=True)
    _check_grads(zip([n for n, _ in params] + ["x0", "x1", "x2", "mask_features"], grads), want, 2e-4)


# --------------------------------------------------------------------------------------------------------------
# losses: batched matcher + criterion against the reference's SetCriterion / HungarianMatcher on recorded points
class _Replay:
    def __init__(self, g):
        self.g = g

    def matcher_points(self, layers, batch, num_points, device):
        assert tuple(self.g["matcher_points"].shape) == (layers, batch, num_points, 2)
        return self.g["matcher_points"].to(device)

    def oversampled_points(self, layers, num_masks, num_sampled, device):
        assert tuple(self.g["candidate_points"].shape) == (layers, num_masks, num_sampled, 2)
        return self.g["candidate_points"].to(device)

    def random_points(self, layers, num_masks, num_random, device):
        assert tuple(self.g["fill_points"].shape) == (layers, num_masks, num_random, 2)
        return self.g["fill_points"].to(device)


def _criterion_case(g, device="cpu"):
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.criterion import SetCriterion
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.matcher import HungarianMatcher
    layers = 3
    P = int(g["num_points"])
    preds = [{"pred_logits": g[f"pred_logits_{l}"].to(device), "pred_masks": g[f"pred_masks_{l}"].to(device)}
             for l in range(layers)]
    targets = [{"labels": g[f"labels_{b}"].to(device), "masks": g[f"masks_{b}"].to(device)} for b in range(2)]
    matcher = HungarianMatcher(cost_class=1.0, cost_mask=20.0, cost_dice=1.0, num_points=P)
    crit = SetCriterion(2, matcher=matcher, weight_dict={}, eos_coef=0.1, losses=["labels", "masks"], num_points=P,
                        oversample_ratio=float(g["oversample_ratio"]),
                        importance_sample_ratio=float(g["importance_sample_ratio"])).to(device)
    return preds, targets, matcher, crit


def test_matcher_assignments_match_reference(golden):
    g, _ = golden("criterion")
    preds, targets, matcher, _ = _criterion_case(g)
    got = matcher.match_layers(preds, targets, _Replay(g))
    for l in range(3):
        for b in range(2):
            assert torch.equal(got[l][b][0], g[f"match_{l}_{b}_pred"]) and got[l][b][0].dtype == torch.int64
            assert torch.equal(got[l][b][1], g[f"match_{l}_{b}_tgt"])
    # the single-layer entry point keeps the reference's signature
    class OneLayer(_Replay):
        def matcher_points(self, layers, batch, num_points, device):
            return self.g["matcher_points"][1:2].to(device)
    one = matcher(preds[1], targets, OneLayer(g))
    assert all(torch.equal(one[b][0], g[f"match_1_{b}_pred"]) for b in range(2))


def test_criterion_losses_and_gradients_match_reference(golden):
    g, _ = golden("criterion")
    preds, targets, _, crit = _criterion_case(g)
    final = {k: v.clone().requires_grad_() for k, v in preds[0].items()}
    losses = crit(dict(final, aux_outputs=preds[1:]), targets, _Replay(g))
    assert list(losses) == g["loss_names"]   # same keys in the same order
    for k, v in losses.items():
        torch.testing.assert_close(v, g["loss::" + k], rtol=1e-5, atol=1e-6)
    gl, gm = torch.autograd.grad(sum(losses.values()), (final["pred_logits"], final["pred_masks"]))
    torch.testing.assert_close(gl, g["grad_pred_logits"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gm, g["grad_pred_masks"], rtol=1e-4, atol=1e-7)
    assert torch.equal(crit.empty_weight, torch.tensor([1.0, 1.0, 0.1]))


def test_criterion_without_targets_and_default_points():
    """No ground truth in the batch: finite losses attached to the graph; default PointSource draws on the device."""
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.criterion import SetCriterion
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.matcher import HungarianMatcher
    torch.manual_seed(0)
    crit = SetCriterion(2, matcher=HungarianMatcher(1.0, 20.0, 1.0, num_points=16), weight_dict={}, eos_coef=0.1,
                        losses=["labels", "masks"], num_points=16, oversample_ratio=3.0, importance_sample_ratio=0.75)
    out = {"pred_logits": torch.randn(2, 5, 3, requires_grad=True), "pred_masks": torch.randn(2, 5, 8, 8, requires_grad=True)}
    empty = [{"labels": torch.zeros(0, dtype=torch.int64), "masks": torch.zeros(0, 16, 16, dtype=torch.bool)}] * 2
    losses = crit(dict(out, aux_outputs=[{k: v.detach() for k, v in out.items()}]), empty)
    assert float(losses["loss_mask"].detach()) == 0.0 and float(losses["loss_dice"].detach()) == 0.0
    assert losses["loss_ce"].isfinite()
    sum(losses.values()).backward()
    some = [{"labels": torch.tensor([1]), "masks": torch.ones(1, 16, 16, dtype=torch.bool)},
            {"labels": torch.zeros(0, dtype=torch.int64), "masks": torch.zeros(0, 16, 16, dtype=torch.bool)}]
    losses = crit(dict(out, aux_outputs=[]), some)
    assert set(losses) == {"loss_ce", "loss_mask", "loss_dice"} and all(v.isfinite() for v in losses.values())


def test_meta_arch_training_branch(ops, monkeypatch):
    """PretrainedMeanShiftMaskFormer.train(): forward returns the reference's weighted loss dict (train branch of
    pretrained_meanshiftformer_model.py:303-334) and gradients reach the backbone, the pixel decoder and the decoder."""
    from torch import nn
    from unseenobjectswithmeanshift_b200 import meanshiftformer as mf
    from unseenobjectswithmeanshift_b200.d2compat import ShapeSpec
    from unseenobjectswithmeanshift_b200.meanshiftformer import modeling
    from unseenobjectswithmeanshift_b200.meanshiftformer.meanshiftformer_model import build_criterion
    for name in ("mask_logits", "mask_to_attn_bits", "dense"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    torch.manual_seed(3)

    class ToyEmbedding(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(3, 64, 3, padding=1)

        def forward(self, image, label=None, depth=None):
            return self.conv(image)

    shapes = {"res5": ShapeSpec(channels=64, stride=1)}
    kw = dict(DEC_KW, dec_layers=2)
    head = modeling.PretrainedMeanShiftMaskFormerHead(
        shapes, num_classes=2, pixel_decoder=modeling.SimpleBasePixelDecoder(shapes, conv_dim=64, mask_dim=32, norm="GN"),
        loss_weight=1.0, ignore_value=255,
        transformer_predictor=modeling.PretrainedMeanShiftTransformerDecoder(64, True, **kw),
        transformer_in_feature="multi_scale_pixel_decoder")
    crit = build_criterion(2, dec_layers=3, train_num_points=32)
    assert list(crit.weight_dict) == ["loss_ce", "loss_mask", "loss_dice", "loss_ce_0", "loss_mask_0", "loss_dice_0",
                                      "loss_ce_1", "loss_mask_1", "loss_dice_1"]
    model = mf.PretrainedMeanShiftMaskFormer(
        backbone=ToyEmbedding(), sem_seg_head=head, criterion=crit, num_queries=10, object_mask_threshold=0.8,
        overlap_threshold=0.8, metadata=None, size_divisibility=0, sem_seg_postprocess_before_inference=True,
        pixel_mean=[0.0, 0.0, 0.0], pixel_std=[1.0, 1.0, 1.0], semantic_on=False, panoptic_on=False, instance_on=True,
        test_topk_per_image=5).train()
    masks = torch.zeros(2, 16, 24, dtype=torch.bool)
    masks[0, 2:9, 3:12] = True
    masks[1, 8:15, 14:22] = True
    batch = [{"image": torch.rand(3, 16, 24), "instances": {"gt_masks": masks, "gt_classes": torch.tensor([0, 1])}},
             {"image": torch.rand(3, 16, 24), "instances": {"gt_masks": masks[:1], "gt_classes": torch.tensor([1])}}]
    losses = model(batch)
    assert list(losses) == list(crit.weight_dict)
    assert all(v.isfinite() and v.requires_grad for v in losses.values())
    sum(losses.values()).backward()
    for prefix in ("pretrained_backbone.", "sem_seg_head.pixel_decoder.", "sem_seg_head.predictor."):
        grads = [p.grad for n, p in model.named_parameters() if n.startswith(prefix) and p.grad is not None]
        assert grads and all(g.isfinite().all() for g in grads) and any(float(g.abs().max()) > 0 for g in grads), prefix
    # loss weighting: mask losses carry the reference's factor 20 relative to the raw criterion output
    torch.manual_seed(9)
    raw = crit({k: (v.detach() if torch.is_tensor(v) else [{a: b.detach() for a, b in d.items()} for d in v])
                for k, v in model._head_outputs(batch)[0].items()}, model.prepare_targets(
                    [x["instances"] for x in batch], (16, 24)))
    torch.manual_seed(9)
    again = crit({k: (v.detach() if torch.is_tensor(v) else [{a: b.detach() for a, b in d.items()} for d in v])
                  for k, v in model._head_outputs(batch)[0].items()}, model.prepare_targets(
                      [x["instances"] for x in batch], (16, 24)))
    assert all(torch.equal(raw[k], again[k]) for k in raw)   # same seed, same points, same losses
    with pytest.raises(RuntimeError, match="training needs a criterion"):
        mf.PretrainedMeanShiftMaskFormer(
            backbone=ToyEmbedding(), sem_seg_head=head, criterion=None, num_queries=10, object_mask_threshold=0.8,
            overlap_threshold=0.8, metadata=None, size_divisibility=0, sem_seg_postprocess_before_inference=True,
            pixel_mean=[0.0] * 3, pixel_std=[1.0] * 3, semantic_on=False, panoptic_on=False, instance_on=True,
            test_topk_per_image=5).train()(batch)


def test_whole_training_step_on_the_r50_style_head(ops, monkeypatch, golden):
    """HeadTrainer (MSDeformAttn pixel decoder + multi-scale decoder + criterion) through training.train_step on CPU,
    every kernel entry point replaced by its contract-level stand-in: the host code of the whole step - which ops
    take the autograd route, what is detached, optimizer groups, clipping - runs end to end, gradients reach every
    trainable tensor and a repeated batch is fitted."""
    from unseenobjectswithmeanshift_b200 import training, workloads
    from unseenobjectswithmeanshift_b200.d2compat import ShapeSpec
    from unseenobjectswithmeanshift_b200.meanshiftformer import modeling as M
    from unseenobjectswithmeanshift_b200.meanshiftformer.meanshiftformer_model import build_criterion
    for name in ("mask_logits", "mask_to_attn_bits", "dense", "ms_deform_attn_forward", "ms_deform_attn_backward"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)  # no driver in the CPU container
    g, sd = golden("head_r50style")
    shapes = {"res2": ShapeSpec(channels=8, stride=4), "res3": ShapeSpec(channels=16, stride=8),
              "res4": ShapeSpec(channels=32, stride=16), "res5": ShapeSpec(channels=64, stride=32)}
    pixel = M.MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=4,
                                       transformer_dim_feedforward=64, transformer_enc_layers=2, conv_dim=32,
                                       mask_dim=32, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                                       common_stride=4)
    head = M.PretrainedMeanShiftMaskFormerHead(shapes, num_classes=2, pixel_decoder=pixel, loss_weight=1.0,
                                               ignore_value=255,
                                               transformer_predictor=M.MeanShiftTransformerDecoder(32, True, **DEC_KW),
                                               transformer_in_feature="multi_scale_pixel_decoder")
    head.load_state_dict(sd, strict=True)
    model = workloads.HeadTrainer(head.train(), build_criterion(2, dec_layers=5, train_num_points=64), 64, 96)
    feats = {k[3:]: v for k, v in g.items() if k.startswith("in_")}
    B = next(iter(feats.values())).shape[0]
    masks = torch.zeros(2, 64, 96, dtype=torch.bool)
    masks[0, 8:30, 10:40] = True
    masks[1, 34:60, 50:90] = True
    targets = [{"labels": torch.tensor([0, 1]), "masks": masks} for _ in range(B)]
    opt = training.build_optimizer(model, lr=1e-3)
    assert {(gr["lr"], gr["weight_decay"]) for gr in opt.param_groups} == {(1e-3, 0.05), (1e-3, 0.0)}
    torch.manual_seed(0)
    history = []
    for _ in range(6):
        losses = training.train_step(model, opt, {"features": feats, "targets": targets}, clip_value=1.0)
        history.append(float(sum(losses.values())))
    assert list(losses) == list(model.criterion.weight_dict)
    assert all(h == h and abs(h) < 1e6 for h in history), history
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing
    assert min(history[3:]) < history[0], history
