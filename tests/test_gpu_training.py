"""GPU tests (marker ``gpu``) of the training-side kernels and of the packed-operand paths:
attention backward, decoder gradients against the reference's autograd, whole training steps, the autograd dense
layer, the weights-epoch cache rule. Written at the end of round 1 as staged tests; green on the B200 since round 2.
Their host wiring is covered on CPU by tests/test_training_wiring.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def msm():
    from unseenobjectswithmeanshift_b200 import ops
    return ops


def _pack_bits(blocked):
    B, Q, S = blocked.shape
    words = (S + 31) // 32
    pad = torch.zeros(B, Q, words * 32, dtype=torch.bool, device=blocked.device)
    pad[..., :S] = blocked
    v = (pad.view(B, Q, words, 32).long() << torch.arange(32, device=blocked.device)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


def _close(got, want, rel):
    err = (got.double() - want.double()).abs().max().item()
    scale = max(want.abs().max().item(), 1e-30)
    assert err <= rel * scale, f"max abs err {err:.3e} vs peak {scale:.3e}"


def test_vmf_attention_bwd_golden(msm, golden):
    """Against torch.autograd through the REFERENCE's hypersphere_attention (tests/golden/make_golden.py)."""
    g, _ = golden("hypersphere_attention_bwd")
    q, k, v = (g[n].cuda().unsqueeze(1) for n in "qkv")   # [G,1,L,E]: batch = G, one head (E = 8: padded head dim)
    gout = g["grad_out"].cuda().unsqueeze(1)
    fmask = torch.zeros(g["blocked"].shape).masked_fill_(g["blocked"], float("-inf")).cuda()
    for tag, mask, kappa in (("masked", fmask, 30.0), ("nomask", None, 30.0), ("kappa10", None, 10.0)):
        out, den = msm.vmf_attention(q, k, v, add_mask=mask, kappa=kappa, return_den=True, save_norm=True)
        _close(out.squeeze(1).cpu(), g[f"out_{tag}"], 2e-5)
        gq, gk, gv = msm.vmf_attention_bwd(q, k, v, out, gout, den, add_mask=mask, kappa=kappa)
        for t, name in ((gq, "gq"), (gk, "gk"), (gv, "gv")):
            _close(t.squeeze(1).cpu(), g[f"{name}_{tag}"], 1e-4)   # fp32: 1e-4 of the peak gradient


@pytest.mark.parametrize("B,H,Q,S,hd,masked", [(1, 1, 100, 64, 32, False), (2, 8, 100, 300, 32, True),
                                               (2, 8, 100, 1200, 32, True), (1, 2, 37, 777, 64, True),
                                               (1, 4, 128, 130, 16, False), (8, 8, 100, 4800, 32, True)])
def test_vmf_attention_autograd_vs_fp64(msm, B, H, Q, S, hd, masked):
    """VmfAttentionFunction on the decoder's strided head views (seq-first projections, fused k|v buffer) against
    fp64 torch.autograd of the same expression; the last case is the training config's largest level."""
    dev = torch.device("cuda")
    gen = torch.Generator(device="cuda").manual_seed(S + hd)
    C = H * hd
    q = torch.randn(B, Q, C, device=dev, generator=gen, requires_grad=True)
    kv = torch.randn(B, S, 2 * C, device=dev, generator=gen, requires_grad=True)
    heads = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S, device=dev, generator=gen) < 0.5
        blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    out = msm.vmf_attention_autograd(heads(q), heads(kv[..., :C]), heads(kv[..., C:]), blocked_bits=bits, row_open=ro)
    gout = torch.randn(B, H, Q, hd, device=dev, generator=gen)
    gq, gkv = torch.autograd.grad(out, (q, kv), gout)

    q2, kv2 = q.detach().double().requires_grad_(), kv.detach().double().requires_grad_()
    qn = torch.nn.functional.normalize(heads(q2), dim=-1, eps=1e-12)
    kn = torch.nn.functional.normalize(heads(kv2[..., :C]), dim=-1, eps=1e-12)
    s = 30.0 * qn @ kn.transpose(-1, -2)
    if eff is not None:
        s = s.masked_fill(eff, float("-inf"))
    ref = torch.nn.functional.normalize(torch.softmax(s, -1) @ heads(kv2[..., C:]), dim=-1, eps=1e-12)
    rq, rkv = torch.autograd.grad(ref, (q2, kv2), gout.double())
    _close(out, ref, 2e-5)
    _close(gq, rq, 1e-4)
    _close(gkv[..., :C], rkv[..., :C], 1e-4)
    _close(gkv[..., C:], rkv[..., C:], 1e-4)
    # deterministic: a second backward gives the same bits
    out2 = msm.vmf_attention_autograd(heads(q), heads(kv[..., :C]), heads(kv[..., C:]), blocked_bits=bits, row_open=ro)
    gq2, gkv2 = torch.autograd.grad(out2, (q, kv), gout)
    assert torch.equal(gq, gq2) and torch.equal(gkv, gkv2)


def test_vmf_attention_bwd_errors_are_loud(msm):
    dev = torch.device("cuda")
    q = torch.randn(1, 1, 200, 32, device=dev)   # more than 128 queries: unsupported by the backward
    k = torch.randn(1, 1, 64, 32, device=dev)
    out, den = msm.vmf_attention(q, k, k, return_den=True, save_norm=True)
    with pytest.raises(RuntimeError, match="at most 128 queries"):
        msm.vmf_attention_bwd(q, k, k, out, torch.ones_like(out), den)
    with pytest.raises(ValueError, match="save_norm"):
        msm.vmf_attention_bwd(q[:, :, :100], k, k, out[:, :, :100].contiguous(), torch.ones(1, 1, 100, 32, device=dev),
                              den[0])


def test_decoder_training_gradients_golden(golden):
    """MeanShiftTransformerDecoder with grad enabled on the device (VmfAttentionFunction + MaskLogitsFunction around
    the kernels, cuBLAS / ATen elsewhere) against torch.autograd through the REFERENCE decoder."""
    from scenes import probe_loss
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder import (
        meanshiftformer_transformer_decoder as dec)
    g, sd = golden("decoder_multiscale")
    want, _ = golden("decoder_multiscale_bwd")
    kw = dict(num_classes=2, hidden_dim=32, num_queries=10, nheads=2, dim_feedforward=64, dec_layers=4,
              pre_norm=False, mask_dim=32, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)
    m = dec.MeanShiftTransformerDecoder(int(g["in_channels"]), True, **kw)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    x = [g[f"x{i}"].cuda().requires_grad_() for i in range(3)]
    mf = g["mask_features"].cuda().requires_grad_()
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = m(x, mf)
        loss = probe_loss(out)
        params = list(m.named_parameters())
        grads = torch.autograd.grad(loss, [p for _, p in params] + x + [mf], allow_unused=True)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    _close(out["pred_masks"].cpu(), g["pred_masks"], 1e-3)
    assert abs(loss.item() - float(want["loss"])) <= 1e-3 * max(1.0, abs(float(want["loss"])))
    seen = 0
    for name, gr in zip([n for n, _ in params] + ["x0", "x1", "x2", "mask_features"], grads):
        key = "grad::" + name
        if gr is None:
            assert key not in want, name
            continue
        _close(gr.cpu(), want[key], 2e-3)   # the hard sigmoid < 0.5 masks make this a same-mask comparison
        seen += 1
    assert seen == sum(k.startswith("grad::") for k in want)


def test_head_training_steps_r50style(golden):
    """Whole training step on the device on the small R50-style head of tests/golden/head_r50style.npz: pixel decoder
    (MSDeformAttn forward / backward kernels), decoder (attention and mask-head Functions), criterion, clipped AdamW.
    Checks plumbing: finite losses with the reference's keys, gradients on every trainable tensor, parameters move,
    and the loss of a repeated batch goes down."""
    from unseenobjectswithmeanshift_b200 import training, workloads
    from unseenobjectswithmeanshift_b200.d2compat import ShapeSpec
    from unseenobjectswithmeanshift_b200.meanshiftformer import modeling as M
    from unseenobjectswithmeanshift_b200.meanshiftformer.meanshiftformer_model import build_criterion
    g, sd = golden("head_r50style")
    shapes = {"res2": ShapeSpec(channels=8, stride=4), "res3": ShapeSpec(channels=16, stride=8),
              "res4": ShapeSpec(channels=32, stride=16), "res5": ShapeSpec(channels=64, stride=32)}
    kw = dict(num_classes=2, hidden_dim=32, num_queries=10, nheads=2, dim_feedforward=64, dec_layers=4,
              pre_norm=False, mask_dim=32, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)
    pixel = M.MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=4,
                                       transformer_dim_feedforward=64, transformer_enc_layers=2, conv_dim=32,
                                       mask_dim=32, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                                       common_stride=4)
    head = M.PretrainedMeanShiftMaskFormerHead(shapes, num_classes=2, pixel_decoder=pixel, loss_weight=1.0,
                                               ignore_value=255,
                                               transformer_predictor=M.MeanShiftTransformerDecoder(32, True, **kw),
                                               transformer_in_feature="multi_scale_pixel_decoder")
    head.load_state_dict(sd, strict=True)
    model = workloads.HeadTrainer(head.train(), build_criterion(2, dec_layers=5, train_num_points=256), 64, 96).cuda()
    feats = {k[3:]: v.cuda() for k, v in g.items() if k.startswith("in_")}
    B = next(iter(feats.values())).shape[0]
    masks = torch.zeros(2, 64, 96, dtype=torch.bool, device="cuda")
    masks[0, 8:30, 10:40] = True
    masks[1, 34:60, 50:90] = True
    targets = [{"labels": torch.tensor([0, 1], device="cuda"), "masks": masks} for _ in range(B)]
    opt = training.build_optimizer(model, lr=1e-3)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    torch.manual_seed(0)
    history = []
    for _ in range(8):
        losses = training.train_step(model, opt, {"features": feats, "targets": targets}, clip_value=1.0)
        history.append(float(sum(losses.values())))
    assert list(losses) == list(model.criterion.weight_dict)
    assert all(h == h and abs(h) < 1e6 for h in history), history
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing
    assert any(not torch.equal(before[n], p) for n, p in model.named_parameters())
    assert min(history[4:]) < history[0], history


@pytest.mark.parametrize("fold", ["kv", "v", "off"])
@pytest.mark.parametrize("decoder,levels", [("MeanShiftTransformerDecoder", 3), ("PretrainedMeanShiftTransformerDecoder", 1)])
def test_decoder_packed_kv_path_matches_default(monkeypatch, decoder, levels, fold):
    """MSM_PACKED_KV=1 (K / V projections write operand images, packed attention kernel) against the default path of
    the same decoder: heads of 32 channels (the packed path's requirement), masks on, key counts with tails.
    ``fold``: input_proj (32 -> 64 channels here) folded into both projections, with the keys' sine embedding as two
    separable tables in the epilogue ("kv", the default), into the value projection only, or not at all; the 35-key
    level takes the token-major input, the 140- and 560-key levels the channel-major map itself."""
    from unseenobjectswithmeanshift_b200.meanshiftformer import modeling as M
    monkeypatch.setenv("MSM_FOLD_V", "0" if fold == "off" else "1")
    monkeypatch.setenv("MSM_FOLD_K", "1" if fold == "kv" else "0")
    torch.manual_seed(3)
    kw = dict(num_classes=2, hidden_dim=64, num_queries=20, nheads=2, dim_feedforward=128, dec_layers=4,
              pre_norm=False, mask_dim=64, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)
    m = getattr(M, decoder)(32, True, **kw).cuda().eval()
    sizes = [(5, 7), (10, 14), (20, 28)][:levels] if levels == 3 else [(20, 28)]
    x = [torch.randn(2, 32, h, w, device="cuda") for h, w in sizes]
    mf = torch.randn(2, 64, 40, 56, device="cuda")
    with torch.no_grad():
        monkeypatch.setenv("MSM_PACKED_KV", "0")
        want = m(x, mf)
        monkeypatch.setenv("MSM_PACKED_KV", "1")
        got = m(x, mf)
        again = m(x, mf)     # the cached image buffer is reused: key tails must still be zero
    for a, b in ((got, want), (again, want)):
        _close(a["pred_masks"], b["pred_masks"], 1e-3)
        _close(a["pred_logits"], b["pred_logits"], 1e-3)


def test_mean_shift_packed_path_matches_default(msm, monkeypatch):
    """MSM_PACKED_MS=1 (X packed once per call, bulk-copy streaming kernel) against the shipped hill climb."""
    torch.manual_seed(0)
    for B, n, m, d in ((2, 5000, 100, 64), (1, 777, 37, 32), (2, 40000, 128, 64)):
        X = torch.nn.functional.normalize(torch.randn(B, n, d, device="cuda"), dim=-1)
        Z = X[:, torch.randperm(n, device="cuda")[:m]].contiguous()
        monkeypatch.setenv("MSM_PACKED_MS", "0")
        want = msm.mean_shift_hill_climb(X, Z, 10.0, 10)
        monkeypatch.setenv("MSM_PACKED_MS", "1")
        got = msm.mean_shift_hill_climb(X, Z, 10.0, 10)
        assert (got - want).abs().max().item() < 1e-4


@pytest.mark.parametrize("B,H,Q,S,masked", [(8, 8, 100, 100, False), (8, 8, 100, 300, True), (1, 1, 128, 70, True),
                                           (2, 2, 37, 1000, True)])
def test_small_attention_kernel_vs_shipped(msm, B, H, Q, S, masked):
    """vmf_small_kernel (single launch, CUDA cores, short key sequences) against the shipped dispatcher."""
    dev = torch.device("cuda")
    gen = torch.Generator(device="cuda").manual_seed(S + Q)
    C = H * 32
    q = torch.randn(B, Q, C, device=dev, generator=gen)
    kv = torch.randn(B, S, 2 * C, device=dev, generator=gen)
    hv = lambda t: t.unflatten(-1, (H, 32)).permute(0, 2, 1, 3)
    bits = ro = None
    if masked:
        blocked = torch.rand(B, Q, S, device=dev, generator=gen) < 0.5
        blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
    want, wden = msm.vmf_attention(hv(q), hv(kv[..., :C]), hv(kv[..., C:]), blocked_bits=bits, row_open=ro,
                                   return_den=True, save_norm=True)
    got, gden = msm.vmf_attention_small(hv(q), hv(kv[..., :C]), hv(kv[..., C:]), blocked_bits=bits, row_open=ro,
                                        return_den=True, save_norm=True)
    assert (got - want).abs().max().item() < 2e-5
    assert ((gden - wden).abs() / wden.abs().clamp_min(1e-30)).max().item() < 1e-4


def test_eval_after_train_step_sees_the_new_weights():
    """AdamW(fused=True) updates parameters in place WITHOUT bumping tensor._version, and every
    inference-side cache of derived weights (prepared 16-bit copies, concatenated K/V weights, row-bias tables, the 3x3
    repack) was keyed on (address, _version) only. eval -> train_step -> eval must equal an eval with fresh caches."""
    from unseenobjectswithmeanshift_b200 import ops, training, workloads
    from unseenobjectswithmeanshift_b200.meanshiftformer import modeling as M
    torch.manual_seed(11)
    kw = dict(num_classes=2, hidden_dim=64, num_queries=20, nheads=2, dim_feedforward=128, dec_layers=3,
              pre_norm=False, mask_dim=64, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)
    dec = M.MeanShiftTransformerDecoder(32, True, **kw).cuda()
    x = [torch.randn(2, 32, h, w, device="cuda") for h, w in ((5, 7), (10, 14), (20, 28))]
    mf = torch.randn(2, 64, 40, 56, device="cuda")

    def evaluate():
        dec.eval()
        with torch.no_grad():
            return dec(x, mf)["pred_masks"].clone()

    before = evaluate()
    dec.train()
    opt = torch.optim.AdamW(dec.parameters(), lr=1e-2, fused=True)
    versions = [p._version for p in dec.parameters()]

    class Wrap(torch.nn.Module):
        def forward(self, _):
            return {"loss": dec(x, mf)["pred_masks"].square().mean()}

    training.train_step(Wrap(), opt, None, clip_value=0)
    after = evaluate()
    assert not torch.equal(before, after), "the optimizer step changed nothing?"
    ops.clear_prepared_weights()          # everything derived from weights is rebuilt from scratch
    fresh = evaluate()
    assert torch.equal(after, fresh), "eval after train_step used stale derived weights"
    # the premise of the fix, recorded: does the fused optimizer bump _version on this torch build?
    print("fused AdamW bumped _version:", [p._version for p in dec.parameters()] != versions)


@pytest.mark.parametrize("M,N,K,relu,bias", [(800, 256, 256, False, True), (1000, 64, 1024, True, True),
                                             (333, 96, 64, False, False), (128, 2048, 256, True, True)])
def test_dense_function_gradients_vs_fp64(M, N, K, relu, bias):
    """ops.dense under autograd (fp32 training): forward and the input gradient in linear_tc_kernel, weight / bias
    gradients in cuBLAS - against fp64 autograd of act(x W^T + b)."""
    from unseenobjectswithmeanshift_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1 if bias else None
    go = torch.randn(M, N, generator=g)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd = torch.relu(yd) if relu else yd
    yd.backward(go.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = b.cuda().requires_grad_(True) if bias else None
    assert ops.dense_autograd_supported(xc, wc)
    y = ops.dense(xc, wc, bc, relu=relu)
    assert y.grad_fn is not None and "DenseFunction" in type(y.grad_fn).__name__
    y.backward(go.cuda())
    rel = lambda a, r: (a.detach().cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 1e-12)  # noqa: E731
    assert rel(y, yd.detach()) < 1e-5
    assert rel(xc.grad, xd.grad) < 1e-5
    assert rel(wc.grad, wd.grad) < 1e-5
    if bias:
        assert rel(bc.grad, bd.grad) < 1e-5


@pytest.mark.parametrize("B,Q,C,H,W", [(2, 100, 256, 24, 32), (1, 10, 32, 12, 20), (2, 37, 64, 9, 12), (1, 128, 160, 8, 8)])
def test_mask_logits_function_gradients_vs_fp64(B, Q, C, H, W):
    """MaskLogitsFunction: forward = mask kernel; g_feat through the same kernel with queries and channels swapped
    (zero-padded to 32 queries, 128 channels per launch), g_embed through cuBLAS fp32 - against fp64 autograd of the
    einsum."""
    from unseenobjectswithmeanshift_b200 import ops
    g = torch.Generator().manual_seed(B + Q + C)
    e, f = torch.randn(B, Q, C, generator=g), torch.randn(B, C, H, W, generator=g)
    go = torch.randn(B, Q, H, W, generator=g)
    ed, fd = e.double().requires_grad_(True), f.double().requires_grad_(True)
    torch.einsum("bqc,bchw->bqhw", ed, fd).backward(go.double())
    ec, fc = e.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    y = ops.mask_logits_autograd(ec, fc)
    y.backward(go.cuda())
    rel = lambda a, r: (a.detach().cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 1e-12)  # noqa: E731
    assert rel(fc.grad, fd.grad) < 1e-5
    assert rel(ec.grad, ed.grad) < 1e-5
