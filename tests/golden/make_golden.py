"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own code on CPU.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is a small .npz holding seeded inputs, the reference module's ``state_dict``
(keys prefixed ``sd::``) and the reference outputs. The oracle (``oracle/``) is pinned against
these files by ``tests/test_oracle_golden.py``; the CUDA path is checked against the same files
by the ``-m gpu`` tests. The reference publishes no stored vectors of its own:
the only known-answer recipe it has is ops/test.py (seed 3, shapes [(6,4),(3,2)]), replayed
below as ``msdeform_core_testpy``.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
from scenes import probe_loss, two_stage_crop_labels, two_stage_scene  # noqa: E402

torch.set_num_threads(4)
torch.backends.mkldnn.enabled = True


def sd_arrays(module):
    return {"sd::" + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


def gen_hypersphere_attention():
    au = ref_shim.ref("modeling.transformer_decoder.attention_util")
    torch.manual_seed(0)
    BH, Q, S, E = 4, 10, 50, 8
    q, k, v = torch.randn(BH, Q, E), torch.randn(BH, S, E), torch.randn(BH, S, E)
    blocked = torch.rand(BH, Q, S) < 0.5
    blocked[:, :, 0] = False  # no fully blocked row (the decoder guarantees this, decoder.py:618)
    fmask = torch.zeros(BH, Q, S).masked_fill_(blocked, float("-inf"))
    out_m, attn_m = au.hypersphere_attention(q, k, v, fmask)
    out_n, attn_n = au.hypersphere_attention(q, k, v, None)
    out_k, attn_k = au.hypersphere_attention(q, k, v, None, 0.0, 10.0)
    save("hypersphere_attention", q=q, k=k, v=v, blocked=blocked, out_masked=out_m, attn_masked=attn_m,
         out_nomask=out_n, attn_nomask=attn_n, out_kappa10=out_k, attn_kappa10=attn_k,
         kappa_default=np.float32(au.KAPPA))


def gen_hypersphere_attention_bwd():
    """Gradients of the reference's own hypersphere_attention (attention_util.py:30-82) by torch.autograd, masked and
    unmasked, for a seeded grad_output: pins the hand-written backward of the oracle (row f4 of SURVEY.md 8)."""
    au = ref_shim.ref("modeling.transformer_decoder.attention_util")
    torch.manual_seed(20)
    BH, Q, S, E = 4, 10, 70, 8
    q, k, v = (torch.randn(BH, n, E).requires_grad_() for n in (Q, S, S))
    blocked = torch.rand(BH, Q, S) < 0.5
    blocked[:, :, 0] = False
    fmask = torch.zeros(BH, Q, S).masked_fill_(blocked, float("-inf"))
    gout = torch.randn(BH, Q, E)
    arrays = dict(q=q, k=k, v=v, blocked=blocked, grad_out=gout)
    for tag, mask, kappa in (("masked", fmask, au.KAPPA), ("nomask", None, au.KAPPA), ("kappa10", None, 10.0)):
        out, _ = au.hypersphere_attention(q, k, v, mask, 0.0, kappa)
        gq, gk, gv = torch.autograd.grad(out, (q, k, v), gout)
        arrays.update({f"out_{tag}": out, f"gq_{tag}": gq, f"gk_{tag}": gk, f"gv_{tag}": gv})
    save("hypersphere_attention_bwd", **arrays)


def gen_meanshift_attention():
    au = ref_shim.ref("modeling.transformer_decoder.attention_util")
    torch.manual_seed(1)
    E, H, L, S, N = 32, 2, 10, 50, 2
    m = au.MeanShiftAttention(E, H).eval()
    with torch.no_grad():
        m.in_proj_bias.normal_(0, 0.1)
        m.out_proj.bias.normal_(0, 0.1)
    query, key, value = torch.randn(L, N, E), torch.randn(S, N, E), torch.randn(S, N, E)
    blocked = torch.rand(N, 1, L, S) < 0.5
    blocked[..., 3] = False
    blocked = blocked.repeat(1, H, 1, 1).flatten(0, 1)
    with torch.no_grad():
        out, w = m(query, key, value, attn_mask=blocked)
        out_self, w_self = m(query, query, query)
    save("meanshift_attention", query=query, key=key, value=value, blocked=blocked, out=out, weights=w,
         out_self=out_self, weights_self=w_self, num_heads=np.int64(H), **sd_arrays(m))


def _decoder_kwargs():
    return dict(num_classes=2, hidden_dim=32, num_queries=10, nheads=2, dim_feedforward=64, dec_layers=4,
                pre_norm=False, mask_dim=32, enforce_input_project=False, use_meanshift_cross_attention=True,
                disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)


def _randomise_biases(module, std=0.05):
    # reference initialisers leave every bias at 0 / LayerNorm at (1, 0); perturb so that the
    # fixtures exercise those terms too
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * std)


def _dump_decoder_out(o):
    d = {"pred_logits": o["pred_logits"], "pred_masks": o["pred_masks"]}
    for i, a in enumerate(o["aux_outputs"]):
        d[f"aux{i}_pred_logits"] = a["pred_logits"]
        d[f"aux{i}_pred_masks"] = a["pred_masks"]
    return d


def gen_decoder_multiscale():
    dec = ref_shim.ref("modeling.transformer_decoder.meanshiftformer_transformer_decoder")
    torch.manual_seed(2)
    m = dec.MeanShiftTransformerDecoder(16, True, **_decoder_kwargs()).eval()
    _randomise_biases(m)
    x = [torch.randn(2, 16, 3, 4), torch.randn(2, 16, 6, 8), torch.randn(2, 16, 12, 16)]
    mf = torch.randn(2, 32, 24, 32)
    with torch.no_grad():
        o = m(x, mf)
    save("decoder_multiscale", x0=x[0], x1=x[1], x2=x[2], mask_features=mf, in_channels=np.int64(16),
         **_dump_decoder_out(o), **sd_arrays(m))


def gen_decoder_multiscale_bwd():
    """Training side: gradients of probe_loss through the REFERENCE decoder (torch.autograd, CPU) wrt every
    parameter, the level features and the mask features, on the decoder_multiscale fixture's weights and inputs."""
    dec = ref_shim.ref("modeling.transformer_decoder.meanshiftformer_transformer_decoder")
    z = np.load(os.path.join(HERE, "decoder_multiscale.npz"))
    m = dec.MeanShiftTransformerDecoder(int(z["in_channels"]), True, **_decoder_kwargs())
    m.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}, strict=True)
    m.train()  # dropout is 0 in every config: train() and eval() compute the same function
    x = [torch.from_numpy(z[f"x{i}"]).requires_grad_() for i in range(3)]
    mf = torch.from_numpy(z["mask_features"]).requires_grad_()
    loss = probe_loss(m(x, mf))
    names = [n for n, _ in m.named_parameters()]
    grads = torch.autograd.grad(loss, [p for _, p in m.named_parameters()] + x + [mf], allow_unused=True)
    arrays = {"loss": loss.detach()}
    for n, g in zip(names + ["x0", "x1", "x2", "mask_features"], grads):
        if g is not None:
            arrays["grad::" + n] = g
    save("decoder_multiscale_bwd", **arrays)


def gen_decoder_pretrained():
    dec = ref_shim.ref("modeling.transformer_decoder.meanshiftformer_transformer_decoder")
    torch.manual_seed(3)
    kw = _decoder_kwargs()
    kw["dec_layers"] = 3
    m = dec.PretrainedMeanShiftTransformerDecoder(16, True, **kw).eval()
    _randomise_biases(m)
    x = [F.normalize(torch.randn(2, 16, 12, 20), dim=1)]
    mf = torch.randn(2, 32, 12, 20)
    with torch.no_grad():
        o = m(x, mf)
    save("decoder_pretrained", x0=x[0], mask_features=mf, in_channels=np.int64(16),
         **_dump_decoder_out(o), **sd_arrays(m))


def gen_posenc():
    pe = ref_shim.ref("modeling.transformer_decoder.position_encoding")
    x = torch.zeros(2, 3, 5, 7)
    save("position_encoding", pos16=pe.PositionEmbeddingSine(16, normalize=True)(x),
         pos128_15x20=pe.PositionEmbeddingSine(128, normalize=True)(torch.zeros(1, 1, 15, 20)),
         shape=np.array([2, 3, 5, 7]))


def gen_msdeform_core():
    fn = ref_shim.ref("modeling.pixel_decoder.ops.functions.ms_deform_attn_func")
    # --- ops/test.py:24-47 recipe (seed 3; generated on the CPU generator, the script uses .cuda()) ---
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    w = torch.rand(N, Lq, M, L, P) + 1e-5
    w /= w.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out64 = fn.ms_deform_attn_core_pytorch(value.double(), shapes, loc.double(), w.double())
    out32 = fn.ms_deform_attn_core_pytorch(value, shapes, loc, w)
    save("msdeform_core_testpy", value=value, spatial_shapes=shapes, level_start_index=lsi, sampling_locations=loc,
         attention_weights=w, out_fp64=out64, out_fp32=out32)

    # --- UOIS-like geometry: 8 heads x 8 channels, 3 levels x 4 points, locations spilling outside [0,1] ---
    torch.manual_seed(4)
    N, M, D, L, P = 2, 8, 8, 3, 4
    shapes = torch.as_tensor([(3, 4), (6, 8), (12, 16)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    Lq = S
    value = torch.randn(N, S, M, D)
    loc = torch.rand(N, Lq, M, L, P, 2) * 1.4 - 0.2
    w = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    out32 = fn.ms_deform_attn_core_pytorch(value, shapes, loc, w)
    out64 = fn.ms_deform_attn_core_pytorch(value.double(), shapes, loc.double(), w.double())
    save("msdeform_core_uois", value=value, spatial_shapes=shapes, level_start_index=lsi, sampling_locations=loc,
         attention_weights=w, out_fp32=out32, out_fp64=out64)


def gen_msdeform_module():
    mod = ref_shim.ref("modeling.pixel_decoder.ops.modules.ms_deform_attn")
    torch.manual_seed(5)
    m = mod.MSDeformAttn(d_model=32, n_levels=3, n_heads=4, n_points=4).eval()
    with torch.no_grad():  # reference init zeroes these weights; make them matter
        m.sampling_offsets.weight.normal_(0, 0.3)
        m.attention_weights.weight.normal_(0, 0.5)
        m.attention_weights.bias.normal_(0, 0.5)
        m.value_proj.bias.normal_(0, 0.1)
        m.output_proj.bias.normal_(0, 0.1)
    shapes = torch.as_tensor([(2, 3), (4, 6), (8, 12)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N = 2
    query, src = torch.randn(N, S, 32), torch.randn(N, S, 32)
    refp = torch.rand(N, S, 3, 2)
    with torch.no_grad():
        out = m(query, refp, src, shapes, lsi, None)
    save("msdeform_module", query=query, input_flatten=src, reference_points=refp, spatial_shapes=shapes,
         level_start_index=lsi, out=out, **sd_arrays(m))


def _pixel_decoder(pd, ShapeSpec):
    input_shape = {"res2": ShapeSpec(channels=8, stride=4), "res3": ShapeSpec(channels=16, stride=8),
                   "res4": ShapeSpec(channels=32, stride=16), "res5": ShapeSpec(channels=64, stride=32)}
    return pd.MSDeformAttnPixelDecoder(
        input_shape, transformer_dropout=0.0, transformer_nheads=4, transformer_dim_feedforward=64,
        transformer_enc_layers=2, conv_dim=32, mask_dim=32, norm="GN",
        transformer_in_features=["res3", "res4", "res5"], common_stride=4)


def _features(seed):
    g = torch.Generator().manual_seed(seed)
    return {"res2": torch.randn(2, 8, 16, 24, generator=g), "res3": torch.randn(2, 16, 8, 12, generator=g),
            "res4": torch.randn(2, 32, 4, 6, generator=g), "res5": torch.randn(2, 64, 2, 3, generator=g)}


def _perturb_msdeform(module):
    mod = ref_shim.ref("modeling.pixel_decoder.ops.modules.ms_deform_attn")
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, mod.MSDeformAttn):
                m.sampling_offsets.weight.normal_(0, 0.2)
                m.attention_weights.weight.normal_(0, 0.5)


def gen_pixel_decoder_msdeform():
    pd = ref_shim.ref("modeling.pixel_decoder.msdeformattn")
    from detectron2.layers import ShapeSpec
    torch.manual_seed(6)
    m = _pixel_decoder(pd, ShapeSpec).eval()
    _perturb_msdeform(m)
    _randomise_biases(m)
    feats = _features(60)
    with torch.no_grad():
        mask_features, enc0, ms = m.forward_features(feats)
    save("pixel_decoder_msdeform", **{"in_" + k: v for k, v in feats.items()}, mask_features=mask_features,
         encoder_first=enc0, ms0=ms[0], ms1=ms[1], ms2=ms[2], **sd_arrays(m))


def gen_pixel_decoder_simple():
    fpn = ref_shim.ref("modeling.pixel_decoder.fpn")
    from detectron2.layers import ShapeSpec
    torch.manual_seed(7)
    m = fpn.SimpleBasePixelDecoder({"res5": ShapeSpec(channels=16, stride=1)}, conv_dim=16, mask_dim=32, norm="GN").eval()
    _randomise_biases(m)
    x = F.normalize(torch.randn(2, 16, 12, 16), dim=1)
    with torch.no_grad():
        mf, none, ms = m.forward_features({"res5": x})
    assert none is None and len(ms) == 1 and ms[0] is x
    save("pixel_decoder_simple", x=x, mask_features=mf, **sd_arrays(m))


def gen_head_r50style():
    """pixel decoder -> multi-scale decoder through the reference head class (meanshift_former_head.py:246-275)."""
    pd = ref_shim.ref("modeling.pixel_decoder.msdeformattn")
    dec = ref_shim.ref("modeling.transformer_decoder.meanshiftformer_transformer_decoder")
    head = ref_shim.ref("modeling.meta_arch.meanshift_former_head")
    from detectron2.layers import ShapeSpec
    torch.manual_seed(8)
    pixel = _pixel_decoder(pd, ShapeSpec)
    _perturb_msdeform(pixel)
    predictor = dec.MeanShiftTransformerDecoder(32, True, **_decoder_kwargs())
    input_shape = {"res2": ShapeSpec(channels=8, stride=4), "res3": ShapeSpec(channels=16, stride=8),
                   "res4": ShapeSpec(channels=32, stride=16), "res5": ShapeSpec(channels=64, stride=32)}
    m = head.PretrainedMeanShiftMaskFormerHead(input_shape, num_classes=2, pixel_decoder=pixel, loss_weight=1.0,
                                               ignore_value=255, transformer_predictor=predictor,
                                               transformer_in_feature="multi_scale_pixel_decoder").eval()
    _randomise_biases(m)
    feats = _features(80)
    with torch.no_grad():
        o, last = m(feats, 64, 96)
    save("head_r50style", **{"in_" + k: v for k, v in feats.items()}, last_feature_map=last,
         **_dump_decoder_out(o), **sd_arrays(m))


def gen_mean_shift():
    ms = ref_shim.ref("modeling.transformer_decoder.mean_shift")
    torch.manual_seed(9)
    n, d, c = 600, 16, 5
    centers = F.normalize(torch.randn(c, d), dim=1)
    X = F.normalize(centers[torch.randint(0, c, (n,))] + 0.15 * torch.randn(n, d), dim=1)
    idx = torch.randperm(n)[:12]
    Z0 = X[idx].clone()
    Z10 = ms.seed_hill_climbing_ball(X, Z0, kappa=10, max_iters=10)
    Z20 = ms.seed_hill_climbing_ball(X, Z0, kappa=20, max_iters=4)
    cc = ms.connected_components(Z10, 0.04)
    labels_ws, Z_ws = ms.mean_shift_with_seeds(X, Z0, 10, max_iters=10)
    np.random.seed(3)  # lib/fcn/config.py:380 RNG_SEED
    seeds, sel = ms.select_smart_seeds(X, 12, return_selected_indices=True)
    np.random.seed(3)
    labels, sel2 = ms.mean_shift_smart_init(X, kappa=20, num_seeds=12, max_iters=10)
    assert torch.equal(sel, sel2)
    save("mean_shift", X=X, seed_indices=idx, Z0=Z0, Z_kappa10_it10=Z10, Z_kappa20_it4=Z20, cc_labels=cc,
         ws_labels=labels_ws, ws_Z=Z_ws, smart_seeds=seeds, smart_indices=sel, smart_init_labels=labels,
         first_seed_index=np.int64(sel[0].item()))


def gen_mean_shift_d64():
    """The whole classical clusterer at the UOIS embedding width (d = 64, 100 seeds, kappa 20, 10 iterations,
    lib/fcn/test_dataset.py:43-59) on a 40x60 synthetic embedding map with 7 objects of unequal size."""
    ms = ref_shim.ref("modeling.transformer_decoder.mean_shift")
    torch.manual_seed(11)
    n, d, c = 2400, 64, 7
    centers = F.normalize(torch.randn(c, d), dim=1)
    which = torch.multinomial(torch.tensor([8.0, 5, 3, 2, 1, 1, 0.5]), n, replacement=True)
    X = F.normalize(centers[which] + 0.04 * torch.randn(n, d), dim=1)
    np.random.seed(3)
    seeds, sel = ms.select_smart_seeds(X, 100, return_selected_indices=True)
    seed_labels, Z = ms.mean_shift_with_seeds(X, seeds.clone(), 20, max_iters=10)
    np.random.seed(3)
    labels, sel2 = ms.mean_shift_smart_init(X, kappa=20, num_seeds=100, max_iters=10)
    assert torch.equal(sel, sel2)
    save("mean_shift_d64", X=X, first_seed_index=np.int64(sel[0].item()), smart_indices=sel, smart_seeds=seeds,
         Z=Z, seed_labels=seed_labels, smart_init_labels=labels)


def gen_instance_inference():
    """Runs the reference's own instance_inference (pretrained_meanshiftformer_model.py:461-497): the method is cut
    out of the file by ast (the module itself needs detectron2 + the UCN networks to import) and executed with the
    three detectron2 structures it touches stubbed: Instances (attribute bag), Boxes (tensor holder), BitMasks
    (get_bounding_boxes restated from detectron2 v0.6 in oracle/instance_inference.py)."""
    import ast
    import types
    from oracle import instance_inference as oii
    path = os.path.join(ref_shim.REF_PKG, "pretrained_meanshiftformer_model.py")
    tree = ast.parse(open(path).read())
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) for n in c.body
              if isinstance(n, ast.FunctionDef) and n.name == "instance_inference")

    class Instances:
        def __init__(self, image_size):
            self.image_size = image_size

    class Boxes:
        def __init__(self, t):
            self.tensor = t

    class BitMasks:
        def __init__(self, t):
            self.tensor = t

        def get_bounding_boxes(self):
            return Boxes(oii.get_bounding_boxes(self.tensor))

    ns = {"torch": torch, "F": F, "np": np, "Instances": Instances, "Boxes": Boxes, "BitMasks": BitMasks}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    torch.manual_seed(21)
    B, Q, K, h, w, H, W, T = 2, 24, 2, 30, 40, 120, 160, 7
    logits = torch.randn(B, Q, K + 1) * 2
    yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    masks = torch.empty(B, Q, h, w)
    for b in range(B):
        for q in range(Q):
            cy, cx, r = torch.rand(1) * h, torch.rand(1) * w, 2 + torch.rand(1) * 8
            masks[b, q] = 4 * (1 - ((yy - cy) ** 2 + (xx - cx) ** 2).sqrt() / r) + 0.5 * torch.randn(h, w)
    masks[1, :3] = -3.0 - torch.rand(3, h, w)      # empty masks: zero box, zero score
    logits[1, :3, 0] += 6                           # ... that are certainly kept
    up = F.interpolate(masks, size=(H, W), mode="bilinear", align_corners=False)   # :337-343
    # get_confident_instances + combine_masks (lib/fcn/test_utils.py:35-52, 93-112), also cut out by ast; the
    # Instances stub supports what they use: attribute access, boolean-mask indexing, .get(), .to()
    class Instances(Instances):
        def __getitem__(self, keep):
            r = Instances(self.image_size)
            for k, v in vars(self).items():
                if k != "image_size":
                    setattr(r, k, (Boxes(v.tensor[keep]) if isinstance(v, Boxes) else v[keep]))
            return r

        def get(self, k):
            return getattr(self, k)

    ns["Instances"] = Instances
    tu = ast.parse(open(os.path.join(ref_shim.REF_ROOT, "lib", "fcn", "test_utils.py")).read())
    exec(compile(ast.Module(body=[n for n in tu.body if isinstance(n, ast.FunctionDef)
                                  and n.name in ("get_confident_instances", "combine_masks")], type_ignores=[]),
                 "test_utils.py", "exec"), ns)
    out = {}
    for b in range(B):
        self_ = types.SimpleNamespace(sem_seg_head=types.SimpleNamespace(num_classes=K), num_queries=Q,
                                      test_topk_per_image=T, panoptic_on=False, device=torch.device("cpu"))
        r = ns["instance_inference"](self_, logits[b], up[b])
        out.update({f"masks_{b}": r.pred_masks.to(torch.uint8), f"boxes_{b}": r.pred_boxes.tensor,
                    f"scores_{b}": r.scores, f"classes_{b}": r.pred_classes})
        for tag, kw in (("score", dict(topk=False, score=0.5)), ("topk", dict(topk=True, low_threshold=0.3))):
            conf = ns["get_confident_instances"]({"instances": r}, num_class=K, **kw)
            out[f"labelmap_{tag}_{b}"] = ns["combine_masks"](conf)
            out[f"kept_{tag}_{b}"] = conf.scores
    save("instance_inference", pred_logits=logits, pred_masks=masks, topk=np.int64(T), height=np.int64(H),
         width=np.int64(W), **out)


def gen_two_stage():
    """Runs the reference's own crop_rois / match_label_crop / filter_labels_depth (lib/fcn/test_dataset.py:62-198)
    and mask_to_tight_box (lib/utils/mask.py:180-195): the functions are cut out of their files by ast (the modules
    import cv2, matplotlib, transforms3d ...) and executed with cfg stubbed to the two fields they read."""
    import ast
    import types
    lib = os.path.join(ref_shim.REF_ROOT, "lib")

    def cut(path, names):
        tree = ast.parse(open(path).read())
        return ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names],
                          type_ignores=[])

    util_ns = {"torch": torch, "np": np}
    exec(compile(cut(os.path.join(lib, "utils", "mask.py"),
                     {"mask_to_tight_box", "mask_to_tight_box_pytorch", "mask_to_tight_box_numpy"}), "mask.py", "exec"),
         util_ns)
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(SYN_CROP_SIZE=64), device="cpu")
    ns = {"torch": torch, "F": F, "np": np, "cfg": cfg, "util_": types.SimpleNamespace(**util_ns)}
    exec(compile(cut(os.path.join(lib, "fcn", "test_dataset.py"),
                     {"crop_rois", "match_label_crop", "filter_labels_depth"}), "test_dataset.py", "exec"), ns)
    import warnings
    warnings.simplefilter("ignore")
    out = {}
    for tag, with_depth in (("d", True), ("n", False)):
        rgb, labels, depth = two_stage_scene(31 if with_depth else 32, with_depth=with_depth)
        out[f"{tag}_rgb"], out[f"{tag}_labels"] = rgb, labels
        if with_depth:
            out["d_depth"] = depth
            filtered = ns["filter_labels_depth"](labels, depth, 0.5)
            out["d_filtered_05"] = filtered
            out["d_filtered_08"] = ns["filter_labels_depth"](labels, depth, 0.8)
            labels = filtered
        rgb_crops, mask_crops, rois, depth_crops = ns["crop_rois"](rgb, labels.clone(), depth)
        labels_crop = two_stage_crop_labels(mask_crops, 7)
        out[f"{tag}_rgb_crops"], out[f"{tag}_mask_crops"], out[f"{tag}_rois"] = rgb_crops, mask_crops, rois
        if with_depth:
            out["d_depth_crops"] = depth_crops
        out[f"{tag}_labels_crop_in"] = labels_crop.clone()
        refined, marked = ns["match_label_crop"](labels, labels_crop, mask_crops, rois, depth_crops)
        out[f"{tag}_refined"], out[f"{tag}_labels_crop_out"] = refined, marked
    save("two_stage", crop_size=np.int64(64), **out)


def gen_criterion():
    """Training losses: the REFERENCE's SetCriterion + HungarianMatcher (criterion.py, matcher.py) on a small
    three-layer prediction, with every random point set recorded in the fixture (the reference draws them per image
    in the matcher and per layer in the criterion; the device mirror draws them once per step, so parity is defined
    on equal points). Also the gradients of the summed losses wrt the final predictions."""
    crit = ref_shim.ref("modeling.criterion")
    mat = ref_shim.ref("modeling.matcher")
    torch.manual_seed(40)
    B, Q, K, h, w, H, W = 2, 10, 2, 24, 32, 96, 128
    layers, P, over, imp = 3, 50, 3.0, 0.75
    counts = [3, 2]
    N, S, R = sum(counts), int(P * over), P - int(imp * P)
    preds = [{"pred_logits": torch.randn(B, Q, K + 1), "pred_masks": torch.randn(B, Q, h, w) * 3} for _ in range(layers)]
    targets = []
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    for b, T in enumerate(counts):
        masks = torch.zeros(T, H, W, dtype=torch.bool)
        for t in range(T):
            cy, cx, r = 20 + 25 * t + 5 * b, 25 + 35 * t, 12 + 4 * t
            masks[t] = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        targets.append({"labels": torch.randint(0, K, (T,)), "masks": masks})
    # make one prediction per target resemble it, so that the assignment is not decided by noise alone
    for l in range(layers):
        for b, T in enumerate(counts):
            for t in range(T):
                small = F.interpolate(targets[b]["masks"][t][None, None].float(), size=(h, w), mode="bilinear")[0, 0]
                preds[l]["pred_masks"][b, (3 * t + l + b) % Q] += 8 * small - 4
    m_pts, cand, fill = torch.rand(layers, B, P, 2), torch.rand(layers, N, S, 2), torch.rand(layers, N, R, 2)
    cursor = {"m": 0, "c": 0, "f": 0}

    def replay(*shape, device=None):
        if shape == (1, P, 2):
            i = cursor["m"]; cursor["m"] += 1
            return m_pts[i // B, i % B][None].clone()
        if shape == (N, S, 2):
            cursor["c"] += 1
            return cand[cursor["c"] - 1].clone()
        if shape == (N, R, 2):
            cursor["f"] += 1
            return fill[cursor["f"] - 1].clone()
        raise AssertionError(f"unexpected draw {shape}")

    mat.torch = ref_shim.TorchProxy(replay)
    ref_shim.RAND[0] = replay
    try:
        matcher = mat.HungarianMatcher(cost_class=1.0, cost_mask=20.0, cost_dice=1.0, num_points=P)
        weight = {"loss_ce": 1.0, "loss_mask": 20.0, "loss_dice": 1.0}
        weight.update({k + f"_{i}": v for i in range(layers - 1) for k, v in list(weight.items())[:3]})
        c = crit.SetCriterion(K, matcher=matcher, weight_dict=weight, eos_coef=0.1, losses=["labels", "masks"],
                              num_points=P, oversample_ratio=over, importance_sample_ratio=imp)
        final = {k: v.clone().requires_grad_() for k, v in preds[0].items()}
        outputs = dict(final, aux_outputs=preds[1:])
        losses = c(outputs, targets)
        assert cursor == {"m": layers * B, "c": layers, "f": layers}, cursor
        g_logits, g_masks = torch.autograd.grad(sum(losses.values()), (final["pred_logits"], final["pred_masks"]))
        # the assignment of every layer, replayed on the same points
        cursor.update(m=0)
        indices = [matcher(p, targets) for p in preds]
    finally:
        mat.torch = torch
        ref_shim.RAND[0] = torch.rand
    arrays = {"matcher_points": m_pts, "candidate_points": cand, "fill_points": fill,
              "num_points": np.int64(P), "oversample_ratio": np.float64(over), "importance_sample_ratio": np.float64(imp),
              "grad_pred_logits": g_logits, "grad_pred_masks": g_masks, "loss_names": np.array(list(losses))}
    for l, p in enumerate(preds):
        arrays[f"pred_logits_{l}"], arrays[f"pred_masks_{l}"] = p["pred_logits"], p["pred_masks"]
        for b in range(B):
            arrays[f"match_{l}_{b}_pred"], arrays[f"match_{l}_{b}_tgt"] = indices[l][b]
    for b, t in enumerate(targets):
        arrays[f"labels_{b}"], arrays[f"masks_{b}"] = t["labels"], t["masks"]
    for k, v in losses.items():
        arrays["loss::" + k] = v
    save("criterion", **arrays)


if __name__ == "__main__":
    gen_hypersphere_attention()
    gen_hypersphere_attention_bwd()
    gen_meanshift_attention()
    gen_decoder_multiscale()
    gen_decoder_multiscale_bwd()
    gen_decoder_pretrained()
    gen_posenc()
    gen_msdeform_core()
    gen_msdeform_module()
    gen_pixel_decoder_msdeform()
    gen_pixel_decoder_simple()
    gen_head_r50style()
    gen_mean_shift()
    gen_mean_shift_d64()
    gen_instance_inference()
    gen_two_stage()
    gen_criterion()
