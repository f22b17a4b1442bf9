"""The REFERENCE's own PyTorch modules, built for a benchmark workload. Test infrastructure only (bench.py's
``--impl reference`` arm and ``cpu_baseline`` leg; tests): never imported by the product package.

The modules are imported by file path from ``baseline/_ref/reference`` (a git-ignored copy of the reference's python
files made by ``oracle/make_ref.py`` in the authoring container; it travels to the GPU box with the snapshot) or, when
present, straight from ``/root/reference`` - under the detectron2 / fvcore stubs of ``tests/golden/ref_shim.py``.
On CPU ``MSDeformAttn.forward`` takes the reference's own pure-PyTorch path (``ms_deform_attn_core_pytorch``,
ops/modules/ms_deform_attn.py:116-121): exactly what the reference does on a machine without its CUDA extension.
The eval tail (upsample + ``instance_inference``) needs detectron2's ``Instances`` / ``BitMasks`` and is therefore the
oracle's restatement (``oracle/head.py::eval_tail``).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """Directory that holds ``MSMFormer/meanshiftformer`` of the reference, or None."""
    for p in (os.path.join(ROOT, "baseline", "_ref", "reference"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "MSMFormer", "meanshiftformer", "modeling")):
            return p
    return None


def _shim():
    root = reference_root()
    if root is None:
        raise FileNotFoundError("no reference copy: run `python oracle/make_ref.py` where /root/reference exists")
    os.environ["MSM_REFERENCE_ROOT"] = root
    g = os.path.join(ROOT, "tests", "golden")
    if g not in sys.path:
        sys.path.insert(0, g)
    import ref_shim
    if os.path.normpath(ref_shim.REF_ROOT) != os.path.normpath(root):  # imported earlier with another root
        ref_shim.REF_ROOT = root
        ref_shim.REF_PKG = os.path.join(root, "MSMFormer", "meanshiftformer")
    return ref_shim


def build_reference_head(kind, state_dict):
    """``PretrainedMeanShiftMaskFormerHead`` of the REFERENCE for workload ``kind`` (r50 / ucn / crop, see
    unseenobjectswithmeanshift_b200/workloads.py) with ``state_dict`` loaded strictly. CPU, eval mode."""
    from unseenobjectswithmeanshift_b200 import workloads
    shim = _shim()
    cfg = workloads.HEAD_CFG[kind]
    dec = shim.ref("modeling.transformer_decoder.meanshiftformer_transformer_decoder")
    head = shim.ref("modeling.meta_arch.meanshift_former_head")
    from detectron2.layers import ShapeSpec  # the stub installed by the shim
    if cfg["pixel_decoder"] == "MSDeformAttnPixelDecoder":
        pd = shim.ref("modeling.pixel_decoder.msdeformattn")
        shapes = {k: ShapeSpec(channels=s.channels, stride=s.stride) for k, s in workloads.R50_SHAPES.items()}
        pixel = pd.MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8,
                                            transformer_dim_feedforward=1024, transformer_enc_layers=6, conv_dim=64,
                                            mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                                            common_stride=4)
    else:
        fpn = shim.ref("modeling.pixel_decoder.fpn")
        shapes = {"res5": ShapeSpec(channels=64, stride=1)}
        pixel = fpn.SimpleBasePixelDecoder(shapes, conv_dim=64, mask_dim=256, norm="GN")
    predictor = getattr(dec, cfg["decoder"])(64, True, **workloads.decoder_kwargs(cfg["dec_layers"]))
    m = head.PretrainedMeanShiftMaskFormerHead(shapes, num_classes=2, pixel_decoder=pixel, loss_weight=1.0,
                                               ignore_value=255, transformer_predictor=predictor,
                                               transformer_in_feature="multi_scale_pixel_decoder")
    m.load_state_dict(state_dict, strict=True)
    return m.eval()


def reference_mean_shift():
    """The reference's ``transformer_decoder/mean_shift.py`` module (needs no stubs)."""
    return _shim().ref("modeling.transformer_decoder.mean_shift")
