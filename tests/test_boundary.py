"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/msmformer_b200.h declares, the ctypes table matches the header, the product package never
imports the oracle, and a missing library / non-CUDA tensor fails loudly (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unseenobjectswithmeanshift_b200")
HEADER = os.path.join(ROOT, "include", "msmformer_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from unseenobjectswithmeanshift_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/msmformer_b200.h but not exported"


def test_ctypes_table_matches_header(built):
    assert sorted(built.SIGNATURES) == header_symbols()
    assert built.lib().msm_abi_version() == 1


def test_argument_errors_need_no_gpu(built):
    L = built.lib()
    # null pointers are rejected before any CUDA call
    rc = L.msm_mask_logits(None, None, None, 1, 1, 1, 1, None)
    assert rc == -1
    assert b"non-null" in L.msm_last_error()
    assert L.msm_vmf_attention_workspace_bytes(1, 8, 100, 4800, 32) > 0
    assert L.msm_mean_shift_workspace_bytes(1, 307200, 100, 64) > 0


def test_product_package_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product code imports the oracle: {bad}"


def test_cpu_tensors_are_rejected(built):
    from unseenobjectswithmeanshift_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.mask_logits(torch.zeros(1, 2, 4), torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.mean_shift_hill_climb(torch.zeros(8, 4), torch.zeros(2, 4), 10.0)
    # the "next" rows (clusterer, eval tail, two-stage glue) have no CPU path either
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.select_smart_seeds(torch.zeros(8, 64), 3, [0])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.seed_connected_components(torch.zeros(4, 64), 0.04)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.instance_topk(torch.zeros(1, 10, 3), 5)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.label_stats(torch.zeros(1, 8, 8), num_ids=4)


def test_new_entry_points_reject_bad_arguments(built):
    """argument checks of the clusterer / eval tail / two-stage entry points run before any CUDA call."""
    L = built.lib()
    assert L.msm_select_smart_seeds(None, None, None, None, 1, 8, 64, 4, None, 0, None) == -1
    assert b"non-null" in L.msm_last_error()
    assert L.msm_seed_connected_components(None, None, None, 1, 4, 64, 0.04, None) == -1
    assert L.msm_instance_topk(None, None, None, None, 1, 10, 3, 5, None) == -1
    assert L.msm_label_stats(None, None, 0, None, 1, 8, 8, 4, None) == -1
    assert L.msm_smart_seeds_workspace_bytes(32, 307200, 100) >= 32 * 307200 * 4 + 32 * 100 * 8 + 32 * 4
    assert L.msm_smart_seeds_workspace_bytes(0, 10, 10) == 0
    assert L.msm_assign_clusters_workspace_bytes(32, 100) >= 32 * 100 * 4
    assert L.msm_instance_masks_workspace_bytes(8, 20, 480) >= 8 * 20 * 30 * 24


def test_missing_library_is_an_import_error(built, monkeypatch):
    monkeypatch.setattr(built, "_lib", None)
    monkeypatch.setattr(built, "LIB_PATH", os.path.join(PKG, "lib", "nope.so"))
    with pytest.raises(ImportError, match="no PyTorch/CPU fallback"):
        built.lib()


def test_registries_and_state_dict_layout():
    from unseenobjectswithmeanshift_b200 import workloads
    from unseenobjectswithmeanshift_b200.d2compat import SEM_SEG_HEADS_REGISTRY
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder.maskformer_transformer_decoder \
        import TRANSFORMER_DECODER_REGISTRY
    for n in ("MeanShiftTransformerDecoder", "PretrainedMeanShiftTransformerDecoder"):
        assert TRANSFORMER_DECODER_REGISTRY.get(n) is not None
    for n in ("PretrainedMeanShiftMaskFormerHead", "MeanShiftMaskFormerHead", "MSDeformAttnPixelDecoder",
              "SimpleBasePixelDecoder"):
        assert SEM_SEG_HEADS_REGISTRY.get(n) is not None
    sd = workloads.build_head("r50").state_dict()
    for k, shape in {
        "predictor.transformer_cross_attention_layers.0.meanshift_attn.in_proj_weight": (768, 256),
        "predictor.transformer_self_attention_layers.8.self_attn.out_proj.weight": (256, 256),
        "predictor.transformer_ffn_layers.0.linear1.weight": (2048, 256),
        "predictor.query_feat.weight": (100, 256),
        "predictor.level_embed.weight": (3, 256),
        "predictor.input_proj.0.weight": (256, 64, 1, 1),
        "predictor.mask_embed.layers.2.weight": (256, 256),
        "pixel_decoder.transformer.encoder.layers.5.self_attn.sampling_offsets.weight": (192, 64),
        "pixel_decoder.input_proj.0.0.weight": (64, 2048, 1, 1),
        "pixel_decoder.mask_features.weight": (256, 64, 1, 1),
    }.items():
        assert tuple(sd[k].shape) == shape, k


def test_meta_arch_registry_and_image_batching():
    """META_ARCH wrappers register under the reference's names; ImageList-style batching pads bottom / right."""
    from unseenobjectswithmeanshift_b200 import meanshiftformer as mf
    from unseenobjectswithmeanshift_b200.d2compat import META_ARCH_REGISTRY
    from unseenobjectswithmeanshift_b200.meanshiftformer.meanshiftformer_model import _batch_images
    from unseenobjectswithmeanshift_b200.meanshiftformer import pretrained_meanshiftformer_model as pm
    assert META_ARCH_REGISTRY.get("MeanShiftMaskFormer") is mf.MeanShiftMaskFormer
    assert pm.PretrainedMeanShiftMaskFormer is META_ARCH_REGISTRY.get("PretrainedMeanShiftMaskFormer")
    a, b = torch.ones(3, 30, 40), 2 * torch.ones(3, 20, 50)
    batch, sizes = _batch_images([a, b], 32)
    assert batch.shape == (2, 3, 32, 64) and sizes == [(30, 40), (20, 50)]
    assert float(batch[0, :, :30, :40].min()) == 1 and float(batch[0, :, 30:].abs().sum()) == 0
    assert float(batch[1, :, :20, :50].min()) == 2 and float(batch[1, :, :, 50:].abs().sum()) == 0
    same, sizes = _batch_images([a, a], 0)
    assert same.shape == (2, 3, 30, 40) and sizes == [(30, 40)] * 2


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/msmformer_b200.h must compile as C99 with nothing but libc headers."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "msmformer_b200.h"\nint main(void) { return msm_abi_version() == MSM_ABI_VERSION ? 0 : 1; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(root, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
