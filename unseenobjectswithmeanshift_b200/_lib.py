"""ctypes binding of libmsmformer_b200.so (the C ABI in include/msmformer_b200.h).

The library is loaded lazily and loudly: a missing .so is an ImportError telling the user to run
``python __graft_entry__.py build`` - the package never falls back to PyTorch or CPU code.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmsmformer_b200.so")

_lib = None

# name -> (restype, argtypes); kept in the order of include/msmformer_b200.h
_P, _I, _L, _F, _Z = c_void_p, c_int, c_int64, c_float, c_size_t
SIGNATURES = {
    "msm_abi_version": (_I, []),
    "msm_last_error": (ctypes.c_char_p, []),
    "msm_device_arch": (_I, []),
    "msm_vmf_attention_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "msm_vmf_attention_fwd": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P,
                                   _P, _I, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "msm_vmf_attention_weights": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _P, _P, _P,
                                       _I, _I, _I, _I, _I, _F, _I, _P]),
    "msm_vmf_attention_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "msm_vmf_attention_bwd": (_I, [_P, _L, _L, _L] * 5 + [_P] + [_P, _L, _L, _L] * 3 +
                              [_P, _I, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "msm_mask_logits": (_I, [_P, _P, _P, _I, _I, _I, _L, _P]),
    "msm_mask_to_attn_bits": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "msm_resample_bilinear_fwd": (_I, [_P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "msm_upsample_add_fwd": (_I, [_P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "msm_linear_weight_bytes": (_Z, [_I, _I]),
    "msm_linear_prepare_weight": (_I, [_P, _L, _P, _I, _I, _P]),
    "msm_linear_prepare_weight_t": (_I, [_P, _L, _P, _I, _I, _P]),
    "msm_linear_fwd": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "msm_linear_ln_fwd": (_I, [_P, _L, _P, _P, _P, _L, _P, _P, _F, _P, _L, _I, _I, _I, _P]),
    "msm_linear_fused_fwd": (_I, [_P, _L, _P, _P, _P, _I, _I, _P, _L, _P, _P, _F, _I, _P, _P, _F, _P, _L, _P, _L,
                                  _I, _I, _I, _P]),
    "msm_ffn_ln_fwd": (_I, [_P, _L, _P, _P, _P, _P, _P, _P, _F, _P, _L, _I, _I, _I, _P]),
    "msm_add_layernorm_fwd": (_I, [_P, _P, _P, _P, _F, _I, _P, _P, _F, _P, _P, _I, _I, _P]),
    "msm_maxpool3x3s2_nhwc_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "msm_conv1x1_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "msm_conv1x1_nhwc_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "msm_conv3x3_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msm_ms_deform_attn_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msm_ms_deform_attn_fused_fwd": (_I, [_P, _P, _P, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msm_ms_deform_attn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msm_decoder_block_weight_bytes": (_Z, [_I]),
    "msm_decoder_block_fwd": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _F, _I, _P, _P, _F,
                                   _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "msm_mean_shift_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "msm_mean_shift_hill_climb": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "msm_smart_seeds_workspace_bytes": (_Z, [_I, _I, _I]),
    "msm_select_smart_seeds": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "msm_seed_connected_components": (_I, [_P, _P, _P, _I, _I, _I, _F, _P]),
    "msm_assign_clusters_workspace_bytes": (_Z, [_I, _I]),
    "msm_assign_clusters": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "msm_instance_topk": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "msm_instance_masks_workspace_bytes": (_Z, [_I, _I, _I]),
    "msm_instance_masks": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "msm_instance_label_map": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "msm_label_stats": (_I, [_P, _P, _L, _P, _I, _I, _I, _I, _P]),
    "msm_relabel_lut": (_I, [_P, _P, _P, _I, _L, _I, _I, _P]),
    "msm_crop_resize": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "msm_crop_label_stats": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "msm_paste_crops": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
}


# EXPERIMENTAL entry points (prefix msmx_): exported by the library but not part of include/msmformer_b200.h; bound on
# first use by ops.py and only reachable with MSM_PACKED_KV=1
X_SIGNATURES = {
    "msmx_vmf_packed_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "msmx_vmf_packed_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "msmx_vmf_pack": (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _I, _I, _I, _I, _I, _P]),
    "msmx_vmf_attention_packed_fwd": (_I, [_P, _L, _L, _L, _P, _P, _L, _L, _L, _P, _I, _P, _I, _I, _I, _I, _I, _F, _I,
                                           _P, _Z, _P]),
    "msmx_linear_packed_kv_fwd": (_I, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msmx_linear_packed_kv_pos_fwd": (_I, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "msmx_set_l2_persisting_window": (_I, [_P, _Z, _P]),
    "msmx_vmf_attention_small_fwd": (_I, [_P, _L, _L, _L] * 4 + [_P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _I, _P]),
    "msmx_mean_shift_packed_bytes": (_Z, [_I, _I, _I]),
    "msmx_mean_shift_packed_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "msmx_mean_shift_pack": (_I, [_P, _P, _I, _I, _I, _P]),
    "msmx_mean_shift_hill_climb_packed": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "msmx_mean_shift_persistent_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "msmx_mean_shift_hill_climb_persistent": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
}
_xbound = False


def xlib():
    """lib() with the experimental entry points bound as well."""
    global _xbound
    handle = lib()
    if not _xbound:
        for name, (res, args) in X_SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _xbound = True
    return handle


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA library first (python __graft_entry__.py build). "
                "unseenobjectswithmeanshift_b200 has no PyTorch/CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and the binding disagree
            fn.restype = res
            fn.argtypes = args
        if handle.msm_abi_version() != 1:
            raise ImportError(f"{LIB_PATH}: ABI version {handle.msm_abi_version()} != 1, rebuild")
        _lib = handle
    return _lib


class MsmError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().msm_last_error().decode("utf-8", "replace")
        raise MsmError(f"{what} failed (code {rc}): {msg}")
