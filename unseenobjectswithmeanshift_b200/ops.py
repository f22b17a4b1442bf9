"""torch.Tensor front-ends of the C ABI (include/msmformer_b200.h).

PyTorch here is plumbing: it owns device memory and the CUDA stream; every function below hands
raw device pointers + sizes to libmsmformer_b200.so and enqueues on torch's current stream.
All functions require CUDA fp32 tensors and raise otherwise - there is no CPU path.
"""
import os

import torch

from . import _lib
from ._lib import check

KAPPA = 30.0  # reference: transformer_decoder/attention_util.py:26


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: unseenobjectswithmeanshift_b200 has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.requires_grad and torch.is_grad_enabled():
        raise RuntimeError(f"{name} requires grad: this op is forward-only, call it under torch.no_grad() (the "
                           "differentiable entry points are vmf_attention_autograd, mask_logits_autograd and "
                           "MSDeformAttnFunction)")
    return t


def _bhld(t, name):
    """4-D view [B, H, L, hd] with unit stride on the last axis -> (tensor, sb, sh, sl)."""
    _require(t, name)
    if t.dim() != 4:
        raise ValueError(f"{name} must be a [B, H, L, hd] view")
    if t.shape[3] > 1 and t.stride(3) != 1:
        raise ValueError(f"{name}: channel axis must be contiguous")
    return t, t.stride(0), t.stride(1), t.stride(2)


def vmf_attention(q, k, v, *, blocked_bits=None, row_open=None, add_mask=None, kappa=KAPPA,
                  normalize_q=True, normalize_k=True, out=None, return_den=False, save_norm=False):
    """vMF attention core on [B, H, L, hd] *views* (any batch/head/row strides, see msm_vmf_attention_fwd).

    blocked_bits int32 [B, Nq, ceil(Ns/32)] (bit set = blocked), row_open int32 [B, Nq];
    or add_mask float [B*H, Nq, Ns]. Returns out [B, H, Nq, hd] view of a [B, Nq, H*hd] buffer
    (and den [B*H, Nq] if asked; with ``save_norm`` den is [2, B*H, Nq]: denominators, then |softmax . v| - what
    vmf_attention_bwd needs)."""
    q, q_sb, q_sh, q_sl = _bhld(q, "q")
    k, k_sb, k_sh, k_sl = _bhld(k, "k")
    v, v_sb, v_sh, v_sl = _bhld(v, "v")
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    if k.shape != (B, H, Ns, hd) or v.shape != (B, H, Ns, hd):
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    if out is None:
        out = torch.empty(B, Nq, H, hd, device=q.device, dtype=torch.float32).permute(0, 2, 1, 3)
    out, o_sb, o_sh, o_sl = _bhld(out, "out")
    if save_norm and not return_den:
        raise ValueError("save_norm needs return_den=True")
    den = torch.empty((2, B * H, Nq) if save_norm else (B * H, Nq), device=q.device,
                      dtype=torch.float32) if return_den else None
    wpr = 0
    if blocked_bits is not None:
        _require(blocked_bits, "blocked_bits", torch.int32)
        if not blocked_bits.is_contiguous() or blocked_bits.shape[:2] != (B, Nq):
            raise ValueError("blocked_bits must be contiguous [B, Nq, words]")
        wpr = blocked_bits.shape[2]
        if row_open is not None:
            _require(row_open, "row_open", torch.int32)
    if add_mask is not None:
        _require(add_mask, "add_mask")
        if not add_mask.is_contiguous() or tuple(add_mask.shape) != (B * H, Nq, Ns):
            raise ValueError("add_mask must be contiguous [B*H, Nq, Ns]")
    L = _lib.lib()
    ws_bytes = L.msm_vmf_attention_workspace_bytes(B, H, Nq, Ns, hd)
    ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
    flags = (1 if normalize_q else 0) | (2 if normalize_k else 0) | (4 if save_norm else 0)
    rc = L.msm_vmf_attention_fwd(
        q.data_ptr(), q_sb, q_sh, q_sl, k.data_ptr(), k_sb, k_sh, k_sl, v.data_ptr(), v_sb, v_sh, v_sl,
        out.data_ptr(), o_sb, o_sh, o_sl, den.data_ptr() if den is not None else None,
        blocked_bits.data_ptr() if blocked_bits is not None else None, wpr,
        row_open.data_ptr() if row_open is not None else None,
        add_mask.data_ptr() if add_mask is not None else None,
        B, H, Nq, Ns, hd, float(kappa), flags, ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msm_vmf_attention_fwd")
    return (out, den) if return_den else out


def vmf_attention_bwd(q, k, v, out, grad_out, den, *, blocked_bits=None, row_open=None, add_mask=None, kappa=KAPPA,
                      normalize_q=True, normalize_k=True):
    """Gradients of the vMF attention core wrt q, k, v (msm_vmf_attention_bwd): [B, H, L, hd] views as in
    vmf_attention, ``out`` / ``den`` from the forward called with return_den=True, save_norm=True and the same
    masks, kappa and flags. Returns (grad_q, grad_k, grad_v) shaped like q, k, v ([B, H, L, hd] views of
    [B, L, H*hd] buffers)."""
    q, q_sb, q_sh, q_sl = _bhld(q, "q")
    k, k_sb, k_sh, k_sl = _bhld(k, "k")
    v, v_sb, v_sh, v_sl = _bhld(v, "v")
    out, o_sb, o_sh, o_sl = _bhld(out, "out")
    if grad_out.dim() == 4 and grad_out.shape[3] > 1 and grad_out.stride(3) != 1:
        grad_out = grad_out.contiguous()
    grad_out, go_sb, go_sh, go_sl = _bhld(grad_out, "grad_out")
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    if k.shape != (B, H, Ns, hd) or v.shape != (B, H, Ns, hd) or out.shape != q.shape or grad_out.shape != q.shape:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)} "
                         f"out {tuple(out.shape)} grad_out {tuple(grad_out.shape)}")
    _require(den, "den")
    if tuple(den.shape) != (2, B * H, Nq) or not den.is_contiguous():
        raise ValueError("den must be the contiguous [2, B*H, Nq] buffer of a forward run with save_norm=True")
    wpr = 0
    if blocked_bits is not None:
        _require(blocked_bits, "blocked_bits", torch.int32)
        if not blocked_bits.is_contiguous() or blocked_bits.shape[:2] != (B, Nq):
            raise ValueError("blocked_bits must be contiguous [B, Nq, words]")
        wpr = blocked_bits.shape[2]
        if row_open is not None:
            _require(row_open, "row_open", torch.int32)
    if add_mask is not None:
        _require(add_mask, "add_mask")
        if not add_mask.is_contiguous() or tuple(add_mask.shape) != (B * H, Nq, Ns):
            raise ValueError("add_mask must be contiguous [B*H, Nq, Ns]")

    def grad_like(L_):
        return torch.empty(B, L_, H, hd, device=q.device, dtype=torch.float32).permute(0, 2, 1, 3)

    gq, gk, gv = grad_like(Nq), grad_like(Ns), grad_like(Ns)
    lib = _lib.lib()
    ws_bytes = lib.msm_vmf_attention_bwd_workspace_bytes(B, H, Nq, Ns, hd)
    ws = torch.empty(max(ws_bytes, 1), device=q.device, dtype=torch.uint8)
    flags = (1 if normalize_q else 0) | (2 if normalize_k else 0)
    rc = lib.msm_vmf_attention_bwd(
        q.data_ptr(), q_sb, q_sh, q_sl, k.data_ptr(), k_sb, k_sh, k_sl, v.data_ptr(), v_sb, v_sh, v_sl,
        out.data_ptr(), o_sb, o_sh, o_sl, grad_out.data_ptr(), go_sb, go_sh, go_sl, den.data_ptr(),
        gq.data_ptr(), gq.stride(0), gq.stride(1), gq.stride(2),
        gk.data_ptr(), gk.stride(0), gk.stride(1), gk.stride(2),
        gv.data_ptr(), gv.stride(0), gv.stride(1), gv.stride(2),
        blocked_bits.data_ptr() if blocked_bits is not None else None, wpr,
        row_open.data_ptr() if row_open is not None else None,
        add_mask.data_ptr() if add_mask is not None else None,
        B, H, Nq, Ns, hd, float(kappa), flags, ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msm_vmf_attention_bwd")
    return gq, gk, gv


class VmfAttentionFunction(torch.autograd.Function):
    """Differentiable vMF attention core: forward = msm_vmf_attention_fwd (keeps two floats per query row),
    backward = msm_vmf_attention_bwd (weights recomputed per key tile). The reference gets the same gradients
    from torch.autograd through attention_util.py:64-82, saving three [G, Nq, Ns] tensors per call."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)  # fp32 kernels inside an autocast region
    def forward(ctx, q, k, v, blocked_bits, row_open, add_mask, kappa, normalize_q, normalize_k):
        q, k, v = q.detach(), k.detach(), v.detach()
        out, den = vmf_attention(q, k, v, blocked_bits=blocked_bits, row_open=row_open, add_mask=add_mask,
                                 kappa=kappa, normalize_q=normalize_q, normalize_k=normalize_k, return_den=True,
                                 save_norm=True)
        ctx.save_for_backward(q, k, v, out, den)
        ctx.masks = (blocked_bits, row_open, add_mask)
        ctx.conf = (float(kappa), bool(normalize_q), bool(normalize_k))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        q, k, v, out, den = ctx.saved_tensors
        bits, row_open, add_mask = ctx.masks
        kappa, nq, nk = ctx.conf
        gq, gk, gv = vmf_attention_bwd(q, k, v, out, grad_out, den, blocked_bits=bits, row_open=row_open,
                                       add_mask=add_mask, kappa=kappa, normalize_q=nq, normalize_k=nk)
        return gq, gk, gv, None, None, None, None, None, None


def vmf_attention_autograd(q, k, v, *, blocked_bits=None, row_open=None, add_mask=None, kappa=KAPPA,
                           normalize_q=True, normalize_k=True):
    """vmf_attention for training: same [B, H, L, hd] views in, out [B, H, Nq, hd] (view of a [B, Nq, H*hd]
    buffer), gradients flow to q, k, v. Masks are constants (the reference detaches its attention mask,
    meanshiftformer_transformer_decoder.py:680)."""
    return VmfAttentionFunction.apply(q, k, v, blocked_bits, row_open, add_mask, kappa, normalize_q, normalize_k)


# ----------------------------------------------------------------------------------------------
# Packed-operand cross-attention (csrc/vmf_attention_packed.cu, linear_tc_kernel<1|2>; the default since round 2): K / V
# projections write 16-bit operand images instead of fp32 rows, the attention kernel streams them with bulk copies
# (DESIGN.md sections 4.2, 4.3). Entry points carry the prefix msmx_ (not in the public header).
# ----------------------------------------------------------------------------------------------
def packed_kv_enabled():
    """Default ON since round 2 (B200: parity 2e-6 against the fp32-row kernel; UCN head step 16.1 -> 13.8 ms, R50-level
    chains 2x): K / V projections write TMA-streamable operand images. MSM_PACKED_KV=0 selects the fp32-row path."""
    return os.environ.get("MSM_PACKED_KV", "1") == "1"


class PackedKV:
    """Operand images of the K and V projections of ONE decoder layer: uint8 [B][H][ceil(S/128)][4][8192] view."""

    def __init__(self, images, batch, heads, num_keys):
        self.images, self.batch, self.heads, self.num_keys = images, batch, heads, num_keys


def packed_kv_alloc(layers, batch, heads, num_keys, device):
    """Image buffer for `layers` decoder layers -> (buffer, bytes/layer). A FRESH buffer per forward: under CUDA-graph
    capture it belongs to that graph's pool, so several graphs of the same module can be in flight at once (a buffer
    cached on the module was shared by all of them - found by test_three_graphs_in_flight_equal_serial_replay). Only the
    last 128-key tile of every (layer, image, head) is zeroed (the projections never write the key tail; it must read
    as zero); with num_keys % 128 == 0 nothing is."""
    per_layer = _lib.xlib().msmx_vmf_packed_bytes(batch, heads, num_keys, 32, 3)
    buf = torch.empty(layers * per_layer, dtype=torch.uint8, device=device)
    tiles = (num_keys + 127) // 128
    if num_keys % 128 and per_layer == batch * heads * tiles * 32768:
        buf.view(layers * batch * heads, tiles, 32768)[:, -1].zero_()
    elif num_keys % 128:
        buf.zero_()
    return buf, per_layer


def linear_packed_kv(x, weight, bias, images, batch, num_keys, channels, which, pos=None):
    """K (which=0: rows L2-normalised per 32-channel head, fp16 halves) or V (which=1: bf16 halves) projection of
    x with weight [layers*C, Cin], written as operand images into `images` (packed_kv_alloc). x: token-major
    [B, S, Cin] or the channel-major map [B, Cin, h, w] itself (h * w = S).
    ``pos`` = (ty [h, layers*C], tx [w, layers*C]): row (b, y*w + x) gets + ty[y] + tx[x] before the normalisation -
    the separable sine embedding pushed through the key weights."""
    x = _require(x, "x")
    nchw = x.dim() == 4
    w = _require(weight, "weight")
    N, K = w.shape
    if nchw:
        ok = x.is_contiguous() and x.shape[0] == batch and x.shape[2] * x.shape[3] == num_keys and num_keys % 4 == 0
        cin = x.shape[1]
    else:
        ok = x.dim() == 3 and x.is_contiguous() and x.shape[0] == batch and x.shape[1] == num_keys
        cin = x.shape[-1]
    if not ok:
        raise ValueError("x must be a contiguous [B, S, Cin] or [B, Cin, h, w] (h * w = S, S % 4 == 0) tensor")
    if K != cin or N % channels or channels % 32 or K % 32:
        raise ValueError(f"weight {tuple(w.shape)} does not fit x {tuple(x.shape)} / {channels} channels per layer")
    b = None if bias is None else _require(bias, "bias").contiguous()
    ty = tx = None
    Wd = 1
    if pos is not None:
        ty, tx = (_require(t, "pos table") for t in pos)
        Wd = tx.shape[0]
        if (not ty.is_contiguous() or not tx.is_contiguous() or ty.shape[1] != N or tx.shape[1] != N
                or ty.shape[0] * Wd != num_keys):
            raise ValueError(f"pos tables {tuple(ty.shape)} / {tuple(tx.shape)} do not fit {num_keys} keys x {N} outputs")
    if pos is None and not nchw:
        rc = _lib.xlib().msmx_linear_packed_kv_fwd(x.data_ptr(), K, prepare_linear_weight(w).data_ptr(),
                                                   b.data_ptr() if b is not None else None, images.data_ptr(), batch,
                                                   num_keys, N, K, channels, int(which), 1, 1, _stream())
        check(rc, "msmx_linear_packed_kv_fwd")
        return
    rc = _lib.xlib().msmx_linear_packed_kv_pos_fwd(x.data_ptr(), K, prepare_linear_weight(w).data_ptr(),
                                                   b.data_ptr() if b is not None else None, images.data_ptr(), batch,
                                                   num_keys, N, K, channels, int(which), 1, 1, 1 if nchw else 0,
                                                   ty.data_ptr() if ty is not None else None,
                                                   tx.data_ptr() if tx is not None else None, Wd, _stream())
    check(rc, "msmx_linear_packed_kv_pos_fwd")


def pack_kv(k, v, normalize_k=True):
    """Stand-alone pack pass: K, V [B, H, S, hd] views (hd 32 or 64) -> PackedKV operand images (K unit-normalised and
    split to fp16 halves, V to bf16 halves). The decoder does not need it - its projections write the images directly
    (linear_packed_kv); this is for callers that hold fp32 K / V rows."""
    k, k_sb, k_sh, k_sl = _bhld(k, "k")
    v, v_sb, v_sh, v_sl = _bhld(v, "v")
    B, H, S, hd = k.shape
    if v.shape != k.shape or hd not in (32, 64):
        raise ValueError(f"k {tuple(k.shape)} / v {tuple(v.shape)}: equal shapes with hd in (32, 64)")
    flags = 1 | (2 if normalize_k else 0)
    X = _lib.xlib()
    images = torch.empty(X.msmx_vmf_packed_bytes(B, H, S, hd, flags), dtype=torch.uint8, device=k.device)
    check(X.msmx_vmf_pack(k.data_ptr(), k_sb, k_sh, k_sl, v.data_ptr(), v_sb, v_sh, v_sl, images.data_ptr(), B, H, S, hd,
                          flags, _stream()), "msmx_vmf_pack")
    return PackedKV(images, B, H, S)


def vmf_attention_packed(q, kv, *, blocked_bits=None, row_open=None, kappa=KAPPA, out=None):
    """vmf_attention with K / V given as operand images (PackedKV): q [B, H, Nq, 32] view, q and k normalised."""
    q, q_sb, q_sh, q_sl = _bhld(q, "q")
    B, H, Nq, hd = q.shape
    if hd != 32 or B != kv.batch or H != kv.heads:
        raise ValueError(f"q {tuple(q.shape)} does not match the packed K / V ({kv.batch} x {kv.heads} heads of 32)")
    if out is None:
        out = torch.empty(B, Nq, H, hd, device=q.device, dtype=torch.float32).permute(0, 2, 1, 3)
    out, o_sb, o_sh, o_sl = _bhld(out, "out")
    wpr = 0
    if blocked_bits is not None:
        _require(blocked_bits, "blocked_bits", torch.int32)
        if not blocked_bits.is_contiguous() or blocked_bits.shape[:2] != (B, Nq):
            raise ValueError("blocked_bits must be contiguous [B, Nq, words]")
        wpr = blocked_bits.shape[2]
        if row_open is not None:
            _require(row_open, "row_open", torch.int32)
    L = _lib.xlib()
    ws_bytes = L.msmx_vmf_packed_workspace_bytes(B, H, Nq, kv.num_keys, hd)
    ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
    rc = L.msmx_vmf_attention_packed_fwd(
        q.data_ptr(), q_sb, q_sh, q_sl, kv.images.data_ptr(), out.data_ptr(), o_sb, o_sh, o_sl,
        blocked_bits.data_ptr() if blocked_bits is not None else None, wpr,
        row_open.data_ptr() if row_open is not None else None, B, H, Nq, kv.num_keys, hd, float(kappa), 3,
        ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msmx_vmf_attention_packed_fwd")
    return out


def vmf_attention_small(q, k, v, *, blocked_bits=None, row_open=None, kappa=KAPPA, normalize_q=True, normalize_k=True,
                        out=None, return_den=False, save_norm=False):
    """Opt-in (measured slower in the R50 step): vmf_attention through the single-launch CUDA-core kernel for short key
    sequences (csrc/vmf_attention_small.cu; hd 32, Nq <= 128, Ns <= 1024). MSM_SMALL_ATTN=1 makes msm_vmf_attention_fwd
    pick it by itself; this front-end calls it directly."""
    q, q_sb, q_sh, q_sl = _bhld(q, "q")
    k, k_sb, k_sh, k_sl = _bhld(k, "k")
    v, v_sb, v_sh, v_sl = _bhld(v, "v")
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    if out is None:
        out = torch.empty(B, Nq, H, hd, device=q.device, dtype=torch.float32).permute(0, 2, 1, 3)
    out, o_sb, o_sh, o_sl = _bhld(out, "out")
    den = torch.empty((2, B * H, Nq) if save_norm else (B * H, Nq), device=q.device,
                      dtype=torch.float32) if return_den else None
    wpr = blocked_bits.shape[2] if blocked_bits is not None else 0
    flags = (1 if normalize_q else 0) | (2 if normalize_k else 0) | (4 if save_norm else 0)
    rc = _lib.xlib().msmx_vmf_attention_small_fwd(
        q.data_ptr(), q_sb, q_sh, q_sl, k.data_ptr(), k_sb, k_sh, k_sl, v.data_ptr(), v_sb, v_sh, v_sl,
        out.data_ptr(), o_sb, o_sh, o_sl, den.data_ptr() if den is not None else None,
        _require(blocked_bits, "blocked_bits", torch.int32).data_ptr() if blocked_bits is not None else None, wpr,
        _require(row_open, "row_open", torch.int32).data_ptr() if row_open is not None else None,
        B, H, Nq, Ns, hd, float(kappa), flags, _stream())
    check(rc, "msmx_vmf_attention_small_fwd")
    return (out, den) if return_den else out


def l2_persist_enabled():
    return os.environ.get("MSM_L2_PERSIST", "0") == "1"


def l2_persist(tensor=None):
    """Opt-in (MSM_L2_PERSIST=1; measured slower): L2 persisting access window on `tensor` for the
    kernels launched on the current stream from now on; None clears it."""
    if tensor is None:
        rc = _lib.xlib().msmx_set_l2_persisting_window(None, 0, _stream())
    else:
        rc = _lib.xlib().msmx_set_l2_persisting_window(tensor.data_ptr(), tensor.numel() * tensor.element_size(),
                                                       _stream())
    check(rc, "msmx_set_l2_persisting_window")


def vmf_attention_weights(q, k, den, *, blocked_bits=None, row_open=None, add_mask=None, kappa=KAPPA,
                          normalize_q=True, normalize_k=True):
    """Attention weights [B*H, Nq, Ns] (the second value hypersphere_attention returns)."""
    q, q_sb, q_sh, q_sl = _bhld(q, "q")
    k, k_sb, k_sh, k_sl = _bhld(k, "k")
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    attn = torch.empty(B * H, Nq, Ns, device=q.device, dtype=torch.float32)
    flags = (1 if normalize_q else 0) | (2 if normalize_k else 0)
    rc = _lib.lib().msm_vmf_attention_weights(
        q.data_ptr(), q_sb, q_sh, q_sl, k.data_ptr(), k_sb, k_sh, k_sl, _require(den, "den").data_ptr(),
        blocked_bits.data_ptr() if blocked_bits is not None else None,
        blocked_bits.shape[2] if blocked_bits is not None else 0,
        row_open.data_ptr() if row_open is not None else None,
        add_mask.data_ptr() if add_mask is not None else None,
        attn.data_ptr(), B, H, Nq, Ns, hd, float(kappa), flags, _stream())
    check(rc, "msm_vmf_attention_weights")
    return attn


def mask_logits(mask_embed, mask_features, out=None):
    """einsum('bqc,bchw->bqhw'): mask_embed [B,Q,C], mask_features [B,C,H,W] -> [B,Q,H,W]."""
    e = _require(mask_embed, "mask_embed").contiguous()
    f = _require(mask_features, "mask_features").contiguous()
    B, Q, C = e.shape
    if f.dim() != 4 or f.shape[0] != B or f.shape[1] != C:
        raise ValueError(f"mask_features {tuple(f.shape)} does not match mask_embed {tuple(e.shape)}")
    H, W = f.shape[2], f.shape[3]
    if out is None:
        out = torch.empty(B, Q, H, W, device=e.device, dtype=torch.float32)
    rc = _lib.lib().msm_mask_logits(e.data_ptr(), f.data_ptr(), out.data_ptr(), B, Q, C, H * W, _stream())
    check(rc, "msm_mask_logits")
    return out


class MaskLogitsFunction(torch.autograd.Function):
    """Differentiable mask head einsum (meanshiftformer_transformer_decoder.py:668): forward = msm_mask_logits.
    Backward: g_feat[b,c,p] = sum_q embed[b,q,c] g[b,q,p] is the SAME contraction with the roles of queries and
    channels swapped, so it runs in the mask kernel too (embed^T as the "embedding", the query-padded gradient as the
    "features"; 128 channels per launch); g_embed = g . feat^T is a 19200-long reduction over the pixels and stays in
    cuBLAS (fp32, TF32 off) until the GEMM kernels have a split-K form."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, mask_embed, mask_features):
        e, f = mask_embed.detach().contiguous(), mask_features.detach().contiguous()
        ctx.save_for_backward(e, f)
        return mask_logits(e, f)

    @staticmethod
    @torch.autograd.function.once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_masks):
        e, f = ctx.saved_tensors
        B, Q, C = e.shape
        g = grad_masks.float().reshape(B, Q, -1)
        ge = gf = None
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            if ctx.needs_input_grad[0]:
                ge = torch.bmm(g, f.reshape(B, C, -1).transpose(1, 2))
            if ctx.needs_input_grad[1]:
                if os.environ.get("MSM_TRAIN_TC_MASK", "1") == "1" and f.dim() == 4 and Q <= 256:
                    Qp = (Q + 31) // 32 * 32
                    gp = g.reshape(B, Q, *f.shape[-2:])
                    et = e.transpose(1, 2)                                   # [B, C, Q]
                    if Qp != Q:                                              # zero query rows / columns: K of the GEMM
                        gp = torch.nn.functional.pad(gp, (0, 0, 0, 0, 0, Qp - Q))
                        et = torch.nn.functional.pad(et, (0, Qp - Q))
                    parts = [mask_logits(et[:, c0:c0 + 128].contiguous(), gp) for c0 in range(0, C, 128)]
                    gf = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
                else:
                    gf = torch.bmm(e.transpose(1, 2), g).reshape(f.shape)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        return ge, gf


def mask_logits_autograd(mask_embed, mask_features):
    """mask_logits for training: gradients flow to the mask embeddings and the mask features."""
    return MaskLogitsFunction.apply(mask_embed, mask_features)


def mask_to_attn_bits(masks, target_size):
    """masks [B,Q,H,W] -> (blocked bits int32 [B,Q,ceil(Ht*Wt/32)], row_open int32 [B,Q])."""
    m = _require(masks, "masks").contiguous()
    B, Q, H, W = m.shape
    Ht, Wt = int(target_size[0]), int(target_size[1])
    words = (Ht * Wt + 31) // 32
    bits = torch.empty(B, Q, words, device=m.device, dtype=torch.int32)
    row_open = torch.empty(B, Q, device=m.device, dtype=torch.int32)
    rc = _lib.lib().msm_mask_to_attn_bits(m.data_ptr(), bits.data_ptr(), row_open.data_ptr(), B, Q, H, W, Ht, Wt,
                                          _stream())
    check(rc, "msm_mask_to_attn_bits")
    return bits, row_open


def resample_bilinear(x, size, align_corners=False):
    """F.interpolate(x, size=size, mode="bilinear", align_corners=align_corners) for x [..., H, W] (inference only)."""
    x = _require(x, "x").contiguous()
    H, W = x.shape[-2:]
    Ht, Wt = int(size[0]), int(size[1])
    y = torch.empty(*x.shape[:-2], Ht, Wt, device=x.device, dtype=torch.float32)
    planes = x.numel() // (H * W)
    rc = _lib.lib().msm_resample_bilinear_fwd(x.data_ptr(), y.data_ptr(), planes, H, W, Ht, Wt, 1 if align_corners else 0,
                                              _stream())
    check(rc, "msm_resample_bilinear_fwd")
    return y


def maxpool3x3s2_channels_last(x):
    """nn.MaxPool2d(3, 2, 1) of a channels_last x [B,C,H,W] (C % 4 == 0) -> channels_last [B,C,Ho,Wo] (inference only)."""
    _require(x, "x")
    B, C, H, W = x.shape
    if not x.permute(0, 2, 3, 1).is_contiguous() or C % 4:
        raise ValueError("x must be a channels_last tensor with C % 4 == 0")
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, Ho, Wo, C, device=x.device, dtype=torch.float32)
    rc = _lib.lib().msm_maxpool3x3s2_nhwc_fwd(x.data_ptr(), y.data_ptr(), B, H, W, C, _stream())
    check(rc, "msm_maxpool3x3s2_nhwc_fwd")
    return y.permute(0, 3, 1, 2)


def upsample_add(x, add):
    """add + F.interpolate(x, size=add.shape[-2:], mode="bilinear", align_corners=False), one pass (inference only)."""
    x = _require(x, "x").contiguous()
    add = _require(add, "add").contiguous()
    H, W = x.shape[-2:]
    Ht, Wt = add.shape[-2:]
    if x.shape[:-2] != add.shape[:-2]:
        raise ValueError(f"x {tuple(x.shape)} and add {tuple(add.shape)} must agree in the leading dimensions")
    y = torch.empty_like(add)
    rc = _lib.lib().msm_upsample_add_fwd(x.data_ptr(), add.data_ptr(), y.data_ptr(), x.numel() // (H * W), H, W, Ht, Wt,
                                         _stream())
    check(rc, "msm_upsample_add_fwd")
    return y


def unpack_attn_bits(bits, row_open, num_keys, num_heads):
    """bits/row_open -> the reference's bool attn_mask [B*heads, Q, S] AFTER its un-mask rule
    (decoder.py:618). For tests and for callers that want the reference representation."""
    B, Q, words = bits.shape
    shifts = torch.arange(32, device=bits.device, dtype=torch.int32)
    blocked = ((bits.unsqueeze(-1) >> shifts) & 1).bool().reshape(B, Q, words * 32)[..., :num_keys]
    blocked = blocked & (row_open != 0).unsqueeze(-1)
    return blocked.unsqueeze(1).repeat(1, num_heads, 1, 1).flatten(0, 1)


# prepared (bf16 hi/lo) copies of weights, keyed by the fp32 tensor's address + shape + version. Each entry
# keeps a reference to its source tensor, so the address cannot be recycled by the allocator while the
# entry lives (a freed-and-reused address with the same version would otherwise be a stale hit).
_PREPARED = {}

# Weights epoch: part of the key of EVERY derived-weight cache of the package (prepared 16-bit copies, concatenated
# projection weights, row-bias tables, the 3x3 repack). Address + tensor ``_version`` catch in-place ops on the
# parameter and load_state_dict, but NOT ``.data`` edits and NOT the fused (multi-tensor) optimizers, which update
# parameters in place without bumping ``_version`` (verified with torch 2.11: AdamW(fused=True) leaves it at 0). Anything
# that changes weights behind autograd's back bumps the epoch: training.train_step does after optimizer.step().
_WEIGHTS_EPOCH = 0


def weights_epoch():
    return _WEIGHTS_EPOCH


def bump_weights_epoch():
    """Invalidate every derived-weight cache (prepared copies, concatenations, tables) of every module."""
    global _WEIGHTS_EPOCH
    _WEIGHTS_EPOCH += 1
    _PREPARED.clear()


def clear_prepared_weights():
    """Drop all derived weight copies - prepared 16-bit copies AND the per-module caches of cached_cat / cached_value
    (concatenated K/V weights, row-bias tables, the conv3x3 repack). Call after editing weights through ``.data`` or
    after an optimizer step taken outside training.train_step (fused optimizers do not bump the tensor version)."""
    bump_weights_epoch()


def prepare_linear_weight(weight):
    """fp32 weight [N, K] (rows 16-byte aligned) -> cached uint8 buffer in the kernel's streaming layout."""
    w = _require(weight, "weight")
    if w.dim() != 2 or w.stride(1) != 1:
        raise ValueError("weight must be a [N, K] matrix with contiguous rows")
    key = (w.data_ptr(), tuple(w.shape), w.stride(0), w._version, w.device.index, _WEIGHTS_EPOCH)
    hit = _PREPARED.get(key)
    if hit is not None:
        return hit[0]
    if True:
        if len(_PREPARED) > 1024:
            _PREPARED.clear()
        N, K = w.shape
        L = _lib.lib()
        buf = torch.empty(L.msm_linear_weight_bytes(N, K), device=w.device, dtype=torch.uint8)
        check(L.msm_linear_prepare_weight(w.data_ptr(), w.stride(0), buf.data_ptr(), N, K, _stream()),
              "msm_linear_prepare_weight")
        _PREPARED[key] = (buf, w)
    return buf


def _prepare_linear_weight_once(weight):
    """prepare_linear_weight without the cache: training (the weight changes every step, and so does its transpose)."""
    w = _require(weight, "weight")
    N, K = w.shape
    L = _lib.lib()
    buf = torch.empty(L.msm_linear_weight_bytes(N, K), device=w.device, dtype=torch.uint8)
    check(L.msm_linear_prepare_weight(w.data_ptr(), w.stride(0), buf.data_ptr(), N, K, _stream()),
          "msm_linear_prepare_weight")
    return buf


def _prepare_linear_weight_t_once(weight):
    """prepared form of weight.T (weight [N, K] -> a [K, N] layer weight) without materialising the transpose"""
    w = _require(weight, "weight")
    N, K = w.shape
    L = _lib.lib()
    buf = torch.empty(L.msm_linear_weight_bytes(K, N), device=w.device, dtype=torch.uint8)
    check(L.msm_linear_prepare_weight_t(w.data_ptr(), w.stride(0), buf.data_ptr(), N, K, _stream()),
          "msm_linear_prepare_weight_t")
    return buf


def tc_linear_enabled():
    """False when MSM_DISABLE_TC_LINEAR is set (dense layers then run in cuBLAS fp32 - a cross-check switch)."""
    return os.environ.get("MSM_DISABLE_TC_LINEAR", "") in ("", "0")


def linear_supported(x, weight):
    """True when msm_linear_fwd takes this layer (N, K multiples of 32, aligned fp32 CUDA rows)."""
    if not tc_linear_enabled():
        return False
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.dim() == 2
            and weight.shape[0] % 32 == 0 and weight.shape[1] % 32 == 0 and weight.stride(1) == 1
            and weight.stride(0) % 4 == 0 and weight.data_ptr() % 16 == 0 and x.shape[-1] == weight.shape[1])


def linear(x, weight, bias=None, relu=False, out=None, _prepared=None):
    """act(x @ weight.T + bias) on the tensor cores (bf16x3 split precision, fp32 accumulate).
    x [..., K] whose leading axes collapse to uniformly strided rows; weight [N, K]; returns [..., N]."""
    _require(x, "x")
    N, K = weight.shape
    if x.shape[-1] != K:
        raise ValueError(f"x {tuple(x.shape)} does not match weight {tuple(weight.shape)}")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)  # a view whenever the rows are uniformly strided
    if x2.stride(1) != 1 or x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    if out is None:
        out = torch.empty(*lead, N, device=x.device, dtype=torch.float32)
    y2 = out.view(-1, N) if out.is_contiguous() else out
    if y2.dim() != 2 or y2.shape != (M, N) or y2.stride(1) != 1:
        raise ValueError("out must be a [M, N] matrix with contiguous rows")
    wp = _prepared if _prepared is not None else prepare_linear_weight(weight)
    b = None
    if bias is not None:
        b = _require(bias, "bias").contiguous()
    if M > 0:
        rc = _lib.lib().msm_linear_fwd(x2.data_ptr(), x2.stride(0), wp.data_ptr(), b.data_ptr() if b is not None else None,
                                       y2.data_ptr(), y2.stride(0), M, N, K, 1 if relu else 0, _stream())
        check(rc, "msm_linear_fwd")
    return out


def linear_ln_supported(x, weight, residual, norm):
    """Shapes msm_linear_ln_fwd takes: N in (32, 64), contiguous fp32 residual rows, an affine LayerNorm over N."""
    N = weight.shape[0]
    return (not torch.is_grad_enabled() and linear_supported(x, weight) and N in (32, 64)
            and isinstance(norm, torch.nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
            and tuple(norm.normalized_shape) == (N,) and residual.is_cuda and residual.dtype == torch.float32
            and residual.is_contiguous() and residual.shape[-1] == N and residual.numel() // N == x.numel() // x.shape[-1])


def linear_ln(x, weight, bias, residual, norm):
    """norm(residual + x @ weight.T + bias) with the add and the LayerNorm fused into the GEMM epilogue."""
    _require(x, "x")
    N, K = weight.shape
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty_like(residual)
    wp = prepare_linear_weight(weight.detach())
    b = None if bias is None else bias.detach().contiguous()
    rc = _lib.lib().msm_linear_ln_fwd(x2.data_ptr(), x2.stride(0), wp.data_ptr(), b.data_ptr() if b is not None else None,
                                      residual.data_ptr(), N, norm.weight.data_ptr(), norm.bias.data_ptr(),
                                      float(norm.eps), out.data_ptr(), N, M, N, K, _stream())
    check(rc, "msm_linear_ln_fwd")
    return out


def linear_fused(x, weight, bias=None, *, rowbias=None, relu=False, residual=None, norm=None, l2_normalize=False,
                 norm2=None):
    """One launch for  v = act(x @ W.T + bias + rowbias[row % len(rowbias)]) + residual;  y = norm(v);
    z = F.normalize(y) if l2_normalize;  returns z, or (z, norm2(z)) when norm2 is given.
    norm / norm2 are affine torch.nn.LayerNorm modules over N; row stages need N <= 256."""
    _require(x, "x")
    N, K = weight.shape
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    lead = x.shape[:-1]
    out = torch.empty(*lead, N, device=x.device, dtype=torch.float32)
    out2 = torch.empty_like(out) if norm2 is not None else None
    wp = prepare_linear_weight(weight.detach())
    b = None if bias is None else bias.detach().contiguous()
    rb = None
    if rowbias is not None:
        rb = _require(rowbias.detach(), "rowbias").contiguous()
        if rb.dim() != 2 or rb.shape[1] != N:
            raise ValueError(f"rowbias must be [period, {N}]")
    res = None
    if residual is not None:
        res = _require(residual.detach(), "residual")
        if not res.is_contiguous() or res.numel() != M * N:
            raise ValueError("residual must be a contiguous tensor with the output's shape")
    for nm in (norm, norm2):
        if nm is not None and not (isinstance(nm, torch.nn.LayerNorm) and nm.elementwise_affine and nm.bias is not None
                                   and tuple(nm.normalized_shape) == (N,)):
            raise ValueError("norm / norm2 must be affine LayerNorm modules over the output features")
    p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    rc = _lib.lib().msm_linear_fused_fwd(
        x2.data_ptr(), x2.stride(0), wp.data_ptr(), p(b), p(rb), rb.shape[0] if rb is not None else 1,
        1 if relu else 0, p(res), N, p(norm.weight) if norm is not None else None,
        p(norm.bias) if norm is not None else None, float(norm.eps) if norm is not None else 0.0,
        1 if l2_normalize else 0, p(norm2.weight) if norm2 is not None else None,
        p(norm2.bias) if norm2 is not None else None, float(norm2.eps) if norm2 is not None else 0.0,
        p(out2), N, out.data_ptr(), N, M, N, K, _stream())
    check(rc, "msm_linear_fused_fwd")
    return out if norm2 is None else (out, out2)


def cached_value(owner, name, deps, fn):
    """fn() cached on ``owner`` until one of the ``deps`` tensors is modified in place or replaced."""
    key = (_WEIGHTS_EPOCH,) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in deps)
    cache = owner.__dict__.setdefault("_msm_cat_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            hit = (key, fn(), list(deps))
        cache[name] = hit
    return hit[1]


def ffn_ln_supported(x, w1, w2, norm):
    D = x.shape[-1]
    return (tc_linear_enabled() and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32
            and x.is_contiguous() and D in (32, 64) and tuple(w1.shape[1:]) == (D,) and tuple(w2.shape) == (D, w1.shape[0])
            and w1.shape[0] % 128 == 0 and w1.shape[0] <= 1792 and isinstance(norm, torch.nn.LayerNorm)
            and norm.elementwise_affine and norm.bias is not None and tuple(norm.normalized_shape) == (D,)
            and os.environ.get("MSM_DISABLE_FFN_FUSION", "") in ("", "0"))


def ffn_ln(x, w1, b1, w2, b2, norm):
    """norm(x + relu(x @ w1.T + b1) @ w2.T + b2) in one kernel; the hidden activation never reaches HBM."""
    _require(x, "x")
    D = x.shape[-1]
    F = w1.shape[0]
    x2 = x.reshape(-1, D)
    M = x2.shape[0]
    out = torch.empty_like(x)
    w1p, w2p = prepare_linear_weight(w1.detach()), prepare_linear_weight(w2.detach())
    rc = _lib.lib().msm_ffn_ln_fwd(x2.data_ptr(), x2.stride(0), w1p.data_ptr(), b1.detach().contiguous().data_ptr(),
                                   w2p.data_ptr(), b2.detach().contiguous().data_ptr(), norm.weight.data_ptr(),
                                   norm.bias.data_ptr(), float(norm.eps), out.data_ptr(), D, M, D, F, _stream())
    check(rc, "msm_ffn_ln_fwd")
    return out


# ----------------------------------------------------------------------------------------------
# one decoder layer after its cross-attention, as ONE cluster kernel per layer (csrc/decoder_block.cu)
# ----------------------------------------------------------------------------------------------
def decoder_block_enabled():
    """Opt-in (MSM_DECODER_BLOCK=1). Measured on B200 (R50 config, batch 8): the cluster kernel takes 135-148 us per layer
    and brings the step from 212 to 97 launches, but the 16 launches it replaces cost ~120 us, so the step is 4 % slower
    (4.61-4.66 vs 4.44 ms) - the all-gathers through distributed shared memory (128 KB into each of 8 CTAs, ~5 us each,
    7 per layer) and the CUDA-core self-attention (22 us) pace it; DESIGN.md section 4.3c has the stage timing."""
    return os.environ.get("MSM_DECODER_BLOCK", "0") == "1"


def _dbk_pieces(w):
    """fp32 [N, 256] weight slice -> bytes of its 8 stream pieces, each [hi | lo][4 k-groups][N][8] fp16 (the layout of
    msm_linear_prepare_weight for a 32-channel K chunk: the UMMA canonical K-major B operand)."""
    N = w.shape[0]
    hi = w.to(torch.float16)
    lo = (w - hi.float()).to(torch.float16)
    t = torch.stack([hi, lo]).view(2, N, 8, 4, 8).permute(2, 0, 3, 1, 4).contiguous()   # [chunk, hi/lo, kg, N, 8]
    return t.view(torch.uint8).reshape(-1)


def decoder_block_pack(w_o1, w_qkv, w_o2, w_f1, w_f2, w_qn, w_m1, w_c, w_m2, w_m3):
    """Weight stream of one decoder layer for the 8 CTAs of a cluster: [8, bytes] uint8. CTA r gets, in the order the
    kernel consumes them: out_proj rows [32r, 32r+32) of the cross-attention; its head's q | k | v rows of the
    self-attention in_proj; self out_proj rows; 2 x 128 hidden units of linear1; linear2 restricted to those 256
    hidden units (outputs 0..127, 128..255); the NEXT layer's cross-attention q in_proj rows (None: last layer);
    mask MLP layer 0 rows + the class head padded to 32 rows; MLP layers 1, 2 rows."""
    C = 256
    if tuple(w_o1.shape) != (C, C) or tuple(w_qkv.shape) != (3 * C, C) or tuple(w_f1.shape) != (2048, C) \
            or tuple(w_f2.shape) != (C, 2048) or w_c.shape[1] != C or w_c.shape[0] > 32:
        raise ValueError("decoder_block: hidden 256, 8 heads, FFN 2048, at most 32 classes")
    wc = torch.cat([w_c, w_c.new_zeros(32 - w_c.shape[0], C)], 0)
    ranks = []
    for r in range(8):
        s = slice(32 * r, 32 * r + 32)
        h = slice(256 * r, 256 * r + 256)
        parts = [w_o1[s], torch.cat([w_qkv[s], w_qkv[C + 32 * r:C + 32 * r + 32], w_qkv[2 * C + 32 * r:2 * C + 32 * r + 32]]),
                 w_o2[s], w_f1[256 * r:256 * r + 128], w_f1[256 * r + 128:256 * r + 256], w_f2[:128, h], w_f2[128:, h]]
        if w_qn is not None:
            parts.append(w_qn[s])
        parts += [torch.cat([w_m1[s], wc]), w_m2[s], w_m3[s]]
        ranks.append(torch.cat([_dbk_pieces(p.detach().float().contiguous()) for p in parts]))
    blob = torch.stack(ranks).contiguous()
    want = _lib.lib().msm_decoder_block_weight_bytes(1 if w_qn is not None else 0)
    if blob.numel() != want:
        raise RuntimeError(f"decoder_block_pack: {blob.numel()} bytes packed, the kernel expects {want}")
    return blob


def decoder_block(o_cross, state, blob, *, b_o1, norm1, b_qkv, t_qk, b_o2, norm2, b_f1, b_f2, norm3, block_norm, normd,
                  b_qn, t_qn, b_m1, b_c32, b_m2, b_m3, kappa=KAPPA):
    """o_cross, state [B, Q, 256] -> (state_out [B,Q,256], logits [B,Q,32] (first K+1 columns valid), embed [B,Q,256],
    q_next [B,Q,256] or None). See csrc/decoder_block.cu."""
    o = _require(o_cross, "o_cross").contiguous()
    st = _require(state, "state").contiguous()
    B, Q, C = o.shape
    if C != 256 or st.shape != o.shape or Q > 128:
        raise ValueError(f"decoder_block: [B, Q <= 128, 256] inputs, got {tuple(o.shape)} / {tuple(st.shape)}")
    dev = o.device
    state_out = torch.empty_like(o)
    logits = torch.empty(B, Q, 32, device=dev, dtype=torch.float32)
    embed = torch.empty_like(o)
    q_next = torch.empty_like(o) if b_qn is not None else None
    p = lambda t: None if t is None else _require(t.detach(), "vector").contiguous().data_ptr()  # noqa: E731
    rc = _lib.lib().msm_decoder_block_fwd(
        o.data_ptr(), st.data_ptr(), blob.data_ptr(), p(b_o1), p(norm1.weight), p(norm1.bias), float(norm1.eps),
        p(b_qkv), p(t_qk), p(b_o2), p(norm2.weight), p(norm2.bias), float(norm2.eps), p(b_f1), p(b_f2),
        p(norm3.weight), p(norm3.bias), float(norm3.eps), 1 if block_norm else 0, p(normd.weight), p(normd.bias),
        float(normd.eps), p(b_qn), p(t_qn), p(b_m1), p(b_c32), p(b_m2), p(b_m3), state_out.data_ptr(),
        logits.data_ptr(), embed.data_ptr(), q_next.data_ptr() if q_next is not None else None, B, Q, float(kappa),
        _stream())
    check(rc, "msm_decoder_block_fwd")
    return state_out, logits, embed, q_next


def add_layernorm(x, y, norm, l2_normalize=False, norm2=None):
    """norm(x + y) [-> F.normalize] [-> also norm2 of the result] in one launch. x, y contiguous [..., C] (y may be
    None); norm / norm2 affine LayerNorm modules over C <= 1024. Returns out, or (out, out2) when norm2 is given."""
    _require(x, "x")
    C = x.shape[-1]
    xc = x.contiguous()
    yc = None if y is None else _require(y, "y").contiguous()
    if yc is not None and yc.shape != xc.shape:
        raise ValueError("x and y must have the same shape")
    for nm in (norm, norm2):
        if nm is not None and not (isinstance(nm, torch.nn.LayerNorm) and nm.elementwise_affine and nm.bias is not None
                                   and tuple(nm.normalized_shape) == (C,)):
            raise ValueError("norm / norm2 must be affine LayerNorm modules over the last axis")
    out = torch.empty_like(xc)
    out2 = torch.empty_like(xc) if norm2 is not None else None
    rc = _lib.lib().msm_add_layernorm_fwd(
        xc.data_ptr(), yc.data_ptr() if yc is not None else None, norm.weight.data_ptr(), norm.bias.data_ptr(),
        float(norm.eps), 1 if l2_normalize else 0, norm2.weight.data_ptr() if norm2 is not None else None,
        norm2.bias.data_ptr() if norm2 is not None else None, float(norm2.eps) if norm2 is not None else 0.0,
        out.data_ptr(), out2.data_ptr() if out2 is not None else None, xc.numel() // C, C, _stream())
    check(rc, "msm_add_layernorm_fwd")
    return out if norm2 is None else (out, out2)


def conv1x1_supported(x, weight):
    if os.environ.get("MSM_DISABLE_TC_LINEAR", "") not in ("", "0"):
        return False
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous() and weight.dim() == 4
            and weight.shape[2] == 1 and weight.shape[3] == 1 and weight.is_contiguous()
            and weight.shape[0] % 32 == 0 and weight.shape[1] % 32 == 0 and x.shape[1] == weight.shape[1]
            and (x.shape[2] * x.shape[3]) % 4 == 0 and x.data_ptr() % 16 == 0)


def _is_channels_last(x):
    """a dense NCHW-shaped tensor whose memory is [B][H][W][C] (torch.channels_last) and not also NCHW-contiguous"""
    return x.dim() == 4 and not x.is_contiguous() and x.permute(0, 2, 3, 1).is_contiguous()


def conv1x1_nhwc_supported(x, weight):
    if os.environ.get("MSM_DISABLE_TC_LINEAR", "") not in ("", "0"):
        return False
    return (x.is_cuda and x.dtype == torch.float32 and _is_channels_last(x) and weight.dim() == 4
            and weight.shape[2] == 1 and weight.shape[3] == 1 and weight.is_contiguous()
            and weight.shape[0] % 32 == 0 and weight.shape[1] % 32 == 0 and x.shape[1] == weight.shape[1]
            and x.data_ptr() % 16 == 0)


def conv1x1_nhwc(x, weight, bias=None, relu=False):
    """kernel_size=1 convolution of a channels_last x [B,K,H,W] (no layout copy): -> [B,N,H,W], NCHW-contiguous when
    H*W % 128 == 0 (msm_conv1x1_nhwc_fwd), else the channels_last view of the token-major result."""
    _require(x, "x")
    B, K, H, W = x.shape
    N = weight.shape[0]
    if (H * W) % 128:
        y = linear(x.permute(0, 2, 3, 1).reshape(B, H * W, K), weight.detach().view(N, K),
                   None if bias is None else bias.detach(), relu=relu)
        return y.transpose(1, 2).reshape(B, N, H, W)
    wp = prepare_linear_weight(weight.detach().view(N, K))
    b = None if bias is None else _require(bias.detach(), "bias").contiguous()
    out = torch.empty(B, N, H, W, device=x.device, dtype=torch.float32)
    rc = _lib.lib().msm_conv1x1_nhwc_fwd(x.data_ptr(), wp.data_ptr(), b.data_ptr() if b is not None else None,
                                         out.data_ptr(), B, H * W, N, K, 1 if relu else 0, _stream())
    check(rc, "msm_conv1x1_nhwc_fwd")
    return out


def conv1x1(x, weight, bias=None, relu=False, tokens_out=False):
    """kernel_size=1 convolution on the tensor cores. x [B,K,H,W] contiguous, weight [N,K,1,1];
    returns [B,N,H,W] (default) or the token-major [B,H*W,N] the decoders consume."""
    _require(x, "x")
    B, K, H, W = x.shape
    N = weight.shape[0]
    wp = prepare_linear_weight(weight.detach().view(N, K))
    b = None if bias is None else _require(bias.detach(), "bias").contiguous()
    out = torch.empty((B, H * W, N) if tokens_out else (B, N, H, W), device=x.device, dtype=torch.float32)
    rc = _lib.lib().msm_conv1x1_fwd(x.data_ptr(), wp.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                                    0 if tokens_out else 1, B, H * W, N, K, 1 if relu else 0, _stream())
    check(rc, "msm_conv1x1_fwd")
    return out


def conv3x3_supported(x, weight):
    return (tc_linear_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()
            and weight.dim() == 4 and weight.shape[2] == 3 and weight.shape[3] == 3 and weight.shape[0] % 32 == 0
            and weight.shape[1] % 32 == 0 and x.shape[1] == weight.shape[1])


def conv3x3(x, weight, bias=None, relu=False):
    """3x3 convolution (padding 1, stride 1) on the tensor cores as an implicit GEMM over 9*C. x [B,C,H,W]
    contiguous, weight [N,C,3,3] -> [B,N,H,W]."""
    _require(x, "x")
    B, C, H, W = x.shape
    N = weight.shape[0]
    w = weight.detach()
    w2 = cached_value(_CONV3_CACHE_OWNER, f"w{w.data_ptr()}_{tuple(w.shape)}", [w],
                      lambda: w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous())
    wp = prepare_linear_weight(w2)
    b = None if bias is None else _require(bias.detach(), "bias").contiguous()
    out = torch.empty(B, N, H, W, device=x.device, dtype=torch.float32)
    Wp = (W + 2 + 3) // 4 * 4  # zero border of one pixel, rows padded to a 16-byte multiple for TMA
    xp = torch.nn.functional.pad(x, (1, Wp - W - 1, 1, 1))
    rc = _lib.lib().msm_conv3x3_fwd(xp.data_ptr(), wp.data_ptr(), b.data_ptr() if b is not None else None,
                                    out.data_ptr(), B, C, H, W, Wp, N, 1 if relu else 0, _stream())
    check(rc, "msm_conv3x3_fwd")
    return out


class _Owner:
    pass


_CONV3_CACHE_OWNER = _Owner()
_PAD_CACHE_OWNER = _Owner()


def conv_layer(conv, x):
    """Forward of a Conv2d module (detectron2-style wrapper with optional .norm / .activation, or plain
    nn.Conv2d): tensor-core path for 1x1 and 3x3/pad-1 convolutions at inference, the module itself otherwise."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad)
    is3 = (conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.dilation == (1, 1)
           and getattr(conv, "padding_mode", "zeros") == "zeros")
    if needs_grad or conv.stride != (1, 1) or conv.groups != 1:
        return conv(x)
    if conv.kernel_size == (1, 1) and conv1x1_supported(x, conv.weight):
        y = conv1x1(x, conv.weight, conv.bias)
    elif is3 and conv3x3_supported(x, conv.weight):
        y = conv3x3(x, conv.weight, conv.bias)
    else:
        return conv(x)
    if getattr(conv, "norm", None) is not None:
        y = conv.norm(y)
    if getattr(conv, "activation", None) is not None:
        y = conv.activation(y)
    return y


def conv1x1_layer(conv, x):
    """Forward of a kernel_size=1 Conv2d module (detectron2-style wrapper with optional .norm / .activation,
    or a plain nn.Conv2d): tensor-core path for inference on shapes it takes, the module itself otherwise."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad)
    plain = conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.groups == 1 and not needs_grad
    if plain and conv1x1_nhwc_supported(x, conv.weight):   # a channels_last backbone feature: no layout copy
        y = conv1x1_nhwc(x, conv.weight, conv.bias)
    elif plain and conv1x1_supported(x, conv.weight):
        y = conv1x1(x, conv.weight, conv.bias)
    else:
        return conv(x)
    if getattr(conv, "norm", None) is not None:
        y = conv.norm(y)
    if getattr(conv, "activation", None) is not None:
        y = conv.activation(y)
    return y


class DenseFunction(torch.autograd.Function):
    """Differentiable dense layer for fp32 training: forward and the input gradient run in
    linear_tc_kernel (split-precision tensor-core GEMM, the same arithmetic as inference; dX = dY . W is the same
    kernel on the transposed weight), the weight gradient dW = dY^T . X - a reduction over all rows - and the bias
    gradient stay in cuBLAS / ATen (fp32, TF32 off)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        xd, wd = x.detach(), weight.detach()
        y = linear(xd, wd, None if bias is None else bias.detach(), relu=relu, _prepared=_prepare_linear_weight_once(wd))
        ctx.save_for_backward(xd, wd, y if relu else None)
        ctx.relu, ctx.has_bias = relu, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        N, K = w.shape
        g2 = gy.reshape(-1, N)
        if ctx.relu:
            g2 = g2 * (y.reshape(-1, N) > 0)
        g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # dX = dY . W = linear(dY, W^T): the kernel's weight layout is written straight from W (no transposed copy)
            gx = linear(g2, w.t(), _prepared=_prepare_linear_weight_t_once(w)).reshape(x.shape)
        if ctx.needs_input_grad[1]:
            tf32 = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                gw = g2.t() @ x.reshape(-1, K)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = tf32
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb, None


def dense_autograd_supported(x, weight):
    """fp32 training outside autocast on shapes the tensor-core kernel takes in BOTH directions (N, K multiples of 32)."""
    return (os.environ.get("MSM_TRAIN_TC_LINEAR", "1") == "1" and linear_supported(x, weight)
            and not torch.is_autocast_enabled() and weight.is_contiguous() and x.numel() > 0)


def dense(x, weight, bias=None, relu=False):
    """The layer call the modules use: tensor-core ``linear`` for inference on shapes it takes,
    torch's F.linear (cuBLAS fp32, autograd-capable) when gradients are needed or N/K are not
    multiples of 32 (e.g. the 3-way class head)."""
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor: unseenobjectswithmeanshift_b200 has no CPU path")
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad
                                              or (bias is not None and bias.requires_grad))
    if not needs_grad and not linear_supported(x, weight) and _paddable(x, weight):
        # narrow heads (the 3-way class head, N = K + 1 classes): zero-padded to 32 output columns once per weight and run
        # through the same tensor-core kernel; the caller gets the [..., :N] view. No cuBLAS launch left in the decoder.
        N = weight.shape[0]
        Np = (N + 31) // 32 * 32
        wp = cached_value(_PAD_CACHE_OWNER, f"w{weight.data_ptr()}_{tuple(weight.shape)}", [weight],
                          lambda: torch.cat([weight.detach(), weight.new_zeros(Np - N, weight.shape[1])], 0).contiguous())
        bp = None
        if bias is not None:
            bp = cached_value(_PAD_CACHE_OWNER, f"b{bias.data_ptr()}_{N}", [bias],
                              lambda: torch.cat([bias.detach(), bias.new_zeros(Np - N)], 0).contiguous())
        return linear(x, wp, bp, relu=relu)[..., :N]
    if needs_grad and dense_autograd_supported(x, weight):
        return DenseFunction.apply(x, weight, bias, relu)
    if needs_grad or not linear_supported(x, weight):
        y = torch.nn.functional.linear(x, weight, bias)
        return torch.relu_(y) if relu else y
    return linear(x, weight.detach(), None if bias is None else bias.detach(), relu=relu)


def _paddable(x, weight):
    return (tc_linear_enabled() and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32
            and weight.dim() == 2 and weight.shape[0] % 32 != 0 and weight.shape[0] <= 1024 and weight.shape[1] % 32 == 0
            and weight.stride(1) == 1 and x.shape[-1] == weight.shape[1])


def cached_cat(owner, name, tensors, dim=0):
    """torch.cat(tensors, dim) cached on ``owner`` until any source tensor is modified in place or
    replaced (keeps concatenated projection weights - and their prepared copies - stable across calls)."""
    key = (_WEIGHTS_EPOCH,) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
    cache = owner.__dict__.setdefault("_msm_cat_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        # the sources are kept referenced so their addresses stay unique while the entry lives
        hit = (key, torch.cat([t.detach() for t in tensors], dim).contiguous(), list(tensors))
        cache[name] = hit
    return hit[1]


def _check_im2col_step(batch, im2col_step):
    """The reference asserts this before chunking the batch (ms_deform_attn_cuda.cu:55-57, :118-120); the kernels here
    do not chunk, the contract is kept."""
    step = min(int(batch), int(im2col_step))
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=128):
    """Same signature and result as MultiScaleDeformableAttention.ms_deform_attn_forward
    (pixel_decoder/ops/src/ms_deform_attn.h:25-45): returns [N, Lq, M*D]."""
    for t, n in ((value, "value"), (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")):
        _require(t, n)
        if not t.is_contiguous():  # the reference asserts this too (ms_deform_attn_cuda.cu:33-37)
            raise RuntimeError(f"{n} tensor has to be contiguous")
    _require(spatial_shapes, "spatial_shapes", torch.int64)
    _require(level_start_index, "level_start_index", torch.int64)
    if not spatial_shapes.is_contiguous() or not level_start_index.is_contiguous():
        raise RuntimeError("spatial_shapes / level_start_index tensor has to be contiguous")
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    _check_im2col_step(N, im2col_step)
    out = torch.empty(N, Lq, M * D, device=value.device, dtype=torch.float32)
    rc = _lib.lib().msm_ms_deform_attn_fwd(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                           sampling_loc.data_ptr(), attn_weight.data_ptr(), out.data_ptr(),
                                           N, S, M, D, L, Lq, P, int(im2col_step), _stream())
    check(rc, "msm_ms_deform_attn_fwd")
    return out


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, offsets_logits, reference_points,
                                 n_levels, n_points):
    """value [N,S,M,D]; offsets_logits [N,Lq,M*L*P*3] (raw sampling offsets | raw attention logits);
    reference_points [N,Lq,L,2] -> [N,Lq,M*D]. Softmax, sampling-location arithmetic and the gather in one kernel."""
    for t, n in ((value, "value"), (offsets_logits, "offsets_logits"), (reference_points, "reference_points")):
        _require(t, n)
    _require(spatial_shapes, "spatial_shapes", torch.int64)
    _require(level_start_index, "level_start_index", torch.int64)
    if not value.is_contiguous():
        raise RuntimeError("value tensor has to be contiguous")
    N, S, M, D = value.shape
    Lq = offsets_logits.shape[1]
    L, P = int(n_levels), int(n_points)
    ol = offsets_logits if offsets_logits.stride(-1) == 1 and offsets_logits.stride(0) == Lq * offsets_logits.stride(1) \
        else offsets_logits.contiguous()
    if ol.shape != (N, Lq, M * L * P * 3) or reference_points.shape != (N, Lq, L, 2):
        raise ValueError(f"offsets_logits {tuple(ol.shape)} / reference_points {tuple(reference_points.shape)} "
                         f"do not match value {tuple(value.shape)} with L={L}, P={P}")
    ref = reference_points.contiguous()
    out = torch.empty(N, Lq, M * D, device=value.device, dtype=torch.float32)
    rc = _lib.lib().msm_ms_deform_attn_fused_fwd(value.data_ptr(), spatial_shapes.data_ptr(),
                                                 level_start_index.data_ptr(), ol.data_ptr(), ol.stride(1),
                                                 ref.data_ptr(), out.data_ptr(), N, S, M, D, L, Lq, P, _stream())
    check(rc, "msm_ms_deform_attn_fused_fwd")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=128):
    """MultiScaleDeformableAttention.ms_deform_attn_backward (ms_deform_attn.h:47-67):
    returns [grad_value, grad_sampling_loc, grad_attn_weight]."""
    go = _require(grad_output, "grad_output").contiguous()
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    _check_im2col_step(N, im2col_step)
    gv = torch.zeros_like(value)
    gl = torch.empty_like(sampling_loc)
    ga = torch.empty_like(attn_weight)
    rc = _lib.lib().msm_ms_deform_attn_bwd(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                           sampling_loc.data_ptr(), attn_weight.data_ptr(), go.data_ptr(),
                                           gv.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                                           N, S, M, D, L, Lq, P, int(im2col_step), _stream())
    check(rc, "msm_ms_deform_attn_bwd")
    return [gv, gl, ga]


def mean_shift_hill_climb(X, Z, kappa, max_iters=10):
    """Batched seed_hill_climbing_ball: X [B,n,d] (or [n,d]), Z [B,m,d] (or [m,d]) -> Z' same shape."""
    squeeze = X.dim() == 2
    Xb = _require(X, "X").contiguous()
    Zb = _require(Z, "Z").contiguous()
    if squeeze:
        Xb, Zb = Xb.unsqueeze(0), Zb.unsqueeze(0)
    B, n, d = Xb.shape
    m = Zb.shape[1]
    if Zb.shape[0] != B or Zb.shape[2] != d:
        raise ValueError(f"X {tuple(X.shape)} / Z {tuple(Z.shape)} mismatch")
    out = torch.empty_like(Zb)
    if os.environ.get("MSM_PACKED_MS", "1") == "1" and d in (32, 64) and m <= 128 and int(max_iters) >= 1:
        # default: X is split into 16-bit operand images ONCE per call instead of in every iteration, and the whole
        # climb is one persistent cooperative kernel (csrc/mean_shift_persistent.cu)
        X_ = _lib.xlib()
        packed = torch.empty(X_.msmx_mean_shift_packed_bytes(B, n, d), device=Xb.device, dtype=torch.uint8)
        check(X_.msmx_mean_shift_pack(Xb.data_ptr(), packed.data_ptr(), B, n, d, _stream()), "msmx_mean_shift_pack")
        if os.environ.get("MSM_MS_PERSISTENT", "1") == "1":
            # ONE cooperative launch for all iterations of all images (csrc/mean_shift_persistent.cu): point tiles split
            # evenly over the SMs, per-image barriers between iterations, seeds re-normalised in the kernel
            ws_bytes = X_.msmx_mean_shift_persistent_workspace_bytes(B, n, m, d)
            ws = torch.empty(ws_bytes, device=Xb.device, dtype=torch.uint8)
            rc = X_.msmx_mean_shift_hill_climb_persistent(packed.data_ptr(), Zb.data_ptr(), out.data_ptr(), B, n, m, d,
                                                          float(kappa), int(max_iters), ws.data_ptr(), ws_bytes,
                                                          _stream())
            check(rc, "msmx_mean_shift_hill_climb_persistent")
            return out[0] if squeeze else out
        ws_bytes = X_.msmx_mean_shift_packed_workspace_bytes(B, n, m, d)
        ws = torch.empty(ws_bytes, device=Xb.device, dtype=torch.uint8)
        rc = X_.msmx_mean_shift_hill_climb_packed(packed.data_ptr(), Zb.data_ptr(), out.data_ptr(), B, n, m, d,
                                                  float(kappa), int(max_iters), ws.data_ptr(), ws_bytes, _stream())
        check(rc, "msmx_mean_shift_hill_climb_packed")
        return out[0] if squeeze else out
    L = _lib.lib()
    ws_bytes = L.msm_mean_shift_workspace_bytes(B, n, m, d)
    ws = torch.empty(ws_bytes, device=Xb.device, dtype=torch.uint8)
    rc = L.msm_mean_shift_hill_climb(Xb.data_ptr(), Zb.data_ptr(), out.data_ptr(), B, n, m, d, float(kappa),
                                     int(max_iters), ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msm_mean_shift_hill_climb")
    return out[0] if squeeze else out


def _points(X, name="X"):
    Xb = _require(X, name).contiguous()
    return (Xb.unsqueeze(0), True) if Xb.dim() == 2 else (Xb, False)


def select_smart_seeds(X, num_seeds, first_index):
    """Farthest-point seeding, batched: X [B,n,d] (or [n,d]) unit rows, first_index [B] (int / sequence / int64
    tensor; the reference draws it with np.random.randint, mean_shift.py:155) ->
    (seeds [B,num_seeds,d], selected int64 [B,num_seeds]). One cooperative launch, no host sync."""
    Xb, squeeze = _points(X)
    B, n, d = Xb.shape
    if torch.is_tensor(first_index):
        first = first_index.to(device=Xb.device, dtype=torch.int64).reshape(-1).contiguous()
    else:
        first = torch.as_tensor(first_index, dtype=torch.int64).reshape(-1).to(Xb.device)
    if first.numel() != B:
        raise ValueError(f"first_index has {first.numel()} entries for {B} images")
    if d not in (16, 32, 64, 128):
        raise ValueError(f"embedding dim {d} not supported (16, 32, 64, 128)")
    seeds = torch.empty(B, num_seeds, d, device=Xb.device, dtype=torch.float32)
    selected = torch.empty(B, num_seeds, device=Xb.device, dtype=torch.int64)
    L = _lib.lib()
    ws_bytes = L.msm_smart_seeds_workspace_bytes(B, n, int(num_seeds))
    ws = torch.empty(ws_bytes, device=Xb.device, dtype=torch.uint8)
    rc = L.msm_select_smart_seeds(Xb.data_ptr(), first.data_ptr(), seeds.data_ptr(), selected.data_ptr(), B, n, d,
                                  int(num_seeds), ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msm_select_smart_seeds")
    return (seeds[0], selected[0]) if squeeze else (seeds, selected)


def seed_connected_components(Z, epsilon):
    """connected_components over converged seeds, batched: Z [B,m,d] (or [m,d]) ->
    (seed_labels int64 [B,m], num_labels int32 [B]) on the device."""
    Zb, squeeze = _points(Z, "Z")
    B, m, d = Zb.shape
    labels = torch.empty(B, m, device=Zb.device, dtype=torch.int64)
    num = torch.empty(B, device=Zb.device, dtype=torch.int32)
    rc = _lib.lib().msm_seed_connected_components(Zb.data_ptr(), labels.data_ptr(), num.data_ptr(), B, m, d,
                                                  float(epsilon), _stream())
    check(rc, "msm_seed_connected_components")
    return (labels[0], num[0]) if squeeze else (labels, num)


def assign_clusters(X, Z, seed_labels, num_labels):
    """Label every point by its closest seed and make the most populous cluster label 0
    (mean_shift.py:207-227), batched: -> labels int64 [B,n]."""
    Xb, squeeze = _points(X)
    Zb, _ = _points(Z, "Z")
    B, n, d = Xb.shape
    m = Zb.shape[1]
    if Zb.shape[0] != B or Zb.shape[2] != d:
        raise ValueError(f"X {tuple(X.shape)} / Z {tuple(Z.shape)} mismatch")
    if d not in (16, 32, 64, 128):
        raise ValueError(f"embedding dim {d} not supported (16, 32, 64, 128)")
    sl = seed_labels.to(device=Xb.device, dtype=torch.int64).reshape(B, m).contiguous()
    nl = num_labels.to(device=Xb.device, dtype=torch.int32).reshape(B).contiguous()
    labels = torch.empty(B, n, device=Xb.device, dtype=torch.int64)
    L = _lib.lib()
    ws_bytes = L.msm_assign_clusters_workspace_bytes(B, m)
    ws = torch.empty(ws_bytes, device=Xb.device, dtype=torch.uint8)
    rc = L.msm_assign_clusters(Xb.data_ptr(), Zb.data_ptr(), sl.data_ptr(), nl.data_ptr(), labels.data_ptr(), B, n, m,
                               d, ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msm_assign_clusters")
    return labels[0] if squeeze else labels


def instance_topk(pred_logits, topk):
    """softmax over classes, drop 'no object', keep the ``topk`` best (query, class) pairs per image:
    pred_logits [B,Q,K+1] -> (query int64 [B,T], class int64 [B,T], score [B,T]), descending score."""
    lg = _require(pred_logits, "pred_logits").contiguous()
    if lg.dim() != 3:
        raise ValueError(f"pred_logits must be [B,Q,K+1], got {tuple(lg.shape)}")
    B, Q, K1 = lg.shape
    T = int(topk)
    if not 1 <= T <= Q * (K1 - 1):
        raise ValueError(f"topk={T} outside [1, {Q * (K1 - 1)}]")
    query = torch.empty(B, T, device=lg.device, dtype=torch.int64)
    cls = torch.empty(B, T, device=lg.device, dtype=torch.int64)
    score = torch.empty(B, T, device=lg.device, dtype=torch.float32)
    rc = _lib.lib().msm_instance_topk(lg.data_ptr(), query.data_ptr(), cls.data_ptr(), score.data_ptr(), B, Q, K1, T,
                                      _stream())
    check(rc, "msm_instance_topk")
    return query, cls, score


def instance_masks(pred_masks, topk_query, topk_score, size, want_masks=True):
    """kept queries only: bilinear upsample to ``size`` + threshold + box + mask score in one pass:
    pred_masks [B,Q,h,w] logits -> (masks 0/1 float [B,T,H,W] or None, boxes [B,T,4], scores [B,T])."""
    pm = _require(pred_masks, "pred_masks").contiguous()
    B, Q, h, w = pm.shape
    H, W = int(size[0]), int(size[1])
    tq = topk_query.to(device=pm.device, dtype=torch.int64).contiguous()
    ts = _require(topk_score, "topk_score").contiguous()
    T = tq.shape[1]
    if tq.shape != (B, T) or ts.shape != (B, T):
        raise ValueError(f"topk_query {tuple(tq.shape)} / topk_score {tuple(ts.shape)} must both be [{B}, T]")
    masks = torch.empty(B, T, H, W, device=pm.device, dtype=torch.float32) if want_masks else None
    boxes = torch.empty(B, T, 4, device=pm.device, dtype=torch.float32)
    scores = torch.empty(B, T, device=pm.device, dtype=torch.float32)
    L = _lib.lib()
    ws_bytes = L.msm_instance_masks_workspace_bytes(B, T, H)
    ws = torch.empty(ws_bytes, device=pm.device, dtype=torch.uint8)
    rc = L.msm_instance_masks(pm.data_ptr(), tq.data_ptr(), ts.data_ptr(), masks.data_ptr() if want_masks else 0,
                              boxes.data_ptr(),
                              scores.data_ptr(), B, Q, h, w, T, H, W, ws.data_ptr(), ws_bytes, _stream())
    check(rc, "msm_instance_masks")
    return masks, boxes, scores


def instance_label_map(pred_masks, topk_query, scores, classes, size, topk_mode=False, num_class=2, score=0.7,
                       low_threshold=0.4):
    """confident instances -> label map without the full-resolution masks: pred_masks [B,Q,h,w] logits, topk_query /
    scores / classes [B,T] -> (label_map fp32 [B,H,W], instance_label int32 [B,T]: 2.. for kept instances, -1 else)."""
    pm = _require(pred_masks, "pred_masks").contiguous()
    B, Q, h, w = pm.shape
    H, W = int(size[0]), int(size[1])
    tq = topk_query.to(device=pm.device, dtype=torch.int64).contiguous()
    sc = _require(scores, "scores").contiguous()
    cl = classes.to(device=pm.device, dtype=torch.int64).contiguous()
    T = tq.shape[1]
    if T > 64:
        raise ValueError(f"at most 64 instances per image, got {T}")
    inst = torch.empty(B, T, device=pm.device, dtype=torch.int32)
    out = torch.empty(B, H, W, device=pm.device, dtype=torch.float32)
    rc = _lib.lib().msm_instance_label_map(pm.data_ptr(), tq.data_ptr(), sc.data_ptr(), cl.data_ptr(), inst.data_ptr(),
                                           out.data_ptr(), B, Q, h, w, T, H, W, 1 if topk_mode else 0, int(num_class),
                                           float(score), float(low_threshold), _stream())
    check(rc, "msm_instance_label_map")
    return out, inst


def label_stats(labels, depth=None, num_ids=None):
    """labels [N,H,W] fp32 ids, depth [N,3,H,W] or None -> int32 [N,L,6] on the device:
    (pixels, pixels with depth z > 0, W - xmin, H - ymin, xmax + 1, ymax + 1), zeros for an absent id."""
    lb = _require(labels, "labels").contiguous()
    if lb.dim() != 3:
        raise ValueError(f"labels must be [N,H,W], got {tuple(lb.shape)}")
    N, H, W = lb.shape
    L = int(num_ids) if num_ids is not None else int(lb.max().item()) + 1
    if not 1 <= L <= 1024:
        raise ValueError(f"label ids must lie in [0, 1024), got max id {L - 1}")
    dz, stride = 0, 0
    if depth is not None:
        dp = _require(depth, "depth").contiguous()
        if dp.shape != (N, 3, H, W):
            raise ValueError(f"depth must be [{N},3,{H},{W}], got {tuple(dp.shape)}")
        dz, stride = dp.data_ptr() + 2 * H * W * 4, 3 * H * W
    stats = torch.empty(N, L, 6, device=lb.device, dtype=torch.int32)
    rc = _lib.lib().msm_label_stats(lb.data_ptr(), dz, stride, stats.data_ptr(), N, H, W, L, _stream())
    check(rc, "msm_label_stats")
    return stats


def relabel_lut(labels, lut, lo=0):
    """out = lut[n][labels - lo] for ids in [lo, lo + L), unchanged elsewhere. labels [N,...] fp32, lut [N,L] fp32."""
    lb = _require(labels, "labels").contiguous()
    lt = _require(lut, "lut").contiguous()
    N = lb.shape[0]
    if lt.dim() != 2 or lt.shape[0] != N:
        raise ValueError(f"lut must be [{N}, L], got {tuple(lt.shape)}")
    out = torch.empty_like(lb)
    rc = _lib.lib().msm_relabel_lut(lb.data_ptr(), lt.data_ptr(), out.data_ptr(), N, lb.numel() // N, lt.shape[1],
                                    int(lo), _stream())
    check(rc, "msm_relabel_lut")
    return out


def crop_resize(rgb, depth, labels, rois, ids, crop_size):
    """every ROI of one image at once: rgb [3,H,W], depth [3,H,W] or None, labels [H,W], rois int32 [num,4]
    (x_min, y_min, x_max, y_max), ids fp32 [num] -> (rgb_crops [num,3,S,S], depth_crops or None, mask_crops [num,S,S])."""
    im = _require(rgb, "rgb").contiguous()
    lb = _require(labels, "labels").contiguous()
    _, H, W = im.shape
    num, S = rois.shape[0], int(crop_size)
    r = rois.to(device=im.device, dtype=torch.int32).contiguous()
    i = ids.to(device=im.device, dtype=torch.float32).contiguous()
    rgb_crops = torch.empty(num, 3, S, S, device=im.device, dtype=torch.float32)
    mask_crops = torch.empty(num, S, S, device=im.device, dtype=torch.float32)
    dp = _require(depth, "depth").contiguous() if depth is not None else None
    depth_crops = torch.empty(num, 3, S, S, device=im.device, dtype=torch.float32) if dp is not None else None
    if num:
        rc = _lib.lib().msm_crop_resize(im.data_ptr(), dp.data_ptr() if dp is not None else 0, lb.data_ptr(),
                                        r.data_ptr(), i.data_ptr(), rgb_crops.data_ptr(),
                                        depth_crops.data_ptr() if dp is not None else 0, mask_crops.data_ptr(), num, H, W,
                                        S, _stream())
        check(rc, "msm_crop_resize")
    return rgb_crops, depth_crops, mask_crops


def crop_label_stats(labels_crop, init_crop, depth_crop=None, num_ids=None):
    """labels_crop / init_crop [num,S,S], depth_crop [num,3,S,S] or None -> (int32 [num,L,4] = pixels, pixels where
    init_crop != 0, pixels with depth z > 0, 0;  float64 [num,L] depth sums)."""
    lb = _require(labels_crop, "labels_crop").contiguous()
    ic = _require(init_crop, "init_crop").contiguous()
    num, S = lb.shape[0], lb.shape[-1]
    L = int(num_ids) if num_ids is not None else int(lb.max().item()) + 1
    if not 1 <= L <= 1024:
        raise ValueError(f"label ids must lie in [0, 1024), got max id {L - 1}")
    dp = _require(depth_crop, "depth_crop").contiguous() if depth_crop is not None else None
    stats = torch.empty(num, L, 4, device=lb.device, dtype=torch.int32)
    dsum = torch.empty(num, L, device=lb.device, dtype=torch.float64)
    rc = _lib.lib().msm_crop_label_stats(lb.data_ptr(), ic.data_ptr(), dp.data_ptr() if dp is not None else 0,
                                         stats.data_ptr(), dsum.data_ptr(), num, S, L, _stream())
    check(rc, "msm_crop_label_stats")
    return stats, dsum


def paste_crops(labels_crop, new_label, order, rois, height, width):
    """refined [H,W]: crops [num,S,S] resized back (nearest) into their ROIs in ``order``; later ones overwrite
    earlier ones wherever their relabelled value new_label[c][id] is non-zero."""
    lb = _require(labels_crop, "labels_crop").contiguous()
    num, S = lb.shape[0], lb.shape[-1]
    nl = _require(new_label, "new_label").contiguous()
    od = order.to(device=lb.device, dtype=torch.int32).contiguous()
    r = rois.to(device=lb.device, dtype=torch.int32).contiguous()
    refined = torch.empty(int(height), int(width), device=lb.device, dtype=torch.float32)
    rc = _lib.lib().msm_paste_crops(lb.data_ptr(), nl.data_ptr(), od.data_ptr(), r.data_ptr(), refined.data_ptr(), num,
                                    int(height), int(width), S, nl.shape[1], _stream())
    check(rc, "msm_paste_crops")
    return refined


# ----------------------------------------------------------------------------------------------
# launch accounting / per-op device timing (used by bench.py; off by default)
# ----------------------------------------------------------------------------------------------
class _Stats:
    launches = 0          # kernels of THIS library enqueued since reset
    timing = False        # when True every op is bracketed by CUDA events on the launching stream
    events = []           # (tag, shape signature, algorithmic bytes, flops, start_event, end_event)


def reset_stats(timing=False):
    _Stats.launches = 0
    _Stats.timing = timing
    _Stats.events = []


def launches():
    return _Stats.launches


def op_times_ms():
    """tag -> (count, total ms); call after torch.cuda.synchronize()."""
    agg = {}
    for tag, _, _, _, a, b in _Stats.events:
        c, t = agg.get(tag, (0, 0.0))
        agg[tag] = (c + 1, t + a.elapsed_time(b))
    return agg


def op_groups():
    """(tag, shape signature) -> dict(count, ms, bytes, flops): launches of one kernel at one shape, with the
    ALGORITHMIC bytes / flops of each launch. Call after torch.cuda.synchronize()."""
    agg = {}
    for tag, sig, by, fl, a, b in _Stats.events:
        g = agg.setdefault((tag, sig), {"count": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
        g["count"] += 1
        g["ms"] += a.elapsed_time(b)
        g["bytes"] += by
        g["flops"] += fl
    return agg


def _instrument(tag, kernels, work=None):
    """work(*args, **kwargs) -> (shape signature, algorithmic bytes, flops) of one call."""
    def deco(fn):
        def wrapped(*args, **kwargs):
            n = kernels(*args, **kwargs) if callable(kernels) else kernels
            _Stats.launches += n
            if not _Stats.timing:
                return fn(*args, **kwargs)
            sig, by, fl = work(*args, **kwargs) if work is not None else ("", 0.0, 0.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn(*args, **kwargs)
            b.record()
            _Stats.events.append((tag, sig, by, fl, a, b))
            return r
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _work_vmf(q, k, v, **kw):
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    shared = k.data_ptr() == v.data_ptr()
    by = 4.0 * B * H * hd * (Ns * (1 if shared else 2) + 2 * Nq) + (B * Nq * Ns / 8.0 if kw.get("blocked_bits") is not None else 0)
    return f"B{B} H{H} Q{Nq} S{Ns} hd{hd}", by, 4.0 * B * H * Nq * Ns * hd


def _work_mask(e, f, out=None):
    B, Q, C = e.shape
    hw = f.shape[2] * f.shape[3]
    return f"B{B} Q{Q} C{C} HW{hw}", 4.0 * B * (C * hw + Q * hw + Q * C), 2.0 * B * Q * C * hw


def _work_bits(m, target):
    B, Q, H, W = m.shape
    return f"B{B} Q{Q} {H}x{W}->{int(target[0])}x{int(target[1])}", 4.0 * B * Q * H * W + B * Q * target[0] * target[1] / 8.0, 0.0


def _work_linear(x, w, bias=None, relu=False, out=None):
    N, K = w.shape
    M = x.numel() // K
    return f"M{M} N{N} K{K}", 4.0 * (M * K + M * N) + 4.0 * N * K, 2.0 * M * N * K


def _work_linear_ln(x, w, bias, residual, norm):
    N, K = w.shape
    M = x.numel() // K
    return f"+res+LN M{M} N{N} K{K}", 4.0 * (M * K + 2 * M * N) + 4.0 * N * K, 2.0 * M * N * K


def _work_linear_fused(x, w, bias=None, **kw):
    N, K = w.shape
    M = x.numel() // K
    extra = (1 if kw.get("residual") is not None else 0) + (1 if kw.get("norm2") is not None else 0)
    return f"fused M{M} N{N} K{K}", 4.0 * (M * K + (1 + extra) * M * N) + 4.0 * N * K, 2.0 * M * N * K


def _work_ffn(x, w1, b1, w2, b2, norm):
    D = x.shape[-1]
    F = w1.shape[0]
    M = x.numel() // D
    return f"ffn+LN M{M} D{D} F{F}", 4.0 * (3 * M * D) + 8.0 * D * F, 4.0 * M * D * F


def _work_conv(x, weight, bias=None, relu=False, tokens_out=False):
    B, K, H, W = x.shape
    N = weight.shape[0]
    M = B * H * W
    return f"conv1x1 M{M} N{N} K{K}", 4.0 * (M * K + M * N) + 4.0 * N * K, 2.0 * M * N * K


def _work_conv3(x, weight, bias=None, relu=False):
    B, C, H, W = x.shape
    N = weight.shape[0]
    M = B * H * W
    return f"conv3x3 M{M} N{N} C{C}", 4.0 * (M * C + M * N) + 36.0 * N * C, 18.0 * M * N * C


def _work_msda(value, shapes, lsi, loc, aw, *a, **k):
    N, S, M, D = value.shape
    Lq = loc.shape[1]
    by = 4.0 * (value.numel() + loc.numel() + aw.numel() + N * Lq * M * D)
    return f"N{N} S{S} M{M} D{D} Lq{Lq}", by, 0.0


def _work_msda_fused(value, shapes, lsi, ol, ref, L, P):
    N, S, M, D = value.shape
    Lq = ol.shape[1]
    by = 4.0 * (value.numel() + ol.numel() + ref.numel() + N * Lq * M * D)
    return f"fused N{N} S{S} M{M} D{D} Lq{Lq}", by, 0.0


def _work_ms(X, Z, kappa, max_iters=10):
    n, d = X.shape[-2], X.shape[-1]
    B = X.shape[0] if X.dim() == 3 else 1
    m = Z.shape[-2]
    return f"B{B} n{n} m{m} d{d} it{max_iters}", 4.0 * B * n * d * max_iters, 4.0 * B * m * n * d * max_iters


vmf_attention = _instrument("vmf_attention", 2, _work_vmf)(vmf_attention)
vmf_attention_weights = _instrument("vmf_attention_weights", 1)(vmf_attention_weights)


def _work_vmf_bwd(q, k, v, out, grad_out, den, **kw):
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    by = 4.0 * B * H * hd * 4 * (Ns + Nq) + (B * Nq * Ns / 8.0 if kw.get("blocked_bits") is not None else 0)
    return f"B{B} H{H} Q{Nq} S{Ns} hd{hd}", by, 10.0 * B * H * Nq * Ns * hd  # two score products, three accumulations


vmf_attention_bwd = _instrument("vmf_attention_bwd", 2, _work_vmf_bwd)(vmf_attention_bwd)
def _work_linear_packed(x, weight, bias, images, batch, num_keys, channels, which, pos=None):
    N, K = weight.shape
    M = batch * num_keys
    # reads the fp32 rows, writes 16-bit hi + lo operand images (4 bytes per element, like fp32 rows)
    return f"{'V' if which else 'K'}-images M{M} N{N} K{K}", 4.0 * (M * K + M * N) + 4.0 * N * K, 2.0 * M * N * K


def _work_vmf_packed(q, kv, **kw):
    B, H, Nq, hd = q.shape
    Ns = kv.num_keys
    by = 4.0 * B * H * hd * (2 * Ns + 2 * Nq) + (B * Nq * Ns / 8.0 if kw.get("blocked_bits") is not None else 0)
    return f"packed B{B} H{H} Q{Nq} S{Ns} hd{hd}", by, 4.0 * B * H * Nq * Ns * hd


linear_packed_kv = _instrument("linear", 1, _work_linear_packed)(linear_packed_kv)
vmf_attention_packed = _instrument("vmf_attention", 2, _work_vmf_packed)(vmf_attention_packed)
mask_logits = _instrument("mask_logits", 1, _work_mask)(mask_logits)
mask_to_attn_bits = _instrument("mask_to_attn_bits", 1, _work_bits)(mask_to_attn_bits)  # one kernel, no memset
maxpool3x3s2_channels_last = _instrument("maxpool", 1, lambda x: (
    f"{tuple(x.shape)}", 4.0 * x.numel() * 1.25, 0.0))(maxpool3x3s2_channels_last)
upsample_add = _instrument("upsample_add", 1, lambda x, add: (
    f"{tuple(x.shape)}->{tuple(add.shape[-2:])}", 4.0 * (x.numel() + 2 * add.numel()), 0.0))(upsample_add)
resample_bilinear = _instrument("resample_bilinear", 1, lambda x, size, align_corners=False: (
    f"{tuple(x.shape)}->{int(size[0])}x{int(size[1])}", 4.0 * (x.numel() // (x.shape[-1] * x.shape[-2])) * int(size[0]) * int(size[1]) * 5, 0.0))(resample_bilinear)
linear = _instrument("linear", 1, _work_linear)(linear)
conv1x1 = _instrument("linear", 1, _work_conv)(conv1x1)
conv1x1_nhwc = _instrument("linear", lambda x, *a, **k: 0 if (x.shape[2] * x.shape[3]) % 128 else 1,
                           _work_conv)(conv1x1_nhwc)   # (H*W % 128 != 0 goes through `linear`, which counts itself)
conv3x3 = _instrument("linear", 1, _work_conv3)(conv3x3)
linear_ln = _instrument("linear", 1, _work_linear_ln)(linear_ln)
ffn_ln = _instrument("ffn", 1, _work_ffn)(ffn_ln)
decoder_block = _instrument("decoder_block", 1, lambda o, st, blob, **kw: (
    f"B{o.shape[0]} Q{o.shape[1]} C256 F2048", 4.0 * o.numel() * 5 + blob.numel() * o.shape[0],
    2.0 * o.shape[0] * o.shape[1] * 256 * (256 * 7.125 + 2 * 2048) + 4.0 * o.shape[0] * o.shape[1] ** 2 * 256))(decoder_block)
add_layernorm = _instrument("add_layernorm", 1, lambda x, y, norm, l2_normalize=False, norm2=None: (
    f"rows{x.numel() // x.shape[-1]} C{x.shape[-1]}", 4.0 * x.numel() * (3 + (1 if norm2 is not None else 0)), 0.0))(
    add_layernorm)
linear_fused = _instrument("linear", 1, _work_linear_fused)(linear_fused)
ms_deform_attn_forward = _instrument("ms_deform_attn_forward", 1, _work_msda)(ms_deform_attn_forward)
ms_deform_attn_fused_forward = _instrument("ms_deform_attn_forward", 1, _work_msda_fused)(ms_deform_attn_fused_forward)
ms_deform_attn_backward = _instrument("ms_deform_attn_backward", 1)(ms_deform_attn_backward)
select_smart_seeds = _instrument("select_smart_seeds", 1, lambda X, num_seeds, first_index: (
    f"B{X.shape[0] if X.dim() == 3 else 1} n{X.shape[-2]} m{num_seeds} d{X.shape[-1]}",
    (4.0 * X.shape[-1] + 8.0) * (X.numel() // X.shape[-1]) * (num_seeds - 1), 2.0 * X.numel() * (num_seeds - 1)))(
    select_smart_seeds)
seed_connected_components = _instrument("seed_connected_components", 1)(seed_connected_components)
assign_clusters = _instrument("assign_clusters", 2, lambda X, Z, seed_labels, num_labels: (
    f"B{X.shape[0] if X.dim() == 3 else 1} n{X.shape[-2]} m{Z.shape[-2]} d{X.shape[-1]}",
    (4.0 * X.shape[-1] + 8.0) * (X.numel() // X.shape[-1]), 2.0 * X.numel() * Z.shape[-2]))(assign_clusters)
instance_label_map = _instrument("instance_label_map", 2)(instance_label_map)
label_stats = _instrument("label_stats", 1)(label_stats)
relabel_lut = _instrument("relabel_lut", 1)(relabel_lut)
crop_resize = _instrument("crop_resize", 1)(crop_resize)
crop_label_stats = _instrument("crop_label_stats", 1)(crop_label_stats)
paste_crops = _instrument("paste_crops", 1)(paste_crops)
instance_topk = _instrument("instance_topk", 1)(instance_topk)
instance_masks = _instrument("instance_masks", 2, lambda pred_masks, topk_query, topk_score, size, want_masks=True: (
    f"B{pred_masks.shape[0]} T{topk_query.shape[1]} {pred_masks.shape[2]}x{pred_masks.shape[3]}->{int(size[0])}x{int(size[1])}",
    4.0 * pred_masks.shape[0] * topk_query.shape[1] * ((int(size[0]) * int(size[1]) if want_masks else 0)
                                                        + pred_masks.shape[2] * pred_masks.shape[3]),
    0.0))(instance_masks)
mean_shift_hill_climb = _instrument("mean_shift_hill_climb", lambda X, Z, kappa, max_iters=10: 2 * int(max_iters),
                                    _work_ms)(mean_shift_hill_climb)
