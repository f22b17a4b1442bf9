"""The TEXT of the CUDA-core training kernel (csrc/vmf_attention_bwd.cu), executed on CPU threads.

The authoring container has no GPU and the round's GPU minutes were spent before this kernel was written, so its
source is compiled as plain C++ against tests/emu/cuda_emu.h (every CUDA thread an OS thread, __syncthreads a barrier,
warp shuffles through a per-warp slot array) and driven through the SAME C entry point, msm_vmf_attention_bwd. This
checks indexing, tiling, masks, strides and the split reduction of the real source against the reference's gradients
(golden) and fp64 autograd; it says nothing about what nvcc / the hardware do with it - that is the staged GPU test
(tests/test_gpu_staged.py)."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(ROOT, "build", "emu", "libemu_vmf_bwd.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    srcs = [os.path.join(EMU_DIR, "emu_vmf_bwd.cpp"), os.path.join(EMU_DIR, "cuda_emu.h"),
            os.path.join(ROOT, "unseenobjectswithmeanshift_b200", "csrc", "vmf_attention_bwd.cu"),
            os.path.join(ROOT, "unseenobjectswithmeanshift_b200", "csrc", "common.cuh")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in srcs):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-w", "-shared", "-fPIC", "-pthread", "-DMSM_EMULATE_ON_HOST",
                               "-I" + CUDA_INC, "-x", "c++", srcs[0], "-o", EMU_LIB])
    from unseenobjectswithmeanshift_b200._lib import SIGNATURES
    h = ctypes.CDLL(EMU_LIB)
    for name in ("msm_vmf_attention_bwd", "msm_vmf_attention_bwd_workspace_bytes"):
        fn = getattr(h, name)
        fn.restype, fn.argtypes = SIGNATURES[name]
    return h


def _bwd(h, q, k, v, out, gout, den, bits=None, row_open=None, add_mask=None, kappa=30.0, flags=3):
    """[B,H,L,hd] CPU views in, gradients as [B,H,L,hd] views of [B,L,H*hd] buffers out (as ops.vmf_attention_bwd)."""
    B, H, Nq, hd = q.shape
    Ns = k.shape[2]
    grad = lambda L: torch.full((B, L, H, hd), float("nan")).permute(0, 2, 1, 3)
    gq, gk, gv = grad(Nq), grad(Ns), grad(Ns)
    wsb = h.msm_vmf_attention_bwd_workspace_bytes(B, H, Nq, Ns, hd)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8)
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    rc = h.msm_vmf_attention_bwd(*st(q), *st(k), *st(v), *st(out), *st(gout), den.data_ptr(), *st(gq), *st(gk), *st(gv),
                                 bits.data_ptr() if bits is not None else None, bits.shape[2] if bits is not None else 0,
                                 row_open.data_ptr() if row_open is not None else None,
                                 add_mask.data_ptr() if add_mask is not None else None,
                                 B, H, Nq, Ns, hd, kappa, flags, ws.data_ptr(), wsb, None)
    assert rc == 0, rc
    return gq, gk, gv


def _forward_planes(q, k, v, blocked, kappa, nq=True, nk=True):
    """What msm_vmf_attention_fwd hands to the backward under MSM_VMF_SAVE_NORM: out and [2,G,Nq] (den, |p.v|)."""
    B, H, Nq, hd = q.shape
    qn = F.normalize(q.double(), dim=-1, eps=1e-12) if nq else q.double()
    kn = F.normalize(k.double(), dim=-1, eps=1e-12) if nk else k.double()
    w = torch.exp(kappa * (qn @ kn.transpose(-1, -2)) - kappa)
    if blocked is not None:
        w = w.masked_fill(blocked, 0.0)
    den = w.sum(-1)
    o = (w @ v.double()) / den.unsqueeze(-1)
    norm = o.norm(dim=-1)
    out = torch.empty(B, Nq, H, hd).permute(0, 2, 1, 3)
    out.copy_(o / norm.clamp_min(1e-12).unsqueeze(-1))
    return out, torch.stack([den.reshape(B * H, Nq), norm.reshape(B * H, Nq)]).float().contiguous()


def _close(got, want, rel):
    err = (got.double() - want.double()).abs().max().item()
    scale = max(want.abs().max().item(), 1e-30)
    assert err <= rel * scale, f"max abs err {err:.3e} vs peak {scale:.3e}"


def test_emulated_kernel_matches_reference_gradients(emu, golden):
    """Golden from torch.autograd through the reference's hypersphere_attention; head dim 8 (padded to 32 lanes),
    additive -inf mask, two kappas."""
    g, _ = golden("hypersphere_attention_bwd")
    q, k, v = (g[n].unsqueeze(1).contiguous() for n in "qkv")
    gout = g["grad_out"].unsqueeze(1).contiguous()
    fmask = torch.zeros(g["blocked"].shape).masked_fill_(g["blocked"], float("-inf")).contiguous()
    for tag, mask, blocked, kappa in (("masked", fmask, g["blocked"].unsqueeze(1), 30.0), ("nomask", None, None, 30.0),
                                      ("kappa10", None, None, 10.0)):
        out, den = _forward_planes(q, k, v, blocked, kappa)
        _close(out.squeeze(1), g[f"out_{tag}"], 1e-5)
        gq, gk, gv = _bwd(emu, q, k, v, out, gout, den, add_mask=mask, kappa=kappa)
        for t, name in ((gq, "gq"), (gk, "gk"), (gv, "gv")):
            _close(t.squeeze(1), g[f"{name}_{tag}"], 1e-4)


def _pack_bits(blocked):
    B, Q, S = blocked.shape
    words = (S + 31) // 32
    pad = torch.zeros(B, Q, words * 32, dtype=torch.bool)
    pad[..., :S] = blocked
    v = (pad.view(B, Q, words, 32).long() << torch.arange(32)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


@pytest.mark.parametrize("B,H,Q,S,hd,masked,flags", [
    (1, 2, 100, 150, 32, True, 3),    # the decoder's geometry: strided heads, bit mask, key tail (150 = 2 x 64 + 22)
    (1, 1, 37, 200, 64, False, 3),    # head dim 64: the other thread mapping
    (2, 1, 128, 64, 16, True, 3),     # maximum number of query rows, padded head dim, exactly one tile
    (1, 1, 20, 100, 32, False, 0),    # mean-shift form: no normalisation of q / k
])
def test_emulated_kernel_vs_fp64_autograd(emu, B, H, Q, S, hd, masked, flags):
    torch.manual_seed(S + hd + Q)
    C = H * hd
    qb, kvb = torch.randn(B, Q, C), torch.randn(B, S, 2 * C)
    if flags == 0:
        qb, kvb = F.normalize(qb, dim=-1), F.normalize(kvb.view(B, S, 2, C), dim=-1).reshape(B, S, 2 * C)
    heads = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    q, k, v = heads(qb), heads(kvb[..., :C]), heads(kvb[..., C:])   # k | v share one buffer: row stride 2C
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S) < 0.5
        blocked[:, 3] = True                                        # a fully blocked row counts as open
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    kappa = 30.0 if flags else 10.0
    out, den = _forward_planes(q, k, v, eff, kappa, nq=bool(flags & 1), nk=bool(flags & 2))
    gout = torch.randn(B, H, Q, hd)
    gq, gk, gv = _bwd(emu, q, k, v, out, gout, den, bits=bits, row_open=ro, kappa=kappa, flags=flags)

    q2, k2, v2 = (t.double().clone().requires_grad_() for t in (q, k, v))
    qn = F.normalize(q2, dim=-1, eps=1e-12) if flags & 1 else q2
    kn = F.normalize(k2, dim=-1, eps=1e-12) if flags & 2 else k2
    s = kappa * qn @ kn.transpose(-1, -2)
    if eff is not None:
        s = s.masked_fill(eff, float("-inf"))
    ref = F.normalize(torch.softmax(s, -1) @ v2, dim=-1, eps=1e-12)
    rq, rk, rv = torch.autograd.grad(ref, (q2, k2, v2), gout.double())
    _close(gq, rq, 1e-4)
    _close(gk, rk, 1e-4)
    _close(gv, rv, 1e-4)
    assert not any(torch.isnan(t).any() for t in (gq, gk, gv))      # every element of the outputs was written


def test_emulated_entry_point_rejects_bad_arguments(emu):
    q = torch.randn(1, 1, 200, 32)
    k = torch.randn(1, 1, 64, 32)
    out, den = _forward_planes(q, k, k, None, 30.0)
    h = emu
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    g = torch.empty_like(q)
    gk = torch.empty_like(k)
    args = (*st(q), *st(k), *st(k), *st(out), *st(out), den.data_ptr(), *st(g), *st(gk), *st(gk), None, 0, None, None,
            1, 1, 200, 64, 32, 30.0, 3, None, 0, None)
    assert h.msm_vmf_attention_bwd(*args) == -2          # MSM_E_UNSUPPORTED: more than 128 query rows
    args = (*st(q), *st(k), *st(k), *st(out), *st(out), den.data_ptr(), *st(g), *st(gk), *st(gk), None, 0, None, None,
            1, 1, 100, 64, 32, 30.0, 3, None, 0, None)
    assert h.msm_vmf_attention_bwd(*args) == -3          # MSM_E_WORKSPACE: no workspace
