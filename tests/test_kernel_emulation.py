"""The TEXT of CUDA kernels executed on CPU threads (tests/emu).

The authoring container has no GPU and the round's GPU minutes were spent before some kernels were written, so their
sources are compiled as plain C++ against tests/emu/cuda_emu.h (every CUDA thread an OS thread, __syncthreads a
barrier, warp shuffles through a per-warp slot array) and tests/emu/tc_emu.h (mbarrier, TMEM, tcgen05.mma with
no-swizzle shared-memory descriptors, bulk copies) and driven through their real host entry points:

* csrc/vmf_attention_bwd.cu (training, CUDA cores): against the reference's gradients (golden) and fp64 autograd;
* csrc/vmf_attention_tc.cu (SHIPPED tcgen05 kernel, parity-green on the B200): the CALIBRATION of the tensor-core
  emulation - it has to reproduce what the hardware is known to produce for this kernel;
* csrc/linear_tc.cu (SHIPPED dense / convolution kernel: tiled TMA with 128B swizzle, operand conversion into TMEM,
  fused epilogues, TMA stores): calibration of the tensor-map emulation, and a CPU development loop for the kernel
  that dominates the headline step;
* csrc/vmf_attention_packed.cu (tcgen05 + bulk copies; written under this emulation in round 1, parity-green on the
  B200 since round 2).

This checks indexing, tiling, masks, strides, descriptors and barrier protocols of the real source. By default
asynchronous operations (TMA, MMA, commits) execute at issue; in LATE mode (emu_set_late) they execute as late as the
barrier protocol allows - when a thread is about to block on the barrier they signal, the tensor pipe in issue order,
TMA stores when their bulk group is waited for - so code that consumes an operand or a result without waiting, or
recycles a buffer too early, computes garbage. Nothing is said about what nvcc / the hardware do with the code - that
is the job of the -m gpu tests."""
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


CSRC = os.path.join(ROOT, "unseenobjectswithmeanshift_b200", "csrc")


def _build(lib, driver, sources):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    deps = [os.path.join(EMU_DIR, f) for f in (driver, "cuda_emu.h", "tc_emu.h")] + \
           [os.path.join(CSRC, f) for f in sources + ["common.cuh", "tc.cuh"]]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-w", "-shared", "-fPIC", "-pthread", "-DMSM_EMULATE_ON_HOST",
                               "-I" + CUDA_INC, "-I" + EMU_DIR, "-x", "c++", deps[0], "-o", lib])
    return ctypes.CDLL(lib)


@pytest.fixture(scope="module")
def emu():
    from unseenobjectswithmeanshift_b200._lib import SIGNATURES
    h = _build(EMU_LIB, "emu_vmf_bwd.cpp", ["vmf_attention_bwd.cu"])
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


def _aligned(nbytes, align=1024):
    buf = torch.zeros(nbytes + align, dtype=torch.uint8)
    off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]


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


# ------------------------------------------------------------------------------------------------------------------
# tensor-core kernels through tests/emu/tc_emu.h
@pytest.fixture(scope="module")
def emu_tc():
    h = _build(os.path.join(ROOT, "build", "emu", "libemu_vmf_tc.so"), "emu_vmf_tc.cpp",
               ["vmf_attention_tc.cu", "vmf_attention_packed.cu"])
    P, I, L, Fl, Z, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_double
    h.emu_last_error.restype = ctypes.c_char_p
    h.emu_set_timeout.argtypes = [D]
    h.emu_vmf_tc_workspace_bytes.restype, h.emu_vmf_tc_workspace_bytes.argtypes = Z, [I, I, I, I]
    h.emu_vmf_attention_tc_partial.restype = I
    h.emu_vmf_attention_tc_partial.argtypes = [P, L, L, L] * 3 + [P, I, P, I, I, I, I, I, Fl, I, P, P, D]
    h.msmx_mean_shift_packed_bytes.restype, h.msmx_mean_shift_packed_bytes.argtypes = Z, [I, I, I]
    h.msmx_mean_shift_packed_workspace_bytes.restype = Z
    h.msmx_mean_shift_packed_workspace_bytes.argtypes = [I, I, I, I]
    h.msmx_mean_shift_pack.restype, h.msmx_mean_shift_pack.argtypes = I, [P, P, I, I, I, P]
    h.msmx_mean_shift_hill_climb_packed.restype = I
    h.msmx_mean_shift_hill_climb_packed.argtypes = [P, P, P, I, I, I, I, Fl, I, P, Z, P]
    h.msmx_vmf_packed_bytes.restype, h.msmx_vmf_packed_bytes.argtypes = Z, [I, I, I, I, I]
    h.msmx_vmf_packed_workspace_bytes.restype, h.msmx_vmf_packed_workspace_bytes.argtypes = Z, [I, I, I, I, I]
    h.msmx_vmf_pack.restype, h.msmx_vmf_pack.argtypes = I, [P, L, L, L, P, L, L, L, P, I, I, I, I, I, P]
    h.msmx_vmf_attention_packed_fwd.restype = I
    h.msmx_vmf_attention_packed_fwd.argtypes = [P, L, L, L, P, P, L, L, L, P, I, P, I, I, I, I, I, Fl, I, P, Z, P]
    return h


@pytest.mark.parametrize("B,H,Q,S,hd,shared,masked", [
    (1, 1, 20, 300, 32, False, False),    # fp16 score operands (q and k normalised by the kernel), 3 key tiles
    (1, 2, 100, 700, 32, False, True),    # bit mask shared by the heads, 6 tiles: the 4-stage ring wraps
    (1, 1, 100, 333, 64, True, False),    # mean-shift form: k == v, one bf16 copy serves both products
])
def test_calibration_shipped_tcgen05_attention_kernel(emu_tc, B, H, Q, S, hd, shared, masked):
    """vmf_attn_tc_kernel is parity-green on the B200 (tests/test_gpu_parity.py); the emulation must agree with it -
    i.e. with the fp64 reference at split-precision accuracy - before it is used to judge anything else."""
    h = emu_tc
    torch.manual_seed(S + hd)
    C = H * hd
    q, k = torch.randn(B, Q, C), torch.randn(B, S, C)
    v = k if shared else torch.randn(B, S, C)
    kappa, flags = (10.0, 0) if shared else (30.0, 3)
    if shared:
        k = v = F.normalize(k.view(B, S, H, hd), dim=-1).view(B, S, C)
        q = F.normalize(q.view(B, Q, H, hd), dim=-1).view(B, Q, C)
    hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    q4, k4, v4 = hv(q), hv(k), hv(v)
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S) < 0.5
        blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    G = B * H
    wsb = h.emu_vmf_tc_workspace_bytes(G, Q, S, hd)
    ws = torch.zeros(wsb // 4)
    ns_max = wsb // 4 // (Q * (hd + 1)) // G
    part_acc, part_den = ws[:G * ns_max * Q * hd], ws[G * ns_max * Q * hd:]
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    ns = h.emu_vmf_attention_tc_partial(*st(q4), *st(k4), *st(v4), bits.data_ptr() if masked else None,
                                        bits.shape[2] if masked else 0, ro.data_ptr() if masked else None,
                                        B, H, Q, S, hd, kappa, flags, part_acc.data_ptr(), part_den.data_ptr(), 120.0)
    assert ns > 0, (ns, h.emu_last_error())
    acc = part_acc[:G * ns * Q * hd].view(G, ns, Q, hd).sum(1)
    den = part_den[:G * ns * Q].view(G, ns, Q).sum(1)
    out = F.normalize(acc / den.unsqueeze(-1), dim=-1).view(B, H, Q, hd)
    qn = F.normalize(q4.double(), dim=-1) if flags & 1 else q4.double()
    kn = F.normalize(k4.double(), dim=-1) if flags & 2 else k4.double()
    s = kappa * qn @ kn.transpose(-1, -2)
    if eff is not None:
        s = s.masked_fill(eff, float("-inf"))
    ref = F.normalize(torch.softmax(s, -1) @ v4.double(), dim=-1)
    assert (out.double() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,n,m,d,kappa,iters", [
    (1, 128, 100, 64, 10.0, 1),     # exactly one tile
    (1, 100, 7, 64, 10.0, 2),       # fewer keys than a tile, few seeds
    (2, 1000, 100, 64, 10.0, 3),    # 8 tiles per CTA: the 6-stage bulk-copy ring wraps; two images
    (1, 777, 37, 32, 20.0, 3),      # d = 32, key tail, the reference's default kappa
    (1, 2000, 128, 32, 10.0, 2),    # all 128 TMEM lanes in use
])
def test_experimental_packed_mean_shift_kernel(emu_tc, B, n, m, d, kappa, iters):
    """csrc/vmf_attention_packed.cu (operands packed once, streamed by bulk copies, 16 softmax warps) through the
    calibrated emulation, against an fp64 restatement of seed_hill_climbing_ball (mean_shift.py:79-109)."""
    h = emu_tc
    torch.manual_seed(n + d)
    X = F.normalize(torch.randn(B, n, d), dim=-1).contiguous()
    idx = torch.stack([torch.randperm(n)[:m] for _ in range(B)])
    Z = torch.gather(X, 1, idx.unsqueeze(-1).expand(B, m, d)).contiguous()
    packed = _aligned(h.msmx_mean_shift_packed_bytes(B, n, d), 128)
    h.emu_set_timeout(120.0)
    assert h.msmx_mean_shift_pack(X.data_ptr(), packed.data_ptr(), B, n, d, None) == 0
    wsb = h.msmx_mean_shift_packed_workspace_bytes(B, n, m, d)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    out = torch.full_like(Z, float("nan"))
    rc = h.msmx_mean_shift_hill_climb_packed(packed.data_ptr(), Z.data_ptr(), out.data_ptr(), B, n, m, d, kappa, iters,
                                             ws.data_ptr(), wsb, None)
    assert rc == 0, (rc, h.emu_last_error())
    Xd, Zd = X.double(), Z.double()
    for _ in range(iters):
        Zd = F.normalize(torch.exp(kappa * (Zd @ Xd.transpose(-1, -2) - 1.0)) @ Xd, dim=-1, eps=1e-12)
    assert (out.double() - Zd).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,H,Q,S,hd,masked,flags,kappa,tol", [
    (1, 2, 100, 700, 32, True, 3, 30.0, 2e-5),    # the decoder's cross-attention: heads by strides, bit mask, 6 tiles
    (1, 1, 37, 333, 64, False, 3, 30.0, 2e-5),    # hd 64, key tail
    (2, 2, 100, 300, 32, False, 3, 30.0, 2e-5),   # two images x two heads
    (1, 1, 100, 1000, 64, True, 3, 30.0, 2e-5),   # hd 64 masked: the 3-stage ring wraps
    (1, 2, 50, 260, 32, True, 1, 5.0, 1e-4),      # k not normalised: bf16 score operands
])
def test_experimental_packed_attention_kernel(emu_tc, B, H, Q, S, hd, masked, flags, kappa, tol):
    """The general form of csrc/vmf_attention_packed.cu: K (normalised, fp16 halves) and V (bf16 halves) packed per
    (batch, head, tile) by vmf_pack_kernel, attention with the decoder's bit masks - what the cross-attention would
    run once the K/V projection writes the images itself."""
    h = emu_tc
    torch.manual_seed(S + hd + Q)
    C = H * hd
    q, kv = torch.randn(B, Q, C), torch.randn(B, S, 2 * C)
    hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    q4, k4, v4 = hv(q), hv(kv[..., :C]), hv(kv[..., C:])   # k | v share one projection buffer
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S) < 0.5
        blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    h.emu_set_timeout(200.0)
    packed = _aligned(h.msmx_vmf_packed_bytes(B, H, S, hd, flags), 128)
    assert h.msmx_vmf_pack(*st(k4), *st(v4), packed.data_ptr(), B, H, S, hd, flags, None) == 0, h.emu_last_error()
    wsb = h.msmx_vmf_packed_workspace_bytes(B, H, Q, S, hd)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    out = torch.full((B, Q, H, hd), float("nan")).permute(0, 2, 1, 3)
    rc = h.msmx_vmf_attention_packed_fwd(*st(q4), packed.data_ptr(), *st(out), bits.data_ptr() if masked else None,
                                         bits.shape[2] if masked else 0, ro.data_ptr() if masked else None,
                                         B, H, Q, S, hd, kappa, flags, ws.data_ptr(), wsb, None)
    assert rc == 0, (rc, h.emu_last_error())
    qn = F.normalize(q4.double(), dim=-1) if flags & 1 else q4.double()
    kn = F.normalize(k4.double(), dim=-1) if flags & 2 else k4.double()
    s = kappa * qn @ kn.transpose(-1, -2)
    if eff is not None:
        s = s.masked_fill(eff, float("-inf"))
    ref = F.normalize(torch.softmax(s, -1) @ v4.double(), dim=-1)
    assert (out.double() - ref).abs().max().item() < tol


# ------------------------------------------------------------------------------------------------------------------
# csrc/linear_tc.cu: tiled TMA (2-D / 3-D / 4-D boxes, 128B swizzle, clipped stores) + every epilogue mode
@pytest.fixture(scope="module")
def emu_lin():
    from unseenobjectswithmeanshift_b200._lib import SIGNATURES
    h = _build(os.path.join(ROOT, "build", "emu", "libemu_linear_tc.so"), "emu_linear_tc.cpp", ["linear_tc.cu"])
    for n in ("msm_linear_weight_bytes", "msm_linear_prepare_weight", "msm_linear_fwd", "msm_linear_ln_fwd",
              "msm_linear_fused_fwd", "msm_conv1x1_fwd", "msm_conv3x3_fwd"):
        f = getattr(h, n)
        f.restype, f.argtypes = SIGNATURES[n]
    h.emu_set_timeout.argtypes = [ctypes.c_double]
    h.emu_set_sms.argtypes = [ctypes.c_int]
    h.emu_last_error.restype = ctypes.c_char_p
    return h


def _prepare(h, W):
    N, K = W.shape
    p = _aligned(h.msm_linear_weight_bytes(N, K))
    assert h.msm_linear_prepare_weight(W.data_ptr(), W.stride(0), p.data_ptr(), N, K, None) == 0
    return p


def _start(h, sms):
    h.emu_set_timeout(300.0)
    h.emu_set_sms(sms)   # few "SMs": persistent CTAs walk several tiles, rings wrap around


@pytest.mark.parametrize("M,N,K,relu,sms", [(300, 64, 64, True, 2), (800, 256, 256, False, 2), (130, 96, 32, False, 1)])
def test_calibration_shipped_linear_kernel(emu_lin, M, N, K, relu, sms):
    h = emu_lin
    torch.manual_seed(M + N + K)
    X, W, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    Y = torch.full((M, N), float("nan"))
    _start(h, sms)
    p = _prepare(h, W)
    rc = h.msm_linear_fwd(X.data_ptr(), K, p.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, int(relu), None)
    assert rc == 0, (rc, h.emu_last_error())
    ref = X.double() @ W.double().t() + b.double()
    assert (Y.double() - (ref.relu() if relu else ref)).abs().max().item() < 1e-5


def test_calibration_shipped_linear_kernel_fused_epilogues(emu_lin):
    h = emu_lin
    torch.manual_seed(7)
    M, N, K, period = 300, 256, 64, 100
    X, W, b, R = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N), torch.randn(M, N)
    rb, g, be, g2, be2 = torch.randn(period, N), torch.randn(N), torch.randn(N), torch.randn(N), torch.randn(N)
    Y, Y2 = torch.full((M, N), float("nan")), torch.full((M, N), float("nan"))
    _start(h, 2)
    p = _prepare(h, W)
    rc = h.msm_linear_fused_fwd(X.data_ptr(), K, p.data_ptr(), b.data_ptr(), rb.data_ptr(), period, 1, R.data_ptr(), N,
                                g.data_ptr(), be.data_ptr(), 1e-5, 1, g2.data_ptr(), be2.data_ptr(), 1e-5,
                                Y2.data_ptr(), N, Y.data_ptr(), N, M, N, K, None)
    assert rc == 0, (rc, h.emu_last_error())
    v = (X.double() @ W.double().t() + b.double() + rb.double().repeat(3, 1)[:M]).relu() + R.double()
    z = F.normalize(F.layer_norm(v, (N,), g.double(), be.double(), 1e-5), dim=-1)
    assert (Y.double() - z).abs().max().item() < 1e-5
    assert (Y2.double() - F.layer_norm(z, (N,), g2.double(), be2.double(), 1e-5)).abs().max().item() < 5e-5
    # narrow form: residual + LayerNorm over N <= 64 outputs
    N = 64
    W, b, R, g, be = torch.randn(N, K) / K ** 0.5, torch.randn(N), torch.randn(M, N), torch.randn(N), torch.randn(N)
    Y = torch.full((M, N), float("nan"))
    p = _prepare(h, W)
    rc = h.msm_linear_ln_fwd(X.data_ptr(), K, p.data_ptr(), b.data_ptr(), R.data_ptr(), N, g.data_ptr(), be.data_ptr(),
                             1e-5, Y.data_ptr(), N, M, N, K, None)
    assert rc == 0, (rc, h.emu_last_error())
    ref = F.layer_norm(R.double() + X.double() @ W.double().t() + b.double(), (N,), g.double(), be.double(), 1e-5)
    assert (Y.double() - ref).abs().max().item() < 1e-5


def test_calibration_shipped_linear_kernel_convolutions(emu_lin):
    h = emu_lin
    torch.manual_seed(11)
    _start(h, 2)
    for B, K, N, H, W, y_nchw in ((2, 64, 64, 10, 20, True), (2, 32, 96, 9, 16, False)):   # 1x1, NCHW input
        x, w, b = torch.randn(B, K, H, W), torch.randn(N, K) / K ** 0.5, torch.randn(N)
        Y = torch.full((B, N, H * W) if y_nchw else (B, H * W, N), float("nan"))
        p = _prepare(h, w)
        rc = h.msm_conv1x1_fwd(x.data_ptr(), p.data_ptr(), b.data_ptr(), Y.data_ptr(), int(y_nchw), B, H * W, N, K, 0, None)
        assert rc == 0, (rc, h.emu_last_error())
        ref = torch.einsum("nk,bkp->bnp", w.double(), x.double().flatten(2)) + b.double()[None, :, None]
        assert (Y.double() - (ref if y_nchw else ref.transpose(1, 2))).abs().max().item() < 1e-5
    B, C, N, H, W = 1, 32, 32, 9, 40                                                          # 3x3 halo tiles, tails
    x, w, b = torch.randn(B, C, H, W), torch.randn(N, C, 3, 3) / (9 * C) ** 0.5, torch.randn(N)
    Y = torch.full((B, N, H, W), float("nan"))
    p = _prepare(h, w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous())
    Wp = (W + 2 + 3) // 4 * 4
    xp = F.pad(x, (1, Wp - W - 1, 1, 1))
    buf = _aligned(xp.numel() * 4, 128).view(torch.float32).view(xp.shape)
    buf.copy_(xp)
    rc = h.msm_conv3x3_fwd(buf.data_ptr(), p.data_ptr(), b.data_ptr(), Y.data_ptr(), B, C, H, W, Wp, N, 1, None)
    assert rc == 0, (rc, h.emu_last_error())
    assert (Y.double() - F.conv2d(x.double(), w.double(), b.double(), padding=1).relu()).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------------------------------
# the fused chain of DESIGN.md section 8, item 1: K / V projections write operand images, packed attention reads them
@pytest.fixture(scope="module")
def emu_chain():
    from unseenobjectswithmeanshift_b200._lib import SIGNATURES
    h = _build(os.path.join(ROOT, "build", "emu", "libemu_kv_chain.so"), "emu_kv_chain.cpp",
               ["linear_tc.cu", "vmf_attention_packed.cu"])
    P, I, L, Fl, Z, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_double
    for n in ("msm_linear_weight_bytes", "msm_linear_prepare_weight"):
        f = getattr(h, n)
        f.restype, f.argtypes = SIGNATURES[n]
    h.msmx_linear_packed_kv_fwd.restype = I
    h.msmx_linear_packed_kv_fwd.argtypes = [P, L, P, P, P, I, I, I, I, I, I, I, I, P]
    h.msmx_linear_packed_kv_pos_fwd.restype = I
    h.msmx_linear_packed_kv_pos_fwd.argtypes = [P, L, P, P, P, I, I, I, I, I, I, I, I, I, P, P, I, P]
    h.msmx_vmf_packed_bytes.restype, h.msmx_vmf_packed_bytes.argtypes = Z, [I, I, I, I, I]
    h.msmx_vmf_packed_workspace_bytes.restype, h.msmx_vmf_packed_workspace_bytes.argtypes = Z, [I, I, I, I, I]
    h.msmx_vmf_attention_packed_fwd.restype = I
    h.msmx_vmf_attention_packed_fwd.argtypes = [P, L, L, L, P, P, L, L, L, P, I, P, I, I, I, I, I, Fl, I, P, Z, P]
    h.emu_set_timeout.argtypes = [D]
    h.emu_set_sms.argtypes = [I]
    h.emu_last_error.restype = ctypes.c_char_p
    return h


@pytest.mark.parametrize("B,S,Cin,C,layers,Q,masked,sms", [
    (1, 300, 64, 64, 2, 50, True, 2),      # two decoder layers of a level projected by one GEMM, two heads, bit masks
    (2, 200, 32, 256, 1, 100, False, 2),   # eight heads; a 128-row GEMM tile straddles the two images
    (1, 130, 64, 32, 3, 20, True, 1),      # three layers, key tail of two keys in the second tile
])
def test_experimental_projection_to_packed_attention_chain(emu_chain, B, S, Cin, C, layers, Q, masked, sms):
    """msmx_linear_packed_kv_fwd (linear_tc_kernel<PACK>: normalise + split in the epilogue, operand images instead of
    fp32 rows) for K = (src + pos) Wk^T and V = src Wv^T, then msmx_vmf_attention_packed_fwd per layer, against the
    fp64 cross-attention of the reference (attention_util.py:64-82 on the projected rows)."""
    h = emu_chain
    torch.manual_seed(S + C + layers)
    H = C // 32
    src, pos = torch.randn(B, S, Cin), torch.randn(B, S, Cin)
    wk, bk = torch.randn(layers * C, Cin) / Cin ** 0.5, torch.randn(layers * C) * 0.1
    wv, bv = torch.randn(layers * C, Cin) / Cin ** 0.5, torch.randn(layers * C) * 0.1
    q = torch.randn(layers, B, Q, C)
    h.emu_set_timeout(600.0)
    h.emu_set_sms(sms)
    per_layer = h.msmx_vmf_packed_bytes(B, H, S, 32, 3)
    packed = _aligned(layers * per_layer, 128)             # zero-initialised: key tails stay zero
    key_in = (src + pos).contiguous()
    pk, pv = _prepare(h, wk), _prepare(h, wv)
    rc = h.msmx_linear_packed_kv_fwd(key_in.data_ptr(), Cin, pk.data_ptr(), bk.data_ptr(), packed.data_ptr(), B, S,
                                     layers * C, Cin, C, 0, 1, 1, None)
    assert rc == 0, h.emu_last_error()
    rc = h.msmx_linear_packed_kv_fwd(src.data_ptr(), Cin, pv.data_ptr(), bv.data_ptr(), packed.data_ptr(), B, S,
                                     layers * C, Cin, C, 1, 0, 0, None)
    assert rc == 0, h.emu_last_error()
    K = (key_in.double() @ wk.double().t() + bk.double()).view(B, S, layers, H, 32)
    V = (src.double() @ wv.double().t() + bv.double()).view(B, S, layers, H, 32)
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    for j in range(layers):
        bits = ro = eff = None
        if masked:
            blocked = torch.rand(B, Q, S) < 0.5
            blocked[:, 3] = True
            ro = (~blocked).any(-1).to(torch.int32).contiguous()
            bits = _pack_bits(blocked)
            eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
        q4 = q[j].unflatten(-1, (H, 32)).permute(0, 2, 1, 3)
        wsb = h.msmx_vmf_packed_workspace_bytes(B, H, Q, S, 32)
        ws = torch.zeros(wsb, dtype=torch.uint8)
        out = torch.full((B, Q, H, 32), float("nan")).permute(0, 2, 1, 3)
        rc = h.msmx_vmf_attention_packed_fwd(*st(q4), packed.data_ptr() + j * per_layer, *st(out),
                                             bits.data_ptr() if masked else None, bits.shape[2] if masked else 0,
                                             ro.data_ptr() if masked else None, B, H, Q, S, 32, 30.0, 3,
                                             ws.data_ptr(), wsb, None)
        assert rc == 0, (rc, h.emu_last_error())
        kj, vj = K[:, :, j].permute(0, 2, 1, 3), V[:, :, j].permute(0, 2, 1, 3)
        s = 30.0 * F.normalize(q4.double(), dim=-1) @ F.normalize(kj, dim=-1).transpose(-1, -2)
        if eff is not None:
            s = s.masked_fill(eff, float("-inf"))
        ref = F.normalize(torch.softmax(s, -1) @ vj, dim=-1)
        err = (out.double() - ref).abs().max().item()
        assert err == err and err < 3e-5, (j, err)


# ------------------------------------------------------------------------------------------------------------------
# the remaining shipped tcgen05 kernels: mask einsum and fused feed-forward block
@pytest.fixture(scope="module")
def emu_gemm():
    from unseenobjectswithmeanshift_b200._lib import SIGNATURES
    h = _build(os.path.join(ROOT, "build", "emu", "libemu_gemm_tc.so"), "emu_gemm_tc.cpp",
               ["linear_tc.cu", "mask_head_tc.cu", "ffn_tc.cu"])
    for n in ("msm_linear_weight_bytes", "msm_linear_prepare_weight", "msm_ffn_ln_fwd"):
        f = getattr(h, n)
        f.restype, f.argtypes = SIGNATURES[n]
    P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    h.emu_mask_logits_tc.restype, h.emu_mask_logits_tc.argtypes = I, [P, P, P, I, I, I, L]
    h.emu_set_timeout.argtypes = [ctypes.c_double]
    h.emu_set_sms.argtypes = [I]
    h.emu_last_error.restype = ctypes.c_char_p
    return h


def _f32_aligned(t):
    b = _aligned(t.numel() * 4, 128).view(torch.float32).view(t.shape)
    b.copy_(t)
    return b


@pytest.mark.parametrize("B,Q,C,H,W,sms", [(2, 100, 256, 12, 16, 2), (1, 10, 32, 24, 32, 1), (1, 100, 64, 30, 44, 2)])
def test_calibration_shipped_mask_head_kernel(emu_gemm, B, Q, C, H, W, sms):
    h = emu_gemm
    torch.manual_seed(Q + C + H)
    e, f = _f32_aligned(torch.randn(B, Q, C)), _f32_aligned(torch.randn(B, C, H, W))
    out = torch.full((B, Q, H, W), float("nan"))
    _start(h, sms)
    rc = h.emu_mask_logits_tc(e.data_ptr(), f.data_ptr(), out.data_ptr(), B, Q, C, H * W)
    assert rc == 0, (rc, h.emu_last_error())
    ref = torch.einsum("bqc,bchw->bqhw", e.double(), f.double())
    assert (out.double() - ref).abs().max().item() < 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize("M,D,Fh,sms", [(300, 64, 256, 2), (520, 32, 128, 1)])
def test_calibration_shipped_ffn_kernel(emu_gemm, M, D, Fh, sms):
    h = emu_gemm
    torch.manual_seed(M + D + Fh)
    x = torch.randn(M, D)
    w1, b1 = torch.randn(Fh, D) / D ** 0.5, torch.randn(Fh) * 0.1
    w2, b2 = torch.randn(D, Fh) / Fh ** 0.5, torch.randn(D) * 0.1
    g, be = torch.randn(D), torch.randn(D)
    y = torch.full((M, D), float("nan"))
    _start(h, sms)
    p1, p2 = _prepare(h, w1), _prepare(h, w2)
    rc = h.msm_ffn_ln_fwd(x.data_ptr(), D, p1.data_ptr(), b1.data_ptr(), p2.data_ptr(), b2.data_ptr(), g.data_ptr(),
                          be.data_ptr(), 1e-5, y.data_ptr(), D, M, D, Fh, None)
    assert rc == 0, (rc, h.emu_last_error())
    xd = x.double()
    ref = F.layer_norm(xd + F.linear(F.relu(F.linear(xd, w1.double(), b1.double())), w2.double(), b2.double()), (D,),
                       g.double(), be.double(), 1e-5)
    assert (y.double() - ref).abs().max().item() < 2e-5


# ------------------------------------------------------------------------------------------------------------------
# late-execution mode: the same kernels with every asynchronous operation delayed as far as their barriers allow
def test_late_execution_mode_linear_and_chain(emu_lin, emu_chain):
    for h in (emu_lin, emu_chain):
        h.emu_set_late.argtypes = [ctypes.c_int]
        h.emu_deferred_ops.restype = ctypes.c_long
        h.emu_set_late(1)
    try:
        before = emu_lin.emu_deferred_ops()
        test_calibration_shipped_linear_kernel(emu_lin, 300, 64, 64, True, 2)
        test_calibration_shipped_linear_kernel_fused_epilogues(emu_lin)
        assert emu_lin.emu_deferred_ops() > before       # the queues were really in use
        test_experimental_projection_to_packed_attention_chain(emu_chain, 1, 300, 64, 64, 2, 50, True, 2)
    finally:
        for h in (emu_lin, emu_chain):
            h.emu_set_late(0)


def test_late_execution_mode_attention_kernels(emu_tc, emu_gemm):
    for h in (emu_tc, emu_gemm):
        h.emu_set_late.argtypes = [ctypes.c_int]
        h.emu_deferred_ops.restype = ctypes.c_long
        h.emu_set_late(1)
    try:
        before = emu_tc.emu_deferred_ops()
        test_calibration_shipped_tcgen05_attention_kernel(emu_tc, 1, 2, 100, 700, 32, False, True)
        test_experimental_packed_mean_shift_kernel(emu_tc, 2, 1000, 100, 64, 10.0, 3)
        test_experimental_packed_attention_kernel(emu_tc, 1, 1, 100, 1000, 64, True, 3, 30.0, 2e-5)
        assert emu_tc.emu_deferred_ops() > before
        test_calibration_shipped_mask_head_kernel(emu_gemm, 1, 100, 64, 30, 44, 2)
        test_calibration_shipped_ffn_kernel(emu_gemm, 300, 64, 256, 2)
    finally:
        for h in (emu_tc, emu_gemm):
            h.emu_set_late(0)


def test_late_mode_detects_a_missing_barrier_wait():
    """Self-test: a consumer that does not wait for its bulk copy is indistinguishable from a correct one when
    asynchronous operations execute at issue, and reads stale shared memory in late mode."""
    lib = os.path.join(ROOT, "build", "emu", "libemu_hazard.so")
    src = os.path.join(EMU_DIR, "emu_hazard.cpp")
    deps = [src, os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "tc_emu.h")]
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-w", "-shared", "-fPIC", "-pthread", "-DMSM_EMULATE_ON_HOST",
                               "-I" + CUDA_INC, "-I" + EMU_DIR, "-x", "c++", src, "-o", lib])
    h = ctypes.CDLL(lib)
    h.emu_hazard_copy.restype = ctypes.c_float
    h.emu_hazard_copy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    data = torch.arange(256, dtype=torch.float32)
    want = float(data.sum())
    assert h.emu_hazard_copy(data.data_ptr(), 1, 0) == want     # correct kernel, synchronous
    assert h.emu_hazard_copy(data.data_ptr(), 1, 1) == want     # correct kernel, late
    assert h.emu_hazard_copy(data.data_ptr(), 0, 1) == -256.0   # missing wait: late mode exposes the stale tile


# ------------------------------------------------------------------------------------------------------------------
# decoder-level wiring of the opt-in packed K / V path, with the REAL kernels running under the emulation
def test_decoder_packed_kv_wiring_with_emulated_kernels(emu_chain, monkeypatch):
    """_MeanShiftDecoderBase with MSM_PACKED_KV=1 on CPU: the K / V projections and the cross-attention go through the
    emulated linear_tc_kernel<PACK> / vmf_attn_packed_kernel (same C entry points as ops.py binds), every other op
    through its contract-level stand-in; the outputs must equal the default path's. Checks the host side of the
    opt-in path: per-level image buffers, layer slicing, key counts, caching across calls."""
    import fake_ops
    from unseenobjectswithmeanshift_b200 import ops
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder import (
        meanshiftformer_transformer_decoder as dec)
    h = emu_chain
    for name in ("vmf_attention", "mask_logits", "mask_to_attn_bits", "dense"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    monkeypatch.setattr(ops, "tc_linear_enabled", lambda: False)      # the un-fused layer sequence: fewer stand-ins

    calls = {"alloc": 0, "project": 0, "attend": 0}

    def packed_kv_alloc(layers, batch, heads, num_keys, device):
        calls["alloc"] += 1
        per_layer = h.msmx_vmf_packed_bytes(batch, heads, num_keys, 32, 3)
        return _aligned(layers * per_layer, 128), per_layer

    def linear_packed_kv(x, weight, bias, images, batch, num_keys, channels, which, pos=None):
        N, K = weight.shape
        calls["project"] += 1
        calls["folded"] = calls.get("folded", 0) + (1 if pos is not None or x.dim() == 4 else 0)
        _start(h, 2)
        w = weight.detach().contiguous()
        p = _prepare(h, w)
        b = bias.detach().contiguous()
        if pos is None and x.dim() == 3:
            rc = h.msmx_linear_packed_kv_fwd(x.data_ptr(), K, p.data_ptr(), b.data_ptr(), images.data_ptr(), batch,
                                             num_keys, N, K, channels, int(which), 1, 1, None)
        else:   # input_proj folded in: the channel-major map (when S % 4 == 0) and the separable positional tables
            ty, tx = pos if pos is not None else (None, None)
            rc = h.msmx_linear_packed_kv_pos_fwd(x.data_ptr(), K, p.data_ptr(), b.data_ptr(), images.data_ptr(), batch,
                                                 num_keys, N, K, channels, int(which), 1, 1, 1 if x.dim() == 4 else 0,
                                                 ty.data_ptr() if ty is not None else None,
                                                 tx.data_ptr() if tx is not None else None,
                                                 tx.shape[0] if tx is not None else 1, None)
        assert rc == 0, h.emu_last_error()

    def vmf_attention_packed(q, kv, *, blocked_bits=None, row_open=None, kappa=30.0, out=None):
        B, H, Nq, hd = q.shape
        calls["attend"] += 1
        st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
        wsb = h.msmx_vmf_packed_workspace_bytes(B, H, Nq, kv.num_keys, hd)
        ws = torch.zeros(wsb, dtype=torch.uint8)
        _start(h, 2)
        rc = h.msmx_vmf_attention_packed_fwd(*st(q), kv.images.data_ptr(), *st(out),
                                             blocked_bits.data_ptr() if blocked_bits is not None else None,
                                             blocked_bits.shape[2] if blocked_bits is not None else 0,
                                             row_open.data_ptr() if row_open is not None else None,
                                             B, H, Nq, kv.num_keys, hd, float(kappa), 3, ws.data_ptr(), wsb, None)
        assert rc == 0, h.emu_last_error()
        return out

    monkeypatch.setattr(ops, "packed_kv_alloc", packed_kv_alloc)
    monkeypatch.setattr(ops, "linear_packed_kv", linear_packed_kv)
    monkeypatch.setattr(ops, "vmf_attention_packed", vmf_attention_packed)

    torch.manual_seed(3)
    kw = dict(num_classes=2, hidden_dim=64, num_queries=12, nheads=2, dim_feedforward=64, dec_layers=4,
              pre_norm=False, mask_dim=32, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)
    m = dec.MeanShiftTransformerDecoder(32, True, **kw).eval()
    x = [torch.randn(2, 32, hh, ww) for hh, ww in ((3, 5), (6, 10), (12, 20))]      # 15 / 60 / 240 keys: tails everywhere
    mf = torch.randn(2, 32, 24, 40)
    with torch.no_grad():
        monkeypatch.setenv("MSM_PACKED_KV", "0")
        want = m(x, mf)
        monkeypatch.setenv("MSM_PACKED_KV", "1")
        got = m(x, mf)
        again = m(x, mf)   # second call reuses the cached image buffers
    # two forwards: 3 levels x (K, V) projections each, 4 cross-attentions each; a fresh image buffer per level and
    # forward (graphs in flight must not share one)
    assert calls == {"alloc": 6, "project": 12, "attend": 8, "folded": 10}, calls   # the 15-key level: token-major V, no table
    for o in (got, again):
        assert (o["pred_masks"] - want["pred_masks"]).abs().max().item() < 1e-3 * want["pred_masks"].abs().max().item()
        assert (o["pred_logits"] - want["pred_logits"]).abs().max().item() < 1e-3


def test_decoder_training_gradients_through_the_emulated_backward_kernel(emu, monkeypatch, golden):
    """The decoder's training path (tests/test_training_wiring.py) with the REAL attention backward kernel: every
    VmfAttentionFunction.backward of the 4-layer decoder (cross- and self-attention, masks, the decoder's strided head
    views) runs csrc/vmf_attention_bwd.cu under the emulation; the gradients of the probe loss must equal
    torch.autograd through the REFERENCE decoder (golden)."""
    import fake_ops
    from scenes import probe_loss
    from unseenobjectswithmeanshift_b200 import ops
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder import (
        meanshiftformer_transformer_decoder as dec)
    for name in ("vmf_attention", "mask_logits", "mask_to_attn_bits", "dense"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    count = {"bwd": 0}

    def vmf_attention_bwd(q, k, v, out, grad_out, den, *, blocked_bits=None, row_open=None, add_mask=None, kappa=30.0,
                          normalize_q=True, normalize_k=True):
        count["bwd"] += 1
        if grad_out.stride(3) != 1:
            grad_out = grad_out.contiguous()
        return _bwd(emu, q, k, v, out, grad_out, den.contiguous(), bits=blocked_bits, row_open=row_open, add_mask=add_mask,
                    kappa=kappa, flags=(1 if normalize_q else 0) | (2 if normalize_k else 0))

    monkeypatch.setattr(ops, "vmf_attention_bwd", vmf_attention_bwd)
    g, sd = golden("decoder_multiscale")
    want, _ = golden("decoder_multiscale_bwd")
    kw = dict(num_classes=2, hidden_dim=32, num_queries=10, nheads=2, dim_feedforward=64, dec_layers=4,
              pre_norm=False, mask_dim=32, enforce_input_project=False, use_meanshift_cross_attention=True,
              disable_attention_mask=False, use_meanshift_self_attention=True, decoder_block_norm=True)
    m = dec.MeanShiftTransformerDecoder(int(g["in_channels"]), True, **kw)
    m.load_state_dict(sd, strict=True)
    m.train()
    x = [g[f"x{i}"].clone().requires_grad_() for i in range(3)]
    mf = g["mask_features"].clone().requires_grad_()
    loss = probe_loss(m(x, mf))
    params = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [p for _, p in params] + x + [mf], allow_unused=True)
    assert count["bwd"] == 8          # four cross- and four self-attention calls
    seen = 0
    for name, gr in zip([n for n, _ in params] + ["x0", "x1", "x2", "mask_features"], grads):
        key = "grad::" + name
        if gr is None:
            assert key not in want, name
            continue
        _close(gr, want[key], 5e-4)
        seen += 1
    assert seen == sum(k.startswith("grad::") for k in want)


# ------------------------------------------------------------------------------------------------------------------
# csrc/vmf_attention_small.cu: single-launch attention for short key sequences (opt-in MSM_SMALL_ATTN=1)
@pytest.fixture(scope="module")
def emu_small():
    h = _build(os.path.join(ROOT, "build", "emu", "libemu_vmf_small.so"), "emu_vmf_small.cpp", ["vmf_attention_small.cu"])
    P, I, L, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    h.emu_vmf_small_supported.restype, h.emu_vmf_small_supported.argtypes = I, [I, I, I]
    h.emu_vmf_attention_small.restype = I
    h.emu_vmf_attention_small.argtypes = [P, L, L, L] * 4 + [P, P, I, P, I, I, I, I, Fl, I]
    return h


@pytest.mark.parametrize("B,H,Q,S,masked,flags,selfattn", [
    (2, 2, 100, 100, False, 3, True),     # the decoder's self-attention: q | k | v slices of one fused projection
    (1, 8, 100, 300, True, 3, False),     # coarsest cross-attention level of the R50 config, bit masks
    (1, 1, 128, 70, True, 3, False),      # all 128 rows, key tail inside the second tile
    (1, 2, 37, 1000, False, 1, False),    # k not normalised, 16 tiles
])
def test_emulated_small_attention_kernel(emu, emu_small, B, H, Q, S, masked, flags, selfattn):
    """vmf_small_kernel against the fp64 reference; with MSM_VMF_SAVE_NORM its (den, |o|) planes then drive the
    emulated BACKWARD kernel: a forward -> backward chain of real kernel sources against fp64 autograd."""
    h = emu_small
    assert h.emu_vmf_small_supported(Q, S, 32) == 1 and h.emu_vmf_small_supported(Q, 5000, 32) == 0
    torch.manual_seed(S + Q)
    hd, C = 32, H * 32
    hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    if selfattn:
        qkv = torch.randn(B, S, 3 * C)
        q4, k4, v4 = hv(qkv[..., :C]), hv(qkv[..., C:2 * C]), hv(qkv[..., 2 * C:])
    else:
        qb, kvb = torch.randn(B, Q, C), torch.randn(B, S, 2 * C)
        if not flags & 2:   # without the kernel's normalisation the rows have to be unit vectors already (the fixed
            kvb = F.normalize(kvb.view(B, S, 2 * H, hd), dim=-1).reshape(B, S, 2 * C)   # shift assumes |cos| <= 1)
        q4, k4, v4 = hv(qb), hv(kvb[..., :C]), hv(kvb[..., C:])
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S) < 0.5
        blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    out = torch.full((B, Q, H, hd), float("nan")).permute(0, 2, 1, 3)
    den = torch.full((2, B * H, Q), float("nan"))
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    rc = h.emu_vmf_attention_small(*st(q4), *st(k4), *st(v4), *st(out), den.data_ptr(),
                                   bits.data_ptr() if masked else None, bits.shape[2] if masked else 0,
                                   ro.data_ptr() if masked else None, B, H, Q, S, 30.0, flags | 4)
    assert rc == 0
    q2, k2, v2 = (t.double().clone().requires_grad_() for t in (q4, k4, v4))
    qn = F.normalize(q2, dim=-1) if flags & 1 else q2
    kn = F.normalize(k2, dim=-1) if flags & 2 else k2
    s = 30.0 * qn @ kn.transpose(-1, -2)
    if eff is not None:
        s = s.masked_fill(eff, float("-inf"))
    w = torch.exp(s - 30.0)
    o = (w @ v2) / w.sum(-1, keepdim=True)
    ref = F.normalize(o, dim=-1)
    assert (out.double() - ref.detach()).abs().max().item() < 2e-6
    rel = lambda a, b_: ((a.double() - b_).abs() / b_.abs().clamp_min(1e-30)).max().item()
    assert rel(den[0].view(B, H, Q), w.sum(-1).detach()) < 1e-5      # softmax denominators with the fixed shift
    assert rel(den[1].view(B, H, Q), o.norm(dim=-1).detach()) < 1e-5  # |softmax . v|
    # backward kernel on the planes the forward kernel wrote
    gout = torch.randn(B, H, Q, hd)
    gq, gk, gv = _bwd(emu, q4, k4, v4, out, gout, den, bits=bits, row_open=ro, kappa=30.0, flags=flags)
    rq, rk, rv = torch.autograd.grad(ref, (q2, k2, v2), gout.double())
    _close(gq, rq, 1e-4)
    _close(gk, rk, 1e-4)
    _close(gv, rv, 1e-4)


def test_regression_layernorm_epilogue_is_never_split_over_column_chunks(emu_lin):
    """Found by a shape sweep under this emulation: with N = 64 and fewer row tiles than half the SMs, pick_bn narrowed
    the column chunk to 32 and the fused residual + LayerNorm normalised the two halves of a row separately (on the
    B200: M <= 9472 rows, e.g. ONE image through the R50 pixel decoder; the GPU test only had M = 12600)."""
    h = emu_lin
    torch.manual_seed(5)
    M, N, K = 100, 64, 64
    X, W, b, R = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N), torch.randn(M, N)
    g, be = torch.randn(N), torch.randn(N)
    Y = torch.full((M, N), float("nan"))
    _start(h, 148)   # as many "SMs" as the B200 has: one row tile leaves room for narrower chunks
    p = _prepare(h, W)
    rc = h.msm_linear_ln_fwd(X.data_ptr(), K, p.data_ptr(), b.data_ptr(), R.data_ptr(), N, g.data_ptr(), be.data_ptr(),
                             1e-5, Y.data_ptr(), N, M, N, K, None)
    assert rc == 0, h.emu_last_error()
    ref = F.layer_norm(R.double() + X.double() @ W.double().t() + b.double(), (N,), g.double(), be.double(), 1e-5)
    assert (Y.double() - ref).abs().max().item() < 1e-5
