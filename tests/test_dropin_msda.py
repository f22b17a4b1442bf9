"""The reference's ONE native interface - the pybind module ``MultiScaleDeformableAttention``
(pixel_decoder/ops/src/vision.cpp:18-21, imported by ops/functions/ms_deform_attn_func.py:21-22) - shipped as a
drop-in python module over the C ABI (unseenobjectswithmeanshift_b200/dropin/MultiScaleDeformableAttention.py).

CPU: the module imports under the pybind module's name, exports the two functions with the reference's argument lists,
and the REFERENCE's own ``MSDeformAttnFunction`` binds to it (when a copy of the reference is reachable).
GPU (-m gpu): forward / backward through the module with the reference's argument list (int64 shape tensors) against
the oracle; and the reference's autograd Function running on top of it, against the reference's pure-PyTorch core.
"""
import importlib
import inspect
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "unseenobjectswithmeanshift_b200", "dropin")


def _import_dropin():
    sys.modules.pop("MultiScaleDeformableAttention", None)   # ref_shim may have installed its empty stand-in
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    return importlib.import_module("MultiScaleDeformableAttention")


def test_dropin_module_exports_reference_signatures():
    m = _import_dropin()
    fwd = inspect.signature(m.ms_deform_attn_forward)
    bwd = inspect.signature(m.ms_deform_attn_backward)
    # ops/src/ms_deform_attn.h:25-31 and :47-54
    assert list(fwd.parameters) == ["value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight",
                                    "im2col_step"]
    assert list(bwd.parameters) == ["value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight",
                                    "grad_output", "im2col_step"]


def test_dropin_module_is_loud_without_cuda():
    m = _import_dropin()
    v = torch.zeros(1, 4, 1, 4)
    shapes = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    w = torch.ones(1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):   # the reference: AT_ERROR("Not implemented on the CPU"), ms_deform_attn.h:43
        m.ms_deform_attn_forward(v, shapes, lsi, loc, w, 128)


def _reference_function():
    """The reference's ops/functions/ms_deform_attn_func.py imported with the drop-in on the path, or None."""
    from oracle import ref_models
    root = ref_models.reference_root()
    if root is None:
        return None
    msda = _import_dropin()
    path = os.path.join(root, "MSMFormer", "meanshiftformer", "modeling", "pixel_decoder", "ops", "functions",
                        "ms_deform_attn_func.py")
    spec = importlib.util.spec_from_file_location("ref_ms_deform_attn_func_dropin", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)     # executes `import MultiScaleDeformableAttention as MSDA` (:21-22)
    assert mod.MSDA is msda
    return mod


def test_reference_function_binds_to_dropin():
    mod = _reference_function()
    if mod is None:
        pytest.skip("no copy of the reference reachable (baseline/_ref or /root/reference)")
    assert hasattr(mod.MSDA, "ms_deform_attn_forward") and hasattr(mod.MSDA, "ms_deform_attn_backward")
    assert hasattr(mod, "MSDeformAttnFunction") and hasattr(mod, "ms_deform_attn_core_pytorch")


def _case(N=2, M=8, D=8, Lq=37, shapes=((15, 20), (30, 40), (60, 80)), P=4, dtype=torch.float32, seed=3):
    g = torch.Generator().manual_seed(seed)
    sh = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    S = int(sh.prod(1).sum())
    L = len(shapes)
    value = (torch.rand(N, S, M, D, generator=g) * 0.01).to(dtype)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g).to(dtype)
    w = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    w = (w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dtype)
    return value, sh, lsi, loc, w


@pytest.mark.gpu
def test_dropin_forward_backward_vs_oracle():
    from oracle import pixel_decoder as opd
    m = _import_dropin()
    value, sh, lsi, loc, w = _case()
    v, l_, w_ = value.cuda(), loc.cuda(), w.cuda()
    out = m.ms_deform_attn_forward(v, sh.cuda(), lsi.cuda(), l_, w_, 128)
    want = opd.ms_deform_attn_core(value.double(), sh, lsi, loc.double(), w.double())
    assert out.shape == (2, 37, 64)
    assert (out.cpu().double() - want).abs().max().item() < 1e-6
    go = torch.randn(2, 37, 64, generator=torch.Generator().manual_seed(1))
    vd, ld, wd = (t.double().requires_grad_(True) for t in (value, loc, w))
    opd.ms_deform_attn_core(vd, sh, lsi, ld, wd).backward(go.double())
    gv, gl, ga = m.ms_deform_attn_backward(v, sh.cuda(), lsi.cuda(), l_, w_, go.cuda(), 128)
    for got, ref in ((gv, vd.grad), (gl, ld.grad), (ga, wd.grad)):
        scale = max(ref.abs().max().item(), 1e-12)
        assert (got.cpu().double() - ref).abs().max().item() / scale < 1e-4
    with pytest.raises(RuntimeError):   # ms_deform_attn_cuda.cu:57: batch % min(batch, im2col_step) == 0
        m.ms_deform_attn_forward(torch.cat([v, v[:1]]), sh.cuda(), lsi.cuda(), torch.cat([l_, l_[:1]]),
                                 torch.cat([w_, w_[:1]]), 2)


@pytest.mark.gpu
def test_reference_autograd_function_on_dropin():
    """The REFERENCE's MSDeformAttnFunction (forward + backward) with the drop-in as its MSDA module, against the
    reference's own ms_deform_attn_core_pytorch - the check of ops/test.py:35-63, at the UOIS geometry."""
    mod = _reference_function()
    if mod is None:
        pytest.skip("no copy of the reference reachable (baseline/_ref or /root/reference)")
    value, sh, lsi, loc, w = _case(seed=5)
    v, l_, w_ = (t.cuda().requires_grad_(True) for t in (value, loc, w))
    out = mod.MSDeformAttnFunction.apply(v, sh.cuda(), lsi.cuda(), l_, w_, 128)
    vr, lr, wr = (t.cuda().double().requires_grad_(True) for t in (value, loc, w))
    ref = mod.ms_deform_attn_core_pytorch(vr, sh.cuda(), lr, wr)
    assert torch.allclose(out.double(), ref, rtol=1e-2, atol=1e-3)   # the reference's own fp32 tolerance (test.py:61)
    assert (out.double() - ref).abs().max().item() < 1e-6
    go = torch.randn_like(out)
    out.backward(go)
    ref.backward(go.double())
    for got, want in ((v.grad, vr.grad), (l_.grad, lr.grad), (w_.grad, wr.grad)):
        scale = max(want.abs().max().item(), 1e-12)
        assert (got.double() - want).abs().max().item() / scale < 1e-4
