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
