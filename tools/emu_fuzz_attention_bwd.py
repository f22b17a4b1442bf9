"""Shape sweep of the attention backward kernel (csrc/vmf_attention_bwd.cu) under the CPU emulation (tests/emu): random
batch / heads / query and key counts / head dims / masks / normalisation flags / SM counts, gradients against fp64
autograd of the reference expression.

    python -m pytest tests/test_kernel_emulation.py -q     # builds build/emu/*.so
    python tools/emu_fuzz_attention_bwd.py [seed] [seconds]
"""
import ctypes, sys, time, random, torch, torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from unseenobjectswithmeanshift_b200._lib import SIGNATURES
h = ctypes.CDLL("/root/repo/build/emu/libemu_vmf_bwd.so")
for name in ("msm_vmf_attention_bwd", "msm_vmf_attention_bwd_workspace_bytes"):
    fn = getattr(h, name); fn.restype, fn.argtypes = SIGNATURES[name]
h.emu_set_sms.argtypes = [ctypes.c_int]
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + float(sys.argv[2] if len(sys.argv) > 2 else 300)
n, worst = 0, 0.0
while time.time() < t_end:
    B, H, Q, hd = rng.randint(1, 2), rng.randint(1, 3), rng.randint(1, 128), rng.choice([4, 8, 16, 32, 64, 20])
    S = rng.choice([rng.randint(1, 70), rng.randint(60, 400)])
    masked, flags = rng.random() < 0.5, rng.choice([3, 3, 3, 0])
    sms = rng.choice([1, 4, 148]); h.emu_set_sms(sms)
    torch.manual_seed(rng.randrange(1 << 30))
    C = H * hd
    q, kv = torch.randn(B, Q, C), torch.randn(B, S, 2 * C)
    if flags == 0:
        kv = F.normalize(kv.view(B, S, 2 * H, hd), dim=-1).reshape(B, S, 2 * C); q = F.normalize(q.view(B, Q, H, hd), dim=-1).reshape(B, Q, C)
    hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    q4, k4, v4 = hv(q), hv(kv[..., :C]), hv(kv[..., C:])
    kappa = 30.0 if flags else 10.0
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S) < rng.choice([0.2, 0.5, 0.9])
        if Q > 3: blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        words = (S + 31) // 32
        pad = torch.zeros(B, Q, words * 32, dtype=torch.bool); pad[..., :S] = blocked
        val = (pad.view(B, Q, words, 32).long() << torch.arange(32)).sum(-1)
        bits = torch.where(val >= 2 ** 31, val - 2 ** 32, val).to(torch.int32).contiguous()
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    q2, k2, v2 = (t.double().clone().requires_grad_() for t in (q4, k4, v4))
    qn = F.normalize(q2, dim=-1) if flags & 1 else q2; kn = F.normalize(k2, dim=-1) if flags & 2 else k2
    s = kappa * qn @ kn.transpose(-1, -2)
    if eff is not None: s = s.masked_fill(eff, float("-inf"))
    w = torch.exp(s - kappa); o = (w @ v2) / w.sum(-1, keepdim=True); ref = F.normalize(o, dim=-1)
    out = torch.empty(B, Q, H, hd).permute(0, 2, 1, 3); out.copy_(ref.detach())
    den = torch.stack([w.sum(-1).detach().reshape(B * H, Q), o.norm(dim=-1).detach().reshape(B * H, Q)]).float().contiguous()
    gout = torch.randn(B, H, Q, hd)
    grad = lambda L_: torch.full((B, L_, H, hd), float("nan")).permute(0, 2, 1, 3)
    gq, gk, gv = grad(Q), grad(S), grad(S)
    wsb = h.msm_vmf_attention_bwd_workspace_bytes(B, H, Q, S, hd); ws = torch.empty(max(wsb, 4), dtype=torch.uint8)
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    rc = h.msm_vmf_attention_bwd(*st(q4), *st(k4), *st(v4), *st(out), *st(gout), den.data_ptr(), *st(gq), *st(gk), *st(gv),
                                 bits.data_ptr() if masked else None, bits.shape[2] if masked else 0, ro.data_ptr() if masked else None, None,
                                 B, H, Q, S, hd, kappa, flags, ws.data_ptr(), wsb, None)
    desc = f"B{B} H{H} Q{Q} S{S} hd{hd} mask{int(masked)} flags{flags} sms{sms}"
    if rc != 0:
        print("FAIL rc", rc, desc, flush=True); continue
    rq, rk, rv = torch.autograd.grad(ref, (q2, k2, v2), gout.double())
    # relative to the peak gradient, with a floor: a single key or a saturated softmax has analytically zero gradients
    # wrt q and k, where only the absolute rounding error (~1e-6) is meaningful
    err = max(((a.double() - b_).abs().max() / b_.abs().max().clamp_min(1e-2)).item() for a, b_ in ((gq, rq), (gk, rk), (gv, rv)))
    n += 1; worst = max(worst, err if err == err else 9e9)
    if not (err == err) or err > 2e-4: print("BAD ", desc, f"rel err {err:.3e}", flush=True)
print("cases", n, f"worst relative-to-peak error {worst:.1e}")
