"""Shape sweep of the attention kernels (shipped vmf_attn_tc_kernel and the packed-operand kernel) under the CPU
emulation (tests/emu): random batch / heads / query and key counts / head dims / masks / execution order, each
against the fp64 reference.

    python -m pytest tests/test_kernel_emulation.py -q     # builds build/emu/*.so
    python tools/emu_fuzz_attention.py [seed] [seconds]
"""
import ctypes, sys, time, random, torch, torch.nn.functional as F
sys.path.insert(0, "/root/repo")
h = ctypes.CDLL("/root/repo/build/emu/libemu_vmf_tc.so")
P, I, L, Fl, Z, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_double
h.emu_vmf_tc_workspace_bytes.restype = Z; h.emu_vmf_tc_workspace_bytes.argtypes = [I, I, I, I]
h.emu_vmf_attention_tc_partial.restype = I
h.emu_vmf_attention_tc_partial.argtypes = [P, L, L, L] * 3 + [P, I, P, I, I, I, I, I, Fl, I, P, P, D]
h.msmx_vmf_packed_bytes.restype = Z; h.msmx_vmf_packed_bytes.argtypes = [I, I, I, I, I]
h.msmx_vmf_packed_workspace_bytes.restype = Z; h.msmx_vmf_packed_workspace_bytes.argtypes = [I, I, I, I, I]
h.msmx_vmf_pack.restype = I; h.msmx_vmf_pack.argtypes = [P, L, L, L, P, L, L, L, P, I, I, I, I, I, P]
h.msmx_vmf_attention_packed_fwd.restype = I
h.msmx_vmf_attention_packed_fwd.argtypes = [P, L, L, L, P, P, L, L, L, P, I, P, I, I, I, I, I, Fl, I, P, Z, P]
h.emu_set_timeout.argtypes = [D]; h.emu_set_late.argtypes = [I]; h.emu_last_error.restype = ctypes.c_char_p
h.emu_set_sms.argtypes = [I]
def aligned(nbytes, align=128):
    buf = torch.zeros(nbytes + align, dtype=torch.uint8); off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + float(sys.argv[2] if len(sys.argv) > 2 else 300)
n = 0; worst = {}
while time.time() < t_end:
    kind = rng.choice(["tc", "packed"])
    B, H, Q, hd = rng.randint(1, 2), rng.randint(1, 3), rng.randint(1, 128), rng.choice([32, 64])
    S = rng.choice([rng.randint(1, 130), rng.randint(100, 700), rng.randint(500, 1500)])
    masked, shared = rng.random() < 0.5, rng.random() < 0.25
    late = rng.choice([0, 1]); h.emu_set_late(late)
    sms = rng.choice([1, 2, 7, 148]); h.emu_set_sms(sms)   # the key-split planner depends on the SM count
    torch.manual_seed(rng.randrange(1 << 30))
    C = H * hd
    q, kv = torch.randn(B, Q, C), torch.randn(B, S, 2 * C)
    hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
    flags, kappa = 3, 30.0
    if shared:
        masked = False; flags, kappa = 0, 10.0
        kv = F.normalize(kv.view(B, S, 2 * H, hd), dim=-1).reshape(B, S, 2 * C); q = F.normalize(q.view(B, Q, H, hd), dim=-1).reshape(B, Q, C)
    q4, k4 = hv(q), hv(kv[..., :C]); v4 = k4 if shared else hv(kv[..., C:])
    bits = ro = eff = None
    if masked:
        blocked = torch.rand(B, Q, S) < rng.choice([0.1, 0.5, 0.95]); 
        if Q > 3: blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        words = (S + 31) // 32
        pad = torch.zeros(B, Q, words * 32, dtype=torch.bool); pad[..., :S] = blocked
        val = (pad.view(B, Q, words, 32).long() << torch.arange(32)).sum(-1)
        bits = torch.where(val >= 2 ** 31, val - 2 ** 32, val).to(torch.int32).contiguous()
        eff = (blocked & (ro != 0).unsqueeze(-1)).unsqueeze(1)
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
    G = B * H
    desc = f"{kind} B{B} H{H} Q{Q} S{S} hd{hd} mask{int(masked)} shared{int(shared)} late{late} sms{sms}"
    h.emu_set_timeout(300.0)
    try:
        if kind == "tc":
            wsb = h.emu_vmf_tc_workspace_bytes(G, Q, S, hd); ws = torch.zeros(wsb // 4)
            ns_max = wsb // 4 // (Q * (hd + 1)) // G
            pa, pd = ws[:G * ns_max * Q * hd], ws[G * ns_max * Q * hd:]
            ns = h.emu_vmf_attention_tc_partial(*st(q4), *st(k4), *st(v4), bits.data_ptr() if masked else None, bits.shape[2] if masked else 0,
                                                ro.data_ptr() if masked else None, B, H, Q, S, hd, kappa, flags, pa.data_ptr(), pd.data_ptr(), 300.0)
            assert ns > 0, (ns, h.emu_last_error())
            acc = pa[:G * ns * Q * hd].view(G, ns, Q, hd).sum(1); den = pd[:G * ns * Q].view(G, ns, Q).sum(1)
            out = F.normalize(acc / den.unsqueeze(-1), dim=-1).view(B, H, Q, hd)
        else:
            pf = 8 if shared else 3
            packed = aligned(h.msmx_vmf_packed_bytes(B, H, S, hd, pf))
            assert h.msmx_vmf_pack(*st(k4), *st(v4), packed.data_ptr(), B, H, S, hd, pf, None) == 0, h.emu_last_error()
            wsb = h.msmx_vmf_packed_workspace_bytes(B, H, Q, S, hd); ws = torch.zeros(wsb, dtype=torch.uint8)
            out = torch.full((B, Q, H, hd), float("nan")).permute(0, 2, 1, 3)
            rc = h.msmx_vmf_attention_packed_fwd(*st(q4), packed.data_ptr(), *st(out), bits.data_ptr() if masked else None, bits.shape[2] if masked else 0,
                                                 ro.data_ptr() if masked else None, B, H, Q, S, hd, kappa, pf, ws.data_ptr(), wsb, None)
            assert rc == 0, (rc, h.emu_last_error())
    except AssertionError as ex:
        print("FAIL", desc, ex, flush=True); continue
    qn = F.normalize(q4.double(), dim=-1) if flags & 1 else q4.double(); kn = F.normalize(k4.double(), dim=-1) if flags & 2 else k4.double()
    s = kappa * qn @ kn.transpose(-1, -2)
    if eff is not None: s = s.masked_fill(eff, float("-inf"))
    ref = F.normalize(torch.softmax(s, -1) @ v4.double(), dim=-1)
    err = (out.double() - ref).abs().max().item(); n += 1
    if not (err == err) or err > 3e-5: print("BAD ", desc, f"err {err:.3e}", flush=True)
    worst[kind] = max(worst.get(kind, 0.0), err if err == err else 9e9)
print("cases", n, {k: f"{v:.1e}" for k, v in worst.items()})
