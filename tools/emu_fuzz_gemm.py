"""Shape sweep of the shipped tcgen05 GEMM-type kernels (linear / conv / fused epilogues / mask einsum / FFN) under the
CPU emulation (tests/emu): random shapes, leading dimensions, epilogue options, SM counts (planning decisions depend on
num_sms) and execution order (synchronous / late), each against an fp64 reference.

    python -m pytest tests/test_kernel_emulation.py -q     # builds build/emu/*.so
    python tools/emu_fuzz_gemm.py [seed] [seconds]

Found the N-split of the fused LayerNorm epilogue (fixed in linear_tc.cu, see the regression tests)."""
import ctypes, sys, time, random, torch, torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from unseenobjectswithmeanshift_b200._lib import SIGNATURES
h = ctypes.CDLL("/root/repo/build/emu/libemu_gemm_tc.so")
P, I, L, Fl, Z, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_double
for n in ("msm_linear_weight_bytes", "msm_linear_prepare_weight", "msm_linear_fwd", "msm_linear_ln_fwd", "msm_linear_fused_fwd",
          "msm_conv1x1_fwd", "msm_conv3x3_fwd", "msm_ffn_ln_fwd"):
    f = getattr(h, n); f.restype, f.argtypes = SIGNATURES[n]
h.emu_mask_logits_tc.restype = I; h.emu_mask_logits_tc.argtypes = [P, P, P, I, I, I, L]
h.emu_set_timeout.argtypes = [D]; h.emu_set_sms.argtypes = [I]; h.emu_last_error.restype = ctypes.c_char_p
h.emu_set_late.argtypes = [I]
def aligned(nbytes, align=1024):
    buf = torch.zeros(nbytes + align, dtype=torch.uint8); off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]
def fl(t, a=128):
    b = aligned(t.numel() * 4, a).view(torch.float32).view(t.shape); b.copy_(t); return b
def prep(W):
    N, K = W.shape; p = aligned(h.msm_linear_weight_bytes(N, K))
    assert h.msm_linear_prepare_weight(W.data_ptr(), W.stride(0), p.data_ptr(), N, K, None) == 0
    return p
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + float(sys.argv[2] if len(sys.argv) > 2 else 300)
n = 0; worst = {}
while time.time() < t_end:
    kind = rng.choice(["linear", "ln", "fused", "conv1", "conv3", "mask", "ffn"])
    sms = rng.choice([1, 2, 5, 148, 148]); h.emu_set_late(rng.choice([0, 1])); h.emu_set_sms(sms); h.emu_set_timeout(300.0)
    torch.manual_seed(rng.randrange(1 << 30))
    try:
        if kind == "linear":
            M, N, K, act = rng.randint(1, 700), 32 * rng.randint(1, 12), 32 * rng.choice([1, 2, 3, 8, 16]), rng.randint(0, 1)
            ldx, ldy = K + 4 * rng.randint(0, 3), N + 4 * rng.randint(0, 3)
            X, W, b = fl(torch.randn(M, ldx)), torch.randn(N, K) / K ** 0.5, torch.randn(N)
            Y = fl(torch.full((M, ldy), float("nan"))); p = prep(W)
            rc = h.msm_linear_fwd(X.data_ptr(), ldx, p.data_ptr(), b.data_ptr() if rng.random() < 0.8 else None, Y.data_ptr(), ldy, M, N, K, act, None)
            b_used = b if True else None
            desc = f"linear M{M} N{N} K{K} act{act} ldx{ldx} ldy{ldy} sms{sms}"
            assert rc == 0, h.emu_last_error()
            # bias may have been None: recompute both and accept either
            r0 = X[:, :K].double() @ W.double().t(); r1 = r0 + b.double()
            if act: r0, r1 = r0.relu(), r1.relu()
            err = min((Y[:, :N].double() - r0).abs().max().item(), (Y[:, :N].double() - r1).abs().max().item())
            assert torch.isnan(Y[:, N:]).all() if ldy > N else True, "wrote outside its columns"
        elif kind == "ln":
            M, N, K = rng.randint(1, 600), rng.choice([32, 64]), 32 * rng.choice([1, 2, 4, 8])
            X, W, b, R = fl(torch.randn(M, K)), torch.randn(N, K) / K ** 0.5, torch.randn(N), fl(torch.randn(M, N))
            g, be = torch.randn(N), torch.randn(N); Y = fl(torch.full((M, N), float("nan"))); p = prep(W)
            rc = h.msm_linear_ln_fwd(X.data_ptr(), K, p.data_ptr(), b.data_ptr(), R.data_ptr(), N, g.data_ptr(), be.data_ptr(), 1e-5, Y.data_ptr(), N, M, N, K, None)
            desc = f"ln M{M} N{N} K{K} sms{sms}"; assert rc == 0, h.emu_last_error()
            err = (Y.double() - F.layer_norm(R.double() + X.double() @ W.double().t() + b.double(), (N,), g.double(), be.double(), 1e-5)).abs().max().item()
        elif kind == "fused":
            M, N, K = rng.randint(1, 500), 32 * rng.randint(1, 8), 32 * rng.choice([1, 2, 8])
            period = rng.randint(1, max(1, min(M, 100)))
            X, W, b, R = fl(torch.randn(M, K)), torch.randn(N, K) / K ** 0.5, torch.randn(N), fl(torch.randn(M, N))
            rb, g, be, g2, be2 = fl(torch.randn(period, N)), torch.randn(N), torch.randn(N), torch.randn(N), torch.randn(N)
            use_rb, act, use_res, use_ln, l2, use_y2 = (rng.random() < 0.5 for _ in range(6))
            use_y2 = use_y2 and use_ln
            Y, Y2 = fl(torch.full((M, N), float("nan"))), fl(torch.full((M, N), float("nan"))); p = prep(W)
            rc = h.msm_linear_fused_fwd(X.data_ptr(), K, p.data_ptr(), b.data_ptr(), rb.data_ptr() if use_rb else None, period, int(act),
                                        R.data_ptr() if use_res else None, N, g.data_ptr() if use_ln else None, be.data_ptr() if use_ln else None, 1e-5,
                                        int(l2), g2.data_ptr() if use_y2 else None, be2.data_ptr() if use_y2 else None, 1e-5,
                                        Y2.data_ptr() if use_y2 else None, N, Y.data_ptr(), N, M, N, K, None)
            desc = f"fused M{M} N{N} K{K} period{period} rb{int(use_rb)} act{int(act)} res{int(use_res)} ln{int(use_ln)} l2{int(l2)} y2{int(use_y2)} sms{sms}"
            if rc != 0:
                print("  (rejected)", desc, h.emu_last_error().decode()); continue
            v = X.double() @ W.double().t() + b.double()
            if use_rb: v = v + rb.double().repeat((M + period - 1) // period, 1)[:M]
            if act: v = v.relu()
            if use_res: v = v + R.double()
            if use_ln: v = F.layer_norm(v, (N,), g.double(), be.double(), 1e-5)
            if l2: v = F.normalize(v, dim=-1)
            err = (Y.double() - v).abs().max().item() / max(1.0, v.abs().max().item())
            if use_y2: err = max(err, (Y2.double() - F.layer_norm(v, (N,), g2.double(), be2.double(), 1e-5)).abs().max().item() / 10)
        elif kind == "conv1":
            B, K, N, HW, nchw = rng.randint(1, 3), 32 * rng.choice([1, 2, 8]), 32 * rng.randint(1, 8), 4 * rng.randint(1, 120), rng.random() < 0.5
            x, w, b = fl(torch.randn(B, K, HW)), torch.randn(N, K) / K ** 0.5, torch.randn(N)
            Y = fl(torch.full((B, N, HW) if nchw else (B, HW, N), float("nan"))); p = prep(w)
            rc = h.msm_conv1x1_fwd(x.data_ptr(), p.data_ptr(), b.data_ptr(), Y.data_ptr(), int(nchw), B, HW, N, K, 0, None)
            desc = f"conv1 B{B} K{K} N{N} HW{HW} nchw{int(nchw)} sms{sms}"; assert rc == 0, h.emu_last_error()
            ref = torch.einsum("nk,bkp->bnp", w.double(), x.double()) + b.double()[None, :, None]
            err = (Y.double() - (ref if nchw else ref.transpose(1, 2))).abs().max().item()
        elif kind == "conv3":
            B, C, N, H, W = rng.randint(1, 2), 32 * rng.choice([1, 2]), 32 * rng.choice([1, 2, 8]), rng.randint(1, 14), rng.randint(1, 70)
            x, w, b = torch.randn(B, C, H, W), torch.randn(N, C, 3, 3) / (9 * C) ** 0.5, torch.randn(N)
            Y = fl(torch.full((B, N, H, W), float("nan"))); p = prep(w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous())
            Wp = (W + 2 + 3) // 4 * 4; xp = fl(F.pad(x, (1, Wp - W - 1, 1, 1)))
            rc = h.msm_conv3x3_fwd(xp.data_ptr(), p.data_ptr(), b.data_ptr(), Y.data_ptr(), B, C, H, W, Wp, N, 1, None)
            desc = f"conv3 B{B} C{C} N{N} {H}x{W} sms{sms}"; assert rc == 0, h.emu_last_error()
            err = (Y.double() - F.conv2d(x.double(), w.double(), b.double(), padding=1).relu()).abs().max().item()
        elif kind == "mask":
            B, Q, C, HW = rng.randint(1, 3), rng.randint(1, 128), 32 * rng.randint(1, 8), 4 * rng.randint(1, 300)
            e, f_ = fl(torch.randn(B, Q, C)), fl(torch.randn(B, C, HW)); out = fl(torch.full((B, Q, HW), float("nan")))
            rc = h.emu_mask_logits_tc(e.data_ptr(), f_.data_ptr(), out.data_ptr(), B, Q, C, HW)
            desc = f"mask B{B} Q{Q} C{C} HW{HW} sms{sms}"
            if rc == -2: continue
            assert rc == 0, h.emu_last_error()
            ref = torch.einsum("bqc,bcp->bqp", e.double(), f_.double()); err = (out.double() - ref).abs().max().item() / max(1.0, ref.abs().max().item()) * 10
        else:
            M, Dm, Fh = rng.randint(1, 600), rng.choice([32, 64]), 128 * rng.randint(1, 8)
            x = fl(torch.randn(M, Dm)); w1, b1 = torch.randn(Fh, Dm) / Dm ** 0.5, torch.randn(Fh) * 0.1
            w2, b2 = torch.randn(Dm, Fh) / Fh ** 0.5, torch.randn(Dm) * 0.1; g, be = torch.randn(Dm), torch.randn(Dm)
            y = fl(torch.full((M, Dm), float("nan"))); p1, p2 = prep(w1), prep(w2)
            rc = h.msm_ffn_ln_fwd(x.data_ptr(), Dm, p1.data_ptr(), b1.data_ptr(), p2.data_ptr(), b2.data_ptr(), g.data_ptr(), be.data_ptr(), 1e-5, y.data_ptr(), Dm, M, Dm, Fh, None)
            desc = f"ffn M{M} D{Dm} F{Fh} sms{sms}"; assert rc == 0, h.emu_last_error()
            xd = x.double()
            err = (y.double() - F.layer_norm(xd + F.linear(F.relu(F.linear(xd, w1.double(), b1.double())), w2.double(), b2.double()), (Dm,), g.double(), be.double(), 1e-5)).abs().max().item()
    except AssertionError as ex:
        print("FAIL", desc, ex, flush=True); continue
    n += 1
    if not (err == err) or err > 5e-5:
        print("BAD ", desc, f"err {err:.3e}", flush=True)
    worst[kind] = max(worst.get(kind, 0.0), err if err == err else 9e9)
print("cases", n, {k: f"{v:.1e}" for k, v in worst.items()})
