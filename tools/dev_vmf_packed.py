"""GPU dev check of the EXPERIMENTAL packed-operand attention / mean-shift kernels (csrc/vmf_attention_packed.cu) and of
the operand-image projection epilogue (linear_tc_kernel<PACK>).

    python tools/dev_vmf_packed.py [quick]    # on the B200 box: build if stale, parity + timing vs the shipped kernels

The kernels live in libmsmformer_b200.so behind msmx_ entry points that nothing calls by default (MSM_PACKED_KV=1 routes
the decoder through them). The check runs in a child process under a timeout: a hung mbarrier protocol must not take the
box with it. Parity: (1) mean-shift seeds after 10 iterations against the shipped msm_mean_shift_hill_climb and, at
small n, an fp64 restatement of seed_hill_climbing_ball (transformer_decoder/mean_shift.py:79-109); (2) the decoder's
cross-attention (heads by strides, bit masks, K normalised + fp16 halves, V bf16 halves) against the shipped
msm_vmf_attention_fwd, with the stand-alone pack kernel timed separately; (3) the fused chain - K / V projections
writing the images, packed attention - against dense projections + the shipped attention.
The same sources already run green on CPU threads under the calibrated emulation (tests/test_kernel_emulation.py).
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build():
    import __graft_entry__
    __graft_entry__.build()


def bind():
    from unseenobjectswithmeanshift_b200 import _lib
    build()
    return _lib.xlib()


def ref64(X, Z, kappa, iters):
    import torch
    X, Z = X.double(), Z.double()
    for _ in range(iters):
        s = kappa * (Z @ X.transpose(-1, -2) - 1.0)  # the common factor exp(-kappa) cancels in the normalisation
        Z = torch.nn.functional.normalize(torch.exp(s) @ X, dim=-1, eps=1e-12)
    return Z


def child(quick):
    import torch
    from unseenobjectswithmeanshift_b200 import ops
    h = bind()
    dev = torch.device("cuda")
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def packed_climb(X, Z, kappa, iters, packed=None):
        B, n, d = X.shape
        m = Z.shape[1]
        if packed is None:
            packed = torch.empty(h.msmx_mean_shift_packed_bytes(B, n, d), dtype=torch.uint8, device=dev)
            rc = h.msmx_mean_shift_pack(X.data_ptr(), packed.data_ptr(), B, n, d, st())
            assert rc == 0, h.msm_last_error()
        wsb = h.msmx_mean_shift_packed_workspace_bytes(B, n, m, d)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        out = torch.empty_like(Z)
        rc = h.msmx_mean_shift_hill_climb_packed(packed.data_ptr(), Z.data_ptr(), out.data_ptr(), B, n, m, d,
                                                 float(kappa), iters, ws.data_ptr(), wsb, st())
        assert rc == 0, h.msm_last_error()
        return out, packed

    def timed(fn, n):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    # B, n, m, d, kappa: tails (n % 128 != 0), one tile, fewer than 128 seeds, both dims, then config #4's geometry
    cases = [(1, 128, 100, 64, 10.0), (1, 100, 7, 64, 10.0), (2, 5000, 100, 64, 10.0), (3, 777, 37, 32, 20.0),
             (2, 40000, 128, 32, 10.0), (4, 307200, 100, 64, 10.0)]
    if not quick:
        cases.append((32, 307200, 100, 64, 10.0))
    for (B, n, m, d, kappa) in cases:
        g = torch.Generator(device="cuda").manual_seed(n + d)
        X = torch.nn.functional.normalize(torch.randn(B, n, d, device=dev, generator=g), dim=-1)
        idx = torch.stack([torch.randperm(n, device=dev, generator=g)[:m] for _ in range(B)])
        Z = torch.gather(X, 1, idx.unsqueeze(-1).expand(B, m, d)).contiguous()
        got, packed = packed_climb(X, Z, kappa, 10)
        want = ops.mean_shift_hill_climb(X, Z, kappa, 10)
        torch.cuda.synchronize()
        e_ship = (got - want).abs().max().item()
        e_ref = float("nan")
        if B * n * m <= 2e8:
            e_ref = (got.double() - ref64(X, Z, kappa, 10)).abs().max().item()
        reps = 3 if n > 100000 else 10
        t_pack = timed(lambda: h.msmx_mean_shift_pack(X.data_ptr(), packed.data_ptr(), B, n, d, st()), reps)
        t_new = timed(lambda: packed_climb(X, Z, kappa, 10, packed), reps)
        t_old = timed(lambda: ops.mean_shift_hill_climb(X, Z, kappa, 10), reps)
        by = 4.0 * B * n * d * 10
        print(f"  B{B} n{n} m{m} d{d}: |packed - shipped| {e_ship:.2e}  |packed - fp64| {e_ref:.2e}   "
              f"pack {t_pack * 1e3:8.1f} us  climb packed {t_new * 1e3:9.1f} us ({by / t_new / 1e6:7.1f} GB/s)  "
              f"shipped {t_old * 1e3:9.1f} us ({by / t_old / 1e6:7.1f} GB/s)", flush=True)
        assert e_ship < 2e-4, "packed kernel disagrees with the shipped kernel"

    # ---- decoder cross-attention on packed K / V images vs the shipped kernel
    def pack_bits(blocked):
        Bq, Q, S = blocked.shape
        words = (S + 31) // 32
        pad = torch.zeros(Bq, Q, words * 32, dtype=torch.bool, device=dev)
        pad[..., :S] = blocked
        v = (pad.view(Bq, Q, words, 32).long() << torch.arange(32, device=dev)).sum(-1)
        return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()

    xcases = [(1, 2, 100, 700, 32, True), (2, 8, 100, 4800, 32, True), (1, 1, 37, 333, 64, False),
              (1, 8, 100, 50176, 32, True), (1, 8, 100, 307200, 32, True)]
    if not quick:
        xcases.append((2, 8, 100, 307200, 32, True))
    for (B, H, Q, S, hd, masked) in xcases:
        g = torch.Generator(device="cuda").manual_seed(S + hd)
        C = H * hd
        q = torch.randn(B, Q, C, device=dev, generator=g)
        kv = torch.randn(B, S, 2 * C, device=dev, generator=g)
        hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
        q4, k4, v4 = hv(q), hv(kv[..., :C]), hv(kv[..., C:])
        bits = ro = None
        if masked:
            blocked = torch.rand(B, Q, S, device=dev, generator=g) < 0.5
            blocked[:, 3] = True
            ro = (~blocked).any(-1).to(torch.int32).contiguous()
            bits = pack_bits(blocked)
        sd = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))
        packed = torch.empty(h.msmx_vmf_packed_bytes(B, H, S, hd, 3), dtype=torch.uint8, device=dev)
        wsb = h.msmx_vmf_packed_workspace_bytes(B, H, Q, S, hd)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        out = torch.empty(B, Q, H, hd, device=dev).permute(0, 2, 1, 3)

        def pack():
            rc = h.msmx_vmf_pack(*sd(k4), *sd(v4), packed.data_ptr(), B, H, S, hd, 3, st())
            assert rc == 0, h.msm_last_error()

        def attend():
            rc = h.msmx_vmf_attention_packed_fwd(*sd(q4), packed.data_ptr(), *sd(out),
                                                 bits.data_ptr() if masked else None, bits.shape[2] if masked else 0,
                                                 ro.data_ptr() if masked else None, B, H, Q, S, hd, 30.0, 3,
                                                 ws.data_ptr(), wsb, st())
            assert rc == 0, h.msm_last_error()

        pack()
        attend()
        want = ops.vmf_attention(q4, k4, v4, blocked_bits=bits, row_open=ro)
        torch.cuda.synchronize()
        err = (out - want).abs().max().item()
        reps = 3 if S > 100000 else 10
        t_pack, t_new = timed(pack, reps), timed(attend, reps)
        t_old = timed(lambda: ops.vmf_attention(q4, k4, v4, blocked_bits=bits, row_open=ro), reps)
        fl = 4.0 * B * H * Q * S * hd
        print(f"  attn B{B} H{H} Q{Q} S{S} hd{hd} mask{int(masked)}: |packed - shipped| {err:.2e}   pack {t_pack * 1e3:8.1f} us  "
              f"packed {t_new * 1e3:9.1f} us ({fl / t_new / 1e9:6.1f} TFLOP/s)  shipped {t_old * 1e3:9.1f} us "
              f"({fl / t_old / 1e9:6.1f} TFLOP/s)", flush=True)
        assert err < 2e-4, "packed attention disagrees with the shipped kernel"

    # ---- the fused chain: K / V projections writing the images (linear_tc_kernel<PACK> in the product library) +
    # packed attention, against dense K, dense V + the shipped attention kernel
    L = h
    ccases = [(2, 300, 256, 256, 3, 100), (1, 4800, 256, 256, 3, 100), (1, 50176, 256, 256, 1, 100),
              (1, 307200, 256, 256, 1, 100)]
    for (B, S, Cin, C, layers, Q) in ccases:
        g = torch.Generator(device="cuda").manual_seed(S + layers)
        H = C // 32
        src = torch.randn(B, S, Cin, device=dev, generator=g)
        key_in = src + torch.randn(B, S, Cin, device=dev, generator=g)
        wk = torch.randn(layers * C, Cin, device=dev, generator=g) / Cin ** 0.5
        wv = torch.randn(layers * C, Cin, device=dev, generator=g) / Cin ** 0.5
        bk, bv = (0.1 * torch.randn(layers * C, device=dev, generator=g) for _ in range(2))
        q = torch.randn(B, Q, C, device=dev, generator=g)
        blocked = torch.rand(B, Q, S, device=dev, generator=g) < 0.5
        blocked[:, 3] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = pack_bits(blocked)
        hv = lambda t: t.unflatten(-1, (H, 32)).permute(0, 2, 1, 3)
        per_layer = h.msmx_vmf_packed_bytes(B, H, S, 32, 3)
        packed = torch.zeros(layers * per_layer, dtype=torch.uint8, device=dev)   # key tails stay zero
        pk, pv = ops.prepare_linear_weight(wk), ops.prepare_linear_weight(wv)
        wsb = h.msmx_vmf_packed_workspace_bytes(B, H, Q, S, 32)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        out = torch.empty(B, Q, H, 32, device=dev).permute(0, 2, 1, 3)
        sd = lambda t: (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))

        def project_packed():
            for which, x, w, b in ((0, key_in, pk, bk), (1, src, pv, bv)):
                rc = L.msmx_linear_packed_kv_fwd(x.data_ptr(), Cin, w.data_ptr(), b.data_ptr(), packed.data_ptr(), B, S,
                                                 layers * C, Cin, C, which, 1, 1, st())
                assert rc == 0, L.msm_last_error()

        def attend_packed(j=0):
            rc = h.msmx_vmf_attention_packed_fwd(*sd(hv(q)), packed.data_ptr() + j * per_layer, *sd(out), bits.data_ptr(),
                                                 bits.shape[2], ro.data_ptr(), B, H, Q, S, 32, 30.0, 3, ws.data_ptr(), wsb,
                                                 st())
            assert rc == 0, h.msm_last_error()

        def shipped(j=0, project=True, attend=True):
            if project:
                shipped.K, shipped.V = ops.dense(key_in, wk, bk), ops.dense(src, wv, bv)
            if attend:
                return ops.vmf_attention(hv(q), hv(shipped.K[..., j * C:(j + 1) * C]), hv(shipped.V[..., j * C:(j + 1) * C]),
                                         blocked_bits=bits, row_open=ro)

        with torch.no_grad():
            project_packed()
            worst = 0.0
            for j in range(layers):
                attend_packed(j)
                want = shipped(j, project=(j == 0))
                torch.cuda.synchronize()
                worst = max(worst, (out - want).abs().max().item())
            reps = 3 if S > 100000 else 10
            t_pp, t_pa = timed(project_packed, reps), timed(attend_packed, reps)
            t_sp = timed(lambda: shipped(project=True, attend=False), reps)
            t_sa = timed(lambda: shipped(project=False, attend=True), reps)
        print(f"  chain B{B} S{S} C{C} layers{layers}: |fused - shipped| {worst:.2e}   projections packed {t_pp * 1e3:9.1f} us "
              f"vs dense {t_sp * 1e3:9.1f} us   attention packed {t_pa * 1e3:9.1f} us vs shipped {t_sa * 1e3:9.1f} us",
              flush=True)
        assert worst < 2e-4, "fused chain disagrees with the shipped path"


if __name__ == "__main__":
    arg = sys.argv[1] if len(sys.argv) > 1 else "full"
    if arg == "build":
        print("built", build())
    elif arg == "child":
        child(len(sys.argv) > 2 and sys.argv[2] == "quick")
    else:
        build()
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), "child", arg], timeout=600)
        sys.exit(rc)
