"""GPU dev check of the EXPERIMENTAL packed-operand mean-shift kernel (csrc/experimental/vmf_packed.cu).

    python tools/dev_vmf_packed.py build      # cross-compile build/experimental/libmsmx_vmf_packed.so (no GPU needed)
    python tools/dev_vmf_packed.py [quick]    # on the B200 box: build if stale, parity + timing vs the shipped kernel

The experimental library is separate from libmsmformer_b200.so (the product build globs csrc/*.cu only) and is bound
here and nowhere else. The check runs in a child process under a timeout: a hung mbarrier protocol must not take the
box with it. Parity: seeds after 10 iterations against the shipped msm_mean_shift_hill_climb and, at small n, an
fp64 restatement of seed_hill_climbing_ball (transformer_decoder/mean_shift.py:79-109).
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "unseenobjectswithmeanshift_b200", "csrc")
XLIB = os.path.join(ROOT, "build", "experimental", "libmsmx_vmf_packed.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build():
    srcs = [os.path.join(CSRC, "experimental", "vmf_packed.cu"), os.path.join(CSRC, "common.cu")]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc.cuh")]
    if os.path.exists(XLIB) and all(os.path.getmtime(d) <= os.path.getmtime(XLIB) for d in deps):
        return XLIB
    os.makedirs(os.path.dirname(XLIB), exist_ok=True)
    subprocess.check_call([NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-Xcompiler", "-fPIC", "-shared", "-o", XLIB] + srcs)
    return XLIB


def bind():
    P, I, F, Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    h = ctypes.CDLL(build())
    h.msm_last_error.restype = ctypes.c_char_p
    h.msmx_mean_shift_packed_bytes.restype = Z
    h.msmx_mean_shift_packed_bytes.argtypes = [I, I, I]
    h.msmx_mean_shift_packed_workspace_bytes.restype = Z
    h.msmx_mean_shift_packed_workspace_bytes.argtypes = [I, I, I, I]
    h.msmx_mean_shift_pack.restype = I
    h.msmx_mean_shift_pack.argtypes = [P, P, I, I, I, P]
    h.msmx_mean_shift_hill_climb_packed.restype = I
    h.msmx_mean_shift_hill_climb_packed.argtypes = [P, P, P, I, I, I, I, F, I, P, Z, P]
    return h


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


if __name__ == "__main__":
    arg = sys.argv[1] if len(sys.argv) > 1 else "full"
    if arg == "build":
        print("built", build())
    elif arg == "child":
        child(len(sys.argv) > 2 and sys.argv[2] == "quick")
    else:
        build()
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), "child", arg], timeout=400)
        sys.exit(rc)
