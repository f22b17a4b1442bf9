"""The K / V operand-image projections of the UCN decoder alone (one image, all 6 layers of the level in one GEMM):
    python tools/prof_kimg.py [k|kpos|v|k256] [reps]     # M 307200, N 1536, K 64 (folded) or 256
ncu: ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 1 -o gpurun_out/prof python tools/prof_kimg.py kpos 3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import ops  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "kpos"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda")
    B, Hh, Ww, C, layers = 1, 480, 640, 256, 6
    S, N = Hh * Ww, layers * C
    K = 256 if what == "k256" else 64
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, K, Hh, Ww, device=dev, generator=g) if K == 64 else torch.randn(B, S, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) * 0.1
    b = torch.randn(N, device=dev, generator=g) * 0.1
    pos = None
    if what == "kpos":
        pos = (torch.randn(Hh, N, device=dev, generator=g).contiguous(), torch.randn(Ww, N, device=dev, generator=g).contiguous())
    images, _ = ops.packed_kv_alloc(layers, B, C // 32, S, dev)
    which = 1 if what == "v" else 0
    fn = lambda: ops.linear_packed_kv(x, w, b, images, B, S, C, which, pos=pos)  # noqa: E731
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = 4.0 * B * S * (K + N)
    print(f"{what}: {ms * 1e3:.1f} us per launch, {by / ms / 1e6:.0f} GB/s algorithmic ({by / 1e6:.0f} MB)")
    if "stamps" in sys.argv[3:]:
        import ctypes
        from unseenobjectswithmeanshift_b200 import _lib
        h = _lib.xlib()
        h.msmx_linear_debug.argtypes = [ctypes.c_void_p]
        h.msmx_linear_debug.restype = None
        buf = torch.zeros(148, 32, dtype=torch.int64, device=dev)
        h.msmx_linear_debug(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        h.msmx_linear_debug(None)
        b = buf.double().cpu()
        names = {0: ("producer", ["wait empty_x", "wait empty_w", "-"]), 4: ("mma", ["wait acc_empty", "wait full_w", "wait full_a"]),
                 8: ("converter w8", ["wait full_x", "wait empty_a", "-"]), 12: ("epilogue w4", ["wait acc_full", "-", "-"]),
                 16: ("epilogue w12", ["wait acc_full", "-", "-"])}
        for base, (role, labels) in names.items():
            tot = b[:, base + 3].mean().item()
            parts = ", ".join(f"{lab} {100 * b[:, base + i].mean().item() / max(tot, 1):.0f}%" for i, lab in enumerate(labels) if lab != "-")
            print(f"  {role:14s} total {tot / 1.9e3:8.1f} us (mean over CTAs): {parts}")


if __name__ == "__main__":
    main()
