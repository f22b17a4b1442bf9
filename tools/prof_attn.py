"""One attention / mean-shift shape, a few launches - the command ncu wraps (B200_PROFILING.md):

    ncu --set full --clock-control none --import-source on -k regex:vmf_attn -s 2 -c 1 -o gpurun_out/prof \
        python tools/prof_attn.py ucn            # B1 H8 Q100 S307200 hd32, bit mask, K/V as packed operand images
    python tools/prof_attn.py ms                 # mean-shift iteration geometry: B4 n307200 m100 d64 (k == v)
    python tools/prof_attn.py r50                # B8 H8 Q100 S4800 hd32
Prints the CUDA-event time of the attention launch (kernel + finalize) when run without ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import _lib, ops  # noqa: E402


def pack_bits(blocked):
    B, Q, S = blocked.shape
    words = (S + 31) // 32
    pad = torch.zeros(B, Q, words * 32, dtype=torch.bool, device=blocked.device)
    pad[..., :S] = blocked
    v = (pad.view(B, Q, words, 32).long() << torch.arange(32, device=blocked.device)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "ucn"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda")
    h = _lib.xlib()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    g = torch.Generator(device="cuda").manual_seed(0)
    if what.startswith("msp"):   # whole hill climb (10 iterations): persistent kernel unless MSM_MS_PERSISTENT=0
        B, n, m, d = int(what[3:] or 4), 307200, 100, 64
        X = torch.nn.functional.normalize(torch.randn(B, n, d, device=dev, generator=g), dim=-1)
        Z = X[:, :m].contiguous()
        fn = lambda: ops.mean_shift_hill_climb(X, Z, 10.0, 10)  # noqa: E731
        flops = 4.0 * B * m * n * d * 10
    elif what == "ms":
        B, n, m, d = 4, 307200, 100, 64
        X = torch.nn.functional.normalize(torch.randn(B, n, d, device=dev, generator=g), dim=-1)
        Z = X[:, :m].contiguous()
        fn = lambda: ops.mean_shift_hill_climb(X, Z, 10.0, 1)  # noqa: E731
        flops = 4.0 * B * m * n * d
    else:
        B, H, Q, S, hd = {"ucn": (1, 8, 100, 307200, 32), "crop": (16, 8, 100, 50176, 32),
                          "r50": (8, 8, 100, 4800, 32)}[what]
        C = H * hd
        q = torch.randn(B, Q, C, device=dev, generator=g)
        kv = torch.randn(B, S, 2 * C, device=dev, generator=g)
        hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)  # noqa: E731
        blocked = torch.rand(B, Q, S, device=dev, generator=g) < 0.5
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = pack_bits(blocked)
        kvp = ops.pack_kv(hv(kv[..., :C]), hv(kv[..., C:]))
        out = torch.empty(B, Q, H, hd, device=dev).permute(0, 2, 1, 3)
        fn = lambda: ops.vmf_attention_packed(hv(q), kvp, blocked_bits=bits, row_open=ro, out=out)  # noqa: E731
        flops = 4.0 * B * H * Q * S * hd
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
    print(f"{what}: {ms * 1e3:.1f} us per call, {flops / ms / 1e9:.1f} TFLOP/s useful, x3 passes "
          f"{3 * flops / ms / 1e9:.1f}")


if __name__ == "__main__":
    main()
