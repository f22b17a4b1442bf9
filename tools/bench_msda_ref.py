"""The reference's OWN MSDeformAttn CUDA kernel (ops/src/cuda/ms_deform_im2col_cuda.cuh, built for sm_100a by
oracle/make_ref.py with the two-token torch-2.x patch) against this library's kernels, same box, same inputs:

    python tools/bench_msda_ref.py            # on the B200 box; prints a markdown table

forward:  reference = softmax + sampling-location arithmetic (5 elementwise torch ops, ops/modules/ms_deform_attn.py:
          103-109) + MSDA.ms_deform_attn_forward;   ours = msm_ms_deform_attn_fused_fwd (one kernel), and the plain
          op-for-op msm_ms_deform_attn_fwd.
backward: reference = MSDA.ms_deform_attn_backward (D = 8 -> 8-thread blocks, .cuh:306-408);  ours = msm_ms_deform_attn_bwd.
Geometry: config #2, N8 Lq6300 M8 D8 L3 P4 (and N1 for the single-image latency)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "msda_build"))
from unseenobjectswithmeanshift_b200 import ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    import MultiScaleDeformableAttention as REF   # the reference's pybind module (baseline/_ref/msda_build)
    dev = torch.device("cuda")
    M, D, L, P = 8, 8, 3, 4
    shapes = torch.tensor([[15, 20], [30, 40], [60, 80]], device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    print("| N | what | reference kernel (us) | ours (us) | speed-up | max abs diff |")
    print("|---|---|---:|---:|---:|---:|")
    for N in (8, 1):
        g = torch.Generator(device="cuda").manual_seed(N)
        value = torch.randn(N, S, M, D, device=dev, generator=g)
        ol = torch.randn(N, S, M * L * P * 3, device=dev, generator=g)      # raw offsets | raw logits, as the module's GEMM emits
        ref_pts = torch.rand(N, S, L, 2, device=dev, generator=g)
        n_off = M * L * P * 2
        wh = torch.stack([shapes[..., 1], shapes[..., 0]], -1)

        def ref_forward():
            offsets = ol[..., :n_off].reshape(N, S, M, L, P, 2)
            weights = torch.softmax(ol[..., n_off:].reshape(N, S, M, L * P), -1).view(N, S, M, L, P)
            loc = ref_pts[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
            return REF.ms_deform_attn_forward(value, shapes, lsi, loc.contiguous(), weights.contiguous(), 128), loc, weights

        out_ref, loc, weights = ref_forward()
        loc, weights = loc.contiguous(), weights.contiguous()
        out_fused = ops.ms_deform_attn_fused_forward(value, shapes, lsi, ol, ref_pts, L, P)
        out_plain = ops.ms_deform_attn_forward(value, shapes, lsi, loc, weights, 128)
        t_ref_full = timed(lambda: ref_forward())
        t_ref_op = timed(lambda: REF.ms_deform_attn_forward(value, shapes, lsi, loc, weights, 128))
        t_fused = timed(lambda: ops.ms_deform_attn_fused_forward(value, shapes, lsi, ol, ref_pts, L, P))
        t_plain = timed(lambda: ops.ms_deform_attn_forward(value, shapes, lsi, loc, weights, 128))
        print(f"| {N} | forward incl. softmax + locations (module math) | {t_ref_full:.1f} | {t_fused:.1f} | "
              f"{t_ref_full / t_fused:.2f}x | {(out_ref - out_fused).abs().max().item():.2e} |")
        print(f"| {N} | forward, the op alone | {t_ref_op:.1f} | {t_plain:.1f} | {t_ref_op / t_plain:.2f}x | "
              f"{(out_ref - out_plain).abs().max().item():.2e} |")
        go = torch.randn_like(out_ref)
        r = REF.ms_deform_attn_backward(value, shapes, lsi, loc, weights, go, 128)
        o = ops.ms_deform_attn_backward(value, shapes, lsi, loc, weights, go, 128)
        diff = max((a - b).abs().max().item() / max(a.abs().max().item(), 1e-12) for a, b in zip(r, o))
        t_rb = timed(lambda: REF.ms_deform_attn_backward(value, shapes, lsi, loc, weights, go, 128))
        t_ob = timed(lambda: ops.ms_deform_attn_backward(value, shapes, lsi, loc, weights, go, 128))
        print(f"| {N} | backward (D = 8) | {t_rb:.1f} | {t_ob:.1f} | {t_rb / t_ob:.2f}x | {diff:.2e} (rel. to peak) |")


if __name__ == "__main__":
    main()
