"""Run one hot-path op at a BASELINE.json size a few times - the command ncu wraps.

    python tools/prof_ops.py mask_r50|mask_ucn|vmf_r50|vmf_ucn|meanshift|msda|linear_kv|linear_ffn1|linear_ffn2|linear_ucn [--iters N]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unseenobjectswithmeanshift_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    with torch.no_grad():
        if a.what in ("mask_r50", "mask_ucn"):
            B, H, W = (8, 120, 160) if a.what == "mask_r50" else (1, 480, 640)
            e, f = rn(B, 100, 256), rn(B, 256, H, W)
            out = torch.empty(B, 100, H, W, device=dev)
            fn = lambda: ops.mask_logits(e, f, out=out)  # noqa: E731
        elif a.what in ("vmf_r50", "vmf_ucn"):
            B, S = (8, 4800) if a.what == "vmf_r50" else (1, 307200)
            q, k, v = rn(B, 100, 256), rn(B, S, 256), rn(B, S, 256)
            hv = lambda t: t.unflatten(-1, (8, 32)).permute(0, 2, 1, 3)  # noqa: E731
            bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, 100, (S + 31) // 32), device=dev, dtype=torch.int32)
            ro = torch.ones(B, 100, device=dev, dtype=torch.int32)
            fn = lambda: ops.vmf_attention(hv(q), hv(k), hv(v), blocked_bits=bits, row_open=ro)  # noqa: E731
        elif a.what == "meanshift":
            B = 4
            X = torch.nn.functional.normalize(rn(B, 307200, 64), dim=-1)
            Z = X[:, :100].contiguous()
            fn = lambda: ops.mean_shift_hill_climb(X, Z, 10.0, 10)  # noqa: E731
        elif a.what == "msda":
            N, M, D, L, P = 8, 8, 8, 3, 4
            shapes = torch.tensor([[15, 20], [30, 40], [60, 80]], device=dev)
            lsi = torch.tensor([0, 300, 1500], device=dev)
            S = 6300
            value = rn(N, S, M, D)
            loc = torch.rand(N, S, M, L, P, 2, device=dev, generator=g)
            aw = torch.softmax(rn(N, S, M, L * P), -1).view(N, S, M, L, P)
            fn = lambda: ops.ms_deform_attn_forward(value, shapes, lsi, loc, aw)  # noqa: E731
        elif a.what in ("linear_kv", "linear_ffn1", "linear_ffn2", "linear_ucn", "linear_ow", "linear_vp", "linear_small",
                        "linear_small_ffn1", "linear_small_ffn2", "linear_mf"):
            M, N, K = {"linear_kv": (38400, 768, 256), "linear_ffn1": (50400, 1024, 64),
                       "linear_ffn2": (50400, 64, 1024), "linear_ucn": (307200, 256, 256),
                       "linear_ow": (50400, 288, 64), "linear_vp": (50400, 64, 64), "linear_small": (800, 256, 256),
                       "linear_small_ffn1": (800, 2048, 256), "linear_small_ffn2": (800, 256, 2048),
                       "linear_mf": (153600, 256, 64)}[a.what]
            x, w, b = rn(M, K), rn(N, K) / K ** 0.5, rn(N)
            out = torch.empty(M, N, device=dev)
            fn = lambda: ops.linear(x, w, b, out=out)  # noqa: E731
        elif a.what == "ffn":
            M, D, F = 50400, 64, 1024
            x = rn(M, D)
            w1, b1, w2, b2 = rn(F, D) / 8, rn(F), rn(D, F) / 32, rn(D)
            norm = torch.nn.LayerNorm(D).cuda()
            fn = lambda: ops.ffn_ln(x, w1, b1, w2, b2, norm)  # noqa: E731
        elif a.what == "msda_fused":
            N, M, D, L, P = 8, 8, 8, 3, 4
            shapes = torch.tensor([[15, 20], [30, 40], [60, 80]], device=dev)
            lsi = torch.tensor([0, 300, 1500], device=dev)
            S = 6300
            value = rn(N, S, M, D)
            ow = rn(N, S, M * L * P * 3)
            ref = torch.rand(N, S, L, 2, device=dev, generator=g)
            fn = lambda: ops.ms_deform_attn_fused_forward(value, shapes, lsi, ow, ref, L, P)  # noqa: E731
        else:
            raise SystemExit(f"unknown op {a.what}")
        for _ in range(a.iters):
            fn()
        torch.cuda.synchronize()
        # device-timed average, L2 flushed between iterations by a 256 MB write
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            s, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e2.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e2) * 1e3)
        print(f"{a.what}: {min(ts):.1f} us best, {sum(ts) / len(ts):.1f} us mean over {a.iters} (events around one call)")


if __name__ == "__main__":
    main()
