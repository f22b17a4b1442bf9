"""GPU dev check of the tcgen05 mask GEMM: error vs fp64 and timing vs the CUDA-core kernel.
Run each configuration in a fresh process (the env switches are read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child():
    import torch
    from unseenobjectswithmeanshift_b200 import ops
    for (B, Q, C, H, W) in [(1, 16, 32, 8, 16), (2, 100, 256, 120, 160), (1, 100, 256, 224, 224), (1, 37, 64, 33, 36), (8, 100, 256, 120, 160)]:
        g = torch.Generator().manual_seed(Q + C)
        e, f = torch.randn(B, Q, C, generator=g), torch.randn(B, C, H, W, generator=g)
        ref = torch.einsum("bqc,bchw->bqhw", e.double(), f.double())
        ec, fc = e.cuda(), f.cuda()
        got = ops.mask_logits(ec, fc)
        torch.cuda.synchronize()
        err = (got.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): ops.mask_logits(ec, fc, out=got)
        a.record()
        for _ in range(20): ops.mask_logits(ec, fc, out=got)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        by = 4.0 * B * (C * H * W + Q * H * W + Q * C)
        print(f"  B{B} Q{Q} C{C} {H}x{W}: peak-rel err {err:.3e}   {us:8.1f} us  {by / us / 1e3:7.1f} GB/s", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for env in ({"MSM_DISABLE_TC": "1"}, {}):
            print("env", env, flush=True)
            r = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, timeout=300)
            print("  rc", r.returncode, flush=True)
