import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def child():
    import torch
    from unseenobjectswithmeanshift_b200 import ops
    for (B, Q, C, H, W) in [(8, 16, 256, 120, 160), (8, 100, 256, 120, 160), (1, 16, 256, 480, 640), (1, 100, 256, 480, 640)]:
        e, f = torch.randn(B, Q, C, device="cuda"), torch.randn(B, C, H, W, device="cuda")
        got = ops.mask_logits(e, f)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): ops.mask_logits(e, f, out=got)
        a.record()
        for _ in range(20): ops.mask_logits(e, f, out=got)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        by = 4.0 * B * (C * H * W + Q * H * W + Q * C)
        print(f"  B{B} Q{Q} C{C} {H}x{W}: {us:8.1f} us  {by / us / 1e3:7.1f} GB/s", flush=True)
if __name__ == "__main__":
    if len(sys.argv) > 1: child()
    else:
        for st in ("2", "3", "4", "6", "8", "10"):
            print("stages", st, flush=True)
            subprocess.run([sys.executable, __file__, "child"], env={**os.environ, "MSM_TC_STAGES": st}, timeout=300)
