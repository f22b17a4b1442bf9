"""cuDNN variants of the R50 backbone at B=8 640x480 (B200 box): math (TF32 / fp32), layout, fused conv+bias+ReLU."""
import sys, os, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import backbones

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

x = torch.randn(8, 3, 480, 640, device="cuda")
ref = None
for tf32, fused in itertools.product((True, False), (False, True)):
    backbones.set_tf32(tf32)
    m = backbones.ResNet50Features(fused_relu=fused).cuda()
    with torch.no_grad():
        out = m(x)
        if ref is None:
            ref = out
        err = max(((out[k] - ref[k]).abs().max() / ref[k].abs().max()).item() for k in out)
        t = timed(lambda: m(x))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(x)
        tg = timed(g.replay)
    print(f"tf32={tf32} fused_relu={fused}: eager {t:.2f} ms, graph {tg:.2f} ms, max rel diff vs first {err:.2e}", flush=True)
