"""cuDNN math / layout variants of the R50 backbone at B=8 640x480 (B200 box): which fp32 path is usable?"""
import sys, os, time, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import backbones

def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

x = torch.randn(8, 3, 480, 640, device="cuda")
for tf32, cl, bench in itertools.product((False, True), (True, False), (False, True)):
    torch.backends.cudnn.benchmark = bench
    backbones.set_tf32(tf32)
    m = backbones.ResNet50Features().cuda()
    if not cl:
        m = m.to(memory_format=torch.contiguous_format)
        fwd = m.forward
        def run(m=m):
            with backbones._conv_math():
                y = m.stem(x); out = {}
                for name in ("res2", "res3", "res4", "res5"):
                    y = getattr(m, name)(y); out[name] = y
            return out
    else:
        run = lambda m=m: m(x)
    with torch.no_grad():
        t = timed(run)
    print(f"tf32={tf32} channels_last={cl} cudnn.benchmark={bench}: {t:.2f} ms", flush=True)
