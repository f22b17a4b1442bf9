"""Why is the clusterer's back-to-back loop slower than the sum of its stages? (B200 box)"""
import os, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import ops
B, n, d, m = 16, 307200, 64, 100
g = torch.Generator(device="cuda").manual_seed(0)
X = F.normalize(torch.randn(B, n, d, device="cuda", generator=g), dim=-1)
first = torch.randint(0, n, (B,), device="cuda")
def step(sync=False):
    seeds, sel = ops.select_smart_seeds(X, m, first)
    if sync: torch.cuda.synchronize()
    Z = ops.mean_shift_hill_climb(X, seeds, 20.0, 10)
    if sync: torch.cuda.synchronize()
    lab, num = ops.seed_connected_components(Z, 0.04)
    return ops.assign_clusters(X, Z, lab, num)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("persistent=%s: loop %.2f ms/step, with syncs between stages %.2f ms/step" % (os.environ.get("MSM_MS_PERSISTENT", "1"), timed(step), timed(lambda: step(True))))
Z0 = X[:, :m].contiguous()
print("hill climb alone, back to back: %.2f ms" % timed(lambda: ops.mean_shift_hill_climb(X, Z0, 20.0, 10)))
print("seeds alone, back to back: %.2f ms" % timed(lambda: ops.select_smart_seeds(X, m, first)))
