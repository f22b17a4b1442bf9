"""GPU dev check of the tcgen05 linear kernel: error vs fp64, time vs torch F.linear (cuBLAS fp32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from unseenobjectswithmeanshift_b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
cases = [(100, 256, 256, False), (130, 96, 64, True), (800, 2048, 256, True), (800, 256, 2048, False),
         (50400, 64, 64, False), (50400, 288, 64, False), (50400, 1024, 64, True), (50400, 64, 1024, False),
         (38400, 768, 256, False), (9600, 768, 256, False), (307200, 256, 256, False)]
for (M, N, K, relu) in cases:
    x = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g)
    y = ops.linear(x, w, b, relu=relu)
    torch.cuda.synchronize()
    ref = x[:4096].double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    err = (y[:4096].double() - ref).abs().max().item() / ref.abs().max().item()
    tail = (y[-300:].double() - ((x[-300:].double() @ w.double().t() + b.double()).clamp_min(0) if relu else (x[-300:].double() @ w.double().t() + b.double()))).abs().max().item()
    def timeit(fn, n=20):
        for _ in range(3): fn()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize()
        return a.elapsed_time(e) / n * 1e3
    out = torch.empty_like(y)
    t_tc = timeit(lambda: ops.linear(x, w, b, relu=relu, out=out))
    t_th = timeit(lambda: F.relu(F.linear(x, w, b)) if relu else F.linear(x, w, b))
    fl = 2.0 * M * N * K
    by = 4.0 * (M * K + M * N)
    print(f"  M{M} N{N} K{K} relu{int(relu)}: peak-rel err {err:.2e} tail abs {tail:.2e}  tc {t_tc:8.1f} us ({fl / t_tc / 1e6:7.1f} TF/s, {by / t_tc / 1e3:7.1f} GB/s)   torch {t_th:8.1f} us", flush=True)
# strided input / output slices
x = torch.randn(1000, 512, device=dev, generator=g)
w = torch.randn(128, 256, device=dev, generator=g)
big = torch.zeros(1000, 384, device=dev)
ops.linear(x[:, 256:], w, None, out=big[:, 128:256])
torch.cuda.synchronize()
ref = x[:, 256:].double() @ w.double().t()
print("  strided: err", ((big[:, 128:256].double() - ref).abs().max() / ref.abs().max()).item(), "untouched", big[:, :128].abs().max().item(), big[:, 256:].abs().max().item())
