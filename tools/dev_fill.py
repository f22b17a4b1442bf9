"""Pure-write and pure-read HBM bandwidth of this box (context for write-bound kernels): fill_ / sum over 2 GB."""
import torch
x = torch.empty(1 << 29, dtype=torch.float32, device="cuda")   # 2 GiB
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
by = x.numel() * 4
t = timed(lambda: x.fill_(1.0)); print(f"fill_  : {t*1e3:.0f} us  {by/t/1e6:.0f} GB/s write")
t = timed(lambda: x.zero_()); print(f"zero_  : {t*1e3:.0f} us  {by/t/1e6:.0f} GB/s write (memset)")
t = timed(lambda: x.sum()); print(f"sum    : {t*1e3:.0f} us  {by/t/1e6:.0f} GB/s read")
y = torch.empty_like(x)
t = timed(lambda: y.copy_(x)); print(f"copy_  : {t*1e3:.0f} us  {2*by/t/1e6:.0f} GB/s read+write")
