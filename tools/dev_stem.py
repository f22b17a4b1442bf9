"""ResNet stem on B200: 7x7/2 conv + ReLU with 3 vs 4 (zero-padded) input channels, and ATen's NHWC max-pool."""
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = True
dev = "cuda"
def timed(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
x3 = torch.randn(8, 3, 480, 640, device=dev).contiguous(memory_format=torch.channels_last)
w3 = torch.randn(64, 3, 7, 7, device=dev).contiguous(memory_format=torch.channels_last)
b = torch.randn(64, device=dev)
x4 = F.pad(x3, (0, 0, 0, 0, 0, 1)).contiguous(memory_format=torch.channels_last)
w4 = F.pad(w3, (0, 0, 0, 0, 0, 1)).contiguous(memory_format=torch.channels_last)
x8 = F.pad(x3, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
w8 = F.pad(w3, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
f = lambda x, w: torch.cudnn_convolution_relu(x, w, b, (2, 2), (3, 3), (1, 1), 1)
y3, y4 = f(x3, w3), f(x4, w4)
print("max diff 3 vs 4 channels", (y3 - y4).abs().max().item(), "scale", y3.abs().max().item())
print("stem conv C=3: %.1f us   C=4: %.1f us   C=8: %.1f us" % (timed(lambda: f(x3, w3)), timed(lambda: f(x4, w4)), timed(lambda: f(x8, w8))))
print("pad 3->4 channels: %.1f us" % timed(lambda: F.pad(x3, (0, 0, 0, 0, 0, 1))))
print("maxpool NHWC: %.1f us" % timed(lambda: F.max_pool2d(y3, 3, 2, 1)))
yn = y3.contiguous()
print("maxpool NCHW: %.1f us" % timed(lambda: F.max_pool2d(yn, 3, 2, 1)))
