"""Persistent mean-shift: the steps of test_mean_shift_config4_size_properties one by one, synchronised."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import ops
small = len(sys.argv) > 1 and sys.argv[1] == "small"
g = torch.Generator().manual_seed(4)
n, d, m = (20000 if small else 307200), 64, 100
X = F.normalize(torch.randn(n, d, generator=g), dim=1).cuda()
Z0 = X[torch.randperm(n, generator=g)[:m].cuda()].clone()
def run(tag, *a):
    z = ops.mean_shift_hill_climb(*a); torch.cuda.synchronize(); print(tag, "ok", float(z.norm(dim=1).mean())); return z
Z = run("10 iters", X, Z0, 10.0, 10)
Z2 = run("doubled", torch.cat([X, X]), Z0, 10.0, 10); print("  diff", (Z2 - Z).abs().max().item())
Za = run("4 iters", X, Z0, 10.0, 4)
Zb = run("6 more", X, Za, 10.0, 6); print("  composition equal:", torch.equal(Zb, Z), (Zb - Z).abs().max().item())
u = F.normalize(torch.randn(1, d, generator=g), dim=1)
Zu = run("fixed point", u.repeat(5000, 1).cuda(), Z0[:7].contiguous(), 10.0, 2); print("  diff", (Zu.cpu() - u).abs().max().item())
for B in (2, 3, 5):
    Xb = torch.stack([X] * B); Zb0 = torch.stack([Z0] * B)
    zb = run(f"B={B}", Xb, Zb0, 10.0, 3)
    print("  images identical:", all(torch.equal(zb[0], zb[i]) for i in range(B)))
