"""Dev check of the clusterer kernels (csrc/cluster.cu) on a small problem against the oracle; run under
compute-sanitizer on the GPU box before the full tests:  python tools/dev_cluster.py [n] [B]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mean_shift as oms  # noqa: E402
from unseenobjectswithmeanshift_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d, m = 64, 30
g = torch.Generator().manual_seed(1)
centers = F.normalize(torch.randn(5, d, generator=g), dim=1)
X = torch.stack([F.normalize(centers[torch.randint(0, 5, (n,), generator=g)] + 0.05 * torch.randn(n, d, generator=g), dim=1)
                 for _ in range(B)])
first = [3 + b for b in range(B)]
seeds, sel = ops.select_smart_seeds(X.cuda(), m, first)
torch.cuda.synchronize()
print("seeds done", flush=True)
Z = ops.mean_shift_hill_climb(X.cuda(), seeds, 20.0, 10)
lab, num = ops.seed_connected_components(Z, 0.04)
torch.cuda.synchronize()
print("cc done", num.tolist(), flush=True)
labels = ops.assign_clusters(X.cuda(), Z, lab, num)
torch.cuda.synchronize()
print("assign done", flush=True)
for b in range(B):
    _, want = oms.select_smart_seeds(X[b], m, first[b])
    print(b, "seed idx equal:", torch.equal(sel[b].cpu(), want), "cc equal:",
          torch.equal(lab[b].cpu(), oms.connected_components(Z[b].cpu(), 0.04)),
          "label hist:", torch.bincount(labels[b].cpu()).tolist())
    ref_labels, _ = oms.mean_shift_smart_init(X[b], 20.0, m, 10, first[b])
    print(b, "label agreement with the oracle pipeline:", float((labels[b].cpu() == ref_labels).float().mean()))

if len(sys.argv) > 3 and sys.argv[3] == "time":
    Bt, nt, mt = 16, 307200, 100
    Xt = F.normalize(torch.randn(Bt, nt, d, device="cuda"), dim=2)
    ft = torch.arange(Bt, device="cuda") * 1000
    for variant in ("seeds", "assign"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Zt = Xt[:, :mt].contiguous()
        lt = torch.arange(mt, device="cuda").repeat(Bt, 1)
        nl = torch.full((Bt,), mt, device="cuda", dtype=torch.int32)
        fn = (lambda: ops.select_smart_seeds(Xt, mt, ft)) if variant == "seeds" else (lambda: ops.assign_clusters(Xt, Zt, lt, nl))
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        by = (4.0 * d + 8) * Bt * nt * ((mt - 1) if variant == "seeds" else 1)
        print(f"{variant}: {ms:.2f} ms  {by / ms / 1e6:.0f} GB/s algorithmic (MSM_SEEDS_VARIANT={os.environ.get('MSM_SEEDS_VARIANT', '1')})",
              flush=True)
