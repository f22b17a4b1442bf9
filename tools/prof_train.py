"""Where the training step's device time goes (config #5 shapes, B = 8, fp32): CUPTI kernel records of one step.
    python tools/prof_train.py [bf16]"""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from unseenobjectswithmeanshift_b200 import training, workloads  # noqa: E402


def main():
    amp = torch.bfloat16 if "bf16" in sys.argv[1:] else None
    dev = torch.device("cuda")
    model = workloads.build_trainer("r50", amp_dtype=amp).to(dev)
    opt = training.build_optimizer(model, lr=1e-4)
    feats = {k: v.to(dev) for k, v in workloads.synthetic_features("r50", 8, seed=0, pin=False).items()}
    targets = [{k: v.to(dev) for k, v in t.items()} for t in workloads.synthetic_targets("r50", 8, seed=0)]
    batch = {"features": feats, "targets": targets}
    for _ in range(3):
        training.train_step(model, opt, batch)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        training.train_step(model, opt, batch)
    b.record()
    torch.cuda.synchronize()
    print(f"train step: {a.elapsed_time(b) / 5:.2f} ms")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        training.train_step(model, opt, batch)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if "cuda" in str(getattr(ev, "device_type", "")).lower():
            n = re.sub(r"\(.*", "", ev.name.replace("void ", "").replace("at::native::", "").replace("(anonymous namespace)::", ""))[:90]
            agg[n][0] += 1
            agg[n][1] += (ev.time_range.end - ev.time_range.start)
    tot = sum(v[1] for v in agg.values())
    print(f"kernel time {tot / 1e3:.2f} ms in {sum(v[0] for v in agg.values())} launches")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t / 1e3:8.3f} ms {c:5d} {100 * t / tot:5.1f}%  {n}")


if __name__ == "__main__":
    main()
