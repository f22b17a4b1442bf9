"""Launch counts and wall time of the training step's phases (config #5 shapes, B = 8): head forward | criterion |
backward | optimizer. Launches from CUPTI records per phase."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from unseenobjectswithmeanshift_b200 import ops, training, workloads  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    dev = torch.device("cuda")
    model = workloads.build_trainer("r50").to(dev)
    opt = training.build_optimizer(model, lr=1e-4)
    feats = {k: v.to(dev) for k, v in workloads.synthetic_features("r50", 8, seed=0, pin=False).items()}
    targets = [{k: v.to(dev) for k, v in t.items()} for t in workloads.synthetic_targets("r50", 8, seed=0)]
    batch = {"features": feats, "targets": targets}
    for _ in range(3):
        training.train_step(model, opt, batch)
    torch.cuda.synchronize()

    def phase(name, fn):
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            t0 = time.perf_counter()
            r = fn()
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
        n = sum(1 for ev in prof.events() if "cuda" in str(getattr(ev, "device_type", "")).lower())
        dev_ms = sum(ev.time_range.end - ev.time_range.start for ev in prof.events()
                     if "cuda" in str(getattr(ev, "device_type", "")).lower()) / 1e3
        print(f"{name:12s} launches {n:5d}  host enqueue {t_host * 1e3:7.2f} ms  wall {t_all * 1e3:7.2f} ms  kernels {dev_ms:7.2f} ms")
        return r

    opt.zero_grad(set_to_none=True)
    outputs = phase("head fwd", lambda: model.sem_seg_head(feats, *model.size)[0])
    losses = phase("criterion", lambda: model.criterion(outputs, targets))
    w = model.criterion.weight_dict
    total = sum(v * w[k] for k, v in losses.items() if k in w)
    phase("backward", lambda: total.backward())
    phase("clip+step", lambda: (torch.nn.utils.clip_grad_norm_(model.parameters(), 0.01), opt.step(), ops.bump_weights_epoch()))


if __name__ == "__main__":
    main()
