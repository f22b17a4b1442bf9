"""Kernel timeline of one graph replay of a forward workload (CUPTI via torch.profiler): span, busy time per stream,
largest gaps. usage: python tools/dev_timeline.py [r50|r50-head|demo] [aux]  ("aux" = eval_aux_masks True)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from unseenobjectswithmeanshift_b200 import backbones, workloads  # noqa: E402
from unseenobjectswithmeanshift_b200.graph import GraphedForward  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "r50"
    aux = "aux" in sys.argv[2:]
    dev = torch.device("cuda")
    full = kind in ("r50", "demo")
    B = 1 if kind == "demo" else 8
    hk = {"r50": "r50", "demo": "ucn", "r50-head": "r50"}[kind]
    H, W = workloads.HEAD_CFG[hk]["height"], workloads.HEAD_CFG[hk]["width"]
    if full:
        backbones.set_tf32(True)
        model = workloads.build_model(kind).to(dev)
        if aux:
            model.sem_seg_head.predictor.eval_aux_masks = True
        inp = {k: v.to(dev) for k, v in workloads.synthetic_images(kind, B, seed=0, pin=False).items()}

        def step(x):
            lm, f = model.label_maps([x])
            return {"label_map": lm, "scores": f["scores"]}
    else:
        head = workloads.build_head(hk).to(dev)
        if not aux:
            head.predictor.eval_aux_masks = False
        inp = {k: v.to(dev) for k, v in workloads.synthetic_features(hk, B, seed=0, pin=False).items()}

        def step(x):
            out, _ = head(x, H, W)
            return {"pred_logits": out["pred_logits"], "pred_masks": out["pred_masks"]}
    with torch.no_grad():
        for _ in range(3):
            step(inp)
        g = GraphedForward(step, inp, warmup=2)
        for _ in range(5):
            g()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g()
        b.record()
        torch.cuda.synchronize()
        print(f"{kind} aux={aux}: graph replay {a.elapsed_time(b) / 50:.3f} ms/step (before CUPTI attaches)")
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            g()
            torch.cuda.synchronize()
        a.record()
        for _ in range(50):
            g()
        b.record()
        torch.cuda.synchronize()
        print(f"after CUPTI attached: {a.elapsed_time(b) / 50:.3f} ms/step")
    evs = []
    for ev in prof.events():
        if "cuda" in str(getattr(ev, "device_type", "")).lower():
            evs.append((ev.time_range.start, ev.time_range.end, ev.name))
    evs.sort()
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    print(f"kernels {len(evs)}  span {(t1 - t0) / 1e3:.3f} ms  sum {sum(e[1] - e[0] for e in evs) / 1e3:.3f} ms")
    # union busy time and the gaps where NO kernel runs
    cur_end, busy, gaps = evs[0][0], 0.0, []
    for s, e, n in evs:
        if s > cur_end:
            gaps.append((s - cur_end, cur_end - t0, n))
            busy += 0
            cur_end_prev = cur_end
        if e > cur_end:
            busy += e - max(s, cur_end)
            cur_end = e
    print(f"union busy {busy / 1e3:.3f} ms, idle {((t1 - t0) - busy) / 1e3:.3f} ms in {len(gaps)} gaps")
    for gap, at, n in sorted(gaps, reverse=True)[:12]:
        print(f"   gap {gap:8.1f} us at +{at / 1e3:7.3f} ms before {n[:80]}")
    # coarse phases: time of first kernel whose name matches
    marks = ["cudnn", "msda", "ffn_tc", "vmf_attn", "mask_gemm", "instance"]
    for m in marks:
        hit = [e for e in evs if m in e[2].lower()]
        if hit:
            print(f"   {m:10s} first +{(hit[0][0] - t0) / 1e3:7.3f} ms  last end +{(max(h[1] for h in hit) - t0) / 1e3:7.3f} ms  n={len(hit)}")
    if "dump" in sys.argv[2:]:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/timeline_{kind}.txt", "w") as f:
            for s_, e_, n_ in evs:
                f.write(f"{(s_ - t0):9.1f} {(e_ - s_):8.1f} {n_[:140]}\n")
    # longest kernels
    for s, e, n in sorted(evs, key=lambda x: x[0] - x[1])[:10]:
        print(f"   {e - s:8.1f} us at +{(s - t0) / 1e3:7.3f} ms {n[:90]}")


if __name__ == "__main__":
    main()
