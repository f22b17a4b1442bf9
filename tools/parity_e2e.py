"""End-to-end parity of the head on a bounded sample vs the CPU oracle, with per-layer error growth.

    python tools/parity_e2e.py [--workload r50] [--images 2] [--seed 0]
Env switches (read at import): MSM_DISABLE_TC=1 (all fp32 CUDA-core kernels), MSM_DISABLE_TC_LINEAR=1.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import head as ohead
from unseenobjectswithmeanshift_b200 import workloads


def stats(got, ref):
    pk = ref.abs().max().item()
    d = (got - ref).abs()
    return {"max_rel_peak": d.max().item() / pk, "mean_rel_peak": d.mean().item() / pk,
            "frac_gt_1e-3": (d > 1e-3 * pk).float().mean().item(), "frac_gt_1e-4": (d > 1e-4 * pk).float().mean().item()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50")
    ap.add_argument("--images", type=int, default=2)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    head = workloads.build_head(a.workload, seed=a.seed)
    sd = {k: v.detach() for k, v in head.state_dict().items()}
    feats = workloads.synthetic_features(a.workload, a.images, seed=a.seed)
    cfg = workloads.HEAD_CFG[a.workload]
    t0 = time.perf_counter()
    with torch.no_grad():
        ref, ref_mf = ohead.head_forward(sd, feats, **workloads.oracle_kwargs(a.workload))
    t_cpu = time.perf_counter() - t0
    head = head.cuda()
    with torch.no_grad():
        out, mf = head({k: v.cuda() for k, v in feats.items()}, cfg["height"], cfg["width"])
    torch.cuda.synchronize()
    res = {"env": {k: os.environ.get(k, "") for k in ("MSM_DISABLE_TC", "MSM_DISABLE_TC_LINEAR")},
           "workload": a.workload, "images": a.images, "cpu_s": t_cpu,
           "mask_features": stats(mf.cpu(), ref_mf), "pred_masks": stats(out["pred_masks"].cpu(), ref["pred_masks"]),
           "pred_logits": stats(out["pred_logits"].cpu(), ref["pred_logits"]),
           "argmax_agreement": (out["pred_masks"].cpu().argmax(1) == ref["pred_masks"].argmax(1)).float().mean().item(),
           "aux_max_rel_peak": [stats(x["pred_masks"].cpu(), y["pred_masks"])["max_rel_peak"]
                                for x, y in zip(out["aux_outputs"], ref["aux_outputs"])]}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
