"""bench.py - images/sec of the MSMFormer segmentation head (the hot path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload r50|ucn|crop]

Workload (BASELINE.json configs[1]): ResNet-50-config head - MSDeformAttn pixel decoder over
res2..res5 of a 640x480 image + 9-layer mean-shift transformer decoder, 100 queries - batch 8 per
GPU, synthetic backbone features, random-init weights, fp32. One step = one forward of the head
over one batch. N > 1: one replica per rank (torchrun), batch-sharded, no data-path collective.

Prints ONE JSON line (rank 0). ``value``: inputs resident in HBM. ``e2e``: the same step through
the public module call with pinned HOST feature tensors copied in and the predictions copied out
inside the timed region. ``roofline``: the dominant kernel of this library, timed with CUDA
events inside the timed region. ``cpu_baseline``: the CPU oracle port on a bounded sample.
``--impl reference`` times that CPU port alone (the reference's PyTorch path restated; the
reference itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# stdout carries exactly ONE JSON line. Libraries (NCCL's version banner, cuDNN notices) write to file descriptor 1
# directly, so the process's fd 1 is pointed at stderr and the JSON line goes to a duplicate of the original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


METRIC = "images/sec MSMFormer head forward 640x480 (R50 config: MSDeformAttn pixel decoder + 9-layer mean-shift decoder, 100 queries)"
PER_GPU_BATCH = 8


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines, self.skip = gpu_index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # nvidia-smi takes 0.1-0.3 s to attach to the driver and stalls kernel launches while it does: wait for
            # its first sample HERE, before the timed region is opened, and count only the samples after it
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 3.0 and self.proc.poll() is None:
                time.sleep(0.01)
            self.skip = len(self.lines)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in (self.lines[self.skip:] or self.lines):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_step(sd, feats, kind):
    from oracle import head as ohead
    from unseenobjectswithmeanshift_b200 import workloads
    with torch.no_grad():
        out, _ = ohead.head_forward(sd, feats, **workloads.oracle_kwargs(kind))
    return out


def run_reference(args, rank, world):
    """The reference's CPU path for the same head (oracle port; all host threads), B=1 per step."""
    if rank != 0:
        return
    from unseenobjectswithmeanshift_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.workload in ("meanshift", "cluster", "tail"):
        run_reference_aux(args, cores)
        return
    if args.workload in ("train", "twostage"):
        run_reference_train_twostage(args, cores)
        return
    kind = args.workload
    full = kind in FULL_MODEL
    # bounded sample per step: the whole --steps K --warmup W run has to end within a few minutes
    sample_b = 1 if (kind in ("demo", "ucn") or args.steps > 50) else 2
    if kind in ("demo", "ucn"):   # 307200-key configurations take ~30 s per image on the CPU
        args.steps, args.warmup = min(args.steps, 3), min(args.warmup, 1)
    host_in = (workloads.synthetic_images(kind, sample_b) if full
               else workloads.synthetic_features(HEAD_KIND[kind], sample_b))
    cpu_step, which, note = _cpu_model(kind, full)
    with torch.no_grad():
        for _ in range(args.warmup):
            cpu_step(host_in)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_step(host_in)
        dt = time.perf_counter() - t0
    val = sample_b * args.steps / dt
    line = {"impl": "reference", "metric": METRICS[kind], "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{kind} 640x480 (CPU reference arm: {note})", "per_step_batch": sample_b,
                       "queries": 100},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": which,
                             "sample": f"{args.steps} steps x {sample_b} image(s) of the same workload; {note}"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_reference_train_twostage(args, cores):
    """--impl reference for the train / twostage workloads: the oracle port on a bounded sample per step.
    train: forward + losses + backward of the R50-config head on 1 image (torch.autograd through the oracle, the
    criterion mirror on CPU; no optimizer update). twostage: stage-1 oracle head on 1 frame + crop oracle head on its
    5 crops (the glue between them is negligible on the CPU)."""
    from oracle import head as ohead
    from unseenobjectswithmeanshift_b200 import workloads
    if args.workload == "train":
        from unseenobjectswithmeanshift_b200.meanshiftformer.meanshiftformer_model import build_criterion
        head = workloads.build_head("r50")
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in head.state_dict().items()}
        feats = workloads.synthetic_features("r50", 1)
        targets = workloads.synthetic_targets("r50", 1)
        crit = build_criterion(2, dec_layers=workloads.HEAD_CFG["r50"]["dec_layers"] + 1)

        def step():
            out, _ = ohead.head_forward(sd, feats, **workloads.oracle_kwargs("r50"))
            losses = crit(out, targets)
            sum(v * crit.weight_dict[k] for k, v in losses.items()).backward()
            for t in sd.values():
                t.grad = None
        sample_b, metric = 1, ("images/sec MSMFormer head training step 640x480 (R50 config: forward + deep-supervision "
                               "losses + backward + clipped AdamW)")
        workload = "train r50-head 640x480, 100 queries, 9 decoder layers, 5 instances/image (no optimizer update)"
    else:
        heads = {k: workloads.build_head(k) for k in ("ucn", "crop")}
        sds = {k: {n: v.detach() for n, v in h.state_dict().items()} for k, h in heads.items()}
        feats = {"ucn": workloads.synthetic_features("ucn", 1), "crop": workloads.synthetic_features("crop", 5)}

        def step():
            with torch.no_grad():
                for k in ("ucn", "crop"):
                    ohead.head_forward(sds[k], feats[k], **workloads.oracle_kwargs(k))
        sample_b, metric = 1, ("frames/sec two-stage RGB-D segmentation 640x480 (stage-1 head + 5 zoom-crops through "
                               "the crop head + paste-back)")
        workload = "twostage 640x480, 5 crops of 224x224 per frame (heads only)"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample_b * args.steps / dt
    emit({"impl": "reference", "metric": metric, "value": val, "unit": "images/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": workload, "per_step_batch": sample_b,
                     "note": "CPU oracle port of the reference PyTorch path, fp32"},
          "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                           "sample": f"{args.steps} steps x {sample_b} image(s) of the same workload"},
          "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def run_reference_aux(args, cores):
    """--impl reference for the stand-alone workloads: the oracle port of the same op on a bounded sample per step
    (1 image for the mean-shift / clusterer workloads, 2 for the eval tail), all host threads."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(4)
    if args.workload == "tail":
        from oracle import instance_inference as oii
        Q, K, h, w, H, W, T, sample_b = 100, 1, 120, 160, 480, 640, 20, 2
        logits = 2 * torch.randn(sample_b, Q, K + 1, generator=g)
        masks = F.interpolate(3 * torch.randn(sample_b, Q, h // 8, w // 8, generator=g), size=(h, w), mode="bicubic")
        step = lambda: oii.inference_tail(logits, masks, (H, W), T)  # noqa: E731
        metric = ("images/sec eval tail: mask upsample + instance_inference (100 queries 120x160 -> top-20 instances "
                  "at 480x640)")
        workload = "tail Q=100 120x160->480x640 top-20"
    else:
        from oracle import mean_shift as oms
        n, d, m, sample_b = 480 * 640, 64, 100, 1
        X = F.normalize(torch.randn(n, d, generator=g), dim=1)
        if args.workload == "meanshift":
            Z0 = X[torch.randperm(n, generator=g)[:m]].clone()
            step = lambda: oms.seed_hill_climbing_ball(X, Z0, 10.0, 10)  # noqa: E731
            metric = ("images/sec standalone vMF mean-shift hill climb (640x480x64-d embeddings, 100 seeds, kappa=10, "
                      "10 iterations)")
            workload = "meanshift n=307200 d=64 m=100 kappa=10 iters=10"
        else:
            step = lambda: oms.mean_shift_smart_init(X, 20.0, m, 10, 0)  # noqa: E731
            metric = ("images/sec classical vMF mean-shift clustering (640x480x64-d embeddings: 100 farthest-point "
                      "seeds, kappa=20, 10 iterations, connected components, nearest-seed labels)")
            workload = "cluster n=307200 d=64 m=100 kappa=20 iters=10"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample_b * args.steps / dt
    emit({"impl": "reference", "metric": metric, "value": val, "unit": "images/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": workload, "per_step_batch": sample_b,
                     "note": "CPU oracle port of the reference PyTorch path, fp32"},
          "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                           "sample": f"{args.steps} steps x {sample_b} image(s) of the same workload"},
          "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def run_meanshift(args, rank, local_rank, world, dev, sharding, ops):
    """BASELINE.json configs[3]: standalone vMF mean-shift hill climb, 640x480x64-d unit embeddings, 100 seeds,
    kappa=10, 10 iterations, `--batch` images per GPU (default 32). One step = the whole 10-iteration climb of
    the batch (20 kernel launches: partial + finalize per iteration); X stays resident in HBM across iterations."""
    import torch.nn.functional as F
    B = args.batch if args.batch > 0 else 32
    n, d, m, kappa, iters = 480 * 640, 64, 100, 10.0, 10
    g = torch.Generator().manual_seed(4 + rank)
    hostX = torch.empty(B, n, d).pin_memory()
    for b in range(B):
        hostX[b] = F.normalize(torch.randn(n, d, generator=g), dim=1)
    idx = torch.stack([torch.randperm(n, generator=g)[:m] for _ in range(B)])
    hostZ = torch.stack([hostX[b, idx[b]] for b in range(B)]).pin_memory()
    X, Z0 = hostX.to(dev), hostZ.to(dev)
    hostOut = torch.empty(B, m, d).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    with torch.no_grad():
        for _ in range(args.warmup):
            ops.mean_shift_hill_climb(X, Z0, kappa, iters)
        torch.cuda.synchronize()
        if rank == 0:
            sampler.start()
        sharding.barrier()
        ops.reset_stats()
        e0.record()
        for _ in range(args.steps):
            Z = ops.mean_shift_hill_climb(X, Z0, kappa, iters)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms_dev = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
        launches = ops.launches()
        # end to end: embeddings from pinned host memory in, modes out, every step
        stage = torch.empty_like(X)
        for _ in range(1):
            stage.copy_(hostX, non_blocking=True)
            hostOut.copy_(ops.mean_shift_hill_climb(stage, Z0, kappa, iters), non_blocking=True)
        torch.cuda.synchronize()
        sharding.barrier()
        n_e2e = max(1, min(args.steps, 5))
        e0.record()
        for _ in range(n_e2e):
            stage.copy_(hostX, non_blocking=True)
            hostOut.copy_(ops.mean_shift_hill_climb(stage, Z0, kappa, iters), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    peaks = load_peaks()
    value = B * world * args.steps / (ms_dev / 1e3)
    by = 4.0 * B * n * d * iters           # X read once per iteration
    fl = 4.0 * B * m * n * d * iters
    t = ms_dev / args.steps / 1e3
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import mean_shift as oms
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t0 = time.perf_counter()
        ref = oms.seed_hill_climbing_ball(hostX[0], hostZ[0], kappa, iters)
        dt = time.perf_counter() - t0
        err = (Z[0].cpu() - ref).abs().max().item()
        cpu_baseline = {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": f"1 image, {iters} iterations (oracle, fp32, {cores} threads)",
                        "parity_on_sample": {"modes_max_abs_err": err}}
    line = {"metric": "images/sec standalone vMF mean-shift hill climb (640x480x64-d embeddings, 100 seeds, kappa=10, "
                      "10 iterations)", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"meanshift n=307200 d=64 m=100 kappa=10 iters=10 batch {B}/GPU",
                       "global_batch": B * world, "parallelism": f"replicas x{world} (batch-sharded, no collective)",
                       "l2_policy": f"inputs_exceed_l2 ({4.0 * B * n * d / 1e6:.0f} MB of embeddings per pass)"},
            "clocks": clocks,
            "e2e": {"value": B * world * n_e2e / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": hostX.numel() * 4, "d2h_bytes_per_step": hostOut.numel() * 4,
                    "ms_per_step": ms_e2e / n_e2e},
            "gpu_launches": launches,
            "roofline": {"kernel": "vmf_attn_tc_kernel<64, shared>", "bound": "hbm", "achieved": by / t / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / t / 1e9 / peaks["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes_per_step": by, "peak_source": peaks["source"],
                         "tensor_TFLOPps_issued": 3.0 * fl / t / 1e12,
                         "tensor_frac_of_bf16_peak": 3.0 * fl / t / 1e12 / peaks["bf16_tflops"]},
            "cpu_baseline": cpu_baseline}
    emit(line)


def run_cluster(args, rank, local_rank, world, dev, sharding, ops):
    """the whole classical clusterer (mean_shift_smart_init / clustering_features,
    lib/fcn/test_dataset.py:43-59) on config #4's geometry: 640x480x64-d unit embeddings, 100 farthest-point seeds,
    kappa=20, 10 hill-climb iterations, connected components, nearest-seed labels. `--batch` images per GPU
    (default 16). One step = the batch through all four stages, no host synchronisation inside."""
    import torch.nn.functional as F
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder import mean_shift as ms
    B = args.batch if args.batch > 0 else 16
    n, d, m, kappa, iters, objects = 480 * 640, 64, 100, 20.0, 10, 12
    g = torch.Generator().manual_seed(40 + rank)
    hostX = torch.empty(B, n, d).pin_memory()
    for b in range(B):
        centers = F.normalize(torch.randn(objects, d, generator=g), dim=1)
        which = torch.multinomial(torch.arange(objects, 0, -1).float(), n, replacement=True, generator=g)
        hostX[b] = F.normalize(centers[which] + 0.04 * torch.randn(n, d, generator=g), dim=1)
    first = torch.randint(0, n, (B,), generator=g)
    X, first_dev = hostX.to(dev), first.to(dev)
    hostOut = torch.empty(B, n, dtype=torch.int64).pin_memory()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    stage_ms = [0.0] * 4
    sampler = ClockSampler(local_rank)

    def step(Xd, record=False):
        if record:
            ev[0].record()
        seeds, selected = ops.select_smart_seeds(Xd, m, first_dev)
        if record:
            ev[1].record()
        Z = ops.mean_shift_hill_climb(Xd, seeds, kappa, iters)
        if record:
            ev[2].record()
        seed_labels, num = ops.seed_connected_components(Z, 0.04)
        if record:
            ev[3].record()
        labels = ops.assign_clusters(Xd, Z, seed_labels, num)
        if record:
            ev[4].record()
        return labels, selected

    with torch.no_grad():
        for _ in range(args.warmup):
            step(X)
        torch.cuda.synchronize()
        for _ in range(3):   # stage split (events between the stages; not part of the timed region)
            step(X, record=True)
            torch.cuda.synchronize()
            for i in range(4):
                stage_ms[i] += ev[i].elapsed_time(ev[i + 1]) / 3
        if rank == 0:
            sampler.start()
        sharding.barrier()
        ops.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            labels, selected = step(X)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms_dev = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
        launches = ops.launches()
        stage = torch.empty_like(X)
        n_e2e = max(1, min(args.steps, 3))
        for timed in (False, True):
            if timed:
                e0.record()
            for _ in range(n_e2e if timed else 1):
                stage.copy_(hostX, non_blocking=True)
                hostOut.copy_(step(stage)[0], non_blocking=True)
            if timed:
                e1.record()
            torch.cuda.synchronize()
            sharding.barrier()
        ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    peaks = load_peaks()
    value = B * world * args.steps / (ms_dev / 1e3)
    by = (4.0 * d + 8.0) * B * n * (m - 1)   # X + the running nearest-seed distance, once per seeding pass
    t_seed = stage_ms[0] / 1e3
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import mean_shift as oms
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t0 = time.perf_counter()
        ref_labels, ref_sel = oms.mean_shift_smart_init(hostX[0], kappa, m, iters, int(first[0]))
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": f"1 image, whole clusterer (oracle, fp32, {cores} threads)",
                        "parity_on_sample": {"seed_indices_equal": bool(torch.equal(selected[0].cpu(), ref_sel)),
                                             "label_agreement": float((labels[0].cpu() == ref_labels).float().mean())}}
    line = {"metric": "images/sec classical vMF mean-shift clustering (640x480x64-d embeddings: 100 farthest-point "
                      "seeds, kappa=20, 10 iterations, connected components, nearest-seed labels)",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cluster n=307200 d=64 m=100 kappa=20 iters=10 batch {B}/GPU",
                       "global_batch": B * world, "parallelism": f"replicas x{world} (batch-sharded, no collective)",
                       "l2_policy": f"inputs_exceed_l2 ({4.0 * B * n * d / 1e6:.0f} MB of embeddings per pass)"},
            "clocks": clocks,
            "e2e": {"value": B * world * n_e2e / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": hostX.numel() * 4, "d2h_bytes_per_step": hostOut.numel() * 8,
                    "ms_per_step": ms_e2e / n_e2e},
            "gpu_launches": launches,
            "stage_ms": {"select_smart_seeds": stage_ms[0], "hill_climb": stage_ms[1],
                         "connected_components": stage_ms[2], "assign_clusters": stage_ms[3]},
            "roofline": {"kernel": "smart_seeds_ring_kernel<64>", "bound": "hbm", "achieved": by / t_seed / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / t_seed / 1e9 / peaks["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes_per_step": by, "peak_source": peaks["source"],
                         "note": "read-only stream (8 B written per 264 B read); the measured peak is a copy "
                                 "(read+write) figure, which a pure read stream can exceed"},
            "cpu_baseline": cpu_baseline}
    emit(line)


def run_tail(args, rank, local_rank, world, dev, sharding, ops):
    """eval-mode tail on config #2's geometry - decoder outputs (100 queries, 120x160 mask logits)
    -> top-20 instances at 480x640 (binary masks, boxes, scores). `--batch` images per GPU (default 8). One step =
    instance_topk + instance_masks (+ finalize): 3 kernels."""
    from unseenobjectswithmeanshift_b200.meanshiftformer import instance_inference as ii
    B = args.batch
    Q, K, h, w, H, W, T = 100, 1, 120, 160, 480, 640, 20
    g = torch.Generator().manual_seed(50 + rank)
    host_logits = (2 * torch.randn(B, Q, K + 1, generator=g)).pin_memory()
    coarse = 3 * torch.randn(B, Q, h // 8, w // 8, generator=g)
    host_masks = (torch.nn.functional.interpolate(coarse, size=(h, w), mode="bicubic")
                  + 0.3 * torch.randn(B, Q, h, w, generator=g)).pin_memory()
    logits, masks = host_logits.to(dev), host_masks.to(dev)
    host_out = torch.empty(B, T, H, W).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from unseenobjectswithmeanshift_b200.graph import GraphedForward

    def step(inp):
        return ii.instance_inference_batched(inp["logits"], inp["masks"], (H, W), T)

    with torch.no_grad():
        for _ in range(args.warmup):
            step({"logits": logits, "masks": masks})
        ops.reset_stats()
        step({"logits": logits, "masks": masks})
        launches_per_step = ops.launches()
        # three small launches: replayed as one CUDA graph on static buffers (resident in HBM)
        graphed = None if args.no_graph else GraphedForward(step, {"logits": logits, "masks": masks}, warmup=2)
        run = (lambda: step({"logits": logits, "masks": masks})) if graphed is None else (lambda: graphed())
        torch.cuda.synchronize()
        if rank == 0:
            sampler.start()
        sharding.barrier()
        total = 0.0
        for _ in range(args.steps):   # outputs (197 MB) would otherwise sit in L2: flush between steps, untimed
            flush.zero_()
            e0.record()
            r = run()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        sharding.barrier()
        ms_dev = sharding.max_over_ranks(total, dev)
        launches = launches_per_step * args.steps
        r = {k: v.clone() for k, v in r.items()}
        n_e2e = max(1, min(args.steps, 5))
        sl, sm = torch.empty_like(logits), torch.empty_like(masks)
        for timed in (False, True):
            if timed:
                e0.record()
            for _ in range(n_e2e if timed else 1):
                sl.copy_(host_logits, non_blocking=True)
                sm.copy_(host_masks, non_blocking=True)
                host_out.copy_(ii.instance_inference_batched(sl, sm, (H, W), T)["pred_masks"], non_blocking=True)
            if timed:
                e1.record()
            torch.cuda.synchronize()
            sharding.barrier()
        ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    peaks = load_peaks()
    by = 4.0 * B * T * (H * W + h * w)          # kept masks written once, their low-resolution logits read once
    t = ms_dev / args.steps / 1e3
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import instance_inference as oii
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t0 = time.perf_counter()
        want = oii.inference_tail(host_logits[:2], host_masks[:2], (H, W), T)
        dt = time.perf_counter() - t0
        agree = []
        for b in range(2):
            go = torch.argsort(r["query_index"][b].cpu())
            wo = torch.argsort(want[b]["query_index"])
            agree.append(float((r["pred_masks"][b].cpu()[go] == want[b]["pred_masks"][wo]).float().mean()))
        cpu_baseline = {"value": 2.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": f"2 images (oracle: upsample all {Q} masks, then instance_inference; fp32, {cores} threads)",
                        "parity_on_sample": {"mask_pixel_agreement": min(agree)}}
    line = {"metric": "images/sec eval tail: mask upsample + instance_inference (100 queries 120x160 -> top-20 "
                      "instances at 480x640)", "value": B * world * args.steps / (ms_dev / 1e3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"tail Q=100 120x160->480x640 top-20 batch {B}/GPU", "global_batch": B * world,
                       "launch": "eager" if args.no_graph else "one CUDA graph per step",
                       "parallelism": f"replicas x{world} (batch-sharded, no collective)",
                       "l2_policy": "l2_flushed_between_steps (256 MB memset, untimed)"},
            "clocks": clocks,
            "e2e": {"value": B * world * n_e2e / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": (host_logits.numel() + host_masks.numel()) * 4,
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": ms_e2e / n_e2e},
            "gpu_launches": launches,
            "roofline": {"kernel": "instance_masks_kernel", "bound": "hbm", "achieved": by / t / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / t / 1e9 / peaks["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes_per_step": by, "peak_source": peaks["source"],
                         "note": "step time includes the top-k and finalize kernels; the reference's op sequence "
                                 "moves ~3.3 GB for the same result (983 MB upsample of all 100 masks + 5 passes)"},
            "cpu_baseline": cpu_baseline}
    emit(line)


def run_twostage(args, rank, local_rank, world, dev, sharding, ops):
    """BASELINE.json config #3: two-stage RGB-D + zoom-crop refinement, `--batch` frames per GPU (default 16 here).
    Per frame: stage-1 head (6 layers, 307200 keys) -> label map; depth filter; 5 padded ROI crops at 224x224;
    stage-2 head (8 layers) on all crops of the frame in one batch; overlap test + paste-back. Random-init weights
    segment noise, so stage 1 runs in full and is timed, but the map handed to the crop stage is a synthetic one with
    5 objects per frame (fixes the unit of work: 5 crops per frame). Embedding backbone: synthetic stand-in."""
    from unseenobjectswithmeanshift_b200 import workloads
    from unseenobjectswithmeanshift_b200.fcn import test_dataset as td
    B = args.batch if args.batch > 0 else 16
    K = 5
    model, model_crop = (m.to(dev) for m in workloads.build_two_stage_models())
    himg, hdepth, hlabels = workloads.synthetic_frames(B, objects=K, seed=rank)
    himg, hdepth = himg.pin_memory(), hdepth.pin_memory()
    img, depth, labels = himg.to(dev), hdepth.to(dev), hlabels.to(dev)
    host_out = torch.empty(B, 480, 640).pin_memory()
    kw = dict(topk=False, score=0.7, low_threshold=0.4)
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def frame(im, dp, lab):
        model.label_maps([{"image": im[0], "depth": dp[0]}], **kw)          # stage 1 (result replaced, see above)
        out_label = td.filter_labels_depth(lab, dp, 0.5)
        rgb_crop, mask_crop, rois, depth_crop = td.crop_rois(im, out_label.clone(), dp, crop_size=224)
        labels_crop, _ = model_crop.label_maps([{"image": rgb_crop, "depth": depth_crop}], **kw)
        refined, _ = td.match_label_crop(out_label, labels_crop, mask_crop, rois, depth_crop)
        return refined, rgb_crop.shape[0]

    def step(im, dp):
        outs, crops = [], 0
        for f in range(B):
            r, n = frame(im[f:f + 1], dp[f:f + 1], labels[f:f + 1])
            outs.append(r)
            crops += n
        return outs, crops

    with torch.no_grad():
        for _ in range(args.warmup):
            step(img, depth)
        ops.reset_stats()
        _, crops = step(img, depth)
        launches_per_step = ops.launches()
        torch.cuda.synchronize()
        if rank == 0:
            sampler.start()
        sharding.barrier()
        e0.record()
        for _ in range(args.steps):   # every frame streams a 79 MB embedding map and 315 MB of mask features
            step(img, depth)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms_dev = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
        n_e2e = max(1, min(args.steps, 3))
        si, sd_ = torch.empty_like(img), torch.empty_like(depth)
        for timed in (False, True):
            if timed:
                e0.record()
            for _ in range(n_e2e if timed else 1):
                si.copy_(himg, non_blocking=True)
                sd_.copy_(hdepth, non_blocking=True)
                outs, _ = step(si, sd_)
                host_out.copy_(torch.cat(outs), non_blocking=True)
            if timed:
                e1.record()
            torch.cuda.synchronize()
            sharding.barrier()
        ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    emit({"metric": "frames/sec two-stage RGB-D segmentation 640x480 (stage-1 head + 5 zoom-crops through the crop head "
                    "+ paste-back)", "value": B * world * args.steps / (ms_dev / 1e3), "unit": "images/s",
          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": f"twostage 640x480 batch {B}/GPU, {crops} crops of 224x224 per step, stage 1: 6 layers, "
                                 f"stage 2: 8 layers", "global_batch": B * world, "launch": "eager, one frame per call",
                     "parallelism": f"replicas x{world} (batch-sharded, no collective)",
                     "l2_policy": "inputs_exceed_l2 (315 MB of mask features per frame)",
                     "backbone": "synthetic 1x1 stand-in: the UCN embedding network is outside the hot path (SURVEY.md 8)"},
          "clocks": clocks,
          "e2e": {"value": B * world * n_e2e / (ms_e2e / 1e3), "unit": "images/s",
                  "h2d_bytes_per_step": (himg.numel() + hdepth.numel()) * 4, "d2h_bytes_per_step": host_out.numel() * 4,
                  "ms_per_step": ms_e2e / n_e2e},
          "gpu_launches": launches_per_step * args.steps, "roofline": None, "cpu_baseline": None,
          "note": "auxiliary workload (config #3 end to end); per-kernel rooflines are on the ucn / crop lines"})


def run_train(args, rank, local_rank, world, dev, sharding, ops):
    """BASELINE.json config #5: training step of the R50-config head - forward, deep-supervision
    losses (one matcher synchronisation), backward (native vMF attention / MSDeformAttn backward kernels, cuBLAS for
    the dense layers), DDP gradient all-reduce over NCCL when world > 1, full-model clipping, fused AdamW. fp32 by
    default; `--amp bf16|fp16` puts the head under autocast like the reference's trainer (the custom kernels and the
    pixel decoder stay fp32; fp16 without a GradScaler is for timing only). `--batch` images per GPU."""
    from unseenobjectswithmeanshift_b200 import training, workloads
    B = args.batch
    amp = {"off": None, "bf16": torch.bfloat16, "fp16": torch.float16}[args.amp]
    model = workloads.build_trainer("r50", amp_dtype=amp).to(dev)
    ddp = training.wrap_ddp(model, local_rank)
    opt = training.build_optimizer(model)
    host_feats = workloads.synthetic_features("r50", B, seed=rank, pin=True)
    feats = {k: v.to(dev) for k, v in host_feats.items()}
    targets = [{k: v.to(dev) for k, v in t.items()} for t in workloads.synthetic_targets("r50", B, seed=rank)]
    stage = {k: torch.empty_like(v) for k, v in feats.items()}
    host_loss = torch.empty(1).pin_memory()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.manual_seed(1234 + rank)
    for _ in range(args.warmup):
        training.train_step(ddp, opt, {"features": feats, "targets": targets})
    ops.reset_stats()
    training.train_step(ddp, opt, {"features": feats, "targets": targets})
    launches_per_step = ops.launches()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    sharding.barrier()
    e0.record()
    for _ in range(args.steps):   # 295 MB of features + every layer's masks and gradients per step: far beyond L2
        losses = training.train_step(ddp, opt, {"features": feats, "targets": targets})
    e1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    ms_dev = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    n_e2e = max(1, min(args.steps, 5))
    for timed in (False, True):
        if timed:
            e0.record()
        for _ in range(n_e2e if timed else 1):
            for k in stage:
                stage[k].copy_(host_feats[k], non_blocking=True)
            out = training.train_step(ddp, opt, {"features": stage, "targets": targets})
            host_loss.copy_(sum(out.values()).reshape(1), non_blocking=True)
        if timed:
            e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
    ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    total_loss = float(sum(losses.values()))
    assert total_loss == total_loss, "training loss is NaN"
    emit({"metric": "images/sec MSMFormer head training step 640x480 (R50 config: forward + deep-supervision losses + "
                    "backward + clipped AdamW)", "value": B * world * args.steps / (ms_dev / 1e3), "unit": "images/s",
          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32" if amp is None else f"{args.amp} autocast (dense layers), f32 kernels", "data": "synthetic",
          "config": {"workload": f"train r50-head 640x480 batch {B}/GPU, 100 queries, 9 decoder layers, 5 instances/image",
                     "global_batch": B * world,
                     "parallelism": f"ddp x{world} (gradient all-reduce over NCCL)" if world > 1 else "single GPU",
                     "l2_policy": "inputs_exceed_l2 (295 MB of backbone features per step)",
                     "backbone": "excluded: the cuDNN backbone is outside the hot path (SURVEY.md 8)"},
          "clocks": clocks,
          "e2e": {"value": B * world * n_e2e / (ms_e2e / 1e3), "unit": "images/s",
                  "h2d_bytes_per_step": sum(v.numel() for v in host_feats.values()) * 4, "d2h_bytes_per_step": 4,
                  "ms_per_step": ms_e2e / n_e2e},
          "gpu_launches": launches_per_step * args.steps, "final_loss": total_loss,
          "roofline": None, "cpu_baseline": None,
          "note": "auxiliary workload (config #5); kernels not yet profiled - no roofline claim"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="r50", choices=["r50", "demo", "r50-head", "ucn", "crop", "meanshift", "cluster",
                                                          "tail", "train", "twostage"],
                    help="r50 (default) = BASELINE.json configs[1], whole model; demo = configs[0] (one RGB-D frame, whole "
                         "model); r50-head / ucn / crop = the head alone on backbone features; the rest: configs[2..4] and "
                         "the SURVEY 8(f) rows")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0 = the workload's default)")
    ap.add_argument("--tail", default="label_map", choices=["label_map", "instances"],
                    help="whole-model workloads: what the step returns (fused get_confident_instances + combine_masks "
                         "label map, or the Instances fields incl. full-resolution masks)")
    ap.add_argument("--skip-profile", action="store_true", help="leave out the CUPTI kernel-share pass")
    ap.add_argument("--backbone-fp32", action="store_true",
                    help="whole-model workloads: strict fp32 cuDNN math for the BACKBONE (the reference's CPU arithmetic). "
                         "Default: PyTorch's own default conv math on a GPU (TF32), i.e. what the reference's stock code "
                         "does there; the head and the tail never use TF32 either way")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=3,
                    help="forward workloads: graph replays in flight at once (each with its own static buffers, on its own "
                         "stream); 1 = one step after another")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--vmf-tflops", action="store_true", help="also time the attention core alone (default with the CPU baseline)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave out the host-buffer leg")
    ap.add_argument("--amp", default="off", choices=["off", "bf16", "fp16"], help="--workload train: autocast dtype")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    from unseenobjectswithmeanshift_b200 import sharding
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return

    import __graft_entry__
    __graft_entry__.build()
    from unseenobjectswithmeanshift_b200 import ops, workloads

    rank, local_rank, world = sharding.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    args.numa = sharding.pin_to_gpu_numa(local_rank) if world > 1 else "single process: not pinned"
    kind = args.workload
    if args.batch <= 0 and kind in ("train", "twostage", "tail"):
        args.batch = 16 if kind == "twostage" else PER_GPU_BATCH
    if kind == "meanshift":
        run_meanshift(args, rank, local_rank, world, dev, sharding, ops)
        return
    if kind == "cluster":
        run_cluster(args, rank, local_rank, world, dev, sharding, ops)
        return
    if kind == "tail":
        run_tail(args, rank, local_rank, world, dev, sharding, ops)
        return
    if kind == "train":
        run_train(args, rank, local_rank, world, dev, sharding, ops)
        return
    if kind == "twostage":
        run_twostage(args, rank, local_rank, world, dev, sharding, ops)
        return

    run_forward(args, rank, local_rank, world, dev, sharding, ops, workloads)


# ------------------------------------------------------------------------------------------------------------------
# forward workloads: whole model at the reference's API boundary (r50 = BASELINE.json configs[1], demo = configs[0]) or
# the head alone on backbone features (r50-head, ucn, crop)
# ------------------------------------------------------------------------------------------------------------------
FULL_MODEL = ("r50", "demo")
HEAD_KIND = {"r50": "r50", "demo": "ucn", "r50-head": "r50", "ucn": "ucn", "crop": "crop"}
METRICS = {
    "r50": "images/sec MSMFormer forward 640x480 (R50 config: ResNet-50 backbone + MSDeformAttn pixel decoder + 9-layer "
           "mean-shift decoder, 100 queries + instance tail)",
    "demo": "images/sec MSMFormer forward 640x480 RGB-D single frame (UCN config: SEGNET RGB-D embedding + "
            "SimpleBasePixelDecoder + 6-layer pretrained mean-shift decoder at full resolution + instance tail)",
    "r50-head": METRIC,
    "ucn": "images/sec MSMFormer head forward 640x480 (UCN RGB-D config: SimpleBasePixelDecoder + 6-layer "
           "pretrained mean-shift decoder on the full-resolution 64-d embedding, 100 queries)",
    "crop": "crops/sec MSMFormer head forward 224x224 (crop config: SimpleBasePixelDecoder + 8-layer "
            "pretrained mean-shift decoder, 100 queries)"}


def _bytes(tensors):
    return sum(v.numel() * v.element_size() for v in tensors.values())


def _kernel_shares(fn, top=12):
    """Device time per kernel NAME over one eager pass of `fn`, from CUPTI activity records (torch.profiler): exact
    kernel durations, no launch gaps and no duration filter. -> (total kernel ms, [(name, launches, ms, share)])."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        agg = {}
        for ev in prof.events():
            if getattr(ev, "device_type", None) is not None and "cuda" in str(ev.device_type).lower():
                name = ev.name
                if name.startswith("Memcpy") or name.startswith("Memset"):
                    continue
                c, t = agg.get(name, (0, 0.0))
                agg[name] = (c + 1, t + ev.device_time_total / 1e3 if hasattr(ev, "device_time_total") else t)
        total = sum(t for _, t in agg.values())
        if total <= 0:
            return None
        rows = sorted(((n, c, t, t / total) for n, (c, t) in agg.items()), key=lambda r: -r[2])[:top]
        return total, rows
    except Exception as e:  # profiling is evidence, not the measurement: never fail the bench on it
        print(f"[bench] kernel shares unavailable: {e!r}", file=sys.stderr)
        return None


def _short(name):
    name = name.replace("void ", "").replace("msm::", "")
    return name[:96]


def run_forward(args, rank, local_rank, world, dev, sharding, ops, workloads):
    from unseenobjectswithmeanshift_b200.graph import GraphedForward
    kind = args.workload
    full = kind in FULL_MODEL
    hk = HEAD_KIND[kind]
    B = args.batch if args.batch > 0 else (1 if kind == "demo" else PER_GPU_BATCH)
    H, W = workloads.HEAD_CFG[hk]["height"], workloads.HEAD_CFG[hk]["width"]

    if full:
        from unseenobjectswithmeanshift_b200 import backbones
        backbones.set_tf32(not args.backbone_fp32)
        model = workloads.build_model(kind).to(dev)
        head = model.sem_seg_head
        host_in = workloads.synthetic_images(kind, B, seed=rank, pin=True)

        if args.tail == "instances":
            def step(inp):   # forward(batched_inputs) of the reference: Instances fields of every image
                outputs, _, padded, _ = model._head_outputs([inp])
                from unseenobjectswithmeanshift_b200.meanshiftformer import instance_inference as ii
                r = ii.instance_inference_batched(outputs["pred_logits"], outputs["pred_masks"], padded,
                                                  model.test_topk_per_image)
                return {k: r[k] for k in ("pred_masks", "pred_boxes", "scores", "pred_classes")}
        else:
            def step(inp):   # what the UOIS scripts consume: get_confident_instances + combine_masks, fused on the device
                label_map, f = model.label_maps([inp])
                return {"label_map": label_map, "scores": f["scores"], "pred_classes": f["pred_classes"],
                        "pred_boxes": f["pred_boxes"], "instance_label": f["instance_label"]}
    else:
        model = None
        head = workloads.build_head(hk).to(dev)
        host_in = workloads.synthetic_features(hk, B, seed=rank, pin=True)

        def step(feats):
            out, _ = head(feats, H, W)
            return {"pred_logits": out["pred_logits"], "pred_masks": out["pred_masks"],
                    "aux_pred_masks": [a["pred_masks"] for a in out["aux_outputs"]]}
    dev_in = {k: v.to(dev) for k, v in host_in.items()}
    out_keys = [k for k in ("label_map", "scores", "pred_classes", "pred_boxes", "instance_label", "pred_masks",
                            "pred_logits")]
    sampler = ClockSampler(local_rank)
    parts_ms, shares = None, None
    with torch.no_grad():
        # ---------------- eager pass: launch count and per-op device time (CUDA events around every library call)
        for _ in range(args.warmup):
            step(dev_in)
        torch.cuda.synchronize()
        ops.reset_stats(timing=True)
        n_eager = 3
        for _ in range(n_eager):
            # park the GPU for ~40 ms so the host enqueues the whole step ahead of it: the kernels then run
            # back to back and the event pairs measure kernel time, not launch gaps
            torch.cuda._sleep(int(0.04 * 1.9e9))
            step(dev_in)
            torch.cuda.synchronize()
        launches_per_step = ops.launches() // n_eager
        op_ms = {k: (c / n_eager, t / n_eager) for k, (c, t) in ops.op_times_ms().items()}
        op_groups = ops.op_groups()
        ops.reset_stats(timing=False)
        if full:   # where the step's time goes: backbone (cuDNN, outside the hot path) | head | tail, eager + events
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            acc = [0.0, 0.0, 0.0]
            for _ in range(5):
                torch.cuda._sleep(int(0.04 * 1.9e9))
                ev[0].record()
                if model.use_other_backbone:
                    feats = model.pretrained_backbone(dev_in["image"])
                else:
                    emb = model.pretrained_backbone(dev_in["image"], None, dev_in.get("depth"))
                    feats = {"res5": torch.nn.functional.normalize(emb, p=2, dim=1)}
                ev[1].record()
                outputs, _ = head(feats, H, W)
                ev[2].record()
                from unseenobjectswithmeanshift_b200.fcn import test_utils as tu
                tu.label_map_from_outputs(outputs["pred_logits"], outputs["pred_masks"], (H, W), 20)
                ev[3].record()
                torch.cuda.synchronize()
                for i in range(3):
                    acc[i] += ev[i].elapsed_time(ev[i + 1]) / 5
            parts_ms = {"backbone_cudnn": acc[0], "head": acc[1], "tail": acc[2], "how": "eager, CUDA events"}

        # ---------------- the step as ONE CUDA graph (static input / output buffers resident in HBM)
        # `--inflight F` (default 3): F graphs with their own static buffers, replayed round-robin on F streams, so
        # that step i+1 starts while step i is still in its latency-bound phases (the 800-row decoder launches fill 64
        # of 148 SMs). Every step is still one whole batch through the whole path; `serial` below is F = 1.
        n_fl = 1 if args.no_graph else max(1, args.inflight)
        graphs = [] if args.no_graph else [GraphedForward(step, dev_in, warmup=2) for _ in range(n_fl)]
        graphed = graphs[0] if graphs else None
        s_main = torch.cuda.current_stream()
        lanes = [s_main] if n_fl == 1 else [torch.cuda.Stream() for _ in range(n_fl)]
        runner = (lambda inp=None: step(dev_in)) if graphed is None else (lambda inp=None: graphed(inp))

        def replay_steps(n, nf):
            """n steps, nf of them in flight; brackets the work on the current stream."""
            if nf == 1:
                for _ in range(n):
                    runner()
                return
            fork = torch.cuda.Event()
            fork.record(s_main)
            for i in range(n):
                with torch.cuda.stream(lanes[i % nf]):
                    if i < nf:
                        lanes[i].wait_event(fork)
                    graphs[i % nf]()
            for ln in lanes[:nf]:
                s_main.wait_stream(ln)

        for _ in range(args.warmup):
            runner()
        replay_steps(2 * n_fl, n_fl)
        torch.cuda.synchronize()
        if rank == 0:
            sampler.start()
        sharding.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms_serial = None
        if n_fl > 1:   # the same steps one after another (step latency), reported beside the pipelined figure
            n_ser = max(3, args.steps // 4)
            torch.cuda.synchronize()
            e0.record()
            replay_steps(n_ser, 1)
            e1.record()
            torch.cuda.synchronize()
            ms_serial = sharding.max_over_ranks(e0.elapsed_time(e1), dev) / n_ser
        torch.cuda.synchronize()
        sharding.barrier()
        e0.record()
        replay_steps(args.steps, n_fl)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms_dev = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
        launches = launches_per_step * args.steps
        # the graphs in flight were fed the same batch: their results must be bit-identical (a replay racing with
        # another one on shared state would show here; tests/test_gpu_config2.py holds the stricter form)
        inflight_ok = None
        if n_fl > 1:
            inflight_ok = all(torch.equal(g.static_out[k], graphs[0].static_out[k])
                              for g in graphs[1:] for k in graphs[0].static_out if torch.is_tensor(graphs[0].static_out[k]))
            if not inflight_ok:
                raise RuntimeError("graphs in flight disagree on identical inputs: concurrent replays share state")

        # ---------------- end to end: pinned host inputs in, results out, every step.
        # Three streams: H2D of step i+1 and D2H of step i-1 overlap the graph replay of step i (PCIe is full
        # duplex); device-side staging copies decouple the graph's static buffers from the transfers.
        out0 = runner()
        out_keys = [k for k in out_keys if k in out0]
        host_out = {k: torch.empty(out0[k].shape, dtype=out0[k].dtype).pin_memory() for k in out_keys}
        h2d, d2h = _bytes(host_in), _bytes(host_out)
        static_ins = [g.static_in for g in graphs] if graphs else [dev_in]
        stage_in = [{k: torch.empty_like(v) for k, v in static_ins[0].items()} for _ in range(n_fl)]
        stage_out = [{k: torch.empty_like(out0[k]) for k in out_keys} for _ in range(n_fl)]
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in_ready, ev_in_free = [torch.cuda.Event() for _ in range(n_fl)], [torch.cuda.Event() for _ in range(n_fl)]
        ev_out_ready, ev_out_free = [torch.cuda.Event() for _ in range(n_fl)], [torch.cuda.Event() for _ in range(n_fl)]

        def e2e_steps(n):
            fork = torch.cuda.Event()
            fork.record(s_main)
            for f in range(n_fl):
                lanes[f].wait_event(fork)
                ev_in_free[f].record(lanes[f])
                s_d2h.wait_event(fork)
                ev_out_free[f].record(s_d2h)
            s_h2d.wait_event(fork)
            for i in range(n):
                f = i % n_fl
                lane = lanes[f]
                with torch.cuda.stream(s_h2d):
                    s_h2d.wait_event(ev_in_free[f])         # previous contents of stage_in[f] consumed
                    for k in stage_in[f]:
                        stage_in[f][k].copy_(host_in[k], non_blocking=True)
                    ev_in_ready[f].record(s_h2d)
                with torch.cuda.stream(lane):
                    lane.wait_event(ev_in_ready[f])
                    for k in static_ins[f]:
                        static_ins[f][k].copy_(stage_in[f][k], non_blocking=True)
                    ev_in_free[f].record(lane)
                    out = graphs[f]() if graphs else step(dev_in)
                    lane.wait_event(ev_out_free[f])         # previous contents of stage_out[f] are on the host
                    for k in out_keys:
                        stage_out[f][k].copy_(out[k], non_blocking=True)
                    ev_out_ready[f].record(lane)
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_out_ready[f])
                    for k in out_keys:
                        host_out[k].copy_(stage_out[f][k], non_blocking=True)
                    ev_out_free[f].record(s_d2h)
            for ln in lanes:
                s_main.wait_stream(ln)
            s_main.wait_stream(s_h2d)
            s_main.wait_stream(s_d2h)

        if not args.skip_e2e:
            e2e_steps(3)
        torch.cuda.synchronize()
        sharding.barrier()
        e0.record()
        e2e_steps(1 if args.skip_e2e else args.steps)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None

    total_images = B * world * args.steps
    value = total_images / (ms_dev / 1e3)
    e2e_value = None if args.skip_e2e else total_images / (ms_e2e / 1e3)
    if rank != 0:
        return

    # ---------------- roofline. (1) `roofline`: the (kernel, shape) group of THIS library with the largest device
    # time per step - every group is a candidate (no duration filter); durations are CUDA events around each library
    # call on the launching stream (GPU parked first, so the pairs bracket kernel time), and the CUPTI kernel list of
    # the same step (`kernel_shares`) is printed beside it so the choice can be checked against exact kernel times.
    # (2) `roofline.step`: the whole step against both roofs - algorithmic bytes and 3 x FLOP (tensor passes of the
    # split-precision products) of all library launches over ms_per_step.
    peaks = load_peaks()
    ridge = peaks["bf16_tflops"] * 1e3 / peaks["hbm_gbs"]
    roofline = None
    step_ms = ms_dev / args.steps
    if op_groups:
        (tag, sig), g = max(op_groups.items(), key=lambda kv: kv[1]["ms"])
        per = g["count"] / n_eager
        avg_ms = g["ms"] / g["count"]
        gbs = g["bytes"] / g["ms"] / 1e6
        tfs = g["flops"] / g["ms"] / 1e9
        tensor_bound = g["bytes"] > 0 and 3.0 * g["flops"] / g["bytes"] > ridge
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of these
        # kernels at these shapes (profiles/): a constant of an earlier run, NOT measured by this process
        # (profiles/r02_ncu_kernels.md, profiles/r02_ncu_operand_images.md)
        traffic_from_profile = {("linear", "M800 N256 K256"): 1.16e6,
                                ("mask_logits", "B8 Q100 C256 HW19200"): 185.1e6,
                                ("ms_deform_attn_forward", "fused N8 S6300 M8 D8 Lq6300"): 75.9e6,
                                ("ffn", "ffn+LN M50400 D64 F1024"): 13.5e6,
                                ("linear", "M50400 N288 K64"): 16.9e6,
                                ("linear", "K-images M307200 N1536 K64"): 1919.2e6,
                                ("linear", "V-images M307200 N1536 K64"): 1908.7e6}
        roofline = {"kernel": tag, "shape": sig, "launches_per_step": per, "avg_launch_ms": avg_ms,
                    "share_of_step": (g["ms"] / n_eager) / step_ms,
                    "algorithmic_bytes_per_launch": g["bytes"] / g["count"],
                    "algorithmic_flops_per_launch": g["flops"] / g["count"], "traffic": None,
                    "traffic_from_profile": traffic_from_profile.get((tag, sig)),
                    "timing": "CUDA events around each library call, eager pass of the same step (same stream); "
                              "largest group by time, no duration filter",
                    "peak_source": peaks["source"]}
        if tensor_bound:
            roofline.update({"bound": "tensor", "achieved": 3.0 * tfs, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": 3.0 * tfs / peaks["bf16_tflops"],
                             "note": "16-bit tensor passes issued (3 per fp32-grade product)"})
        else:
            roofline.update({"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": gbs / peaks["hbm_gbs"]})
        top = sorted(op_groups.items(), key=lambda kv: -kv[1]["ms"])[:10]
        roofline["top_groups"] = [{"kernel": t, "shape": sg, "launches_per_step": v["count"] / n_eager,
                                   "ms_per_step": v["ms"] / n_eager,
                                   "GBps": v["bytes"] / v["ms"] / 1e6 if v["ms"] else None,
                                   "TFLOPps": v["flops"] / v["ms"] / 1e9 if v["ms"] else None} for (t, sg), v in top]
        lib_bytes = sum(v["bytes"] for v in op_groups.values()) / n_eager
        lib_flops = sum(v["flops"] for v in op_groups.values()) / n_eager
        lib_ms = sum(v["ms"] for v in op_groups.values()) / n_eager
        bb_flops = workloads.backbone_flops_per_image(kind) * B if full else 0.0
        roofline["step"] = {
            "ms_per_step": step_ms, "library_ms_per_step_eager_events": lib_ms,
            "library_algorithmic_bytes_per_step": lib_bytes, "library_flops_per_step": lib_flops,
            "backbone_flops_per_step": bb_flops,
            "hbm_GBps": lib_bytes / step_ms / 1e6, "hbm_frac": lib_bytes / step_ms / 1e6 / peaks["hbm_gbs"],
            "tensor_TFLOPps_x3": 3.0 * lib_flops / step_ms / 1e9,
            "tensor_frac": 3.0 * lib_flops / step_ms / 1e9 / peaks["bf16_tflops"],
            "note": "library launches only (the cuDNN backbone is fp32 CUDA-core work); a step far below both roofs "
                    "is latency / launch bound"}
    op_summary = {k: {"calls_per_step": c, "ms_per_step": t} for k, (c, t) in op_ms.items()}

    # ---------------- vMF attention TFLOP/s (second half of BASELINE.json's metric): the attention core alone at the
    # full-resolution key grid of the UCN config (1 image, 8 heads, 100 queries, 307200 keys, hd 32, bit mask), K / V
    # as the operand images the decoder's projections write, L2 flushed between launches, CUDA events
    vmf = None
    if not args.no_cpu_baseline or args.vmf_tflops:
        with torch.no_grad():
            gq = torch.Generator(device=dev).manual_seed(5)
            Sk, Hh, Qn, hd = 307200, 8, 100, 32
            q = torch.randn(1, Qn, Hh * hd, device=dev, generator=gq)
            k = torch.randn(1, Sk, Hh * hd, device=dev, generator=gq)
            v = torch.randn(1, Sk, Hh * hd, device=dev, generator=gq)
            bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (1, Qn, Sk // 32), device=dev, dtype=torch.int32)
            ro = torch.ones(1, Qn, device=dev, dtype=torch.int32)
            hv = lambda t: t.unflatten(-1, (Hh, hd)).permute(0, 2, 1, 3)  # noqa: E731
            kvp = ops.pack_kv(hv(k), hv(v))
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            for _ in range(3):
                ops.vmf_attention_packed(hv(q), kvp, blocked_bits=bits, row_open=ro)
            ts = []
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.vmf_attention_packed(hv(q), kvp, blocked_bits=bits, row_open=ro)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t_ms = statistics.median(ts)
            fl = 4.0 * Hh * Qn * Sk * hd
            vmf = {"shape": f"B1 H{Hh} Q{Qn} S{Sk} hd{hd} masked, packed K/V images (TMA bulk stream)", "ms": t_ms,
                   "tflops_useful": fl / t_ms / 1e9, "tflops_tensor_passes": 3 * fl / t_ms / 1e9,
                   "frac_of_bf16_peak": 3 * fl / t_ms / 1e9 / peaks["bf16_tflops"],
                   "GBps": 4.0 * 2 * Sk * Hh * hd / t_ms / 1e6,
                   "frac_of_hbm_peak": 4.0 * 2 * Sk * Hh * hd / t_ms / 1e6 / peaks["hbm_gbs"]}
            del q, k, v, bits, flush, kvp

    # ---------------- exact kernel durations of one eager step (CUPTI). LAST of the GPU measurements: kineto leaves
    # CUPTI attached to the process, so nothing timed follows this pass
    if not args.skip_profile:
        with torch.no_grad():
            shares = _kernel_shares(lambda: step(dev_in))
    if shares is not None:
        roofline = roofline or {}
        roofline["kernel_shares"] = {"source": "CUPTI kernel records of one eager step (torch.profiler), taken after "
                                               "all timed regions",
                                     "kernel_ms_per_step": shares[0],
                                     "top": [{"kernel": _short(n), "launches": c, "ms": t, "share": sh}
                                             for n, c, t, sh in shares[1]]}

    # ---------------- CPU baseline on a bounded sample, rank 0, N == 1 only: the REFERENCE's own modules when the
    # vendored copy is present (baseline/_ref), else the oracle port
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(kind, model if full else head, host_in, step, dev, full)

    hk_cfg = workloads.HEAD_CFG[hk]
    config = {"workload": (f"{kind} 640x480 batch {B}/GPU: " if kind != "crop" else f"crop 224x224 batch {B}/GPU: ")
                          + ("whole model, images in -> label maps + instance fields out" if full and args.tail != "instances"
                             else "whole model, images in -> Instances fields out" if full
                             else "segmentation head on backbone features"),
              "queries": 100, "decoder_layers": hk_cfg["dec_layers"],
              "launch": "eager" if args.no_graph else "one CUDA graph per step" + (
                  f", {n_fl} steps in flight on {n_fl} streams (each graph has its own static buffers)" if n_fl > 1 else ""),
              "inflight": n_fl, "inflight_results_identical": inflight_ok,
              "global_batch": B * world, "parallelism": f"replicas x{world} (batch-sharded, no collective)",
              "numa": args.numa,
              "l2_policy": "inputs_exceed_l2 (every decoder layer streams the mask features - 157 MB at batch 8 - and "
                           "writes a fresh logits tensor; backbone activations exceed L2)",
              "backbone": (f"included: torchvision ResNet-50, cuDNN {'fp32 (TF32 off, NCHW)' if args.backbone_fp32 else 'at PyTorch default conv math (TF32), channels_last'}, channels_last" if kind == "r50"
                           else f"included: SEGNET RGB-D (two ResNet34-8s streams), cuDNN {'fp32 (TF32 off, NCHW)' if args.backbone_fp32 else 'at PyTorch default conv math (TF32), channels_last'}" if kind == "demo"
                           else "excluded: head-only workload on synthetic backbone features"),
              "gflop_per_image": {"head": workloads.head_flops_per_image(hk) / 1e9,
                                  "backbone": workloads.backbone_flops_per_image(kind) / 1e9 if full else 0.0}}
    line = {"metric": METRICS[kind], "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "pinned host inputs -> device every step, results -> pinned host every step; transfers of "
                           "neighbouring steps overlap the graph replay on separate streams"},
            "serial": None if ms_serial is None else {"ms_per_step": ms_serial, "value": B * world / (ms_serial / 1e3),
                                                       "note": "the same graph replayed one step after another "
                                                               "(step latency)"},
            "gpu_launches": launches, "parts_ms": parts_ms,
            "roofline": roofline, "vmf_attention": vmf, "op_ms": op_summary, "cpu_baseline": cpu_baseline}
    emit(line)


def _cpu_head(kind_head, sd):
    """-> (callable(features) -> (outputs, mask_features), kind): the reference's own head modules when available."""
    from unseenobjectswithmeanshift_b200 import workloads
    try:
        from oracle import ref_models
        if ref_models.reference_root() is not None:
            ref_head = ref_models.build_reference_head(kind_head, sd)
            return (lambda feats: ref_head(feats, workloads.HEAD_CFG[kind_head]["height"],
                                           workloads.HEAD_CFG[kind_head]["width"])), "reference"
    except Exception as e:
        print(f"[bench] reference modules unavailable ({e!r}); timing the oracle port", file=sys.stderr)
    from oracle import head as ohead
    return (lambda feats: ohead.head_forward(sd, feats, **workloads.oracle_kwargs(kind_head))), "port"


def _cpu_model(kind, full, seed=0):
    """CPU arm for workload `kind`: -> (step(inputs) -> outputs dict with pred_masks, kind_of_baseline, note)."""
    from oracle import head as ohead
    from unseenobjectswithmeanshift_b200 import workloads
    hk = HEAD_KIND[kind]
    cfg = workloads.HEAD_CFG[hk]
    if full:
        model = workloads.build_model(kind, seed)   # same seeds -> same weights as the GPU arm
        from unseenobjectswithmeanshift_b200 import backbones
        bb_name = workloads.MODEL_CFG[kind]["backbone"]
        backbone = (backbones.ResNet50Features(seed=seed, fold_bn=False) if bb_name == "ResNet50Features"
                    else model.pretrained_backbone)   # plain eval-mode BatchNorm on the CPU side
        sd = {k: v.detach() for k, v in model.sem_seg_head.state_dict().items()}
        head_fn, which = _cpu_head(hk, sd)

        def step(inp):
            if workloads.MODEL_CFG[kind]["use_other_backbone"]:
                feats = backbone(inp["image"])
            else:
                feats = {"res5": torch.nn.functional.normalize(backbone(inp["image"], None, inp.get("depth")), p=2, dim=1)}
            out, _ = head_fn(feats)
            ohead.eval_tail(out, (cfg["height"], cfg["width"]), 2, 20)   # upsample + instance_inference (restated)
            return out, feats
        note = ("reference modules (head) + the same torchvision/torch backbone on CPU + restated eval tail"
                if which == "reference" else "oracle port of the head + backbone + restated eval tail")
        return step, which, note
    head = workloads.build_head(hk, seed)
    sd = {k: v.detach() for k, v in head.state_dict().items()}
    head_fn, which = _cpu_head(hk, sd)
    return (lambda feats: (head_fn(feats)[0], feats)), which, (
        "reference modules imported from baseline/_ref" if which == "reference" else "oracle port")


def cpu_baseline_leg(kind, gpu_module, host_in, gpu_step, dev, full):
    from unseenobjectswithmeanshift_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hk = HEAD_KIND[kind]
    sb = min(2, next(iter(host_in.values())).shape[0])
    sample = {k: v[:sb].clone() for k, v in host_in.items()}
    cpu_step, which, note = _cpu_model(kind, full)
    with torch.no_grad():
        t0 = time.perf_counter()
        ref, feats = cpu_step(sample)
        dt = time.perf_counter() - t0
        # parity of the hot path on the sample: the GPU head on the SAME backbone features the CPU arm saw (with the
        # auxiliary full-resolution masks the timed eval path skips, so that the first prediction can be compared too)
        head = gpu_module.sem_seg_head if full else gpu_module
        predictor = getattr(head, "predictor", None)
        keep = getattr(predictor, "eval_aux_masks", True)
        if predictor is not None:
            predictor.eval_aux_masks = True
        got, _ = head({k: v.to(dev) for k, v in feats.items()}, workloads.HEAD_CFG[hk]["height"],
                      workloads.HEAD_CFG[hk]["width"])
        if predictor is not None:
            predictor.eval_aux_masks = keep
    pk = ref["pred_masks"].abs().max().item()
    err = (got["pred_masks"].cpu() - ref["pred_masks"]).abs().max().item() / pk
    agree = (got["pred_masks"].cpu().argmax(1) == ref["pred_masks"].argmax(1)).float().mean().item()
    first = (got["aux_outputs"][0]["pred_masks"].cpu() - ref["aux_outputs"][0]["pred_masks"]).abs().max().item() / \
        ref["aux_outputs"][0]["pred_masks"].abs().max().item()
    return {"value": sb / dt, "unit": "images/s", "cores": cores, "kind": which,
            "sample": f"1 run x {sb} image(s) of the same workload ({note}; fp32, {cores} threads)",
            "parity_on_sample": {"pred_masks_max_err_rel_to_peak": err, "argmax_label_agreement": agree,
                                 "first_prediction_max_err_rel_to_peak": first,
                                 "note": "GPU head vs the CPU arm on identical backbone features; later layers "
                                         "inherit mask-bit flips (DESIGN.md section 2)"}}


if __name__ == "__main__":
    main()
