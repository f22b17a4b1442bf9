#!/usr/bin/env bash
# round-2 call G: new tests, MSDA reference-kernel comparison, whole-model bench fp32 / TF32 backbone, training step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python tools/bench_msda_ref.py 2>&1 | tee gpurun_out/r2g_msda_ref.md | tail -12
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_r50.json 2>gpurun_out/r2g_bench_r50.err; echo rc=$?; tail -3 gpurun_out/r2g_bench_r50.err; cut -c1-330 gpurun_out/r2g_bench_r50.json
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --backbone-tf32 --skip-profile > gpurun_out/r2g_bench_r50_tf32bb.json 2>/dev/null; cut -c1-330 gpurun_out/r2g_bench_r50_tf32bb.json
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r2g_bench_train.json 2>gpurun_out/r2g_bench_train.err; echo rc=$?; cut -c1-330 gpurun_out/r2g_bench_train.json
timeout 600 python bench.py --workload train --amp bf16 --steps 10 --warmup 3 > gpurun_out/r2g_bench_train_bf16.json 2>gpurun_out/r2g_bench_train_bf16.err; echo rc=$?; cut -c1-330 gpurun_out/r2g_bench_train_bf16.json
timeout 300 python bench.py --workload demo --steps 30 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2g_bench_demo.json 2>/dev/null; cut -c1-330 gpurun_out/r2g_bench_demo.json
