#!/usr/bin/env bash
# round-2 call H: persistent mean-shift, backbone cuDNN variants, full suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "mean_shift or clusterer" 2>&1 | tail -6
MSM_MS_PERSISTENT=0 timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_ms_launches.json 2>/dev/null; cut -c1-300 gpurun_out/r2h_bench_ms_launches.json
timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 > gpurun_out/r2h_bench_ms.json 2>gpurun_out/r2h_bench_ms.err; tail -3 gpurun_out/r2h_bench_ms.err; cut -c1-300 gpurun_out/r2h_bench_ms.json
timeout 300 python bench.py --workload meanshift --batch 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_ms_b1.json 2>/dev/null; cut -c1-300 gpurun_out/r2h_bench_ms_b1.json
MSM_MS_PERSISTENT=0 timeout 300 python bench.py --workload meanshift --batch 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_ms_b1_launches.json 2>/dev/null; cut -c1-300 gpurun_out/r2h_bench_ms_b1_launches.json
timeout 600 python tools/dev_backbone.py 2>&1 | tail -10
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
