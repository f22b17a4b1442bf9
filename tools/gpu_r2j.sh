#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "mean_shift or clusterer" 2>&1 | tail -4
for b in 1 4 32; do timeout 120 python tools/prof_attn.py msp$b 3; MSM_MS_PERSISTENT=0 timeout 120 python tools/prof_attn.py msp$b 3; done
timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_ms.json 2>/dev/null; cut -c1-300 gpurun_out/r2j_bench_ms.json
timeout 300 python bench.py --workload meanshift --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_ms20.json 2>/dev/null; cut -c1-300 gpurun_out/r2j_bench_ms20.json
MSM_MS_PERSISTENT=0 timeout 300 python bench.py --workload meanshift --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_ms20_launches.json 2>/dev/null; cut -c1-300 gpurun_out/r2j_bench_ms20_launches.json
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_r50.json 2>/dev/null; cut -c1-330 gpurun_out/r2j_bench_r50.json
