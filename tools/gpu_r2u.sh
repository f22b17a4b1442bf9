#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/dev_cluster_loop.py
MSM_MS_PERSISTENT=0 timeout 120 python tools/dev_cluster_loop.py
timeout 600 python -m pytest tests/test_gpu_config2.py -q -x -k "lean" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/final_bench_r50.json 2> gpurun_out/final_bench_r50.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench_r50.json
timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --no-cpu-baseline --skip-profile > gpurun_out/r2u_bench_ucn_b2.json 2>/dev/null; cut -c1-330 gpurun_out/r2u_bench_ucn_b2.json
MSM_FOLD_V=0 timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 300 python bench.py --workload demo --steps 50 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2u_bench_demo.json 2>/dev/null; cut -c1-330 gpurun_out/r2u_bench_demo.json
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
