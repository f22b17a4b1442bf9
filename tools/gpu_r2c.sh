#!/usr/bin/env bash
# round-2 call C: cheaper MMA issue loop of the packed attention kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 120 python tools/prof_attn.py ucn
timeout 120 python tools/prof_attn.py r50
timeout 120 python tools/prof_attn.py crop
timeout 120 python tools/prof_attn.py ms
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vmf_attn_packed -s 2 -c 1 -f -o gpurun_out/r2c_attn_ucn python tools/prof_attn.py ucn 1 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vmf_attn_packed -s 2 -c 1 -f -o gpurun_out/r2c_attn_ms python tools/prof_attn.py ms 1 2>&1 | tail -2
timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_ms.json 2>/dev/null; cut -c1-300 gpurun_out/r2c_bench_ms.json
timeout 300 python bench.py --workload ucn --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_ucn.json 2>/dev/null; cut -c1-300 gpurun_out/r2c_bench_ucn.json
