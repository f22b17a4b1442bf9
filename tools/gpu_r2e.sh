#!/usr/bin/env bash
# round-2 call E: packed attention with three score buffers + two-wide fp32 softmax math
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 120 python tools/prof_attn.py ucn
timeout 120 python tools/prof_attn.py r50
timeout 120 python tools/prof_attn.py crop
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vmf_attn_packed -s 2 -c 1 -f -o gpurun_out/r2e_attn_ucn python tools/prof_attn.py ucn 1 2>&1 | tail -1
MSM_PACKED_MS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:vmf_attn_packed -s 2 -c 1 -f -o gpurun_out/r2e_attn_ms python tools/prof_attn.py ms 1 2>&1 | tail -1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_r50.json 2>gpurun_out/r2e_bench_r50.err; cut -c1-300 gpurun_out/r2e_bench_r50.json
timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_ms.json 2>/dev/null; cut -c1-300 gpurun_out/r2e_bench_ms.json
timeout 300 python bench.py --workload ucn --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_ucn.json 2>/dev/null; cut -c1-300 gpurun_out/r2e_bench_ucn.json
