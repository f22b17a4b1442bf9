#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_config2.py -q -x -k "decoder_block or teacher or graph" 2>&1 | tail -5
timeout 60 python tools/prof_block.py 50
timeout 300 python bench.py --workload r50-head --steps 100 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2o_bench_r50head.json 2>/dev/null; cut -c1-330 gpurun_out/r2o_bench_r50head.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_block -s 3 -c 1 -f -o gpurun_out/r2o_block python tools/prof_block.py 1 2>&1 | tail -1
