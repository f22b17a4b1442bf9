#!/usr/bin/env bash
# round-2 call L: decoder-layer cluster kernel - unit test (short timeout: a protocol bug hangs), then the suite
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_config2.py -q -x -k "decoder_block" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then
  timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_config2.py -q -x -k "decoder_block and 1-100" > gpurun_out/r2l_sanitizer.log 2>&1; tail -30 gpurun_out/r2l_sanitizer.log
  exit 0
fi
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python bench.py --workload r50-head --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_r50head.json 2>gpurun_out/r2l_bench_r50head.err; tail -2 gpurun_out/r2l_bench_r50head.err; cut -c1-330 gpurun_out/r2l_bench_r50head.json
MSM_DECODER_BLOCK=0 timeout 300 python bench.py --workload r50-head --steps 100 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2l_bench_r50head_noblock.json 2>/dev/null; cut -c1-330 gpurun_out/r2l_bench_r50head_noblock.json
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2l_bench_r50.json 2>/dev/null; cut -c1-330 gpurun_out/r2l_bench_r50.json
for b in 1 4; do timeout 120 python tools/prof_attn.py msp$b 3; done
