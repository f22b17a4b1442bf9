#!/usr/bin/env bash
# round-2 call F: whole-model bench (images in -> label maps out), head-only and single-frame lines
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2f_bench_r50.json 2>gpurun_out/r2f_bench_r50.err; echo rc=$?; tail -5 gpurun_out/r2f_bench_r50.err; cut -c1-1500 gpurun_out/r2f_bench_r50.json
timeout 300 python bench.py --workload r50-head --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_r50head.json 2>gpurun_out/r2f_bench_r50head.err; echo rc=$?; tail -3 gpurun_out/r2f_bench_r50head.err; cut -c1-400 gpurun_out/r2f_bench_r50head.json
timeout 300 python bench.py --workload demo --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_demo.json 2>gpurun_out/r2f_bench_demo.err; echo rc=$?; tail -3 gpurun_out/r2f_bench_demo.err; cut -c1-600 gpurun_out/r2f_bench_demo.json
timeout 300 python bench.py --tail instances --steps 50 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2f_bench_r50_inst.json 2>/dev/null; cut -c1-300 gpurun_out/r2f_bench_r50_inst.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_ref_r50.json 2>/dev/null; cut -c1-400 gpurun_out/r2f_ref_r50.json
