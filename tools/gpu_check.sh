#!/usr/bin/env bash
# One-call GPU validation + measurement of a build (on the B200 box: `gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'`):
# parity tests, smoke, every bench workload (JSON lines under gpurun_out/final_*.json), launch list of the default step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/final_bench_r50.json 2> gpurun_out/final_bench_r50.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/final_ref_r50.json 2>/dev/null
timeout 300 python bench.py --backbone-fp32 --steps 50 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/final_bench_r50_fp32bb.json 2>/dev/null
timeout 300 python bench.py --workload r50-head --steps 200 --warmup 10 > gpurun_out/final_bench_r50head.json 2>/dev/null
timeout 300 python bench.py --inflight 1 --no-cpu-baseline --skip-profile > gpurun_out/final_bench_r50_serial.json 2>/dev/null
timeout 300 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_train.json 2>/dev/null
MSM_DECODER_BLOCK=1 timeout 300 python bench.py --workload r50-head --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/final_bench_r50head_block.json 2>/dev/null
timeout 300 python bench.py --workload demo --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_demo.json 2>/dev/null
timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_ucn_b2.json 2>/dev/null
timeout 300 python bench.py --workload crop --batch 16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_crop_b16.json 2>/dev/null
timeout 300 python bench.py --workload meanshift --steps 20 --warmup 3 > gpurun_out/final_bench_meanshift.json 2>/dev/null
timeout 300 python bench.py --workload cluster --steps 20 --warmup 3 > gpurun_out/final_bench_cluster.json 2>/dev/null
timeout 300 python bench.py --workload tail --steps 20 --warmup 3 > gpurun_out/final_bench_tail.json 2>/dev/null
timeout 400 python bench.py --workload twostage --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_twostage.json 2>/dev/null
for f in gpurun_out/final_*.json; do echo "$f: $(cut -c1-260 $f)"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-e2e --no-cpu-baseline --skip-profile > /dev/null 2>&1; echo "launch list rc=$?"
