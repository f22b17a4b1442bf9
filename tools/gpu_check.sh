#!/usr/bin/env bash
# One-call GPU validation of a build (run on the B200 box, e.g. `gpurun --timeout 900 -- 'bash tools/gpu_check.sh'`):
# parity tests, smoke, the headline bench and the two stand-alone workloads. Every step runs under `timeout`.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 300 python bench.py > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "bench rc=$?"
timeout 200 python bench.py --workload cluster --steps 5 --warmup 3 > gpurun_out/bench_cluster.json 2>/dev/null
timeout 200 python bench.py --workload tail --steps 10 --warmup 3 > gpurun_out/bench_tail.json 2>/dev/null
cut -c1-400 gpurun_out/bench_r50.json gpurun_out/bench_cluster.json gpurun_out/bench_tail.json
