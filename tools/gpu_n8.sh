#!/usr/bin/env bash
# 8-GPU check (one box): whole-model replicas and the DDP training step. `gpurun --gpus 8 -- bash tools/gpu_n8.sh`
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/n8_bench_r50.json 2> gpurun_out/n8_bench_r50.err; echo "rc=$?"; cut -c1-420 gpurun_out/n8_bench_r50.json; tail -3 gpurun_out/n8_bench_r50.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n8_bench_train.json 2> gpurun_out/n8_bench_train.err; echo "rc=$?"; cut -c1-420 gpurun_out/n8_bench_train.json; tail -3 gpurun_out/n8_bench_train.err
if [ "$1" != "short" ]; then
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/n8_bench_r50_n1.json 2>/dev/null; cut -c1-300 gpurun_out/n8_bench_r50_n1.json
fi
