#!/usr/bin/env bash
# 2 GPUs: replica scaling of the default bench (device-timed and end to end), DDP training step
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2s_bench_n2.json 2>gpurun_out/r2s_bench_n2.err; echo rc=$?; tail -3 gpurun_out/r2s_bench_n2.err; cut -c1-900 gpurun_out/r2s_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 10 --warmup 3 > gpurun_out/r2s_train_n2.json 2>gpurun_out/r2s_train_n2.err; echo rc=$?; tail -3 gpurun_out/r2s_train_n2.err; cut -c1-700 gpurun_out/r2s_train_n2.json
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --skip-profile > gpurun_out/r2s_bench_n1.json 2>/dev/null; cut -c1-330 gpurun_out/r2s_bench_n1.json
