#!/usr/bin/env bash
# First GPU call of the next round (run on the B200 box: `gpurun --timeout 1500 -- 'bash tools/gpu_next.sh'`):
# everything that was written after this round's GPU minutes ran out, in order of risk. Every step runs under
# `timeout`; logs go to gpurun_out/next_*.log.
set -x
mkdir -p gpurun_out
# 1. the green suite must still be green (the attention finalize kernel gained an optional output)
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
# 1b. ONE image through the R50 head against the CPU oracle: the shape of the LayerNorm-epilogue bug fixed on CPU
timeout 400 python tools/parity_e2e.py --images 1 2>&1 | tail -6
# 2. staged parity tests of the training side: attention backward kernel, decoder gradients, whole training steps
timeout 600 python -m pytest tests -q -m gpu_staged -x 2>&1 | tee gpurun_out/next_staged.log | tail -15
# 3. the same backward kernel under compute-sanitizer (memcheck), smallest cases only
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_staged.py -q -x \
    -k "golden and vmf" > gpurun_out/next_sanitizer.log 2>&1; tail -5 gpurun_out/next_sanitizer.log
# 4. experimental packed-operand attention / mean-shift kernels and the operand-image projection epilogue: parity + timing
#    against the shipped kernels (the staged test of step 2 already ran the decoder with MSM_PACKED_KV=1)
timeout 660 python tools/dev_vmf_packed.py quick 2>&1 | tee gpurun_out/next_vmf_packed.log | tail -12
# 5. training workload (config #5), one GPU
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/next_bench_train.json \
    2> gpurun_out/next_bench_train.err; echo "train bench rc=$?"; cut -c1-600 gpurun_out/next_bench_train.json
tail -3 gpurun_out/next_bench_train.err
# 6. two-stage pipeline end to end (config #3), one GPU
timeout 600 python bench.py --workload twostage --steps 3 --warmup 3 > gpurun_out/next_bench_twostage.json \
    2> gpurun_out/next_bench_twostage.err; echo "twostage bench rc=$?"; cut -c1-600 gpurun_out/next_bench_twostage.json
tail -3 gpurun_out/next_bench_twostage.err
# 7. UCN head step with the packed K / V path against the default (config #1 / #3 stage 1)
timeout 400 python bench.py --workload ucn --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench_ucn_default.json 2>/dev/null
MSM_PACKED_KV=1 timeout 400 python bench.py --workload ucn --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench_ucn_packed.json 2>/dev/null
cut -c1-300 gpurun_out/next_bench_ucn_default.json gpurun_out/next_bench_ucn_packed.json
# 8. mean-shift workload (config #4) with the packed path against the default
timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench_ms_default.json 2>/dev/null
MSM_PACKED_MS=1 timeout 300 python bench.py --workload meanshift --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench_ms_packed.json 2>/dev/null
cut -c1-300 gpurun_out/next_bench_ms_default.json gpurun_out/next_bench_ms_packed.json
# 9. headline R50 step with the L2 persisting window on the mask features against the default
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/next_bench_r50_default.json 2>/dev/null
MSM_L2_PERSIST=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/next_bench_r50_l2persist.json 2> gpurun_out/next_bench_r50_l2persist.err
cut -c1-260 gpurun_out/next_bench_r50_default.json gpurun_out/next_bench_r50_l2persist.json; tail -2 gpurun_out/next_bench_r50_l2persist.err
# 10. headline R50 step with the single-launch kernel for the short attention problems (self-attention, 300-key level)
MSM_SMALL_ATTN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/next_bench_r50_smallattn.json 2> gpurun_out/next_bench_r50_smallattn.err
cut -c1-260 gpurun_out/next_bench_r50_smallattn.json; tail -2 gpurun_out/next_bench_r50_smallattn.err
