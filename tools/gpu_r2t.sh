#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for w in linear_ow linear_vp linear_mf linear_kv linear_small linear_small_ffn1 linear_small_ffn2 msda_fused ffn; do timeout 60 python tools/prof_ops.py $w --iters 10; done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 3 -c 1 -f -o gpurun_out/r2t_linear_ow python tools/prof_ops.py linear_ow --iters 2 2>&1 | tail -1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 3 -c 1 -f -o gpurun_out/r2t_linear_small python tools/prof_ops.py linear_small --iters 2 2>&1 | tail -1
