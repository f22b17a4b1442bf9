#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for b in 1 4 32; do timeout 120 python tools/prof_attn.py msp$b 3; MSM_MS_PERSISTENT=0 timeout 120 python tools/prof_attn.py msp$b 3; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mean_shift_persistent -s 1 -c 1 -f -o gpurun_out/r2i_msp python tools/prof_attn.py msp4 1 2>&1 | tail -1
