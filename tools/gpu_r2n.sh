#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/dev_ms_persist.py 2>&1 | tail -16
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 60 python tools/prof_block.py 50
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decoder_block -s 3 -c 1 -f -o gpurun_out/r2n_block python tools/prof_block.py 1 2>&1 | tail -1
