#!/usr/bin/env bash
set -x
timeout 120 python tools/dev_fill.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 1 -f -o gpurun_out/r3d_v python tools/prof_kimg.py v 3 > /dev/null 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 1 -f -o gpurun_out/r3d_kpos python tools/prof_kimg.py kpos 3 > /dev/null 2>&1; echo "ncu rc=$?"
