#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for w in v k kpos k256; do timeout 120 python tools/prof_kimg.py $w 10; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 1 -f -o gpurun_out/r2z_kpos python tools/prof_kimg.py kpos 3 > /dev/null 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 1 -f -o gpurun_out/r2z_v python tools/prof_kimg.py v 3 > /dev/null 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
