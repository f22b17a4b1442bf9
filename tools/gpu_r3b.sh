#!/usr/bin/env bash
set -x
for w in v kpos k256; do timeout 120 python tools/prof_kimg.py $w 10 stamps; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 2 -c 1 -f -o gpurun_out/r3b_v python tools/prof_kimg.py v 3 > /dev/null 2>&1; echo "ncu rc=$?"
