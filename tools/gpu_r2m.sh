#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/dev_ms_persist.py 2>&1 | tail -25
timeout 300 compute-sanitizer --tool memcheck python tools/dev_ms_persist.py small > gpurun_out/r2m_sanitizer.log 2>&1; grep -v "^$" gpurun_out/r2m_sanitizer.log | head -60
