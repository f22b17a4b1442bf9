#!/usr/bin/env bash
set -x
timeout 120 python -m pytest tests/test_gpu_config2.py -q -x -k "decoder_block or teacher or graph" 2>&1 | tail -5
timeout 60 python tools/prof_block.py 50 stamps
MSM_DECODER_BLOCK=1 timeout 300 python bench.py --workload r50-head --steps 100 --warmup 5 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
