#!/usr/bin/env bash
set -x
timeout 300 python tools/dev_backbone.py 2>&1 | tail -8
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
