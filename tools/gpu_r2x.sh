#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config2.py -q -x -k "lean or resample" 2>&1 | tail -3
timeout 300 python tools/dev_timeline.py r50 2>&1 | grep -v Warning | tail -32
for f in 1 2 3 4; do
timeout 600 python bench.py --inflight $f --no-cpu-baseline --skip-profile > gpurun_out/r2x_r50_f$f.json 2>gpurun_out/r2x_err.log || tail -20 gpurun_out/r2x_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/r2x_r50_f$f.json').read().strip().splitlines()[-1]); print('r50 inflight $f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'), d['clocks'])
PY
done
MSM_DECODER_BLOCK=1 timeout 600 python bench.py --inflight 2 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
