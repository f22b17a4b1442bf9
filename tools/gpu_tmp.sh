#!/usr/bin/env bash
set -x
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -5
for f in 1 3; do
timeout 600 python bench.py --inflight $f --no-cpu-baseline --skip-profile > gpurun_out/tmp_r50_f$f.json 2>gpurun_out/tmp_err.log || tail -20 gpurun_out/tmp_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/tmp_r50_f$f.json').read().strip().splitlines()[-1]); print('r50 inflight $f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'))
PY
done
timeout 600 python bench.py --workload r50-head --inflight 3 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-300
