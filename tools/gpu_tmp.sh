#!/usr/bin/env bash
set -x
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -5
for f in 1 3; do
timeout 600 python bench.py --inflight $f --no-cpu-baseline --skip-profile > gpurun_out/tmp_r50_f$f.json 2>gpurun_out/tmp_err.log || tail -20 gpurun_out/tmp_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/tmp_r50_f$f.json').read().strip().splitlines()[-1]); print('r50 inflight $f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'), d['parts_ms'])
PY
done
timeout 300 python bench.py --workload tail --steps 20 --warmup 3 > gpurun_out/final_bench_tail.json 2> gpurun_out/final_bench_tail.err; echo "tail rc=$?"; cut -c1-300 gpurun_out/final_bench_tail.json
timeout 300 python tools/dev_timeline.py r50 2>&1 | grep -v Warning | tail -30
