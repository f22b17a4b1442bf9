#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for f in 1 2 3; do
timeout 600 python bench.py --inflight $f --no-cpu-baseline --skip-profile > gpurun_out/r2v_r50_f$f.json 2>gpurun_out/r2v_err.log || tail -20 gpurun_out/r2v_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/r2v_r50_f$f.json').read().strip().splitlines()[-1]); print('r50 inflight $f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'), d['clocks'])
PY
done
for f in 2 3; do
timeout 600 python bench.py --workload r50-head --inflight $f --no-cpu-baseline --skip-profile > gpurun_out/r2v_r50head_f$f.json 2>gpurun_out/r2v_err.log || tail -20 gpurun_out/r2v_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/r2v_r50head_f$f.json').read().strip().splitlines()[-1]); print('r50-head inflight $f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'), d['clocks'])
PY
done
timeout 600 python bench.py --workload demo --steps 50 --inflight 2 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 600 python bench.py --workload ucn --batch 2 --steps 20 --inflight 2 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 600 python bench.py --workload cluster --steps 10 --no-cpu-baseline 2>/dev/null | cut -c1-400
timeout 600 python bench.py --workload meanshift --steps 10 --no-cpu-baseline 2>/dev/null | cut -c1-400
timeout 900 python bench.py --steps 100 > gpurun_out/r2v_r50_default.json 2>gpurun_out/r2v_err.log; echo "rc=$?"; cut -c1-420 gpurun_out/r2v_r50_default.json
