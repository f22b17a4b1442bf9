#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
for w in v kpos k256; do timeout 120 python tools/prof_kimg.py $w 10 stamps; done
timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --inflight 1 --no-cpu-baseline --skip-profile > gpurun_out/r3c_ucn.json 2>gpurun_out/r3c_err.log || tail -20 gpurun_out/r3c_err.log
python - <<PY
import json; d=json.loads(open('gpurun_out/r3c_ucn.json').read().strip().splitlines()[-1]); print('ucn', d['value'], d['ms_per_step'])
for g in d['roofline']['top_groups'][:8]: print("   %-26s %-44s n=%5.1f %7.3f ms  %7.1f GB/s %7.1f TF"%(g['kernel'],g['shape'],g['launches_per_step'],g['ms_per_step'],g['GBps'] or 0,g['TFLOPps'] or 0))
PY
for f in 1 3; do
timeout 600 python bench.py --workload r50-head --inflight $f --no-cpu-baseline --skip-profile > gpurun_out/r3c_r50head_f$f.json 2>gpurun_out/r3c_err.log || tail -20 gpurun_out/r3c_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/r3c_r50head_f$f.json').read().strip().splitlines()[-1]); print('r50-head inflight $f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'))
for g in d['roofline']['top_groups'][:6]: print("   %-26s %-44s n=%5.1f %7.3f ms  %7.1f GB/s %7.1f TF"%(g['kernel'],g['shape'],g['launches_per_step'],g['ms_per_step'],g['GBps'] or 0,g['TFLOPps'] or 0))
PY
done
timeout 600 python bench.py --inflight 3 --no-cpu-baseline --skip-profile > gpurun_out/r3c_r50_f3.json 2>gpurun_out/r3c_err.log || tail -20 gpurun_out/r3c_err.log; python - <<PY
import json; d=json.loads(open('gpurun_out/r3c_r50_f3.json').read().strip().splitlines()[-1]); print('r50 inflight 3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('serial'))
PY
