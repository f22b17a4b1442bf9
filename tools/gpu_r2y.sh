#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -x -k "packed_kv" 2>&1 | tail -5
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
for fk in 1 0; do
MSM_FOLD_K=$fk timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --inflight 1 --no-cpu-baseline --skip-profile > gpurun_out/r2y_ucn_fk$fk.json 2>gpurun_out/r2y_err.log || tail -20 gpurun_out/r2y_err.log
python - <<PY
import json; d=json.loads(open('gpurun_out/r2y_ucn_fk$fk.json').read().strip().splitlines()[-1]); print('ucn foldk $fk', d['value'], d['ms_per_step'])
for g in d['roofline']['top_groups'][:8]: print("   %-26s %-44s n=%5.1f %7.3f ms  %7.1f GB/s %7.1f TF"%(g['kernel'],g['shape'],g['launches_per_step'],g['ms_per_step'],g['GBps'] or 0,g['TFLOPps'] or 0))
PY
done
timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --inflight 2 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 300 python bench.py --workload demo --steps 50 --inflight 1 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 300 python bench.py --workload demo --steps 50 --inflight 3 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 300 python bench.py --workload crop --batch 16 --steps 20 --inflight 1 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
timeout 300 python bench.py --workload crop --batch 16 --steps 20 --inflight 2 --no-cpu-baseline --skip-profile 2>/dev/null | cut -c1-330
