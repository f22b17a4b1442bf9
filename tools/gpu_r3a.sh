#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_config2.py -q -x 2>&1 | tail -3
for w in v k kpos k256; do timeout 120 python tools/prof_kimg.py $w 10; done
timeout 300 python bench.py --workload ucn --batch 2 --steps 20 --warmup 3 --inflight 1 --no-cpu-baseline --skip-profile > gpurun_out/r3a_ucn.json 2>gpurun_out/r3a_err.log || tail -20 gpurun_out/r3a_err.log
python - <<PY
import json; d=json.loads(open('gpurun_out/r3a_ucn.json').read().strip().splitlines()[-1]); print('ucn', d['value'], d['ms_per_step'])
for g in d['roofline']['top_groups'][:8]: print("   %-26s %-44s n=%5.1f %7.3f ms  %7.1f GB/s %7.1f TF"%(g['kernel'],g['shape'],g['launches_per_step'],g['ms_per_step'],g['GBps'] or 0,g['TFLOPps'] or 0))
PY
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
