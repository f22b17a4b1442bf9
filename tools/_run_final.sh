set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 200 python bench.py --no-cpu-baseline --skip-e2e --steps 30 > gpurun_out/bench_a.json 2>/dev/null; cut -c150-330 gpurun_out/bench_a.json
MSM_FOLD_POS=1 timeout 200 python bench.py --no-cpu-baseline --skip-e2e --steps 30 > gpurun_out/bench_b.json 2>/dev/null; cut -c150-330 gpurun_out/bench_b.json
timeout 200 python bench.py --no-cpu-baseline --skip-e2e --steps 30 > gpurun_out/bench_c.json 2>/dev/null; cut -c150-330 gpurun_out/bench_c.json
