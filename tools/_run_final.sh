set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c150-330 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 200 python bench.py --workload cluster --steps 5 --warmup 3 > gpurun_out/bench_cluster.json 2>/dev/null; cut -c150-400 gpurun_out/bench_cluster.json
