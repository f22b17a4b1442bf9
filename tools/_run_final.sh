set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
