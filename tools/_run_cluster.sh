set -x
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck python tools/dev_cluster.py 3000 2 > gpurun_out/cl_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -9 gpurun_out/cl_sanitizer.log
timeout 120 python tools/dev_cluster.py 60000 3 time 2>&1 | tail -9
MSM_SEEDS_VARIANT=0 timeout 120 python tools/dev_cluster.py 60000 3 time 2>&1 | tail -3
timeout 400 python -m pytest tests -q -m gpu -x -k "clusterer or mean_shift" 2>&1 | tail -5
timeout 300 python bench.py --workload cluster --steps 5 --warmup 3 > gpurun_out/bench_cluster.json 2> gpurun_out/bench_cluster.err; echo "bench rc=$?"; cat gpurun_out/bench_cluster.json; tail -3 gpurun_out/bench_cluster.err
