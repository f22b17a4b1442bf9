set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x -k "instance_inference" 2>&1 | tail -5
timeout 300 python bench.py --workload tail --steps 10 --warmup 3 > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; echo "bench rc=$?"; cat gpurun_out/bench_tail.json; tail -3 gpurun_out/bench_tail.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tail.csv python bench.py --workload tail --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu tail rc=$?"
timeout 280 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cluster.csv python bench.py --workload cluster --batch 4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu cluster rc=$?"
