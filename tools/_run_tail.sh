set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -x -k "instance_inference_golden or (instance_inference_vs_oracle and 3-17)" > gpurun_out/tail_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -6 gpurun_out/tail_sanitizer.log
timeout 400 python -m pytest tests -q -m gpu -x -k "instance_inference" 2>&1 | tail -15
timeout 300 python bench.py --workload tail --steps 10 --warmup 3 > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; echo "bench rc=$?"; cat gpurun_out/bench_tail.json; tail -3 gpurun_out/bench_tail.err
