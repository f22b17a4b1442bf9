set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x -k "decoder or head or msdeform or pixel" 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_fold.json; tail -3 gpurun_out/bench_fold.err
