set -x
mkdir -p gpurun_out
timeout 200 python bench.py --workload tail --steps 10 --warmup 3 > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; echo "tail rc=$?"; cut -c1-700 gpurun_out/bench_tail.json; tail -3 gpurun_out/bench_tail.err
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:instance_masks_kernel -c 1 -o gpurun_out/c40_instance_masks python bench.py --workload tail --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu1 rc=$?"
timeout 280 ncu --set full --clock-control none --import-source on -k regex:"smart_seeds_ring_kernel|assign_kernel" -c 2 -o gpurun_out/c41_cluster python bench.py --workload cluster --batch 4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep
