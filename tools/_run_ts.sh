set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -x -k "two_stage_golden or two_stage_no_objects or (two_stage_vs_oracle and 61)" > gpurun_out/ts_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -25 gpurun_out/ts_sanitizer.log
timeout 300 python -m pytest tests -q -m gpu -x -k "two_stage" 2>&1 | tail -25
