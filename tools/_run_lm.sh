set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -x -k "label_map_from_outputs or instance_inference_golden" > gpurun_out/lm_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -25 gpurun_out/lm_sanitizer.log
timeout 300 python -m pytest tests -q -m gpu -x -k "label_map or instance_inference" 2>&1 | tail -15
