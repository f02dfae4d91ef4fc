#!/bin/bash
# One compute-sanitizer tool over a small slice of the GPU parity tests (one tool per gpurun call):
#   scripts/sanitize.sh memcheck|racecheck|synccheck|initcheck  ->  gpurun_out/sanitizer_<tool>.log
tool=${1:-memcheck}
sel=${2:-"(test_forward_rotmat_surface and fp32 and (5 or 130)) or (test_backward_full and fp32 and not bf16) or test_fused_forward_equals_two_kernel_forward and fp32 and 130"}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel" > gpurun_out/sanitizer_$tool.log 2>&1
rc=$?
echo "compute-sanitizer --tool $tool exit code $rc" >> gpurun_out/sanitizer_$tool.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" gpurun_out/sanitizer_$tool.log | tail -5
