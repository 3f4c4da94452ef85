#!/bin/bash
# compute-sanitizer over a representative slice of the GPU suites (memcheck on the launch and persistent-kernel paths,
# racecheck and initcheck on the bit-exact kernel tests); logs go to gpurun_out/ and, summarised, to profiles/.
# Usage (GPU box): bash tools/sanitize.sh
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="tests/test_gpu_kernels.py tests/test_gpu_cycles.py -k corner_nc8_l4"
for tool in memcheck racecheck initcheck; do
  timeout 900 $CS --tool $tool --error-exitcode 86 --print-limit 20 python -m pytest $SEL -m gpu -q -x -p no:cacheprovider \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" gpurun_out/sanitizer_$tool.log | tail -4
done
timeout 900 $CS --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_stencils.py tests/test_gpu_builders_device.py tests/test_gpu_state_changes.py -k "lsf_sphere_corner or eps_smooth_corner or rod-linear or set_rhs or interior_download" -m gpu -q -x -p no:cacheprovider \
  > gpurun_out/sanitizer_memcheck_stencils.log 2>&1
echo "memcheck (stencils) exit code $?" >> gpurun_out/sanitizer_memcheck_stencils.log
grep -E "ERROR SUMMARY|passed|failed|exit code" gpurun_out/sanitizer_memcheck_stencils.log | tail -3
# the kernels the large nc = 8 levels take (one-wave half-sweep, four boxes per CTA in the prolongation) on S2 at full size
timeout 900 $CS --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_fullsize.py -k "s2" -m gpu -q -x -p no:cacheprovider \
  > gpurun_out/sanitizer_memcheck_s2.log 2>&1
echo "memcheck (S2) exit code $?" >> gpurun_out/sanitizer_memcheck_s2.log
grep -E "ERROR SUMMARY|passed|failed|exit code" gpurun_out/sanitizer_memcheck_s2.log | tail -3
