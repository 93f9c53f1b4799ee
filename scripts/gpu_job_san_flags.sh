#!/bin/bash
# compute-sanitizer over the handle-owned NaN-exception flag state (cleaned by the head kernel, no per-call
# memset): the two tests that flag windows, change the flagged set between calls, switch streams and batch sizes.
mkdir -p gpurun_out
SEL='flag_state_is_clean or nan_inf_windows_are_recomputed'
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout -k 10 ${SAN_TIMEOUT:-330} compute-sanitizer --tool $tool --print-limit 20 \
    python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "$SEL" > gpurun_out/sanflags_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -c "Invalid\|out of bounds\|misaligned\|hazard\|Barrier error" gpurun_out/sanflags_$tool.log
  tail -4 gpurun_out/sanflags_$tool.log
done
