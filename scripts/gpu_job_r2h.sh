#!/bin/bash
# round 2, re-entry: full GPU test suite of the current build, then A/B of the per-step overhead variants on this box
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -4 gpurun_out/h_pytest.log
VARIANTS="${VARIANTS:-memset carve}" bash scripts/gpu_job_ab.sh
