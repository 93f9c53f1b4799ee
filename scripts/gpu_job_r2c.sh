#!/bin/bash
# round 2, job C: parity tests of the current build, then A/B of projection lag / spin-vs-park variants on one box
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
VARIANTS="lag6 lag7 mmaspin epispin spin7" bash scripts/gpu_job_ab.sh
