#!/bin/bash
# round 2, job B: all GPU tests; A/B of the compile-time-unrolled epilogue; ncu full capture of the fused kernel
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -12 gpurun_out/b_pytest.log
VARIANTS="nounroll" bash scripts/gpu_job_ab.sh
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -o gpurun_out/b_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --sustained-seconds 0 --parity-windows 0 > gpurun_out/b_ncu_full.log 2>&1
tail -2 gpurun_out/b_ncu_full.log | cut -c1-300
