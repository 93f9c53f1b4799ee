#!/bin/bash
# quick iteration on the tensor-core kernels: TC tests (short timeout), bench, one ncu capture
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1
timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/c_tc.log 2>&1; rc=$?; echo "tc pytest rc=$rc" >> gpurun_out/c_tc.log
tail -8 gpurun_out/c_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
cat gpurun_out/c_bench.json; tail -5 gpurun_out/c_bench.err
if [ "$1" == "ncu" ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -o gpurun_out/c_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/c_ncu_full.log 2>&1
fi
