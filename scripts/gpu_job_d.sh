#!/bin/bash
# compare the two fused-kernel epilogue variants
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout -k 10 240 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/d_tc.log 2>&1; rc=$?; echo "tc pytest rc=$rc" >> gpurun_out/d_tc.log
tail -4 gpurun_out/d_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
for v in ${VARIANTS:-1 2}; do
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tc-variant $v > gpurun_out/d_bench_v$v.json 2> gpurun_out/d_bench.err
grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/d_bench_v$v.json | tr '\n' ' '; echo " variant $v"
done
if [ "$1" != "" ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -o gpurun_out/d_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --tc-variant $1 > gpurun_out/d_ncu_full.log 2>&1
fi
