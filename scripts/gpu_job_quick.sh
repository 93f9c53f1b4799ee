#!/bin/bash
# quick loop: tensor-core parity tests + bench of the current build; NCU=1 adds a full ncu capture of the fused kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
timeout -k 10 400 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/q_tc.log 2>&1; rc=$?; echo "tc pytest rc=$rc" >> gpurun_out/q_tc.log
tail -5 gpurun_out/q_tc.log
for sp in ${SPL:-3}; do
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tc-splits $sp > gpurun_out/q_bench_s$sp.json 2>> gpurun_out/k.err
echo "splits $sp: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/q_bench_s$sp.json | tr '\n' ' ')"
done
if [ "$NCU" != "" ] && [ $rc -eq 0 ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -o gpurun_out/q_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/q_ncu_full.log 2>&1
tail -2 gpurun_out/q_ncu_full.log | cut -c1-200
fi
tail -3 gpurun_out/k.err
