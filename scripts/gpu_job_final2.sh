#!/bin/bash
# final evidence of round 2 (re-entry session): full GPU tests, smoke, both bench arms, fp32 and B=32768 lines, launch list, ncu full of the fused kernel
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z_build.log 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
tail -4 gpurun_out/z_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/z_smoke.log
cat gpurun_out/z_smoke.log
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench.err
cut -c1-300 gpurun_out/z_bench_ref.json
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/z_bench.json 2>> gpurun_out/z_bench.err; echo "bench rc=$?"
cat gpurun_out/z_bench.json; tail -5 gpurun_out/z_bench.err
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dtype f32 > gpurun_out/z_bench_f32.json 2>> gpurun_out/z_bench.err; echo "bench f32 rc=$?"
cut -c1-200 gpurun_out/z_bench_f32.json
timeout -k 10 300 python bench.py --batch 32768 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench_b32768.json 2>> gpurun_out/z_bench.err; echo "bench B=32768 rc=$?"
cut -c1-200 gpurun_out/z_bench_b32768.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/z_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/z_ncu_launch.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -f -o gpurun_out/z_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/z_ncu_full.log 2>&1
ls gpurun_out | grep "^z_"
