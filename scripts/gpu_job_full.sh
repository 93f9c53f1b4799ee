#!/bin/bash
# milestone job: full GPU test suite, smoke, bench (both arms), ncu launch list + full capture
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f_build.log 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
cat gpurun_out/f_smoke.log
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench.err
cat gpurun_out/f_bench_ref.json
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2>> gpurun_out/f_bench.err; echo "bench rc=$?"
cat gpurun_out/f_bench.json; tail -5 gpurun_out/f_bench.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -o gpurun_out/f_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/f_ncu_full.log 2>&1
ls -la gpurun_out | head -40
