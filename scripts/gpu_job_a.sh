#!/bin/bash
# First GPU job: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -30 gpurun_out/a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
cat gpurun_out/a_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
cat gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2>> gpurun_out/a_bench.err
cat gpurun_out/a_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/a_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 2 -c 1 -o gpurun_out/a_front \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/a_ncu_full.log 2>&1
ls -la gpurun_out
