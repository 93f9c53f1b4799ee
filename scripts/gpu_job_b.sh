#!/bin/bash
# tcgen05 path bring-up: TC tests first, diagnostics on failure, then everything + bench + ncu.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/b_tc.log 2>&1; rc=$?; echo "tc pytest rc=$rc" >> gpurun_out/b_tc.log
tail -40 gpurun_out/b_tc.log
if [ $rc -ne 0 ]; then
  timeout 600 python scripts/tc_diag.py > gpurun_out/b_diag.log 2>&1; cat gpurun_out/b_diag.log | tail -40
  exit 0
fi
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
cat gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/b_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/b_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_frontend -s 2 -c 1 -o gpurun_out/b_tc \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out
