#!/bin/bash
# last evidence of round 2: full GPU tests, smoke and both bench arms on the final tree
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/y_build.log 2>&1
timeout -k 10 600 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
tail -4 gpurun_out/y_pytest.log
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/y_smoke.log
cat gpurun_out/y_smoke.log
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/y_bench_ref.json 2> gpurun_out/y_bench.err
cut -c1-200 gpurun_out/y_bench_ref.json
timeout -k 10 400 python bench.py > gpurun_out/y_bench.json 2>> gpurun_out/y_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/y_bench.json; tail -3 gpurun_out/y_bench.err
