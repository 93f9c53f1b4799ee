#!/bin/bash
# round 2, job A: all GPU tests, smoke, both bench arms (new bench.py blocks)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt
timeout -k 10 1200 python -m pytest tests -m gpu -q -x > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -8 gpurun_out/a_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
tail -4 gpurun_out/a_smoke.log
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/a_bench_ref.json; grep -o '"cpu_baseline".*' gpurun_out/a_bench_ref.json | cut -c1-1200
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_ref2.json 2>> gpurun_out/a_bench.err; echo "ref2 rc=$?"
cut -c1-200 gpurun_out/a_bench_ref2.json
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2>> gpurun_out/a_bench.err; echo "bench rc=$?"
cat gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
