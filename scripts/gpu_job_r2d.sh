#!/bin/bash
# round 2, job D: all GPU tests (batch kernel, padded rows), bench, call overhead, sweep
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -15 gpurun_out/d_pytest.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/d_bench.json')); print({k: j[k] for k in ('value','ms_per_step','gpu_launches')}, j['roofline']['kernel_ms'], j['roofline']['head_ms'], j['roofline']['whole_step_frac'], j['sustained']['ms_per_step'], j['parity'])"
timeout -k 10 300 python scripts/call_overhead.py > gpurun_out/d_call_overhead.txt 2>&1; head -3 gpurun_out/d_call_overhead.txt | cut -c1-900
timeout -k 10 900 python scripts/sweep.py > gpurun_out/d_sweep.jsonl 2> gpurun_out/d_sweep.err; echo "sweep rc=$?"; cut -c1-420 gpurun_out/d_sweep.jsonl; tail -3 gpurun_out/d_sweep.err
