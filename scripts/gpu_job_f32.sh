#!/bin/bash
# fp32 windows: bench line of the streaming kernel, the generic kernel beside it, and a full ncu capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --dtype f32 > gpurun_out/s_bench_f32.json 2> gpurun_out/s.err
echo "f32 stream : $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"path": "[a-z]*"' gpurun_out/s_bench_f32.json | tr '\n' ' ')"
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --dtype f32 --path generic > gpurun_out/s_bench_f32_generic.json 2>> gpurun_out/s.err
echo "f32 generic: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"frac": [0-9.]*\|"path": "[a-z]*"' gpurun_out/s_bench_f32_generic.json | tr '\n' ' ')"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:stream_f32 -s 2 -c 1 -o gpurun_out/s_stream \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --dtype f32 > gpurun_out/s_ncu_full.log 2>&1
tail -2 gpurun_out/s_ncu_full.log | cut -c1-160
tail -3 gpurun_out/s.err
