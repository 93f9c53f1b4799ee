#!/bin/bash
# wire-format decoders after the frame-row bound: GPU tests plain and under compute-sanitizer memcheck
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_wire.py tests/test_stream.py -x -q -m gpu > gpurun_out/x_wire.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_wire.log
tail -3 gpurun_out/x_wire.log
timeout -k 10 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_wire.py -x -q -m gpu -k "sample_messages or decimal or array" > gpurun_out/x_wire_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/x_wire_memcheck.log
