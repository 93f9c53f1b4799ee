#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
echo "== odd ranges (expected to fail if TMA needs 16B-aligned box starts)" > gpurun_out/b2_exp.log
B2CNN_TC_ODD_RANGES=1 timeout 300 python -m pytest "tests/test_gpu_tc.py::test_tc_features_and_logits" -q -x -k "7504-128" >> gpurun_out/b2_exp.log 2>&1
echo "== even ranges" >> gpurun_out/b2_exp.log
timeout 300 python -m pytest "tests/test_gpu_tc.py::test_tc_features_and_logits" -q -x -k "7504-128" >> gpurun_out/b2_exp.log 2>&1
echo "== even ranges, 3 tiles per CTA" >> gpurun_out/b2_exp.log
B2CNN_TC_TILES=3 timeout 300 python -m pytest "tests/test_gpu_tc.py::test_tc_features_and_logits" -q -x >> gpurun_out/b2_exp.log 2>&1
grep -E "==|passed|failed|Error|error" gpurun_out/b2_exp.log | head -30
bash scripts/gpu_job_b.sh
