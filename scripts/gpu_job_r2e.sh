#!/bin/bash
# round 2, job E: GPU tests (wire formats, fixes), ncu launch lists of the bench step and of the short-window calls
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -12 gpurun_out/e_pytest.log
timeout -k 10 300 python scripts/small_calls.py > gpurun_out/e_small_calls.txt 2>&1; cat gpurun_out/e_small_calls.txt | tail -3
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_small_launches.csv \
    python scripts/small_calls.py > gpurun_out/e_ncu_small.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/e_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --sustained-seconds 0 --parity-windows 0 > gpurun_out/e_ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/e_small_launches.csv", "gpurun_out/e_launches.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5 and r[0].isdigit()]
    agg = collections.defaultdict(list)
    for r in rows:
        agg[r[4][:60]].append(float(r[-1]))
    print(f)
    for k, v in agg.items():
        print(f"  {k:60s} n={len(v):3d} mean={sum(v)/len(v):9.1f} min={min(v):9.1f}")
PY
