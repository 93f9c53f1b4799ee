#!/usr/bin/env python
"""Copy the judged evidence from gpurun_out/ (scratch) to profiles/ (tracked):
   python scripts/save_profiles.py <tag> <prefix>     e.g.  r01_fused f_"""
import csv, io, json, os, shutil, subprocess, sys
tag, pre = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed']
for f in os.listdir(G):
    if not f.startswith(pre):
        continue
    src = os.path.join(G, f)
    if f.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, unit = rows[0], rows[1]
        with open(os.path.join(P, f"{tag}_{f[len(pre):-8]}_ncu_full.txt"), "w") as o:
            o.write(f"# ncu --set full --clock-control none (one launch, ~40 replay passes; read with ncu -i {f} --page raw --csv)\n")
            for r in rows[2:]:
                o.write(f"kernel: {r[hdr.index('Kernel Name')]}\n")
                for i, h in enumerate(hdr):
                    if h in KEYS or 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                        o.write(f"  {h} [{unit[i]}] = {r[i]}\n")
    elif f.endswith("launches.csv"):
        rows = list(csv.reader(open(src)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
        hdr, data = rows[hi], rows[hi + 1:]
        ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
        agg = {}
        for r in data:
            if len(r) <= vi: continue
            v = float(r[vi].replace(',', '')) / {'us': 1e3, 'ns': 1e6, 'ms': 1.0}.get(r[ui], 1.0)
            agg.setdefault(r[ki], []).append(v)
        ours = {k: v for k, v in agg.items() if 'b2cnn' in k}
        tot = sum(sum(v) for v in ours.values())
        with open(os.path.join(P, f"{tag}_launches.txt"), "w") as o:
            o.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
            o.write("# only this repo's kernels (torch's data-generation kernels of bench.py omitted)\n")
            o.write("# mean | max (= the full 4096-window launches; the smaller ones are bench.py's e2e chunks) | count | share | kernel\n")
            for k, v in ours.items():
                o.write(f"{sum(v)/len(v):9.4f} ms  max {max(v):8.4f} ms  x{len(v):3d}  share {100*sum(v)/tot:5.1f}%  {k[:110]}\n")
        shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
    elif f.endswith(".json"):
        shutil.copy(src, os.path.join(P, f"{tag}_{f[len(pre):]}"))
print(sorted(os.listdir(P)))
