#!/usr/bin/env python
"""Per-instruction warp-stall samples of one kernel from an .ncu-rep captured with --import-source on:
   python scripts/ncu_source_stalls.py report.ncu-rep [top_n]
prints (1) stall reasons summed over the kernel, (2) the same by opcode, (3) the top_n instructions by samples."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[hdr_i:]))))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
def num(v):
    try: return float(v)
    except Exception: return 0.0
tot = collections.Counter(); by_op = collections.defaultdict(collections.Counter); n_samples = 0
for r in rows:
    op = r["Source"].split()[0] if r["Source"] else "?"
    if op.startswith("@"): op = r["Source"].split()[1]
    op = op.split(".")[0]
    s = num(r["# Samples"]); n_samples += s
    for c in stall_cols:
        v = num(r[c]); tot[c] += v; by_op[op][c] += v
    by_op[op]["#"] += s; by_op[op]["n"] += num(r["Instructions Executed"])
print(f"samples {n_samples:.0f}")
print("stall reasons:", ", ".join(f"{k[6:]} {v / n_samples * 100:.1f}%" for k, v in tot.most_common() if v))
print("by opcode (share of samples | warp-instructions executed | top reasons):")
for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["#"])[:25]:
    rs = sorted(((k, v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:4]
    print(f"  {op:10s} {c['#'] / n_samples * 100:5.1f}%  n={c['n']:.3g}  " + ", ".join(f"{k[6:]} {v / max(c['#'], 1) * 100:.0f}%" for k, v in rs if v))
print(f"top {top} instructions:")
for r in sorted(rows, key=lambda r: -num(r["# Samples"]))[:top]:
    rs = sorted(((c, num(r[c])) for c in stall_cols), key=lambda kv: -kv[1])[:3]
    print(f"  {r['Address'][-5:]} {num(r['# Samples']):7.0f}  {r['Source'][:70]:70s} " + ", ".join(f"{k[6:]} {v:.0f}" for k, v in rs if v))
