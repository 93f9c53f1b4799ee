#!/usr/bin/env python
"""Opcode histogram of the hottest loop (the backward branch whose body holds the most FFMA2) of a kernel, from
cuobjdump -sass.  Offline check of what a source change does to the epilogue's instruction count:
   python scripts/sass_loop_stats.py [kernel-name-substring]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "time-series-kafka-demo_b200", "csrc")
pat = sys.argv[1] if len(sys.argv) > 1 else "tc_fused_kernelILi3ELi3ELi0"
with tempfile.TemporaryDirectory() as td:
    cub = os.path.join(td, "tc.cubin")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-cubin", "-o", cub,
                    os.path.join(CS, "b2cnn_tc.cu")] + sys.argv[2:], check=True, stderr=subprocess.DEVNULL)
    out = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True).stdout
fn, ins = None, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
cands = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:[!U]*P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        n2 = sum("FFMA2" in x[1] for x in body)
        if n2 >= 80:
            cands.append((n2, body))
# every loop that holds the packed math, innermost first (the unrolled main loop and the run-time tail loop)
seen = []
for n2, body in sorted(cands, key=lambda c: len(c[1])):
    if any(body[0][0] <= b[0][0] and b[-1][0] <= body[-1][0] for b in seen):
        continue                                        # an outer loop around one already printed
    seen.append(body)
    ops = collections.Counter()
    for a, t in body:
        t = re.sub(r"^@!?U?P\d\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    steps = max(1, round(n2 / 57))
    print(f"loop body @0x{body[0][0]:x}: {len(body)} instructions, {n2} FFMA2  (~{steps} steps -> {len(body) / steps:.0f} instructions per step)")
    print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
