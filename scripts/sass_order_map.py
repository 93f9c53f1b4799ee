#!/usr/bin/env python
"""One character per SASS instruction of the fused kernel's hottest loop, in program order -- shows how ptxas interleaved
the pipes:  M = MUFU (XU), F = packed/scalar fp32 FMA-pipe op, a = ALU-pipe op (FMNMX, LOP3, SHF, ...), T = LDTM/STTM,
S = SYNCS (mbarrier), B = branch, W = WARPSYNC/BSYNC, u = uniform-datapath op, . = other
   python scripts/sass_order_map.py [extra nvcc flags ...]"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "time-series-kafka-demo_b200", "csrc")
pat = os.environ.get("KERNEL", "tc_fused_kernelILi3ELi3ELi0")
with tempfile.TemporaryDirectory() as td:
    cub = os.path.join(td, "tc.cubin")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-cubin", "-o", cub,
                    os.path.join(CS, "b2cnn_tc.cu")] + sys.argv[1:], check=True, stderr=subprocess.DEVNULL)
    out = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", cub], capture_output=True, text=True).stdout
fn, ins = None, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m: fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def cls(t):
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    op = op.split(".")[0]
    if op == "MUFU": return "M"
    if op in ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "HFMA2", "IMAD"): return "F"
    if op in ("FMNMX", "FMNMX3", "LOP3", "SHF", "IADD3", "ISETP", "LEA", "MOV", "F2FP", "PRMT", "SEL", "VIMNMX3", "PLOP3"): return "a"
    if op in ("LDTM", "STTM"): return "T"
    if op == "SYNCS": return "S"
    if op in ("BRA", "EXIT", "CALL", "RET"): return "B"
    if op in ("WARPSYNC", "BSYNC", "BSSY", "NANOSLEEP", "BAR"): return "W"
    if op.startswith("U") or op in ("S2UR", "R2UR", "VOTEU", "LDCU"): return "u"
    if op in ("LDS", "LDC", "LDG", "STG", "STS"): return "L"
    return "."
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:[!U]*P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        n2 = sum("FFMA2" in x[1] for x in body)
        if n2 >= 300 and (best is None or len(body) < len(best)): best = body
for l in res.splitlines():
    if pat in l: print(l.strip()[:200])
    elif "REG:" in l and fn_seen: print(l.strip()[:200]); fn_seen = False
    fn_seen = pat in l
s = "".join(cls(t) for a, t in best)
print(f"main loop: {len(best)} instructions, {s.count('M')} MUFU")
for i in range(0, len(s), 120): print(s[i:i + 120])
# MUFU gap histogram
pos = [i for i, c in enumerate(s) if c == "M"]
gaps = [b - a for a, b in zip(pos, pos[1:])]
import collections
h = collections.Counter(min(g, 12) for g in gaps)
print("gaps between consecutive MUFU (instructions, 12 = 12+):", dict(sorted(h.items())))
