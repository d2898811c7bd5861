#!/usr/bin/env python
"""Per-kernel SASS evidence from the built runtime library (no GPU needed):
python tools/sass_summary.py > profiles/rN_sass_summary.txt
Counts the mnemonics that show what the kernels are made of: UTMALDG (TMA tile loads), SYNCS
(mbarrier), LDS/STS, SHFL, STG.E.128 (128-bit stores), FADD2 (packed adds), UTMACCTL.PF / UBLKPF (tensor-map /
bulk L2 prefetch), PREEXIT (programmatic dependent launch), and FFMA / DFMA, whose absence is the
bit-exactness argument (every multiply and add rounds separately)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "physis_b200", "lib", "libphysis_rt_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["total"] += 1
        for key in ("UTMALDG", "SYNCS", "SHFL", "FADD2", "FFMA", "DFMA", "FMUL", "FADD", "DMUL", "DADD", "LDS", "STS",
                    "LDG", "UBLKPF", "MEMBAR", "PREEXIT"):
            if op == key or op.startswith(key + "."):
                kernels[cur][key] += 1
        if op.startswith("STG.E.128") or op.startswith("STG.E.EF.128"):
            kernels[cur]["STG.128"] += 1
        if op.startswith("UTMACCTL.PF") or op.startswith("UTMAPF"):
            kernels[cur]["TMAPF"] += 1
print(f"library: {os.path.relpath(lib, ROOT)}   cubin architectures: {', '.join(arch)}   kernels: {len(kernels)}")
cols = ["total", "UTMALDG", "TMAPF", "UBLKPF", "PREEXIT", "SYNCS", "LDS", "STS", "SHFL", "LDG", "STG.128", "FMUL", "FADD", "FADD2", "DMUL",
        "DADD", "FFMA", "DFMA"]
print(f"{'kernel':78s} " + " ".join(f"{c:>7s}" for c in cols))
tot_fma = 0
for name, c in kernels.items():
    d = demangle(name)
    d = re.sub(r"physis_b200::\(anonymous namespace\)::|physis_b200::", "", d)
    d = re.sub(r"\(CUtensorMap_st.*", "", d)
    d = re.sub(r"\((int|bool|PSReduceOp)\)", "", d)
    tot_fma += c["FFMA"] + c["DFMA"]
    print(f"{d[:78]:78s} " + " ".join(f"{c[k]:7d}" for k in cols))
print(f"\nFFMA + DFMA over all kernels: {tot_fma}")
