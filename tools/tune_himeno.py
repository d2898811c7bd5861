#!/usr/bin/env python
"""Sweep the Himeno kernel's tile height / ring depth / z-chunk / occupancy on the GPU box.
Writes gpurun_out/tune_himeno.csv.  Tuning tool only — not on any product path."""
import ctypes as C
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physis_b200
from physis_b200 import api

size = sys.argv[1] if len(sys.argv) > 1 else "XL"
mi, mj, mk = {"XL": (1024, 512, 512), "L": (512, 256, 256)}[size]
nn = 10
lib = physis_b200.load_programs()
lib.himeno_init.argtypes = [C.c_int] * 3
lib.himeno_init(mi, mj, mk)
lib.himeno_sweeps_only.argtypes = [C.c_int, C.c_int]
r = api.rt()
pts = (mi - 2) * (mj - 2) * (mk - 2)
rows = []
for by, st, zc, occ, gosa in itertools.product([7, 15], [4, 6, 8], [0, 32, 64], [0, 1], [0]):
    api.set_option(f"himeno_by={by}")
    api.set_option(f"himeno_stages={st}")
    api.set_option(f"himeno_zc={zc}")
    api.set_option(f"himeno_occ={occ}")
    lib.himeno_sweeps_only(2, gosa)
    r.__PSB200TimerStart()
    lib.himeno_sweeps_only(nn, gosa)
    ms = r.__PSB200TimerStopMs() / nn
    rows.append((by, st, zc, occ, gosa, ms, pts * 56 / ms / 1e6))
rows.sort(key=lambda x: -x[-1])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"tune_himeno_{size}.csv"), "w") as f:
    f.write("by,stages,zc,occ,gosa,ms_per_sweep,alg_GBps\n")
    for row in rows:
        f.write(",".join(str(x) for x in row) + "\n")
for row in rows[:12]:
    print(row)
lib.himeno_finalize()
