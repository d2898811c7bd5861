#!/usr/bin/env python
"""Sweep the star-7 kernel's tile shapes / ring depth / z-chunking on the GPU box.
Writes gpurun_out/tune_star7.csv.  Tuning tool only — not on any product path."""
import ctypes as C
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physis_b200
from physis_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dtype = sys.argv[2] if len(sys.argv) > 2 else "f32"
sweeps = 40
lib = physis_b200.load_programs()
lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.initialize_physis(0, None, n, n, n)
lib.initialize_benchmark_physis(n, n, n)
f0 = np.random.default_rng(0).random(n ** 3, dtype=np.float32)
lib.copyin_physis.argtypes = [C.c_void_p]
lib.copyin_physis(f0.ctypes.data)
lib.run_sweeps_only_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
co = [0.1] * 6 + [0.4]
r = api.rt()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
rows = []
variants = [int(v) for v in os.environ.get('TUNE_VARIANTS', ','.join(str(i) for i in range(6))).split(',')]
stages = [4, 5, 6, 8]
zcs = [16, 32, 64, 128]
hints = [(0, 0), (1, 0), (0, 1), (1, 1)]


def measure(v, s, zc, l2, st, occ=0):
    api.set_option(f"star7_variant={v}")
    api.set_option(f"star7_stages={s}")
    api.set_option(f"star7_zc={zc}")
    api.set_option(f"star7_sthint={st}")
    api.set_option(f"star7_occ={occ}")
    lib.run_sweeps_only_physis(4, n, n, n, *co)
    r.__PSB200TimerStart()
    lib.run_sweeps_only_physis(sweeps, n, n, n, *co)
    ms = r.__PSB200TimerStopMs() / sweeps
    gbs = 8.0 * n ** 3 / ms / 1e6
    return ms, gbs


best = None
for v, s, zc in itertools.product(variants, stages, zcs):
    try:
        ms, gbs = measure(v, s, zc, 0, 0)
    except Exception as e:  # noqa
        print("fail", v, s, zc, e)
        continue
    rows.append((v, s, zc, 0, 0, 0, ms, gbs))
    if best is None or gbs > best[-1]:
        best = rows[-1]
rows.sort(key=lambda x: -x[-1])
top = rows[:6]
for (v, s, zc, _, _, _, _, _) in top:
    for l2, st in hints[1:]:
        ms, gbs = measure(v, s, zc, l2, st)
        rows.append((v, s, zc, l2, st, 0, ms, gbs))
    for occ in (1, 2, 3):
        ms, gbs = measure(v, s, zc, 0, 0, occ)
        rows.append((v, s, zc, 0, 0, occ, ms, gbs))
rows.sort(key=lambda x: -x[-1])
with open(os.path.join(ROOT, "gpurun_out", f"tune_star7_{n}.csv"), "w") as f:
    f.write("variant,stages,zc,l2hint,sthint,occ,ms_per_sweep,alg_GBps\n")
    for row in rows:
        f.write(",".join(str(x) for x in row) + "\n")
for row in rows[:15]:
    print(row)
lib.finalize_benchmark_physis()
