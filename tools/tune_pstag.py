#!/usr/bin/env python
"""Sweep the config-5 kernel's tile shape / ring depth on the GPU box.  Tuning tool only."""
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
lib = physis_b200.load_programs()
lib.pstag_init.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.pstag_init(0, None, n, n, n)
u = np.zeros((n ** 3, 2))
u[:, 0] = np.random.default_rng(0).random(n ** 3)
kap = np.full((n + 1) ** 3, 0.05)
lib.pstag_copyin_local.argtypes = [C.c_void_p, C.c_void_p]
lib.pstag_copyin_local(u.ctypes.data, kap.ctypes.data)
lib.pstag_sweeps_only.argtypes = [C.c_int] * 4
r = api.rt()
rows = []
VARIANTS = [int(x) for x in os.environ.get("PSTAG_VARIANTS", "0,1").split(",")]
STAGES = [int(x) for x in os.environ.get("PSTAG_STAGES", "3,6").split(",")]
OCCS = [int(x) for x in os.environ.get("PSTAG_OCCS", "0,1,2,3,4").split(",")]
for v, st, occ in itertools.product(VARIANTS, STAGES, OCCS):
    api.set_option(f"pstag_variant={v}")
    api.set_option(f"pstag_stages={st}")
    api.set_option(f"pstag_occ={occ}")
    try:
        lib.pstag_sweeps_only(4, n, n, n)
        r.__PSB200TimerStart()
        lib.pstag_sweeps_only(20, n, n, n)
        ms = r.__PSB200TimerStopMs() / 20
    except Exception as e:  # noqa
        print("fail", v, st, e)
        continue
    rows.append((v, st, occ, ms, n ** 3 * 24 / ms / 1e6))
rows.sort(key=lambda x: -x[-1])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"tune_pstag_{n}.csv"), "w") as f:
    f.write("variant,stages,occ,ms_per_sweep,alg_GBps\n")
    for row in rows:
        f.write(",".join(str(x) for x in row) + "\n")
for row in rows[:14]:
    print(row)
lib.pstag_finalize()
