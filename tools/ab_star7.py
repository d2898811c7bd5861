#!/usr/bin/env python
"""A/B of star-7 kernel forms on one GPU: interleaved repeats so power/clock drift hits all alike.
usage: ab_star7.py n key=v1,v2,... [fixed=opt ...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physis_b200
from physis_b200 import api

n = int(sys.argv[1])
key, vals = sys.argv[2].split("=")
vals = vals.split(",")
fixed = sys.argv[3:]
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
for kv in fixed:
    api.set_option(kv)
sweeps = 200
res = {v: [] for v in vals}
for rep in range(6):
    for v in vals:
        api.set_option(f"{key}={v}")
        lib.run_sweeps_only_physis(10, n, n, n, *co)
        r.__PSB200TimerStart()
        lib.run_sweeps_only_physis(sweeps, n, n, n, *co)
        ms = r.__PSB200TimerStopMs() / sweeps
        res[v].append(ms)
for v in vals:
    a = np.array(res[v])
    print(f"{key}={v}: ms/sweep median {np.median(a):.4f} min {a.min():.4f} max {a.max():.4f}  "
          f"-> {8.0 * n ** 3 / np.median(a) / 1e6:.0f} GB/s")
lib.finalize_benchmark_physis()
