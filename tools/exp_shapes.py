#!/usr/bin/env python
"""Time the star-7 sweep on a few grid shapes / tile variants (tuning tool, GPU box only)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physis_b200
from physis_b200 import api

lib = physis_b200.load_programs()
lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.copyin_physis.argtypes = [C.c_void_p]
lib.run_sweeps_only_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
co = [0.1] * 6 + [0.4]
shapes = [(1024, 1024, 128), (1024, 1024, 512), (512, 512, 512), (2048, 512, 128)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
first = True
for (nx, ny, nz) in shapes:
    if first:
        lib.initialize_physis(0, None, nx, ny, nz)
        first = False
    r = api.rt()
    for variant in [-1] + [int(v) for v in os.environ.get("EXP_VARIANTS", "0,2,3,5").split(",")]:
        for zc in (0, 16, 32, 64):
            api.set_option(f"star7_variant={variant}")
            api.set_option(f"star7_zc={zc}")
            lib.initialize_benchmark_physis(nx, ny, nz)
            f0 = np.random.default_rng(0).random(nx * ny * nz, dtype=np.float32)
            lib.copyin_physis(f0.ctypes.data)
            try:
                lib.run_sweeps_only_physis(4, nx, ny, nz, *co)
                r.__PSB200TimerStart()
                lib.run_sweeps_only_physis(40, nx, ny, nz, *co)
                ms = r.__PSB200TimerStopMs() / 40
                print(f"{nx}x{ny}x{nz} variant={variant} zc={zc}: {ms:.4f} ms/sweep "
                      f"{8.0 * nx * ny * nz / ms / 1e6:.0f} GB/s", flush=True)
            finally:
                lib.finalize_benchmark_physis()
