#!/usr/bin/env python
"""Time the fused two-sweep pass against the single-sweep kernel (tuning tool, GPU box only)."""
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
shapes = [(512, 512, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
count = int(os.environ.get("EXP_COUNT", "200"))
configs = [("star7_fuse=0",)]
if os.environ.get("EXP_CONFIGS") == "short":
    configs += [("star7_fuse=1",), ("star7_fuse=1", "star7_pair_zc=64")]
else:
    for zc in (0, 32, 64, 128, 256):
        configs.append(("star7_fuse=1", f"star7_pair_zc={zc}"))
for (nx, ny, nz) in shapes:
    f0 = np.random.default_rng(0).random(nx * ny * nz, dtype=np.float32)
    for cfg in configs:
        lib.initialize_physis(0, None, nx, ny, nz)
        for kv in cfg:
            api.set_option(kv)
        lib.initialize_benchmark_physis(nx, ny, nz)
        lib.copyin_physis(f0.ctypes.data)
        r = api.rt()
        lib.run_sweeps_only_physis(20, nx, ny, nz, *co)
        r.__PSB200TimerStart()
        lib.run_sweeps_only_physis(count, nx, ny, nz, *co)
        ms = r.__PSB200TimerStopMs() / count
        print(f"{nx}x{ny}x{nz} {' '.join(cfg)}: {ms:.4f} ms/sweep "
              f"{nx * ny * nz / ms / 1e6:.0f} GLUP/s {8.0 * nx * ny * nz / ms / 1e6:.0f} GB/s(alg)", flush=True)
        lib.finalize_benchmark_physis()
