#!/usr/bin/env python
"""Per-rank timing of the fused pass on z-slabs (run under torchrun; tuning tool only)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physis_b200
from physis_b200 import api

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
lib = physis_b200.load_programs()
lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.copyin_local_physis.argtypes = [C.c_void_p]
lib.run_sweeps_only_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
co = [0.1] * 6 + [0.4]
n = 512
count = 400
configs = [("star7_fuse=0",), ("star7_fuse=1",), ("star7_fuse=1", "star7_iso=0")]
for cfg in configs:
    lib.initialize_physis(0, None, n, n, n * world)
    for kv in cfg:
        api.set_option(kv)
    lib.initialize_benchmark_physis(n, n, n * world)
    f0 = np.random.default_rng(rank).random(n * n * n, dtype=np.float32)
    lib.copyin_local_physis(f0.ctypes.data)
    r = api.rt()
    lib.run_sweeps_only_physis(20, n, n, n * world, *co)
    r.__PSB200Synchronize()
    r.__PSB200TimerStart()
    lib.run_sweeps_only_physis(count, n, n, n * world, *co)
    ms = r.__PSB200TimerStopMs() / count
    print(f"rank {rank}/{world} {' '.join(cfg)}: {ms:.4f} ms/sweep {n ** 3 / ms / 1e6:.0f} GLUP/s per GPU", flush=True)
    lib.finalize_benchmark_physis()
