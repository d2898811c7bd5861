#!/usr/bin/env python
"""Config-5 kernel on z-slabs with the halo-exchange profile (run under torchrun; tuning tool)."""
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
n = int(os.environ.get("EXP_N", "512"))
count = int(os.environ.get("EXP_COUNT", "100"))
lib = physis_b200.load_programs()
lib.pstag_init.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.pstag_copyin_local.argtypes = [C.c_void_p, C.c_void_p]
lib.pstag_sweeps_only.argtypes = [C.c_int] * 4
configs = [c.split("+") if c else [] for c in os.environ.get("EXP_CONFIGS", "|debug_slab=1|debug_slab=2|debug_slab=3|early_signal=0").split("|")]
for cfg in configs:
    lib.pstag_init(0, None, n, n, n * world)
    for kv in cfg:
        api.set_option(kv)
    api.set_option("halo_profile=1")
    uo, ul, ko, kl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.pstag_local_size(C.byref(uo), C.byref(ul), C.byref(ko), C.byref(kl))
    u = np.zeros((ul.value * n * n, 2))
    u[:, 0] = np.random.default_rng(rank).random(ul.value * n * n)
    kap = np.full(kl.value * (n + 1) * (n + 1), 0.05)
    lib.pstag_copyin_local(u.ctypes.data, kap.ctypes.data)
    r = api.rt()
    lib.pstag_sweeps_only(20, n, n, n * world)
    r.__PSB200Synchronize()
    r.__PSB200ResetStats()
    r.__PSB200TimerStart()
    lib.pstag_sweeps_only(count, n, n, n * world)
    ms = r.__PSB200TimerStopMs()
    st = api.stats()
    ctas = max(int(st.halo_wait_ctas), 1)
    launches = max(int(st.halo_wait_launches), 1)
    print(f"rank {rank}/{world} pstag {' '.join(cfg) or 'default'}: {ms / count:.4f} ms/sweep {n ** 3 * count / ms / 1e6:.1f} GLUP/s per GPU | "
          f"wait mean/CTA {st.halo_wait_ns_sum / ctas / 1e3:.2f} us max {st.halo_wait_ns_max / 1e3:.1f} us, waiting CTAs/launch {ctas / launches:.0f}", flush=True)
    lib.pstag_finalize()
