#!/usr/bin/env python
"""Per-rank timing of the fused pass on z-slabs with the halo-exchange profile (how long the
sweeps' CTAs wait for the ring neighbours).  Run under torchrun; tuning tool only."""
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
n = int(os.environ.get("EXP_N", "512"))
nzr = int(os.environ.get("EXP_NZ", str(n)))   # planes per rank
count = int(os.environ.get("EXP_COUNT", "400"))
configs = [c.split("+") for c in os.environ.get(
    "EXP_CONFIGS", "star7_fuse=1|star7_fuse=1+star7_pair_zbl=0|star7_fuse=1+early_signal=0|star7_fuse=0").split("|")]
for cfg in configs:
    lib.initialize_physis(0, None, n, n, nzr * world)
    for kv in cfg:
        api.set_option(kv)
    api.set_option("halo_profile=1")
    lib.initialize_benchmark_physis(n, n, nzr * world)
    f0 = np.random.default_rng(rank).random(n * n * nzr, dtype=np.float32)
    lib.copyin_local_physis(f0.ctypes.data)
    r = api.rt()
    lib.run_sweeps_only_physis(40, n, n, nzr * world, *co)
    r.__PSB200Synchronize()
    r.__PSB200ResetStats()
    r.__PSB200TimerStart()
    lib.run_sweeps_only_physis(count, n, n, nzr * world, *co)
    ms = r.__PSB200TimerStopMs()
    st = api.stats()
    launches = max(int(st.halo_wait_launches), 1)
    ctas = max(int(st.halo_wait_ctas), 1)
    print(f"rank {rank}/{world} {' '.join(cfg)}: {ms / count:.4f} ms/sweep {n * n * nzr * count / ms / 1e6:.0f} GLUP/s per GPU | "
          f"launches {int(st.kernel_launches)} fused {int(st.fused_pairs)} | wait: mean/CTA {st.halo_wait_ns_sum / ctas / 1e3:.2f} us, "
          f"max {st.halo_wait_ns_max / 1e3:.1f} us, waiting CTAs/launch {ctas / launches:.0f}, "
          f"launch ms {ms / max(int(st.kernel_launches), 1):.4f}", flush=True)
    lib.finalize_benchmark_physis()
