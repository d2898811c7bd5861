#!/usr/bin/env python
"""Himeno XL timing under a list of option sets (tuning tool, GPU box only):
EXP_CONFIGS="a=1+b=2|c=3" EXP_MODES=sweeps,each python tools/exp_himeno.py [XL|L|M|S] [nn]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physis_b200
from physis_b200 import api

size = sys.argv[1] if len(sys.argv) > 1 else "XL"
nn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mi, mj, mk = {"XL": (1024, 512, 512), "L": (512, 256, 256), "M": (256, 128, 128), "S": (128, 64, 64)}.get(size) or tuple(int(v) for v in size.split("x"))
lib = physis_b200.load_programs()
lib.himeno_init_local.argtypes = [C.c_int] * 3
lib.himeno_sweeps_only.argtypes = [C.c_int, C.c_int]
lib.himeno_jacobi_gosa_each.argtypes = [C.c_int]
lib.himeno_jacobi_gosa_each.restype = C.c_float
pts = (mi - 2) * (mj - 2) * (mk - 2)
configs = [c.split("+") if c else [] for c in os.environ.get(
    "EXP_CONFIGS", "|himeno_fuse=0|himeno_pair_pf=1|himeno_pair_pf=4|himeno_pair_zc=64|himeno_pair_zc=32").split("|")]
for cfg in configs:
    lib.himeno_init_local(mi, mj, mk)
    for kv in cfg:
        api.set_option(kv)
    r = api.rt()
    for mode in os.environ.get("EXP_MODES", "sweeps,each").split(","):
        run = (lambda: lib.himeno_sweeps_only(nn, 0)) if mode == "sweeps" else (lambda: lib.himeno_jacobi_gosa_each(nn))
        run()
        r.__PSB200Synchronize()
        r.__PSB200ResetStats()
        r.__PSB200TimerStart()
        run()
        ms = r.__PSB200TimerStopMs()
        st = api.stats()
        print(f"{size} {' '.join(cfg) or 'default'} [{mode}]: {ms / nn:.4f} ms/sweep {pts * nn / ms / 1e6:.1f} GLUP/s "
              f"fused passes {int(st.fused_pairs)} launches {int(st.kernel_launches)}", flush=True)
    lib.himeno_finalize()
