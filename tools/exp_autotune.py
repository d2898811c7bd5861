"""What option autotune=1 settles on for a few shapes, and the rate after it (one GPU).
Usage: python tools/exp_autotune.py  (under gpurun)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from physis_b200 import api  # noqa: E402


def run(shape, count, tune):
    nx, ny, nz = shape
    api.PSInit(["t"], 3, shape)
    api.set_option(f"autotune={tune}")
    a, b = api.Grid(shape, api.PS_FLOAT), api.Grid(shape, api.PS_FLOAT)
    a.copyin(np.random.default_rng(1).random(nx * ny * nz, dtype=np.float32))
    dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
    co = [float(np.float32(0.1))] * 6 + [float(np.float32(0.4))]
    d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
    d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
    api.stencil_run(count // 2, [d0, d1])   # tunes (or warms up)
    best = 1e9
    for _ in range(3):
        api.rt().__PSB200Synchronize()
        t0 = time.perf_counter()
        api.stencil_run(count // 2, [d0, d1])
        api.rt().__PSB200Synchronize()
        best = min(best, time.perf_counter() - t0)
    what = api.last_tuning()
    api.PSFinalize()
    return nx * ny * nz * count / best / 1e9, what


def run_himeno(dims, nn, tune):
    import ctypes as C
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import helpers as H
    lib = H.b200_programs()
    os.environ["PHYSIS_B200_OPTIONS"] = f"autotune={tune}"
    lib.himeno_init_local.argtypes = [C.c_int] * 3
    lib.himeno_init_local(*dims)
    lib.himeno_jacobi.argtypes = [C.c_int]
    lib.himeno_jacobi.restype = C.c_float
    lib.himeno_jacobi(100)
    api.rt().__PSB200Synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        lib.himeno_jacobi(100)
        api.rt().__PSB200Synchronize()
        best = min(best, time.perf_counter() - t0)
    what = api.last_tuning()
    lib.himeno_finalize()
    return (dims[0] - 2) * (dims[1] - 2) * (dims[2] - 2) * 100 / best / 1e9, what


if __name__ == "__main__":
    for dims in [] if os.environ.get("EXP_SHAPES") else [(128, 64, 64), (256, 128, 128), (512, 256, 256), (1024, 512, 512)]:
        g0, _ = run_himeno(dims, 100, 0)
        g1, what = run_himeno(dims, 100, 1)
        print(f"himeno {dims}: defaults {g0:.1f} GLUP/s, autotune {g1:.1f} GLUP/s  [{what}]", flush=True)
    shapes = os.environ.get("EXP_SHAPES")
    if shapes:
        for sh in shapes.split("|"):
            shape = tuple(int(v) for v in sh.split("x"))
            g0, _ = run(shape, 400, 0)
            g1, what = run(shape, 400, 1)
            print(f"{shape}: defaults {g0:.1f} GLUP/s, autotune {g1:.1f} GLUP/s  [{what}]", flush=True)
        sys.exit(0)
    for shape in [(64, 64, 64), (128, 128, 128), (256, 256, 256), (384, 384, 384), (512, 512, 512), (256, 512, 1024),
                  (640, 640, 640), (1024, 256, 256)]:
        g0, _ = run(shape, 1000, 0)
        g1, what = run(shape, 1000, 1)
        print(f"{shape}: defaults {g0:.1f} GLUP/s, autotune {g1:.1f} GLUP/s  [{what}]", flush=True)
