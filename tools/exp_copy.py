#!/usr/bin/env python
"""PSGridCopyin / PSGridCopyout throughput: pageable host memory through the
pinned double-buffered staging, pinned host memory directly, and user types (host AoS <-> device SoA,
transposed on the GPU).  Measurement tool, GPU box only."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from physis_b200 import api

api.PSInit(dims=(512, 512, 512))
r = api.rt()
for kv in sys.argv[1:]:
    api.set_option(kv)
    print("option", kv)


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def report(name, nbytes, t_in, t_out):
    print(f"{name:44s} copyin {nbytes / t_in / 1e9:6.1f} GB/s   copyout {nbytes / t_out / 1e9:6.1f} GB/s", flush=True)


# primitive grid, 512 MiB
g = api.Grid((512, 512, 512), api.PS_FLOAT)
nb = 512 ** 3 * 4
page = np.random.default_rng(0).random(512 ** 3, dtype=np.float32)
out = np.empty_like(page)
report("float 512^3, pageable host (staged)", nb,
       timed(lambda: r.PSGridCopyin(g.ptr, page.ctypes.data)),
       timed(lambda: r.PSGridCopyout(g.ptr, out.ctypes.data)))
assert np.array_equal(page, out)
pin, pin_ptr = api.pinned_empty(nb, np.float32)
pin[:] = page
report("float 512^3, pinned host (direct DMA)", nb,
       timed(lambda: r.PSGridCopyin(g.ptr, pin.ctypes.data)),
       timed(lambda: r.PSGridCopyout(g.ptr, pin.ctypes.data)))
g.free()

# user type {double p, q}: host AoS, device SoA; 512x512x256 cells = 1 GiB
u = api.Grid((512, 512, 256), members=[(api.PS_DOUBLE, ()), (api.PS_DOUBLE, ())])
nb = 512 * 512 * 256 * 16
aos = np.random.default_rng(1).random(512 * 512 * 256 * 2)
back = np.empty_like(aos)
report("struct{double p,q} 512x512x256, pageable", nb,
       timed(lambda: r.PSGridCopyin(u.ptr, aos.ctypes.data)),
       timed(lambda: r.PSGridCopyout(u.ptr, back.ctypes.data)))
assert np.array_equal(aos, back)
pin2, pin2_ptr = api.pinned_empty(nb, np.float64)
pin2[:] = aos
report("struct{double p,q} 512x512x256, pinned", nb,
       timed(lambda: r.PSGridCopyin(u.ptr, pin2.ctypes.data)),
       timed(lambda: r.PSGridCopyout(u.ptr, pin2.ctypes.data)))
u.free()
api.PSFinalize()
