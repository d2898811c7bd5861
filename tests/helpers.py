"""Shared test plumbing: oracle loading (CPU checker) and program drivers.

Only tests (and bench.py's cpu_baseline leg) touch oracle/.  The same driver
functions run a program from either side — `lib` is the oracle port, the
reference build (oracle/_ref) or the b200 programs library — because all three
export the same entry points (they are translations of the same DSL source).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

_cache = {}


def oracle_port():
    """oracle/liboracle.so — the CPU restatement; built on demand (gcc only)."""
    if "port" not in _cache:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
        _cache["port"] = C.CDLL(path)
    return _cache["port"]


def oracle_ref():
    """oracle/_ref/libphysis_ref.so — the reference's own runtime; None if never built."""
    if "ref" not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libphysis_ref.so")
        if not os.path.exists(path) and os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
        _cache["ref"] = C.CDLL(path) if os.path.exists(path) else None
    return _cache["ref"]


def b200_programs():
    import physis_b200
    return physis_b200.load_programs()


# ---- diffusion benchmark (examples/diffusion-benchmark) ----------------------

def diffusion_params(nx, ny, nz):
    out = np.zeros(15, np.float32)
    f = oracle_port().oracle_diffusion3d_params
    f.argtypes = [C.c_int] * 3 + [C.c_void_p]
    f(nx, ny, nz, out.ctypes.data)
    return out


def diffusion_initial(nx, ny, nz, p=None, time=0.0):
    p = diffusion_params(nx, ny, nz) if p is None else p
    buf = np.zeros(nx * ny * nz, np.float32)
    f = oracle_port().oracle_diffusion3d_initialize
    f.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_float] * 8
    f(buf.ctypes.data, nx, ny, nz, p[12], p[13], p[14], p[7], p[8], p[9], p[11], time)
    return buf


def run_diffusion(lib, field, nx, ny, nz, count, coeffs, entry="run_kernel_physis"):
    """PSInit .. run_kernel_physis(count) .. PSFinalize; returns the final field."""
    f = np.array(field, dtype=np.float32, copy=True)
    lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.initialize_physis(0, None, nx, ny, nz)
    lib.initialize_benchmark_physis.argtypes = [C.c_int] * 3
    lib.initialize_benchmark_physis(nx, ny, nz)
    fn = getattr(lib, entry)
    fn.argtypes = [C.c_int, C.c_void_p] + [C.c_int] * 3 + [C.c_float] * 7
    fn(count, f.ctypes.data, nx, ny, nz, *[float(c) for c in coeffs[:7]])
    lib.finalize_benchmark_physis()
    return f


# ---- Himeno -------------------------------------------------------------------

HIMENO_GRIDS = ["P0", "P1", "BND", "WRK1", "A0", "A1", "A2", "A3", "B0", "B1", "B2", "C0", "C1", "C2", "GOSA"]


def run_himeno(lib, dims, nn, gosa=False, seed=None, omega=None, each=False, before_finalize=None,
               p1_differs=False):
    """himeno_init + optional random coefficient fields + jacobi(nn); returns (p0, p1, gosa)."""
    mi, mj, mk = dims
    ne = mi * mj * mk
    lib.himeno_init.argtypes = [C.c_int] * 3
    lib.himeno_init(mi, mj, mk)
    lib.himeno_set_grid.argtypes = [C.c_int, C.c_void_p]
    lib.himeno_get_grid.argtypes = [C.c_int, C.c_void_p]
    if seed is not None:
        rng = np.random.default_rng(seed)
        for g in range(14):
            b = rng.random(ne, dtype=np.float32)
            if g in (0, 1) and not p1_differs:   # the benchmark starts p0 and p1 identical
                if g == 0:                        # (boundary cells persist)
                    p_init = b
                b = p_init
            lib.himeno_set_grid(g, b.ctypes.data)
    if omega is not None:
        lib.himeno_set_omega.argtypes = [C.c_float]
        lib.himeno_set_omega(omega)
    # each: the original benchmark's structure, a PSReduce after every iteration of the pair
    f = (lib.himeno_jacobi_gosa_each if each else lib.himeno_jacobi_gosa) if gosa else lib.himeno_jacobi
    f.argtypes = [C.c_int]
    f.restype = C.c_float
    g = f(nn)
    p0 = np.zeros(ne, np.float32)
    p1 = np.zeros(ne, np.float32)
    lib.himeno_get_grid(0, p0.ctypes.data)
    lib.himeno_get_grid(1, p1.ctypes.data)
    gg = np.zeros(ne, np.float32)
    lib.himeno_get_grid(14, gg.ctypes.data)
    if before_finalize is not None:
        before_finalize()
    lib.himeno_finalize()
    return p0, p1, g, gg


# ---- config 5: periodic staggered user-type diffusion ---------------------------

def pstag_inputs(nx, ny, nz, seed=3):
    rng = np.random.default_rng(seed)
    u = rng.random((nx * ny * nz, 2))
    kap = rng.random((nx + 1) * (ny + 1) * (nz + 1)) * 0.1
    return u, kap


def run_pstag(lib, u, kap, nx, ny, nz, count):
    u = np.array(u, dtype=np.float64, copy=True)
    lib.pstag_init.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.pstag_init(0, None, nx, ny, nz)
    lib.pstag_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.pstag_run(count, u.ctypes.data, np.ascontiguousarray(kap).ctypes.data, nx, ny, nz)
    lib.pstag_finalize()
    return u


# ---- numpy restatement of the clamped 7-point update (fp32 / fp64) ------------

def diffusion7_numpy(field, shape, coeffs, steps):
    """`steps` sweeps of kernel_physis (examples/diffusion-benchmark/diffusion3d_physis.c:29-58)
    in the REFERENCE target's evaluation order: separately rounded products summed left to
    right, a neighbour outside the grid replaced by the centre.  numpy's elementwise
    multiply and add round once each (no FMA contraction), so this is bit-identical to the
    C oracle (checked on CPU by tests/test_oracle.py).  coeffs: ce, cw, cn, cs, ct, cb, cc."""
    nx, ny, nz = shape
    dt = np.asarray(field).dtype.type
    ce, cw, cn, cs, ct, cb, cc = [dt(c) for c in coeffs[:7]]
    f = np.array(field, copy=True).reshape(nz, ny, nx)
    for _ in range(steps):
        w = np.concatenate([f[:, :, :1], f[:, :, :-1]], axis=2)
        e = np.concatenate([f[:, :, 1:], f[:, :, -1:]], axis=2)
        n = np.concatenate([f[:, :1, :], f[:, :-1, :]], axis=1)
        s = np.concatenate([f[:, 1:, :], f[:, -1:, :]], axis=1)
        b = np.concatenate([f[:1], f[:-1]], axis=0)
        t = np.concatenate([f[1:], f[-1:]], axis=0)
        r = cc * f
        r = r + cw * w
        r = r + ce * e
        r = r + cs * s
        r = r + cn * n
        r = r + cb * b
        r = r + ct * t
        f = r
    return f.reshape(-1)
