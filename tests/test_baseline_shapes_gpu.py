"""GPU parity at the BASELINE.json shapes (not only at toy sizes): every hand-written sweep,
called through the C ABI, against the CPU oracle on the same inputs -- bit for bit.

  config 2   7-pt fp32 512^3, 100 sweeps, the benchmark's own coefficients: against the
             reference's OpenMP form of the sweep (examples/diffusion-benchmark/
             diffusion3d_openmp.cc compiled unmodified into oracle/_ref; pinned bit-identical to
             the oracle by tests/test_oracle.py) -- or, where oracle/_ref does not exist, against
             the oracle port at a reduced sweep count
  config 4   rows of 1024 floats (the x-tiled fused pass and the 8-box single sweep)
  config 3   Himeno size L (512x256x256) and an XL-wide case (1024 floats = 8 boxes per row)
  config 5   fp64 periodic staggered user type at 128^3 and at 512x512 planes
  PSReduce   all 4 operators x 4 element types, lengths 4k+1 / 4k+2 / 4k+3

The way the reference's own system tests compare (stdout of the translated program diffed
against a plain-C twin, tests/system_tests/run_system_tests.sh.cmake:833-883), at size.
"""
import ctypes as C

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64)


def _assert_same(want, got, what):
    w, g = _bits(want).ravel(), _bits(got).ravel()
    if not np.array_equal(w, g):
        bad = np.nonzero(w != g)[0]
        raise AssertionError(f"{what}: {bad.size} of {w.size} elements differ, first at {bad[:5]}")


def test_config2_512cubed_100_sweeps_vs_reference():
    n, count = 512, 100
    p = H.diffusion_params(n, n, n)
    f0 = H.diffusion_initial(n, n, n, p)
    ref = H.oracle_ref()
    if ref is not None:
        want = f0.copy()
        ref.ref_openmp_load.argtypes = [C.c_int] * 3 + [C.c_void_p]
        ref.ref_openmp_store.argtypes = [C.c_void_p]
        ref.ref_openmp_load(n, n, n, want.ctypes.data)
        ref.ref_openmp_sweeps(count)
        ref.ref_openmp_store(want.ctypes.data)
    else:
        count = 6
        want = H.run_diffusion(H.oracle_port(), f0, n, n, n, count, p)
    from physis_b200 import api
    got = H.run_diffusion(H.b200_programs(), f0, n, n, n, count, p)
    _assert_same(want, got, f"7-pt 512^3 x {count}")
    # the benchmark's accuracy figure against the analytic solution (baseline.cc:51-60)
    exact = H.diffusion_initial(n, n, n, p, time=float(np.float32(p[10]) * np.float32(count)))
    rms = float(np.sqrt(np.mean((got.astype(np.float64) - exact.astype(np.float64)) ** 2)))
    assert rms < 1e-5, rms
    del api


@pytest.mark.parametrize("shape,count", [((1024, 64, 64), 1), ((1024, 64, 64), 8), ((1024, 40, 21), 5),
                                          ((768, 33, 18), 6), ((640, 20, 12), 4), ((2048, 24, 9), 4)])
def test_wide_rows_match_oracle(shape, count):
    """Rows wider than one fused tile (config 4: 1024 floats)."""
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + ny)
    f0 = rng.random(nx * ny * nz, dtype=np.float32)
    for co in (np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44], np.float32),
               np.array([0.1234567] * 6 + [0.2592598], np.float32)):
        want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, co)
        got = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, count, co)
        _assert_same(want, got, f"7-pt {shape} x {count}")


@pytest.mark.parametrize("dims,nn", [((512, 256, 256), 4), ((1024, 32, 32), 4), ((1024, 47, 19), 2)])
def test_himeno_at_size(dims, nn):
    a = H.run_himeno(H.oracle_port(), dims, nn, gosa=True, seed=5)
    b = H.run_himeno(H.b200_programs(), dims, nn, gosa=True, seed=5)
    _assert_same(a[0], b[0], f"himeno p0 {dims}")
    _assert_same(a[1], b[1], f"himeno p1 {dims}")
    _assert_same(a[3], b[3], f"himeno ss^2 grid {dims}")
    exact = float(np.sum(a[3].astype(np.float64)))
    assert abs(b[2] - exact) <= 8e-6 * abs(exact), (b[2], exact)
    # and without the residual emit
    a = H.run_himeno(H.oracle_port(), dims, 2, gosa=False, seed=6)
    b = H.run_himeno(H.b200_programs(), dims, 2, gosa=False, seed=6)
    _assert_same(a[0], b[0], f"himeno p0 {dims}")
    _assert_same(a[1], b[1], f"himeno p1 {dims}")


@pytest.mark.parametrize("shape,count", [((128, 128, 128), 4), ((512, 512, 16), 3), ((256, 64, 40), 5)])
def test_periodic_staggered_at_size(shape, count):
    nx, ny, nz = shape
    u, kap = H.pstag_inputs(nx, ny, nz)
    want = H.run_pstag(H.oracle_port(), u, kap, nx, ny, nz, count)
    got = H.run_pstag(H.b200_programs(), u, kap, nx, ny, nz, count)
    _assert_same(want, got, f"config 5 {shape} x {count}")


# ---- PSReduce: every (type, operator), ragged lengths -----------------------------------------

_TYPES = [("Float", 2, np.float32), ("Double", 3, np.float64), ("Int", 0, np.int32), ("Long", 1, np.int64)]


def _reduce_data(dtype, op, n, rng):
    """Data whose reduction is exact in every association order, so the GPU tree and the
    REFERENCE target's sequential fold (runtime/libphysis_rt_ref.cc:19-30) agree bit for bit."""
    if op == 3:  # PROD: powers of two with bounded running exponent (floats), +-1 and a few 2s (ints)
        if np.issubdtype(dtype, np.floating):
            e = np.zeros(n, np.int64)
            k = min(n // 2, 40)
            idx = rng.permutation(n)[:2 * k]
            e[idx[:k]] = 1
            e[idx[k:]] = -1
            sign = np.where(rng.random(n) < 0.3, -1.0, 1.0)
            return (sign * np.exp2(e)).astype(dtype)
        d = np.where(rng.random(n) < 0.5, -1, 1).astype(dtype)
        d[rng.permutation(n)[:min(n, 20)]] = 2
        return d
    if np.issubdtype(dtype, np.floating):
        return rng.integers(-1000, 1000, n).astype(dtype) * dtype(0.25)   # sums stay exact
    return rng.integers(-100000, 100000, n).astype(dtype)


def _oracle_reduce(data, ptype, op, name):
    """PSInit / __PSGridNew / PSGridCopyin / __PSReduceGrid<T> on the oracle port (REF ABI)."""
    class TI(C.Structure):
        _fields_ = [("type", C.c_int), ("size", C.c_int), ("num_members", C.c_int), ("members", C.c_void_p)]
    lib = H.oracle_port()
    argc = C.c_int(1)
    argv = (C.c_char_p * 2)(b"test", None)
    pargv = C.pointer(argv)
    lib.PSInit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.PSInit(C.byref(argc), C.byref(pargv), 1)
    ti = TI(ptype, data.itemsize, 0, None)
    dims = (C.c_int * 3)(data.size, 0, 0)
    lib.__PSGridNew.restype = C.c_void_p
    lib.__PSGridNew.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    g = lib.__PSGridNew(C.byref(ti), 1, dims)
    lib.PSGridCopyin.argtypes = [C.c_void_p, C.c_void_p]
    lib.PSGridCopyin(g, data.ctypes.data)
    out = np.zeros(1, data.dtype)
    f = getattr(lib, "__PSReduceGrid" + name)
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    f(out.ctypes.data, op, g)
    lib.PSGridFree.argtypes = [C.c_void_p]
    lib.PSGridFree(g)
    lib.PSFinalize()
    return out[0]


@pytest.mark.parametrize("name,ptype,dtype", _TYPES)
@pytest.mark.parametrize("op", [0, 1, 2, 3])
def test_reduce_every_type_and_operator(name, ptype, dtype, op):
    from physis_b200 import api
    rng = np.random.default_rng(17 * op + ptype)
    for n in (1, 2, 3, 5, 1021, 4098, 65539, 1 << 20):
        data = _reduce_data(dtype, op, n, rng)
        want = _oracle_reduce(data, ptype, op, name)
        api.PSInit(["t"], 1, (n,))
        g = api.Grid((n,), ptype)
        g.copyin(data)
        got = g.reduce(op)
        g.free()
        api.PSFinalize()
        assert np.asarray(got, dtype).tobytes() == np.asarray(want, dtype).tobytes(), (name, op, n, got, want)


def test_reduce_3d_grids_odd_extents():
    from physis_b200 import api
    rng = np.random.default_rng(3)
    for shape in [(7, 5, 3), (33, 17, 9), (129, 3, 11)]:
        n = int(np.prod(shape))
        for name, ptype, dtype in _TYPES:
            data = _reduce_data(dtype, 2, n, rng)
            api.PSInit(["t"], 3, shape)
            g = api.Grid(shape, ptype)
            g.copyin(data)
            got = [g.reduce(op) for op in (0, 1, 2)]
            g.free()
            api.PSFinalize()
            assert got[0] == data.max() and got[1] == data.min()
            assert np.asarray(got[2], dtype).tobytes() == np.asarray(data.sum(dtype=dtype), dtype).tobytes()
