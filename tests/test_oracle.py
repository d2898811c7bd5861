"""CPU: pins the oracle (oracle/) against the reference's own code and golden outputs.

  * fixtures in tests/golden/reference_golden.json were produced by the reference itself
    (tests/golden/make_golden.py); where oracle/_ref exists they are re-derived and compared
  * the plain-C port of the REF runtime + hand-emitted `physisc --ref` programs
    (oracle/liboracle.so) must equal, bit for bit,
      - the reference's own diffusion `Baseline` (examples/diffusion-benchmark/baseline.cc)
      - the reference's original Himeno (examples/himeno/himenobmtxpa_original.c)
      - the same programs linked against the reference's unmodified libphysis_rt_ref sources
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import helpers as H

with open(os.path.join(H.GOLDEN_DIR, "reference_golden.json")) as f:
    GOLD = json.load(f)

REF_GOLDEN_BIN = os.path.join(H.ORACLE_DIR, "_ref", "golden")
has_ref_bins = os.path.isdir(REF_GOLDEN_BIN)
needs_ref = pytest.mark.skipif(H.oracle_ref() is None, reason="oracle/_ref not built (no /root/reference)")


def fmt_f(a):
    """printf("%f\\n") of every element, as the reference's dump() does."""
    return "".join("%f\n" % float(v) for v in a)


def sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


@pytest.mark.skipif(not has_ref_bins, reason="reference golden programs not built here")
def test_committed_fixtures_match_reference_programs():
    names = sorted(n for n in os.listdir(REF_GOLDEN_BIN) if n.startswith("test_"))
    assert len(names) == 36 == len(GOLD["system_tests"])
    for n in names:
        out = subprocess.run([os.path.join(REF_GOLDEN_BIN, n)], capture_output=True, text=True, check=True).stdout
        assert sha(out) == GOLD["system_tests"][n]["sha256"], n


@pytest.mark.parametrize("n,count", [(32, 10), (64, 20)])
def test_port_diffusion_equals_reference_baseline(n, count):
    p = H.diffusion_params(n, n, n)
    f0 = H.diffusion_initial(n, n, n, p)
    got = H.run_diffusion(H.oracle_port(), f0, n, n, n, count, p)
    g = GOLD["diffusion_baseline"][f"{n}x{count}"]
    assert hashlib.sha256(got.tobytes()).hexdigest() == g["sha256"]
    assert "%.9e" % float(np.sum(got, dtype=np.float64)) == g["sum"]


@needs_ref
def test_port_setup_equals_reference_setup():
    ref = H.oracle_ref()
    for n in (16, 64, 100):
        p = H.diffusion_params(n, n, n)
        q = np.zeros(15, np.float32)
        ref.ref_diffusion3d_params.argtypes = [C.c_int] * 3 + [C.c_void_p]
        ref.ref_diffusion3d_params(n, n, n, q.ctypes.data)
        assert np.array_equal(p.view(np.uint32), q.view(np.uint32))
        a = H.diffusion_initial(n, n, n, p)
        b = np.zeros(n ** 3, np.float32)
        ref.ref_diffusion3d_initialize.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_float] * 8
        ref.ref_diffusion3d_initialize(b.ctypes.data, n, n, n, p[12], p[13], p[14], p[7], p[8], p[9], p[11], 0.0)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@needs_ref
@pytest.mark.parametrize("shape,count", [((32, 32, 32), 6), ((40, 12, 7), 4), ((5, 3, 2), 2)])
def test_port_equals_ref_runtime_diffusion(shape, count):
    nx, ny, nz = shape
    rng = np.random.default_rng(1)
    f0 = rng.random(nx * ny * nz, dtype=np.float32)
    co = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44], np.float32)
    a = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, co)
    b = H.run_diffusion(H.oracle_ref(), f0, nx, ny, nz, count, co)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # and the reference's Baseline on the same arbitrary field (symmetric coefficients only:
    # Baseline takes its own) -- covered by test_port_diffusion_equals_reference_baseline


@needs_ref
@pytest.mark.parametrize("dims,nn,gosa", [((32, 16, 16), 4, True), ((20, 9, 6), 2, False)])
def test_port_equals_ref_runtime_himeno(dims, nn, gosa):
    a = H.run_himeno(H.oracle_port(), dims, nn, gosa=gosa, seed=2)
    b = H.run_himeno(H.oracle_ref(), dims, nn, gosa=gosa, seed=2)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x, np.float32).view(np.uint32), np.asarray(y, np.float32).view(np.uint32))


@needs_ref
def test_port_equals_ref_runtime_periodic_staggered():
    nx, ny, nz = 12, 8, 5
    u, kap = H.pstag_inputs(nx, ny, nz)
    a = H.run_pstag(H.oracle_port(), u, kap, nx, ny, nz, 3)
    b = H.run_pstag(H.oracle_ref(), u, kap, nx, ny, nz, 3)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_port_himeno_equals_reference_original_benchmark():
    """himenobmtxpa_original.c XS (32x32x64, k fastest) dumps p after 4 sweeps; the Physis
    version is the same arithmetic with x = k (64x32x32, x fastest)."""
    p0, p1, g, _ = H.run_himeno(H.oracle_port(), (64, 32, 32), 4)
    assert sha(fmt_f(p0)) == GOLD["himeno_original_XS"]["sha256"]


def test_port_himeno_gosa_close_to_original():
    # the original accumulates gosa sequentially in fp32 inside the sweep; the DSL form emits
    # ss*ss and PSReduce()s it sequentially: the same addends in the same order
    _, _, g, _ = H.run_himeno(H.oracle_port(), (64, 32, 32), 4, gosa=True)
    want = float(GOLD["himeno_original_XS"]["rehearsal_gosa"])
    assert abs(g - want) <= 5e-7 * want  # printed with 7 significant digits


def _runtime_reduce(lib, data, ptype, op, name, ref_abi):
    """PSInit / __PSGridNew / PSGridCopyin / __PSReduceGrid<T> through a REF-ABI runtime."""
    class TI(C.Structure):
        _fields_ = [("type", C.c_int), ("size", C.c_int), ("num_members", C.c_int), ("members", C.c_void_p)]
    argc = C.c_int(1)
    argv = (C.c_char_p * 2)(b"test", None)
    pargv = C.pointer(argv)
    lib.PSInit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.PSInit(C.byref(argc), C.byref(pargv), 3)
    ti = TI(ptype, data.itemsize, 0, None)
    dims = (C.c_int * 3)(*data.shape[::-1])
    lib.__PSGridNew.restype = C.c_void_p
    lib.__PSGridNew.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    g = lib.__PSGridNew(C.byref(ti), 3, dims)
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


def test_port_reduce_matches_reference_golden_sum():
    n = 16
    data = np.arange(n ** 3, dtype=np.float32).reshape(n, n, n)
    v = _runtime_reduce(H.oracle_port(), data, 2, 2, "Float", True)
    assert "%f\n" % float(v) == "\n".join(GOLD["system_tests"]["test_reduction-3d-sum"]["head"]) + "\n"


@needs_ref
@pytest.mark.parametrize("name,ptype,dtype", [("Float", 2, np.float32), ("Double", 3, np.float64),
                                              ("Int", 0, np.int32), ("Long", 1, np.int64)])
@pytest.mark.parametrize("op", [0, 1, 2, 3])
def test_port_reduce_equals_ref_runtime(name, ptype, dtype, op):
    rng = np.random.default_rng(op * 7 + ptype)
    if np.issubdtype(dtype, np.floating):
        data = (rng.random((4, 5, 6)) * 2 - 0.7).astype(dtype)
    else:
        data = rng.integers(-3, 4, (4, 5, 6)).astype(dtype)
        data[data == 0] = 1
    a = _runtime_reduce(H.oracle_port(), data, ptype, op, name, True)
    b = _runtime_reduce(H.oracle_ref(), data, ptype, op, name, True)
    assert a.tobytes() == b.tobytes()
    # sequential left fold in T (libphysis_rt_ref.cc:19-30)
    flat = data.ravel()
    acc = flat[0]
    np.seterr(over='ignore')
    for x in flat[1:]:
        if op == 0:
            acc = acc if acc > x else x
        elif op == 1:
            acc = acc if acc < x else x
        elif op == 2:
            acc = dtype(acc + x)
        else:
            acc = dtype(acc * x)
    assert np.asarray(acc, dtype).tobytes() == a.tobytes()


def test_diffusion_accuracy_figure_matches_reference():
    """The benchmark's own accuracy check (RMS error vs the analytic solution) on the port's result
    reproduces the figure the reference's Baseline printed (fixture)."""
    n, count = 64, 20
    p = H.diffusion_params(n, n, n)
    f0 = H.diffusion_initial(n, n, n, p)
    got = H.run_diffusion(H.oracle_port(), f0, n, n, n, count, p)
    exact = H.diffusion_initial(n, n, n, p, time=float(np.float32(p[10]) * np.float32(count)))
    # GetAccuracy: sqrt(sum((a-b)^2)/n) accumulated in double (diffusion3d.h:113-120)
    err = np.sqrt(np.sum((got.astype(np.float64) - exact.astype(np.float64)) ** 2) / got.size)
    want = float(GOLD["diffusion_baseline"][f"{n}x{count}"]["accuracy"])
    assert abs(err - want) <= 1e-4 * want  # the reference accumulates the squares in its own order


@pytest.mark.parametrize("shape,count", [((32, 17, 9), 6), ((64, 8, 5), 4), ((4, 2, 2), 6), ((20, 1, 3), 2)])
def test_numpy_restatement_equals_c_oracle(shape, count):
    """helpers.diffusion7_numpy (the checker of the fused two-sweep GPU tests, also used for
    fp64) is bit-identical to the C restatement of the REFERENCE target."""
    nx, ny, nz = shape
    rng = np.random.default_rng(11)
    f0 = rng.random(nx * ny * nz, dtype=np.float32)
    co = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44], np.float32)
    want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, co)
    got = H.diffusion7_numpy(f0, shape, co, count)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_reference_openmp_form_equals_oracle():
    """The reference's own multi-threaded CPU form of the sweep (diffusion3d_openmp.cc, built
    unmodified into oracle/_ref and timed by bench.py's reference arm) computes the same bits
    as the oracle."""
    ref = H.oracle_ref()
    if ref is None:
        pytest.skip("oracle/_ref was never built (needs /root/reference)")
    nx, ny, nz, count = 48, 20, 13, 6
    p = H.diffusion_params(nx, ny, nz)
    f0 = H.diffusion_initial(nx, ny, nz, p)
    want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, p)
    got = f0.copy()
    ref.ref_openmp_load.argtypes = [C.c_int] * 3 + [C.c_void_p]
    ref.ref_openmp_store.argtypes = [C.c_void_p]
    ref.ref_openmp_load(nx, ny, nz, got.ctypes.data)
    ref.ref_openmp_sweeps(count)
    ref.ref_openmp_store(got.ctypes.data)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
