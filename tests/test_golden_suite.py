"""The reference's own system tests (tests/system_tests/test_cases/test_*.c) as known-answer
tests.  The programs are hand-emitted translations shared by both targets
(examples/golden/golden_suite.inc); their expected stdout comes from the reference's
hand-written twins (*.manual.ref.c -> tests/golden/reference_golden.json).

  CPU : the oracle (REF-target shape on the plain-C port of the REF runtime, and on the
        reference's real runtime where built) prints exactly what the reference expects
  GPU : the b200 target (generic per-point kernels through __PSB200StencilRun, device SoA for
        user types, PSReduce) returns the oracle's bytes and the reference's stdout
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import helpers as H

with open(os.path.join(H.GOLDEN_DIR, "reference_golden.json")) as f:
    GOLD = json.load(f)["system_tests"]

F, D, I = np.float32, np.float64, np.int32


def _fmt_cols(a, cols, fmt):
    a = a.reshape(-1, cols)
    line = " ".join([fmt] * cols) + "\n"
    return "".join(line % tuple(r) for r in a.tolist())


# name -> (dtype, values per element in the copied-out struct, columns printed, printf format)
SUITE = {
    "test_7-pt": (F, 1, [0], "%f"),
    "test_7-pt-multi-iterations": (F, 1, [0], "%f"),
    "test_7-pt-double-type": (D, 1, [0], "%f"),
    "test_7-pt-int-type": (I, 1, [0], "%d"),
    "test_7-pt-periodic": (F, 1, [0], "%f"),
    "test_3-pt-periodic": (F, 1, [0], "%f"),
    "test_16": (F, 1, [0], "%f"),
    "test_15": (F, 1, [0], "%f"),
    "test_27-pt": (F, 1, [0], "%f"),
    "test_27-pt-periodic": (F, 1, [0], "%f"),
    "test_asymmetric": (F, 1, [0], "%f"),
    "test_asymmetric-periodic": (F, 1, [0], "%f"),
    "test_stencil-hole": (F, 1, [0], "%f"),
    "test_7-pt-neumann-cond": (F, 1, [0], "%f"),
    "test_7-pt-type-mix": (D, 1, [0], "%f"),
    "test_mixed-dim": (F, 1, [0], "%f"),
    "test_27-pt-reduction": (I, 1, [0], "%d"),
    "test_reduction-3d-sum": (F, 1, [0], "%f"),
    "test_user-defined-type-7-pt": (F, 2, [0, 1], "%f"),
    "test_user-defined-type-7-pt-periodic": (F, 2, [0, 1], "%f"),
    "test_user-defined-type-7-pt-periodic-complex": (F, 2, [0, 1], "%f"),
    "test_user-defined-type1": (F, 3, [0, 1, 2], "%f"),
    "test_user-defined-type3": (F, 3, [0, 1, 2], "%f"),
    "test_user-defined-type5": (F, 2, [1], "%f"),
    "test_user-defined-type-multi-members": (F, 4, [3], "%f"),
    "test_user-defined-type-multi-dim-member": (F, 6, [2], "%f"),
    "test_3-pt-1d": (F, 1, [0], "%f"),
    "test_5-pt-2d": (F, 1, [0], "%f"),
    "test_5-pt-periodic": (F, 1, [0], "%f"),
    "test_9-pt-2d": (F, 1, [0], "%f"),
    "test_9-pt-reduction": (I, 1, [0], "%d"),
    "test_9-pt-periodic-reduction": (I, 1, [0], "%d"),
    "test_redblack": (F, 1, [0], "%f"),
    "test_redblack-periodic": (F, 1, [0], "%f"),
    "test_mixed-dim2": (F, 1, [0], "%f"),
    "test_mixed-dim3": (F, 1, [0], "%f"),
}


def run_golden(lib, name):
    fn = getattr(lib, "golden_" + name.replace("-", "_"))
    fn.argtypes = [C.c_void_p]
    fn.restype = C.c_size_t
    buf = np.zeros(32 ** 3 * 6 * 4, np.uint8)   # the largest dump: 6 floats per point
    n = fn(buf.ctypes.data)
    return buf[:n].copy()


def stdout_of(name, raw):
    dtype, width, cols, fmt = SUITE[name]
    a = raw.view(dtype).reshape(-1, width)[:, cols]
    return _fmt_cols(np.ascontiguousarray(a), len(cols), fmt)


def sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


@pytest.mark.parametrize("name", sorted(SUITE))
def test_oracle_prints_what_the_reference_expects(name):
    out = stdout_of(name, run_golden(H.oracle_port(), name))
    assert out.count("\n") == GOLD[name]["lines"]
    assert out.splitlines()[:4] == GOLD[name]["head"]
    assert sha(out) == GOLD[name]["sha256"]


@pytest.mark.skipif(H.oracle_ref() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("name", sorted(SUITE))
def test_real_ref_runtime_agrees_with_port(name):
    a = run_golden(H.oracle_port(), name)
    b = run_golden(H.oracle_ref(), name)
    assert a.tobytes() == b.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SUITE))
def test_b200_matches_oracle_and_reference_stdout(name):
    want = run_golden(H.oracle_port(), name)
    got = run_golden(H.b200_programs(), name)
    assert got.tobytes() == want.tobytes()
    assert sha(stdout_of(name, got)) == GOLD[name]["sha256"]


def test_suite_covers_every_reference_golden():
    # all 36 reference system tests that ship an expected-output twin (*.manual.ref.c)
    assert set(GOLD) == set(SUITE) and len(SUITE) == 36
