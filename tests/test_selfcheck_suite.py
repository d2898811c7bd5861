"""The reference's system tests that ship WITHOUT an expected-output twin
(tests/system_tests/test_cases/: test_01/02/03/08/09/10, test_multi-kernels, test_param_name,
test_cplusplus, test_redblack-separated, test_7-pt.module, test_reduction-{2d,3d-int,3d-long,
3d-max,3d-min,3d-prod}, test_user-defined-type{2,-array-member-copy,-copyin-copyout,
-copyin-copyout-two-members,-kernel-copy,-transpose}) as hand-emitted translations for both
targets (examples/golden/selfcheck_suite.inc).  The reference's driver only checks that these
run; most check themselves.  Here: the program's own check must pass on every target, and the
b200 target must return the REFERENCE target's bytes (parity unpinned by a reference golden --
the REFERENCE-target emission on the real REF runtime is the oracle).
"""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import test_golden_suite as G

FAILED = C.c_size_t(-1).value

# name -> element dtype of what the program reads back
SUITE = {
    "test_01": np.float32, "test_02": np.float32, "test_03": np.float32, "test_cplusplus": np.float32,
    "test_08": np.float32, "test_09": np.float32, "test_10": np.float32,
    "test_multi-kernels": np.float32, "test_param_name": np.float32,
    "test_redblack-separated": np.float32, "test_7-pt.module": np.float32,
    "test_reduction-2d": np.float64, "test_reduction-3d-int": np.int32, "test_reduction-3d-long": np.int64,
    "test_reduction-3d-max": np.float32, "test_reduction-3d-min": np.float32,
    "test_reduction-3d-prod": np.float32,
    "test_user-defined-type2": np.float32, "test_user-defined-type-array-member-copy": np.float32,
    "test_user-defined-type-copyin-copyout": np.float32,
    "test_user-defined-type-copyin-copyout-two-members": np.float32,
    "test_user-defined-type-kernel-copy": np.float32, "test_user-defined-type-transpose": np.float32,
}
# values the reference programs print / check against, computed independently here
EXPECTED_SCALARS = {
    "test_reduction-2d": float(sum(range(16))), "test_reduction-3d-int": sum(range(512)),
    "test_reduction-3d-long": sum(range(512)), "test_reduction-3d-max": 63.0, "test_reduction-3d-min": 0.0,
}


def run(lib, name):
    fn = getattr(lib, "selfcheck_" + name.replace("-", "_").replace(".", "_"))
    fn.argtypes = [C.c_void_p]
    fn.restype = C.c_size_t
    buf = np.zeros(32 ** 3 * 6 * 4, np.uint8)
    n = fn(buf.ctypes.data)
    assert n != FAILED, f"{name}: the program's own check failed"
    return buf[:n].copy()


def test_suite_is_the_reference_tests_without_a_twin():
    # 60 reference programs = 36 with a twin (tests/test_golden_suite.py) + 24 without;
    # test_7-pt.module_base.c is the second translation unit of test_7-pt.module.c
    assert len(SUITE) == 23 and not (set(SUITE) & set(G.SUITE))


@pytest.mark.parametrize("name", sorted(SUITE))
def test_oracle_passes_the_programs_own_check(name):
    out = run(H.oracle_port(), name).view(SUITE[name])
    if name in EXPECTED_SCALARS:
        assert out[0] == EXPECTED_SCALARS[name]
    if name == "test_reduction-3d-prod":
        assert abs(float(out[0]) - 1.1 ** 64) <= 1e-5 * 1.1 ** 64
    if name == "test_redblack-separated":   # the reference diffs it against test_redblack's twin
        assert G.sha(G.stdout_of("test_redblack", out.view(np.uint8))) == G.GOLD["test_redblack"]["sha256"]
    if name == "test_7-pt.module":
        assert G.sha(G.stdout_of("test_7-pt", out.view(np.uint8))) == G.GOLD["test_7-pt"]["sha256"]


@pytest.mark.skipif(H.oracle_ref() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("name", sorted(SUITE))
def test_real_ref_runtime_agrees_with_port(name):
    assert run(H.oracle_port(), name).tobytes() == run(H.oracle_ref(), name).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SUITE))
def test_b200_matches_the_reference_target(name):
    want = run(H.oracle_port(), name)
    got = run(H.b200_programs(), name)
    if name == "test_reduction-3d-prod":
        # 64 factors of 1.1f: the GPU tree and the sequential fold differ by reassociation
        # only; the test's own tolerance is 1e-5 relative (test_reduction-3d-prod.c:40)
        w, g = float(want.view(np.float32)[0]), float(got.view(np.float32)[0])
        assert abs(w - g) <= 1e-5 * w
    else:
        assert got.tobytes() == want.tobytes()
