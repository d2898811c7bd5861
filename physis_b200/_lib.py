"""Loading (and, on request, building) the native libraries."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_RT = None
_PROGRAMS = None


def lib_dir():
    return os.path.join(_HERE, "lib")


def build_native(verbose=False):
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=out)
    subprocess.check_call(["make", "-C", os.path.join(_ROOT, "examples", "b200")], stdout=out)


class NativeLibraryMissing(RuntimeError):
    pass


def _load(name):
    path = os.path.join(lib_dir(), name)
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(physis_b200 has no CPU fallback; the CUDA library is the product)")
    return ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)


def load_runtime():
    """libphysis_rt_b200.so — the runtime + kernels.  Loading needs no GPU; calling PSInit does."""
    global _RT
    if _RT is None:
        _RT = _load("libphysis_rt_b200.so")
    return _RT


def load_programs():
    """libphysis_b200_programs.so — hand-emitted b200 translations of the benchmark programs."""
    global _PROGRAMS
    if _PROGRAMS is None:
        load_runtime()
        _PROGRAMS = _load("libphysis_b200_programs.so")
    return _PROGRAMS
