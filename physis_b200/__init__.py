"""physis_b200 — B200-native runtime + sm_100a sweep kernels for the Physis stencil DSL.

The product is the C-ABI library ``physis_b200/lib/libphysis_rt_b200.so``
(sources in ``physis_b200/csrc``, interface in ``include/physis/physis_b200.h``).
This Python package is only the thin ctypes host-side mirror used by the
tests and ``bench.py``; it never computes anything itself and raises if the
CUDA library is missing (there is no CPU fallback).
"""
from ._lib import load_runtime, load_programs, lib_dir, build_native  # noqa: F401
from . import api  # noqa: F401

__all__ = ["load_runtime", "load_programs", "lib_dir", "build_native", "api"]
