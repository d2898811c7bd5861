#!/usr/bin/env python
"""Generates tests/golden/reference_golden.json from the REFERENCE ITSELF.

Runs, in the build container (where /root/reference exists):
  * the reference's 36 hand-written expected-output programs
    tests/system_tests/test_cases/*.manual.ref.c   (compiled by `make -C oracle ref`
    from the sources where they lie into oracle/_ref/golden/)
  * the reference's original Himeno benchmark examples/himeno/himenobmtxpa_original.c
    (size XS; it dumps p after its 4-sweep rehearsal to himeno.original.dat)
  * the reference's own diffusion `Baseline` (examples/diffusion-benchmark/baseline.cc)
    through oracle/_ref/libphysis_ref.so
and records sha256 digests (+ a few head/tail lines for humans).  Only digests and a
handful of output lines are committed — no reference source.

    python tests/golden/make_golden.py
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def digest_text(text):
    lines = text.splitlines()
    return {"sha256": hashlib.sha256(text.encode()).hexdigest(), "lines": len(lines),
            "head": lines[:4], "tail": lines[-2:]}


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    out = {"_generated_by": "tests/golden/make_golden.py", "system_tests": {}}
    gdir = os.path.join(REFDIR, "golden")
    for name in sorted(os.listdir(gdir)):
        if not name.startswith("test_"):
            continue
        text = subprocess.run([os.path.join(gdir, name)], capture_output=True, text=True, check=True).stdout
        out["system_tests"][name] = digest_text(text)
    # original Himeno, XS (32x32x64, k fastest)
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([os.path.join(gdir, "himeno_original"), "XS"], cwd=td, capture_output=True, text=True)
        with open(os.path.join(td, "himeno.original.dat")) as f:
            dat = f.read()
        gosa_line = [l for l in r.stdout.splitlines() if "MFLOPS" in l][0]
        out["himeno_original_XS"] = dict(digest_text(dat), sweeps=4,
                                         rehearsal_gosa=gosa_line.split()[-1])
    # the reference's own Baseline diffusion: 64^3 x 20 steps and 32^3 x 10
    lib = C.CDLL(os.path.join(REFDIR, "libphysis_ref.so"))
    lib.ref_baseline_run.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    out["diffusion_baseline"] = {}
    for n, count in ((64, 20), (32, 10)):
        buf = np.zeros(n ** 3, np.float32)
        acc = C.c_float(0)
        lib.ref_baseline_run(n, n, n, count, buf.ctypes.data, C.byref(acc))
        out["diffusion_baseline"][f"{n}x{count}"] = {
            "sha256": hashlib.sha256(buf.tobytes()).hexdigest(), "accuracy": "%.6e" % acc.value,
            "sum": "%.9e" % float(np.sum(buf, dtype=np.float64))}
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out["system_tests"]), "system-test digests")


if __name__ == "__main__":
    sys.exit(main())
