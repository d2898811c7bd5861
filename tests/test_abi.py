"""CPU: the drop-in boundary.  The C-ABI library loads without a GPU, exports every symbol
include/physis/physis_b200.h declares, its structs have the layout the ctypes mirror (and
therefore generated code) assumes, and the product path fails loudly without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

import helpers as H

HEADER = os.path.join(H.ROOT, "include", "physis", "physis_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    for m in re.finditer(r"^[A-Za-z_][\w \*]*?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M):
        line = m.group(0)
        if "typedef" in line or "static" in line:
            continue
        names.add(m.group(1))
    return names


def test_library_loads_and_exports_every_declared_symbol():
    import physis_b200
    from physis_b200 import api
    lib = physis_b200.load_runtime()
    declared = _declared_functions()
    assert {"PSInit", "PSFinalize", "__PSGridNew", "__PSGridCopyin", "__PSGridCopyout",
            "__PSReduceGridFloat", "__PSB200StencilRun", "PSDomain3DNew"} <= declared
    for name in sorted(declared | set(api.EXPORTED_SYMBOLS)):
        assert hasattr(lib, name), f"{name} declared in physis_b200.h but not exported"
    assert set(api.EXPORTED_SYMBOLS) - {"__ps_trace"} <= declared


def test_programs_library_exports_reference_benchmark_entry_points():
    import physis_b200
    lib = physis_b200.load_programs()
    for name in ["initialize_physis", "initialize_benchmark_physis", "run_kernel_physis",
                 "finalize_benchmark_physis", "himeno_init", "himeno_jacobi", "himeno_jacobi_gosa",
                 "pstag_init", "pstag_run"]:
        assert hasattr(lib, name)


def test_struct_layouts_match_ctypes_mirror():
    from physis_b200 import api
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "physis/physis_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(__PSDomain), sizeof(__PSGridTypeMemberInfo),
         sizeof(__PSGridTypeInfo), sizeof(__PSGrid), sizeof(__PSB200StencilDesc), sizeof(__PSB200Stats));
  printf("%zu %zu %zu %zu\n", offsetof(__PSGrid, dim), offsetof(__PSGrid, num_elms),
         offsetof(__PSGrid, dev), offsetof(__PSB200StencilDesc, scalars));
  printf("%zu %zu %zu\n", offsetof(__PSB200StencilDesc, grids), offsetof(__PSB200StencilDesc, launch),
         offsetof(__PSB200StencilDesc, name));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "layout.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "layout")
        subprocess.check_call(["gcc", "-std=gnu11", "-I", os.path.join(H.ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    got = [int(x) for x in out]
    want = [C.sizeof(api.PSDomain), C.sizeof(api.MemberInfo), C.sizeof(api.TypeInfo),
            C.sizeof(api.PSGridStruct), C.sizeof(api.StencilDesc), C.sizeof(api.Stats),
            api.PSGridStruct.dim.offset, api.PSGridStruct.num_elms.offset, api.PSGridStruct.dev.offset,
            api.StencilDesc.scalars.offset, api.StencilDesc.grids.offset, api.StencilDesc.launch.offset,
            api.StencilDesc.name.offset]
    assert got == want


def test_header_is_plain_c_and_cxx():
    with tempfile.TemporaryDirectory() as td:
        for ext, cc in (("c", "gcc"), ("cc", "g++")):
            src = os.path.join(td, "inc." + ext)
            open(src, "w").write('#define PHYSIS_B200\n#include "physis/physis.h"\nint main(void){return 0;}\n')
            subprocess.check_call([cc, "-fsyntax-only", "-Wall", "-I", os.path.join(H.ROOT, "include"), src])


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_product_path_fails_loudly_without_gpu():
    code = ("import sys; sys.path.insert(0, %r); from physis_b200 import api; api.PSInit()" % H.ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_missing_library_raises(monkeypatch, tmp_path):
    import physis_b200._lib as L
    monkeypatch.setattr(L, "_RT", None)
    monkeypatch.setattr(L, "lib_dir", lambda: str(tmp_path))
    with pytest.raises(L.NativeLibraryMissing):
        L.load_runtime()


def test_product_sources_do_not_touch_the_oracle():
    bad = []
    for base in ("physis_b200", "include", os.path.join("examples", "b200")):
        for dp, _, files in os.walk(os.path.join(H.ROOT, base)):
            for fn in files:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"oracle[/_.]|liboracle|/root/reference/.*\.(so|a)\b", text) and "ORACLE" not in text:
                    if re.search(r"(import|include|CDLL|dlopen|open)\W.*oracle", text):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_fused_schedule_leaves_grids_as_the_reference_schedule_does():
    """Host logic of the fused two-sweep schedule (no GPU): simulate which time level each grid
    holds.  A fused pass reads one grid and writes level+2 into the other; single sweeps
    ping-pong.  After `iter` iterations A must hold level 2*iter and B level 2*iter-1."""
    import ctypes as C
    import physis_b200
    lib = physis_b200.load_runtime()
    lib.__PSB200FusedPassCount.argtypes = [C.c_int]
    lib.__PSB200FusedPassCount.restype = C.c_int
    for it in range(0, 200):
        passes = lib.__PSB200FusedPassCount(it)
        assert passes % 2 == 0 and 0 <= passes <= max(it - 1, 0)
        level = {"A": 0, "B": None}
        src, dst = "A", "B"
        for _ in range(passes):
            level[dst] = level[src] + 2
            src, dst = dst, src
        assert src == "A"
        for _ in range(it - passes):
            level["B"] = level["A"] + 1
            level["A"] = level["B"] + 1
        if it > 0:
            assert level == {"A": 2 * it, "B": 2 * it - 1}, (it, passes, level)
    assert lib.__PSB200FusedPassCount(500) == 498


def test_every_runtime_option_is_documented():
    """INTEGRATION.md §6 lists every key `PHYSIS_B200_OPTIONS` / `__PSB200SetOption` understands."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    keys = re.findall(r'k == "([a-z0-9_]+)"', open(os.path.join(root, "physis_b200", "csrc", "runtime.cu")).read())
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert len(keys) > 30
    missing = [k for k in keys if f"`{k}`" not in doc]
    assert not missing, missing


def test_kernels_carry_no_fused_multiply_add_and_are_sm100a_only():
    """The bit-exactness argument, checked on the built library: every multiply and add of the
    sweep kernels rounds separately (no FFMA / DFMA in any kernel's SASS), the sweep kernels move
    their tiles with TMA (UTMALDG), and the only architecture in the library is sm_100a."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "physis_b200", "lib", "libphysis_rt_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    assert set(re.findall(r"arch = (sm_\w+)", sass)) == {"sm_100a"}
    ops = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            ops[cur] = set()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            ops[cur].add(m.group(1))
    assert len(ops) > 50
    # (HFMA2 does appear: the compiler's idiom for materialising a constant, not arithmetic on data)
    fused = sorted(k for k, v in ops.items() if v & {"FFMA", "DFMA", "FFMA2"})
    assert not fused, fused[:3]
    for family in ("Star7KernelV2", "Star7PairKernel", "HimenoKernel", "HimenoPairKernel", "PstagKernel"):
        ks = [k for k in ops if family in k]
        assert ks and all("UTMALDG" in ops[k] for k in ks), family
