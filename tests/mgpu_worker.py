"""One rank of a multi-GPU parity run (launched by tests/test_multigpu.py or by hand with
torchrun-style environment).  Every rank runs the SAME Physis programs (SPMD) on global
arrays, and compares what it gets back with the CPU oracle run on the same inputs."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402


def check(name, want, got, view):
    ok = np.array_equal(np.ascontiguousarray(want).view(view), np.ascontiguousarray(got).view(view))
    if not ok:
        w, g = np.ascontiguousarray(want).ravel(), np.ascontiguousarray(got).ravel()
        bad = np.nonzero(w.view(view) != g.view(view))[0]
        raise AssertionError(f"{name}: {bad.size} of {w.size} elements differ, first at {bad[:5]}: "
                             f"want {w[bad[:3]]} got {g[bad[:3]]}")


def case_diffusion():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the last shape has one plane per rank: thinner than the default two-plane halo
    for (nx, ny, nz), count in [((128, 32, 16), 6), ((64, 48, 37), 4), ((256, 16, 9), 2), ((30, 17, 11), 4),
                                ((128, 8, max(world, 2)), 4)]:
        rng = np.random.default_rng(nx + nz)
        f0 = rng.random(nx * ny * nz, dtype=np.float32)
        co = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44], np.float32)
        want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, co)
        got = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, count, co)
        check(f"diffusion {nx}x{ny}x{nz}", want, got, np.uint32)
        got = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, count, co, entry="run_kernel_physis_generic")
        check(f"diffusion generic {nx}x{ny}x{nz}", want, got, np.uint32)


def case_pair():
    """The fused two-sweep pass on z-slabs: two halo planes per side, delivered by the kernel."""
    from physis_b200 import api
    co64 = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    iso64 = np.array([0.1234567] * 6 + [0.2592598])   # equal neighbour coefficients: shared-product form
    cases = [((128, 32, 16), 3, np.float32, (), co64), ((256, 20, 33), 5, np.float32, ("star7_pair_zc=3",), co64),
             ((512, 17, 24), 4, np.float32, (), iso64), ((64, 30, 19), 6, np.float64, ("star7_pair_zc=2",), co64),
             ((384, 9, 41), 3, np.float32, (), co64), ((128, 21, 26), 5, np.float64, (), iso64),
             ((256, 12, 35), 4, np.float32, ("star7_pair_zc=6",), iso64),
             # rows wider than one fused tile: x tiles
             ((1024, 14, 32), 3, np.float32, (), iso64), ((768, 11, 40), 4, np.float32, ("star7_pair_zc=3",), co64),
             # slabs of 4+ planes on up to 3 ranks (fused), of 3 planes on 4 ranks (sweep by sweep)
             ((128, 18, 13), 3, np.float32, (), co64)]
    for shape, iters, dtype, opts, co64 in cases:
        nx, ny, nz = shape
        api.PSInit(["t"], 3, shape)
        for kv in opts:
            api.set_option(kv)
        pt = api.PS_FLOAT if dtype == np.float32 else api.PS_DOUBLE
        a, b = api.Grid(shape, pt), api.Grid(shape, pt)
        rng = np.random.default_rng(nx + 3 * nz)
        f0 = rng.random(nx * ny * nz).astype(dtype)
        a.copyin(f0)
        b.copyin(rng.random(nx * ny * nz).astype(dtype))
        dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
        co = [float(dtype(c)) for c in co64]
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co, elm_type=pt)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co, elm_type=pt)
        api.rt().__PSB200ResetStats()
        for rep in range(2):   # the second run starts from halos left by single sweeps
            api.stencil_run(iters, [d0, d1])
        pairs = int(api.stats().fused_pairs)
        envopt = os.environ.get("PHYSIS_B200_OPTIONS", "")
        in_kernel_exchange = "halo_push=0" not in envopt and "sync_mode=0" not in envopt and "sync_mode=1" not in envopt
        if world == 1 or (nz // world >= 4 and in_kernel_exchange):
            assert pairs == 2 * ((iters - 1) & ~1), (shape, pairs)
        else:
            assert pairs == 0, (shape, pairs)
        fa, fb = a.copyout(), b.copyout()
        view = np.uint32 if dtype == np.float32 else np.uint64
        check(f"pair A {shape}", H.diffusion7_numpy(f0, shape, co64.astype(dtype), 4 * iters), fa, view)
        check(f"pair B {shape}", H.diffusion7_numpy(f0, shape, co64.astype(dtype), 4 * iters - 1), fb, view)
        a.free()
        b.free()
        api.PSFinalize()


def case_pair_tail():
    """Early signal with a one-plane tail chunk (nz_loc % zc == 1) and more work items than CTA
    slots: the boundary items must include the second-last chunk (star7_pair.cu, sweep_common.cuh
    SlabSyncSetBoundary).  Repeated, since a race shows only sometimes."""
    from physis_b200 import api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    co64 = np.array([0.1234567] * 6 + [0.2592598])
    for shape, iters, opts, reps in [((512, 512, 21 * world), 4, ("star7_pair_zc=5",), 12),
                                     ((256, 64, 9 * world), 5, ("star7_pair_zc=1",), 6),
                                     ((128, 40, 13 * world), 5, ("star7_pair_zc=4",), 6),
                                     # the boundary-first schedule (short end chunks in the first wave,
                                     # unequal interior chunks): 37 y tiles x 4 groups fill the 148 SMs
                                     ((128, 512, 33 * world), 4, ("star7_pair_zbl=2",), 3),
                                     ((64, 512, 61 * world), 5, (), 3),
                                     ((64, 512, 57 * world + 1), 4, ("star7_pair_zbl=3",), 2)]:
        nx, ny, nz = shape
        api.PSInit(["t"], 3, shape)
        for kv in opts:
            api.set_option(kv)
        a, b = api.Grid(shape, api.PS_FLOAT), api.Grid(shape, api.PS_FLOAT)
        rng = np.random.default_rng(nx + nz)
        f0 = rng.random(nx * ny * nz, dtype=np.float32)
        g0 = rng.random(nx * ny * nz, dtype=np.float32)
        dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
        co = [float(np.float32(c)) for c in co64]
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
        want_a = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), 2 * iters)
        want_b = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), 2 * iters - 1)
        for rep in range(reps):
            a.copyin(f0)
            b.copyin(g0)
            api.rt().__PSB200ResetStats()
            api.stencil_run(iters, [d0, d1])
            assert int(api.stats().fused_pairs) == ((iters - 1) & ~1), shape
            check(f"tail A {shape} rep {rep}", want_a, a.copyout(), np.uint32)
            check(f"tail B {shape} rep {rep}", want_b, b.copyout(), np.uint32)
        a.free()
        b.free()
        api.PSFinalize()


def case_himeno():
    for dims, nn in [((64, 32, 32), 4), ((128, 20, 13), 2)]:
        a = H.run_himeno(H.oracle_port(), dims, nn, gosa=True, seed=5)
        b = H.run_himeno(H.b200_programs(), dims, nn, gosa=True, seed=5)
        check(f"himeno p0 {dims}", a[0], b[0], np.uint32)
        check(f"himeno p1 {dims}", a[1], b[1], np.uint32)
        check(f"himeno gosa grid {dims}", a[3], b[3], np.uint32)
        exact = float(np.sum(a[3].astype(np.float64)))
        assert abs(b[2] - exact) <= 8e-6 * abs(exact), (b[2], exact)


def case_himeno_pair():
    """Fused two-sweep Himeno passes on z-slabs (two halo planes per side, forwarded by the
    kernel): both p grids and the residual grid bit-identical to the oracle's sweep-by-sweep run."""
    from physis_b200 import api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    envopt = os.environ.get("PHYSIS_B200_OPTIONS", "")
    in_kernel_exchange = "halo_push=0" not in envopt and "sync_mode=0" not in envopt and "sync_mode=1" not in envopt
    for dims, nn in [((64, 32, 32), 8), ((128, 20, 13), 6), ((256, 30, 40), 10), ((136, 17, 33), 6)]:
        seen = {}
        for gosa in (False, True):
            a = H.run_himeno(H.oracle_port(), dims, nn, gosa=gosa, seed=11)
            b = H.run_himeno(H.b200_programs(), dims, nn, gosa=gosa, seed=11,
                             before_finalize=lambda: seen.update(n=int(api.stats().fused_pairs)))
            want = ((nn // 2 - 1) & ~1) if (world == 1 or (dims[2] // world >= 4 and in_kernel_exchange)) else 0
            assert seen["n"] == want, (dims, seen, want)
            check(f"himeno pair p0 {dims}", a[0], b[0], np.uint32)
            check(f"himeno pair p1 {dims}", a[1], b[1], np.uint32)
            if gosa:
                check(f"himeno pair ss^2 {dims}", a[3], b[3], np.uint32)
                exact = float(np.sum(a[3].astype(np.float64)))
                assert abs(b[2] - exact) <= 8e-6 * abs(exact), (b[2], exact)


def case_pstag():
    for (nx, ny, nz), count in [((64, 16, 8), 4), ((128, 32, 11), 3)]:
        u, kap = H.pstag_inputs(nx, ny, nz)
        want = H.run_pstag(H.oracle_port(), u, kap, nx, ny, nz, count)
        got = H.run_pstag(H.b200_programs(), u, kap, nx, ny, nz, count)
        check(f"pstag {nx}x{ny}x{nz}", want, got, np.uint64)


def case_api():
    """Runtime calls through the ctypes mirror: slab-local copies, PSReduce, PSGridSet."""
    from physis_b200 import api
    rank = int(os.environ.get("RANK", "0"))
    api.PSInit(["t"], 3, (32, 8, 21))
    r = api.rt()
    world = r.__PSB200WorldSize()
    assert r.__PSB200Rank() == rank
    g = api.Grid((32, 8, 21), api.PS_INT)
    full = (np.arange(32 * 8 * 21, dtype=np.int64) % 1000 - 300).astype(np.int32)
    g.copyin(full)
    assert int(g.reduce(api.PS_SUM)) == int(full.sum())
    assert int(g.reduce(api.PS_MAX)) == int(full.max())
    assert int(g.reduce(api.PS_MIN)) == int(full.min())
    back = g.copyout()
    check("copyout gather", full, back, np.uint32)
    off, ln = C.c_int(), C.c_int()
    r.__PSB200GridLocalSize(g.ptr, C.byref(off), C.byref(ln))
    plane = 32 * 8
    mine = (full[off.value * plane:(off.value + ln.value) * plane] * 2).astype(np.int32)
    r.__PSB200GridCopyinLocal(g.ptr, C.c_void_p(mine.ctypes.data))
    assert int(g.reduce(api.PS_SUM)) == 2 * int(full.sum())
    out = np.zeros_like(mine)
    r.__PSB200GridCopyoutLocal(g.ptr, C.c_void_p(out.ctypes.data))
    check("local round trip", mine, out, np.uint32)
    g.set((3, 2, 20), np.int32(77777).tobytes())
    g.set((1, 1, 0), np.int32(-5).tobytes())
    back = g.copyout()
    want = (full * 2).astype(np.int32)
    want[3 + 2 * 32 + 20 * plane] = 77777
    want[1 + 1 * 32] = -5
    check("set", want, back, np.uint32)
    g.free()
    api.PSFinalize()
    assert world == int(os.environ.get("WORLD_SIZE", "1"))


def case_golden():
    """A subset of the reference's system tests through the generic path on z-slabs (halo=2 for
    the asymmetric stencil); case_golden_all runs every one."""
    import test_golden_suite as G
    names = ["test_7-pt", "test_7-pt-multi-iterations", "test_7-pt-double-type", "test_7-pt-int-type",
             "test_16", "test_15", "test_27-pt", "test_asymmetric", "test_stencil-hole",
             "test_7-pt-neumann-cond", "test_7-pt-type-mix", "test_mixed-dim", "test_mixed-dim2",
             "test_mixed-dim3", "test_27-pt-reduction", "test_reduction-3d-sum",
             "test_user-defined-type-7-pt", "test_user-defined-type1", "test_user-defined-type3",
             "test_redblack", "test_3-pt-1d", "test_5-pt-2d", "test_5-pt-periodic", "test_9-pt-2d",
             "test_9-pt-reduction", "test_9-pt-periodic-reduction"]
    for n in names:
        want = G.run_golden(H.oracle_port(), n)
        got = G.run_golden(H.b200_programs(), n)
        assert got.tobytes() == want.tobytes(), n
        assert G.sha(G.stdout_of(n, got)) == G.GOLD[n]["sha256"], n


def case_golden_all():
    """All 36 reference system tests with a twin on z-slabs: also the kernels that wrap
    periodically in z (the wrap is the ring exchange: rank 0's lower halo holds the last plane)
    and user types with array members (component stride = the rank's allocation)."""
    import test_golden_suite as G
    for n in sorted(G.SUITE):
        want = G.run_golden(H.oracle_port(), n)
        got = G.run_golden(H.b200_programs(), n)
        assert got.tobytes() == want.tobytes(), n
        assert G.sha(G.stdout_of(n, got)) == G.GOLD[n]["sha256"], n


def case_selfcheck():
    """The reference system tests without a twin on z-slabs: each program's own check must pass and
    the bytes must equal the REFERENCE target's (test_reduction-3d-prod within its own tolerance)."""
    import test_selfcheck_suite as S
    for n in sorted(S.SUITE):
        want = S.run(H.oracle_port(), n)
        got = S.run(H.b200_programs(), n)
        if n == "test_reduction-3d-prod":
            w, g = float(want.view(np.float32)[0]), float(got.view(np.float32)[0])
            assert abs(w - g) <= 1e-5 * w, n
        else:
            assert got.tobytes() == want.tobytes(), n


def case_autotune():
    """Option autotune=1 on z-slabs: the ranks try the same forms on the run's own first iterations
    and agree on the slowest rank's times; the result stays the oracle's bits."""
    from physis_b200 import api
    co64 = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44])
    co = [float(np.float32(c)) for c in co64]
    # (the last shape: uneven slabs, 16 and 17 planes on two ranks -- the list of forms, which includes z chunks
    # derived from the slab thickness, must still be the same on every rank)
    for shape in [(128, 32, 16), (512, 12, 24), (128, 24, 33)]:
        nx, ny, nz = shape
        api.PSInit(["t"], 3, shape)
        api.set_option("autotune=1")
        a, b = api.Grid(shape, api.PS_FLOAT), api.Grid(shape, api.PS_FLOAT)
        f0 = np.random.default_rng(nx).random(nx * ny * nz, dtype=np.float32)
        a.copyin(f0)
        dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
        api.rt().__PSB200ResetStats()
        api.stencil_run(60, [d0, d1])
        assert int(api.stats().autotune_trials) >= 3
        api.stencil_run(7, [d0, d1])
        want = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), 2 * 67)
        check(f"autotune {shape}", want, a.copyout(), np.uint32)
        api.PSFinalize()
    dims, nn = (64, 20, 24), 80
    os.environ["PHYSIS_B200_OPTIONS"] = "autotune=0"
    ref = H.run_himeno(H.oracle_port(), dims, nn, seed=4, omega=0.1)
    os.environ["PHYSIS_B200_OPTIONS"] = "autotune=1"
    seen = {}
    got = H.run_himeno(H.b200_programs(), dims, nn, seed=4, omega=0.1,
                       before_finalize=lambda: seen.update(n=int(api.stats().autotune_trials)))
    assert seen["n"] == 6, seen
    for i in (0, 1):
        check(f"autotune himeno p{i}", ref[i], got[i], np.uint32)


CASES = {"autotune": case_autotune, "selfcheck": case_selfcheck, "himeno_pair": case_himeno_pair, "golden": case_golden, "golden_all": case_golden_all, "pair_tail": case_pair_tail, "diffusion": case_diffusion, "pair": case_pair, "himeno": case_himeno, "pstag": case_pstag, "api": case_api}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
    print("ok rank", os.environ.get("RANK", "0"), names, flush=True)
