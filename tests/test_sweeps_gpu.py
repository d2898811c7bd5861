"""GPU parity: the hand-written sm_100a sweeps, called through the C ABI by the
hand-emitted b200 programs, against the CPU oracle on identical inputs.
Bit-exact (fp32 and fp64): the kernels round every multiply/add separately in
the reference's evaluation order."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,count", [
    ((32, 32, 32), 4), ((64, 64, 64), 20), ((128, 32, 16), 6), ((256, 48, 9), 4),
    ((132, 20, 7), 2), ((4, 4, 4), 2), ((8, 3, 1), 2), ((512, 64, 40), 2),
])
def test_diffusion7_matches_oracle(shape, count):
    nx, ny, nz = shape
    p = H.diffusion_params(nx, ny, nz)
    rng = np.random.default_rng(nx * 7 + ny)
    f0 = (H.diffusion_initial(nx, ny, nz, p) + rng.random(nx * ny * nz, dtype=np.float32)).astype(np.float32)
    # asymmetric coefficients so a swapped neighbour would show
    co = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44], np.float32)
    want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, co)
    got = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, count, co)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("variant", range(6))
def test_diffusion7_all_tile_variants(variant):
    from physis_b200 import api
    nx, ny, nz = 256, 72, 21
    p = H.diffusion_params(nx, ny, nz)
    f0 = H.diffusion_initial(nx, ny, nz, p)
    want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, 4, p)
    lib = H.b200_programs()
    f = f0.copy()
    lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.initialize_physis(0, None, nx, ny, nz)
    api.set_option(f"star7_variant={variant}")
    api.set_option("star7_zc=5")
    lib.initialize_benchmark_physis(nx, ny, nz)
    lib.run_kernel_physis.argtypes = [C.c_int, C.c_void_p] + [C.c_int] * 3 + [C.c_float] * 7
    lib.run_kernel_physis(4, f.ctypes.data, nx, ny, nz, *[float(c) for c in p[:7]])
    lib.finalize_benchmark_physis()
    assert np.array_equal(want.view(np.uint32), f.view(np.uint32))


def test_diffusion7_generic_path_and_fallback():
    # nx not a multiple of 4: the specialised kernel declines, the program's generic
    # per-point kernel runs instead (still on the GPU)
    for shape in [(30, 17, 5), (64, 64, 8)]:
        nx, ny, nz = shape
        p = H.diffusion_params(nx, ny, nz)
        f0 = H.diffusion_initial(nx, ny, nz, p)
        want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, 4, p)
        got = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, 4, p)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
        got2 = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, 4, p, entry="run_kernel_physis_generic")
        assert np.array_equal(want.view(np.uint32), got2.view(np.uint32))


def test_diffusion7_config1_accuracy_and_parity():
    # BASELINE config 1 shape at reduced step count: 256^3 would take the oracle ~12 s per
    # 100 steps; 128^3 x 20 keeps the CPU side in seconds.  Also checks the benchmark's own
    # accuracy figure (RMS vs analytic solution) agrees.
    n, count = 128, 20
    p = H.diffusion_params(n, n, n)
    f0 = H.diffusion_initial(n, n, n, p)
    want = H.run_diffusion(H.oracle_port(), f0, n, n, n, count, p)
    got = H.run_diffusion(H.b200_programs(), f0, n, n, n, count, p)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("dims,nn,gosa", [
    ((64, 32, 32), 4, False), ((64, 32, 32), 4, True), ((128, 20, 11), 2, True),
    ((136, 9, 5), 2, False), ((8, 4, 3), 2, True),
])
def test_himeno_matches_oracle(dims, nn, gosa):
    a = H.run_himeno(H.oracle_port(), dims, nn, gosa=gosa, seed=5)
    b = H.run_himeno(H.b200_programs(), dims, nn, gosa=gosa, seed=5)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    if gosa:
        assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))  # emitted ss*ss grid
        # PSReduce: REF folds sequentially in fp32, the GPU folds as a tree; both are within
        # fp32 summation error of the fp64 sum of the same (bit-identical) addends
        exact = float(np.sum(a[3].astype(np.float64)))
        assert abs(b[2] - exact) <= 1e-6 * abs(exact) * 8
        assert abs(a[2] - exact) <= 1e-4 * abs(exact)


@pytest.mark.parametrize("dims,nn", [((64, 32, 32), 6), ((136, 21, 9), 4), ((1024, 20, 11), 2)])
def test_himeno_residual_reduced_from_the_sweeps_partials(dims, nn):
    """PSReduce(PS_SUM) right after the residual form of the sweep folds the per-CTA partial sums
    the sweep left (no pass over the grid): same value as the full reduction to fp32 accuracy,
    grids unchanged, and the statistics say which path answered."""
    from physis_b200 import api
    seen = {}
    a = H.run_himeno(H.oracle_port(), dims, nn, gosa=True, seed=5, each=True)
    b = H.run_himeno(H.b200_programs(), dims, nn, gosa=True, seed=5, each=True,
                     before_finalize=lambda: seen.update(n=int(api.stats().reduces_from_partials)))
    assert seen["n"] == nn // 2
    for i in (0, 1, 3):
        assert np.array_equal(a[i].view(np.uint32), b[i].view(np.uint32))
    exact = float(np.sum(a[3].astype(np.float64)))
    assert abs(b[2] - exact) <= 2e-7 * abs(exact), (b[2], exact)   # fp64 partials: one rounding to fp32
    assert abs(a[2] - exact) <= 1e-3 * abs(exact)                  # REF's own sequential fp32 fold


def test_reduce_partials_are_dropped_when_the_grid_changes():
    """The partial sums stand for the grid only while nothing else wrote it: a copyin, a
    PSGridSet, a sweep over a smaller domain (stale values outside it) or non-zero contents
    outside the domain all send PSReduce back to the pass over the grid."""
    from physis_b200 import api
    dims = (64, 24, 16)
    ne = int(np.prod(dims))
    api.PSInit(["t"], 3, dims)
    rng = np.random.default_rng(2)
    names = ["p0", "p1", "a0", "a1", "a2", "a3", "b0", "b1", "b2", "c0", "c1", "c2", "bnd", "wrk1", "gosa"]
    g = {n: api.Grid(dims, api.PS_FLOAT) for n in names}
    for n in names[:-1]:
        g[n].copyin(rng.random(ne, dtype=np.float32))
    order = [g[n] for n in names]

    def sweep(lo, hi):
        dom = api.PSDomain3DNew(lo, dims[0] - lo, lo, dims[1] - lo, lo, hi)
        api.stencil_run(1, [api.stencil_desc(api.KIND_HIMENO19_GOSA, dom, order, [0.8])])

    def check(expect_fused):
        api.rt().__PSB200ResetStats()
        v = float(g["gosa"].reduce(api.PS_SUM))
        assert int(api.stats().reduces_from_partials) == (1 if expect_fused else 0)
        exact = float(np.sum(g["gosa"].copyout().astype(np.float64)))
        assert abs(v - exact) <= 8e-6 * abs(exact), (v, exact)

    sweep(1, dims[2] - 1)
    check(True)
    check(True)                       # still valid: nothing wrote the grid
    g["gosa"].set((3, 3, 3), np.float32(5.0).tobytes())
    check(False)                      # PSGridSet
    sweep(1, dims[2] - 1)
    check(False)                      # contents outside the domain are no longer known to be zero
    g["gosa"].free()
    g["gosa"] = api.Grid(dims, api.PS_FLOAT)
    order[-1] = g["gosa"]
    sweep(1, dims[2] - 1)
    check(True)
    sweep(2, dims[2] - 2)             # smaller domain: the shell written before is stale data
    check(False)
    sweep(1, dims[2] - 1)
    check(True)                       # covers everything ever emitted again
    for n in names:
        g[n].free()
    api.PSFinalize()


@pytest.mark.parametrize("dims,nn,opts", [
    ((64, 32, 32), 6, ()), ((128, 40, 21), 8, ()), ((136, 17, 9), 6, ()), ((256, 33, 19), 10, ("himeno_pair_zc=5",)),
    ((1024, 20, 13), 6, ()), ((520, 45, 12), 8, ("himeno_pair_zc=3",)), ((8, 5, 4), 6, ()), ((64, 16, 3), 6, ()),
    # both ways of pulling the coefficient rows towards L2: bulk prefetches per row, tensor-map boxes
    ((1024, 20, 13), 6, ("himeno_pair_pfmode=1",)), ((1024, 20, 13), 6, ("himeno_pair_pfmode=2", "himeno_pair_pf=2")),
    ((136, 29, 11), 8, ("himeno_pair_pfmode=1", "himeno_pair_pf=3")), ((136, 29, 11), 8, ("himeno_pair_pfmode=2",)),
])
def test_himeno_fused_two_sweep_passes_match_oracle(dims, nn, opts):
    """A ping-pong run of nn/2 >= 3 iterations executes fused two-sweep passes (himeno_pair.cu):
    both p grids bit-identical to the sweep-by-sweep schedule of the oracle, with and without the
    residual emit (whose grid must hold the last sweep's ss*ss)."""
    import os
    from physis_b200 import api
    for gosa in (False, True):
        seen = {}
        a = H.run_himeno(H.oracle_port(), dims, nn, gosa=gosa, seed=7)
        lib = H.b200_programs()
        # the program calls PSInit itself: options reach it through the environment
        os.environ["PHYSIS_B200_OPTIONS"] = ",".join(opts)
        try:
            b = H.run_himeno(lib, dims, nn, gosa=gosa, seed=7,
                             before_finalize=lambda: seen.update(n=int(api.stats().fused_pairs)))
        finally:
            os.environ.pop("PHYSIS_B200_OPTIONS", None)
        assert seen["n"] == ((nn // 2 - 1) & ~1) > 0, seen
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        if gosa:
            assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))
            exact = float(np.sum(a[3].astype(np.float64)))
            assert abs(b[2] - exact) <= 8e-6 * abs(exact)


def test_himeno_pair_not_fused_when_boundary_cells_differ():
    """The fused pass takes the intermediate field's boundary cells from the grid it reads; when
    the two p grids differ there, the run goes sweep by sweep (and still matches the oracle)."""
    from physis_b200 import api
    seen = {}
    dims, nn = (64, 24, 16), 8
    a = H.run_himeno(H.oracle_port(), dims, nn, seed=9, p1_differs=True)
    b = H.run_himeno(H.b200_programs(), dims, nn, seed=9, p1_differs=True,
                     before_finalize=lambda: seen.update(n=int(api.stats().fused_pairs)))
    assert seen["n"] == 0
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_himeno_default_initial_condition():
    a = H.run_himeno(H.oracle_port(), (64, 32, 32), 4)
    b = H.run_himeno(H.b200_programs(), (64, 32, 32), 4)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))


@pytest.mark.parametrize("shape,count", [((64, 16, 8), 4), ((128, 32, 5), 2), ((256, 16, 3), 2),
                                          ((130, 16, 4), 2), ((2, 16, 1), 2)])
def test_periodic_staggered_matches_oracle(shape, count):
    nx, ny, nz = shape
    u, kap = H.pstag_inputs(nx, ny, nz)
    want = H.run_pstag(H.oracle_port(), u, kap, nx, ny, nz, count)
    got = H.run_pstag(H.b200_programs(), u, kap, nx, ny, nz, count)
    assert np.array_equal(want.view(np.uint64), got.view(np.uint64))


def test_periodic_staggered_generic_fallback():
    nx, ny, nz = 10, 6, 4   # ny not a multiple of the tile height -> generic kernel
    u, kap = H.pstag_inputs(nx, ny, nz)
    want = H.run_pstag(H.oracle_port(), u, kap, nx, ny, nz, 2)
    got = H.run_pstag(H.b200_programs(), u, kap, nx, ny, nz, 2)
    assert np.array_equal(want.view(np.uint64), got.view(np.uint64))


def test_plan_cache_iter1_loop_and_invalidation():
    """`for (...) PSStencilRun(..., 1)` -- the common Physis idiom -- reuses the prepared plans
    (TMA descriptors, launch shape) of the first call; freeing a grid or changing an option drops
    them.  Results stay the oracle's bits throughout."""
    from physis_b200 import api
    co64 = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44])
    co = [float(np.float32(c)) for c in co64]
    api.PSInit(["t"], 3, (256, 40, 24))
    for round_, shape in enumerate([(256, 40, 24), (128, 40, 24), (256, 40, 24)]):
        nx, ny, nz = shape
        a, b = api.Grid(shape, api.PS_FLOAT), api.Grid(shape, api.PS_FLOAT)
        rng = np.random.default_rng(round_)
        f0 = rng.random(nx * ny * nz, dtype=np.float32)
        a.copyin(f0)
        dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
        api.rt().__PSB200ResetStats()
        for _ in range(5):
            api.stencil_run(1, [d0, d1])
        st = api.stats()
        assert int(st.kernel_launches) == 10 and int(st.plan_cache_hits) == 8
        api.stencil_run(4, [d0, d1])       # the same descriptors, now with fused passes
        assert int(api.stats().plan_cache_hits) == 10
        if round_ == 1:
            api.set_option("star7_zc=3")   # drops every plan
            api.rt().__PSB200ResetStats()
            api.stencil_run(1, [d0, d1])
            assert int(api.stats().plan_cache_hits) == 0
            api.set_option("star7_zc=0")
            steps = 20
        else:
            steps = 18
        want = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), steps)
        assert np.array_equal(a.copyout().view(np.uint32), want.view(np.uint32))
        a.free()                           # the next round's grids may reuse these addresses
        b.free()
    api.PSFinalize()


def test_residual_emission_is_kept_when_another_stencil_of_the_run_reads_the_grid():
    """Inside one PSStencilRun a residual-emitting Himeno sweep whose grid a later iteration
    overwrites runs in its plain form -- but not when another stencil of the same run uses that
    grid (here a 7-point sweep reads it and feeds p0): the run of 3 iterations must equal three
    runs of one iteration, bit for bit."""
    from physis_b200 import api
    dims = (64, 24, 16)
    ne = int(np.prod(dims))
    names = ["p0", "p1", "a0", "a1", "a2", "a3", "b0", "b1", "b2", "c0", "c1", "c2", "bnd", "wrk1", "gosa"]
    co = [float(np.float32(c)) for c in (0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44)]
    results = []
    for split in (False, True):
        api.PSInit(["t"], 3, dims)
        rng = np.random.default_rng(12)
        g = {n: api.Grid(dims, api.PS_FLOAT) for n in names}
        for n in names[:-1]:
            g[n].copyin(rng.random(ne, dtype=np.float32) * 0.1)
        inner = api.PSDomain3DNew(1, dims[0] - 1, 1, dims[1] - 1, 1, dims[2] - 1)
        whole = api.PSDomain3DNew(0, dims[0], 0, dims[1], 0, dims[2])
        d = [api.stencil_desc(api.KIND_HIMENO19_GOSA, inner, [g[n] for n in names], [0.8]),
             api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, whole, [g["gosa"], g["p0"]], co)]
        if split:
            for _ in range(3):
                api.stencil_run(1, d)
        else:
            api.stencil_run(3, d)
        results.append([g[n].copyout() for n in ("p0", "p1", "gosa")])
        for n in names:
            g[n].free()
        api.PSFinalize()
    for a, b in zip(*results):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
