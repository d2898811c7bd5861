"""GPU parity of the fused two-sweep pass (physis_b200/csrc/star7_pair.cu): a ping-pong
pair of whole-grid 7-point sweeps run as one kernel must leave BOTH grids bit-identical to
the sweep-by-sweep schedule (the reference's PSStencilRun semantics,
translator/reference_runtime_builder.cc:837-893) — checked against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

CO = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44])  # ce, cw, cn, cs, ct, cb, cc
# equal neighbour coefficients (the benchmark's isotropic case) select the shared-product form
CO_ISO = np.array([0.1234567] * 6 + [0.2592598])


def _run_pair(shape, iters, dtype, options=(), coeffs=None):
    """PSStencilRun(map(A->B), map(B->A), iters) through the C ABI; returns (A, B, stats)."""
    from physis_b200 import api
    nx, ny, nz = shape
    api.PSInit(dims=shape)
    try:
        for kv in options:
            api.set_option(kv)
        pt = api.PS_FLOAT if dtype == np.float32 else api.PS_DOUBLE
        a, b = api.Grid(shape, pt), api.Grid(shape, pt)
        rng = np.random.default_rng(nx * 31 + ny * 7 + nz)
        f0 = rng.random(nx * ny * nz).astype(dtype)
        g0 = rng.random(nx * ny * nz).astype(dtype)   # B's old contents must not leak
        a.copyin(f0)
        b.copyin(g0)
        dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
        co = [float(dtype(c)) for c in (CO if coeffs is None else coeffs)]
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co, elm_type=pt)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co, elm_type=pt)
        api.rt().__PSB200ResetStats()
        api.stencil_run(iters, [d0, d1])
        st = api.stats()
        fa, fb = a.copyout(), b.copyout()
        a.free()
        b.free()
        return f0, fa, fb, int(st.kernel_launches), int(st.fused_pairs)
    finally:
        api.PSFinalize()


@pytest.mark.parametrize("shape,iters,opts", [
    ((512, 40, 24), 3, ()),
    ((512, 33, 19), 4, ("star7_pair_zc=5",)),
    ((384, 30, 17), 5, ("star7_pair_zc=4",)),
    ((256, 72, 21), 6, ("star7_pair_zc=7",)),
    ((128, 14, 40), 3, ("star7_pair_zc=16",)),
    ((200, 29, 9), 3, ("star7_pair_zc=1",)),
    ((64, 5, 6), 4, ()),
    ((4, 2, 2), 3, ()),
    ((512, 128, 64), 3, ()),
    ((256, 15, 33), 7, ("star7_pair_zc=8",)),
    # rows wider than one tile: x tiles with one-vector seams (BASELINE config 4's 1024 floats)
    ((1024, 40, 24), 3, ()),
    ((1024, 64, 64), 5, ()),
    ((768, 33, 19), 4, ("star7_pair_zc=5",)),
    ((1100, 22, 9), 5, ()),
    ((2048, 21, 11), 3, ("star7_pair_zc=3",)),
    ((640, 41, 10), 3, ("star7_pair_variant=0",)),
    ((5000, 6, 5), 3, ()),
])
def test_fused_pair_fp32_matches_oracle(shape, iters, opts):
    f0, fa, fb, launches, pairs = _run_pair(shape, iters, np.float32, opts)
    assert pairs == ((iters - 1) & ~1) and pairs > 0, "the fused pass did not run"
    assert launches == pairs + 2 * (iters - pairs)
    co = CO.astype(np.float32)
    want_a = H.diffusion7_numpy(f0, shape, co, 2 * iters)
    want_b = H.diffusion7_numpy(f0, shape, co, 2 * iters - 1)
    assert np.array_equal(fa.view(np.uint32), want_a.view(np.uint32))
    assert np.array_equal(fb.view(np.uint32), want_b.view(np.uint32))


def test_fused_pair_fp32_matches_c_oracle():
    # the same through the translated benchmark program, against oracle/liboracle.so
    nx, ny, nz, count = 512, 48, 30, 12
    p = H.diffusion_params(nx, ny, nz)
    f0 = (H.diffusion_initial(nx, ny, nz, p)
          + np.random.default_rng(5).random(nx * ny * nz, dtype=np.float32)).astype(np.float32)
    co = CO.astype(np.float32)
    want = H.run_diffusion(H.oracle_port(), f0, nx, ny, nz, count, co)
    got = H.run_diffusion(H.b200_programs(), f0, nx, ny, nz, count, co)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("shape,iters,opts", [
    ((256, 40, 24), 3, ()),
    ((192, 33, 19), 4, ("star7_pair_zc=5",)),
    ((128, 30, 17), 5, ("star7_pair_zc=4",)),
    ((64, 14, 12), 3, ()),
    ((50, 9, 7), 3, ()),
    ((520, 21, 12), 3, ()),
    ((1024, 19, 9), 4, ("star7_pair_zc=4",)),
])
def test_fused_pair_fp64_matches_oracle(shape, iters, opts):
    f0, fa, fb, launches, pairs = _run_pair(shape, iters, np.float64, opts)
    assert pairs == ((iters - 1) & ~1) and pairs > 0, "the fused pass did not run"
    want_a = H.diffusion7_numpy(f0, shape, CO, 2 * iters)
    want_b = H.diffusion7_numpy(f0, shape, CO, 2 * iters - 1)
    assert np.array_equal(fa.view(np.uint64), want_a.view(np.uint64))
    assert np.array_equal(fb.view(np.uint64), want_b.view(np.uint64))


def test_unfused_schedule_unchanged():
    # star7_fuse=0 keeps the sweep-by-sweep schedule; so do wide rows with x tiling switched off
    f0, fa, fb, launches, pairs = _run_pair((256, 20, 12), 4, np.float32, ("star7_fuse=0",))
    assert pairs == 0 and launches == 8
    want = H.diffusion7_numpy(f0, (256, 20, 12), CO.astype(np.float32), 8)
    assert np.array_equal(fa.view(np.uint32), want.view(np.uint32))
    f0, fa, fb, launches, pairs = _run_pair((1024, 12, 6), 3, np.float32, ("star7_pair_xtile=0",))
    assert pairs == 0 and launches == 6
    want = H.diffusion7_numpy(f0, (1024, 12, 6), CO.astype(np.float32), 6)
    assert np.array_equal(fa.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,iters,opts", [
    ((512, 40, 24), 3, ()),
    ((256, 33, 19), 4, ("star7_pair_zc=5",)),
    ((128, 30, 17), 5, ("star7_pair_zc=4",)),
    ((200, 29, 9), 3, ("star7_pair_zc=1",)),
    ((64, 5, 6), 4, ()),
    ((256, 15, 33), 7, ("star7_pair_zc=8",)),
    ((256, 47, 20), 3, ("star7_iso=0",)),
    ((1024, 31, 20), 3, ()),
    ((1536, 18, 7), 4, ("star7_pair_zc=2",)),
])
def test_fused_pair_equal_coefficients(shape, iters, opts, dtype):
    if dtype == np.float64:
        shape = (shape[0] // 2, shape[1], shape[2])
    f0, fa, fb, launches, pairs = _run_pair(shape, iters, dtype, opts, coeffs=CO_ISO)
    assert pairs == ((iters - 1) & ~1) and pairs > 0, "the fused pass did not run"
    co = CO_ISO.astype(dtype)
    view = np.uint32 if dtype == np.float32 else np.uint64
    want_a = H.diffusion7_numpy(f0, shape, co, 2 * iters)
    want_b = H.diffusion7_numpy(f0, shape, co, 2 * iters - 1)
    assert np.array_equal(fa.view(view), want_a.view(view))
    assert np.array_equal(fb.view(view), want_b.view(view))


def test_full_size_properties_512cubed():
    """BASELINE config 2's grid (512^3 fp32) at a reduced sweep count, through size-independent
    properties: the fused schedule equals the sweep-by-sweep schedule bit for bit on both grids
    (the latter is pinned to the oracle at small sizes), the benchmark's accuracy figure
    (RMS error against the analytic solution, examples/diffusion-benchmark/baseline.cc:51-60)
    is at its expected level, and a constant field is a fixed point (cc + 6 c = 1 up to
    rounding: every point must come out identical)."""
    from physis_b200 import api
    n, iters = 512, 12
    p = H.diffusion_params(n, n, n)
    co = [float(c) for c in p[:7]]
    f0 = H.diffusion_initial(n, n, n, p)
    out = {}
    for fuse in (0, 1):
        api.PSInit(dims=(n, n, n))
        api.set_option(f"star7_fuse={fuse}")
        a, b = api.Grid((n, n, n), api.PS_FLOAT), api.Grid((n, n, n), api.PS_FLOAT)
        a.copyin(f0)
        dom = api.PSDomain3DNew(0, n, 0, n, 0, n)
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
        api.rt().__PSB200ResetStats()
        api.stencil_run(iters, [d0, d1])
        assert int(api.stats().fused_pairs) == (10 if fuse else 0)
        out[fuse] = (a.copyout(), b.copyout())
        if fuse:
            const = np.full(n ** 3, 0.3125, np.float32)
            a.copyin(const)
            api.stencil_run(iters, [d0, d1])
            c = a.copyout()
            assert np.all(c == c[0]) and abs(float(c[0]) - 0.3125) < 1e-5
        a.free()
        b.free()
        api.PSFinalize()
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))
    # accuracy against the analytic solution after 2*iters steps
    t = 2 * iters * float(p[10])
    exact = H.diffusion_initial(n, n, n, p, time=t)
    rms = float(np.sqrt(np.mean((out[1][0].astype(np.float64) - exact.astype(np.float64)) ** 2)))
    assert rms < 1e-5, rms


def test_config2_full_run_fused_equals_sweep_by_sweep():
    """BASELINE config 2 in full -- 512^3 fp32, 1000 sweeps, the benchmark's own coefficients
    (equal neighbour coefficients: the shared-product form) -- the fused schedule (498 two-sweep
    passes + 4 single sweeps) against the sweep-by-sweep schedule: both grids bit-identical."""
    from physis_b200 import api
    n, iters = 512, 500
    p = H.diffusion_params(n, n, n)
    co = [float(c) for c in p[:7]]
    f0 = H.diffusion_initial(n, n, n, p)
    out = {}
    for fuse in (0, 1):
        api.PSInit(dims=(n, n, n))
        api.set_option(f"star7_fuse={fuse}")
        a, b = api.Grid((n, n, n), api.PS_FLOAT), api.Grid((n, n, n), api.PS_FLOAT)
        a.copyin(f0)
        dom = api.PSDomain3DNew(0, n, 0, n, 0, n)
        d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
        d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
        api.rt().__PSB200ResetStats()
        api.stencil_run(iters, [d0, d1])
        st = api.stats()
        assert int(st.fused_pairs) == (498 if fuse else 0)
        assert int(st.kernel_launches) == (502 if fuse else 1000)
        out[fuse] = (a.copyout(), b.copyout())
        a.free()
        b.free()
        api.PSFinalize()
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))
    # the field is still the smooth decaying mode: compare with the analytic solution
    exact = H.diffusion_initial(n, n, n, p, time=2 * iters * float(p[10]))
    rms = float(np.sqrt(np.mean((out[1][0].astype(np.float64) - exact.astype(np.float64)) ** 2)))
    assert rms < 1e-4, rms
