"""Option autotune=1: the first long PSStencilRun of a shape tries the forms the hand-written
kernels come in (tile shapes, fused passes or single sweeps) on its own first iterations and
keeps the fastest -- the reference's AUTO_TUNING trial iterations (include/physis/runtime.h:32-52,
translator/configuration.cc:27-57) with kernel forms in place of CUDA_BLOCK_SIZE patterns.  All
forms compute the same bits, so a tuned run must still match the oracle exactly."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(256, 40, 24), (512, 16, 20), (1024, 12, 9)])
def test_diffusion_tuned_on_its_own_iterations(shape):
    from physis_b200 import api
    co64 = np.array([0.11, 0.07, 0.13, 0.05, 0.17, 0.03, 0.44])
    co = [float(np.float32(c)) for c in co64]
    nx, ny, nz = shape
    api.PSInit(["t"], 3, shape)
    api.set_option("autotune=1")
    a, b = api.Grid(shape, api.PS_FLOAT), api.Grid(shape, api.PS_FLOAT)
    f0 = np.random.default_rng(7).random(nx * ny * nz, dtype=np.float32)
    a.copyin(f0)
    dom = api.PSDomain3DNew(0, nx, 0, ny, 0, nz)
    d0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [a, b], co)
    d1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [b, a], co)
    api.rt().__PSB200ResetStats()
    api.stencil_run(5, [d0, d1])                  # too short to spend iterations on trials
    assert int(api.stats().autotune_trials) == 0 and api.last_tuning() == ""
    api.stencil_run(60, [d0, d1])                 # up to 9 forms x 6 iterations fit
    trials = int(api.stats().autotune_trials)
    assert 3 <= trials <= 9, trials               # defaults, z chunks, single sweeps, the tile shapes that fit
    assert "ms per iteration" in api.last_tuning()
    want = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), 2 * 65)
    assert np.array_equal(a.copyout().view(np.uint32), want.view(np.uint32))
    prev = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), 2 * 65 - 1)
    assert np.array_equal(b.copyout().view(np.uint32), prev.view(np.uint32))
    api.stencil_run(60, [d0, d1])                 # the shape is settled: no more trials
    assert int(api.stats().autotune_trials) == trials
    want = H.diffusion7_numpy(want, shape, co64.astype(np.float32), 2 * 60)
    assert np.array_equal(a.copyout().view(np.uint32), want.view(np.uint32))
    # a new pair of grids of the same shape starts from the settled form
    c, d = api.Grid(shape, api.PS_FLOAT), api.Grid(shape, api.PS_FLOAT)
    c.copyin(f0)
    e0 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [c, d], co)
    e1 = api.stencil_desc(api.KIND_DIFFUSION7_CLAMP, dom, [d, c], co)
    api.stencil_run(60, [e0, e1])
    assert int(api.stats().autotune_trials) == trials
    want = H.diffusion7_numpy(f0, shape, co64.astype(np.float32), 2 * 60)
    assert np.array_equal(c.copyout().view(np.uint32), want.view(np.uint32))
    # changing an option forgets what was settled
    api.set_option("star7_zc=0")
    api.stencil_run(60, [e0, e1])
    assert int(api.stats().autotune_trials) > trials
    api.PSFinalize()


@pytest.mark.parametrize("gosa", [False, True])
def test_himeno_tuned(monkeypatch, gosa):
    from physis_b200 import api
    dims, nn = (136, 21, 14), 80
    a = H.run_himeno(H.oracle_port(), dims, nn, gosa=gosa, seed=11, omega=0.1)
    monkeypatch.setenv("PHYSIS_B200_OPTIONS", "autotune=1")
    seen = {}
    b = H.run_himeno(H.b200_programs(), dims, nn, gosa=gosa, seed=11, omega=0.1,
                     before_finalize=lambda: seen.update(n=int(api.stats().autotune_trials), s=api.last_tuning()))
    assert seen["n"] == 6, seen   # the defaults, fused regardless, two prefetch forms, single sweeps with 7- and 15-row tiles
    for i in (0, 1, 3):
        assert np.array_equal(a[i].view(np.uint32), b[i].view(np.uint32))
    if gosa:
        assert abs(a[2] - b[2]) <= 1e-3 * abs(a[2])   # the oracle sums sequentially in fp32


def test_config5_tuned(monkeypatch):
    nx, ny, nz = 128, 16, 12
    u, kap = H.pstag_inputs(nx, ny, nz)
    want = H.run_pstag(H.oracle_port(), u, kap, nx, ny, nz, 30)
    monkeypatch.setenv("PHYSIS_B200_OPTIONS", "autotune=1")
    got = H.run_pstag(H.b200_programs(), u, kap, nx, ny, nz, 30)
    assert np.array_equal(want.view(np.uint64), got.view(np.uint64))
