/*
 * Hand-emitted `b200`-target translation of
 *   /root/reference/examples/diffusion-benchmark/diffusion3d_physis.c
 * i.e. what a B200RuntimeBuilder (a CUDARuntimeBuilder subclass overriding
 * BuildRunKernelFunc / BuildRunFuncBody, translator/cuda_runtime_builder.cc:
 * 1392-1410,1465-1583) would emit.  Host-side shape is the CUDA target's
 * (stencil-map struct, __PSStencilMap_<k>, __PSStencilRun_<id>); the
 * differences are confined to the run function:
 *   - the kernel body was recognised as the 7-pt clamped diffusion family, so
 *     the descriptor names PSB200_KIND_DIFFUSION7_CLAMP (hand-written sm_100a
 *     sweep in the runtime) ...
 *   - ... and still carries the generic per-point kernel + launch stub, used
 *     when a grid shape is outside the specialised kernel's preconditions.
 * The translator (ROSE) cannot be built here, hence hand emission.
 */
#define PHYSIS_B200
#include "physis/physis.h"
#include "physis/physis_b200_generic.cuh"

#define REAL float

static __PSGrid *f1g;
static __PSGrid *f2g;

extern "C" {

void initialize_physis(int argc, char **argv, int nx, int ny, int nz) {
  PSInit(&argc, &argv, 3, nx, ny, nz);
}

void initialize_benchmark_physis(int nx, int ny, int nz) {
  {
    PSVectorInt dims = {nx, ny, nz};
    __PSGridTypeInfo type_info = {PS_FLOAT, sizeof(float), 0, NULL};
    f1g = __PSGridNew(&type_info, 3, dims, NULL);
  }
  {
    PSVectorInt dims = {nx, ny, nz};
    __PSGridTypeInfo type_info = {PS_FLOAT, sizeof(float), 0, NULL};
    f2g = __PSGridNew(&type_info, 3, dims, NULL);
  }
}

void finalize_benchmark_physis(void) {
  __PSGridFree(f1g, NULL);
  __PSGridFree(f2g, NULL);
  PSFinalize();
}

}  // extern "C"

/* user kernel, Get/Emit rewritten to device offsets (cuda_runtime_builder.cc:152-186) */
__device__ static inline void kernel_physis(const int x, const int y, const int z,
                                            __PSGrid3DFloat_dev *g1, __PSGrid3DFloat_dev *g2,
                                            REAL ce, REAL cw, REAL cn, REAL cs,
                                            REAL ct, REAL cb, REAL cc) {
  int nx, ny, nz;
  nx = __PSGridDimDev(g1, 0);
  ny = __PSGridDimDev(g1, 1);
  nz = __PSGridDimDev(g1, 2);

  REAL c, w, e, n, s, b, t;
  c = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)];
  if (x == 0)    w = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)]; else w = g1->p[__PSGridGetOffset3DDev(g1, x-1, y, z)];
  if (x == nx-1) e = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)]; else e = g1->p[__PSGridGetOffset3DDev(g1, x+1, y, z)];
  if (y == 0)    n = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)]; else n = g1->p[__PSGridGetOffset3DDev(g1, x, y-1, z)];
  if (y == ny-1) s = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)]; else s = g1->p[__PSGridGetOffset3DDev(g1, x, y+1, z)];
  if (z == 0)    b = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)]; else b = g1->p[__PSGridGetOffset3DDev(g1, x, y, z-1)];
  if (z == nz-1) t = g1->p[__PSGridGetOffset3DDev(g1, x, y, z)]; else t = g1->p[__PSGridGetOffset3DDev(g1, x, y, z+1)];
  g2->p[__PSGridGetOffset3DDev(g2, x, y, z)] =
      cc*c + cw*w + ce*e + cs*s
      + cn*n + cb*b + ct*t;
  return;
}

struct __PSStencil_kernel_physis {
  PSDomain3D dom;
  __PSGrid *g1;
  int g1_index;
  __PSGrid *g2;
  int g2_index;
  REAL ce, cw, cn, cs, ct, cb, cc;
};

static struct __PSStencil_kernel_physis __PSStencilMap_kernel_physis(
    PSDomain3D dom, __PSGrid *g1, __PSGrid *g2,
    REAL ce, REAL cw, REAL cn, REAL cs, REAL ct, REAL cb, REAL cc) {
  struct __PSStencil_kernel_physis stencil = {
      dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), ce, cw, cn, cs, ct, cb, cc};
  return stencil;
}

__global__ void __PSStencilRun_kernel_physis(__PSDomain dom, int zchunk,
                                             __PSGrid3DFloat_dev g1, __PSGrid3DFloat_dev g2,
                                             REAL ce, REAL cw, REAL cn, REAL cs,
                                             REAL ct, REAL cb, REAL cc) {
  __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)
    kernel_physis(x, y, z, &g1, &g2, ce, cw, cn, cs, ct, cb, cc);
  __PSB200_FOREACH_POINT_END
}

static void __PSStencilLaunch_kernel_physis(const void *sv, const __PSDomain *dom,
        __PSB200Stream stream) {
  const struct __PSStencil_kernel_physis *s = (const struct __PSStencil_kernel_physis *)sv;
  __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);
  __PSStencilRun_kernel_physis<<<sh.grid, sh.block, 0, (cudaStream_t)stream>>>(
      *dom, sh.zchunk, *((__PSGrid3DFloat_dev *)(s->g1->dev)),
      *((__PSGrid3DFloat_dev *)(s->g2->dev)), s->ce, s->cw, s->cn, s->cs, s->ct, s->cb, s->cc);
}

static void __PSStencilDescribe_kernel_physis(const struct __PSStencil_kernel_physis *s,
                                              __PSB200StencilDesc *d) {
  memset(d, 0, sizeof(*d));
  d->kind = PSB200_KIND_DIFFUSION7_CLAMP;
  d->elm_type = PS_FLOAT;
  d->dom = s->dom;
  d->num_grids = 2;
  d->grids[0] = s->g1; d->members[0] = -1;
  d->grids[1] = s->g2; d->members[1] = -1;
  d->num_scalars = 7;
  d->scalars[0] = s->ce; d->scalars[1] = s->cw; d->scalars[2] = s->cn; d->scalars[3] = s->cs;
  d->scalars[4] = s->ct; d->scalars[5] = s->cb; d->scalars[6] = s->cc;
  d->stencil = s;
  d->launch = __PSStencilLaunch_kernel_physis;
  d->name = "kernel_physis";
}

static float __PSStencilRun_0(int iter, struct __PSStencil_kernel_physis s0,
                              struct __PSStencil_kernel_physis s1) {
  __PSB200StencilDesc d[2];
  __PSStencilDescribe_kernel_physis(&s0, &d[0]);
  __PSStencilDescribe_kernel_physis(&s1, &d[1]);
  return __PSB200StencilRun(iter, 2, d);
}

extern "C" {

void run_kernel_physis(int count, REAL *f1_host,
                       int nx, int ny, int nz,
                       REAL ce, REAL cw, REAL cn, REAL cs,
                       REAL ct, REAL cb, REAL cc) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  __PSGridCopyin(f1g, f1_host, NULL);

  __PSStencilRun_0(count/2,
                   __PSStencilMap_kernel_physis(dom, f1g, f2g,
                                                ce, cw, cn, cs, ct, cb, cc),
                   __PSStencilMap_kernel_physis(dom, f2g, f1g,
                                                ce, cw, cn, cs, ct, cb, cc));

  __PSGridCopyout(f1g, f1_host, NULL);
}

/* bench hooks (not part of the translated program): sweeps on resident data */
void run_sweeps_only_physis(int count, int nx, int ny, int nz,
                            REAL ce, REAL cw, REAL cn, REAL cs,
                            REAL ct, REAL cb, REAL cc) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  __PSStencilRun_0(count/2,
                   __PSStencilMap_kernel_physis(dom, f1g, f2g,
                                                ce, cw, cn, cs, ct, cb, cc),
                   __PSStencilMap_kernel_physis(dom, f2g, f1g,
                                                ce, cw, cn, cs, ct, cb, cc));
}
/* the common Physis idiom `for (i < n) PSStencilRun(..., 1);` (one run call per iteration) */
void run_sweeps_iter1_physis(int count, int nx, int ny, int nz,
                             REAL ce, REAL cw, REAL cn, REAL cs,
                             REAL ct, REAL cb, REAL cc) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  for (int i = 0; i < count / 2; ++i)
    __PSStencilRun_0(1,
                     __PSStencilMap_kernel_physis(dom, f1g, f2g,
                                                  ce, cw, cn, cs, ct, cb, cc),
                     __PSStencilMap_kernel_physis(dom, f2g, f1g,
                                                  ce, cw, cn, cs, ct, cb, cc));
}
void copyin_physis(const REAL *f1_host) { __PSGridCopyin(f1g, f1_host, NULL); }
/* multi-GPU bench hooks: this rank's slab only (see __PSB200GridCopyinLocal) */
void copyin_local_physis(const REAL *slab) { __PSB200GridCopyinLocal(f1g, slab); }
void copyout_local_physis(REAL *slab) { __PSB200GridCopyoutLocal(f1g, slab); }
void local_size_physis(int *z_off, int *z_len) { __PSB200GridLocalSize(f1g, z_off, z_len); }
void run_kernel_local_physis(int count, REAL *slab, int nx, int ny, int nz,
                             REAL ce, REAL cw, REAL cn, REAL cs, REAL ct, REAL cb, REAL cc) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  __PSB200GridCopyinLocal(f1g, slab);
  __PSStencilRun_0(count/2,
                   __PSStencilMap_kernel_physis(dom, f1g, f2g, ce, cw, cn, cs, ct, cb, cc),
                   __PSStencilMap_kernel_physis(dom, f2g, f1g, ce, cw, cn, cs, ct, cb, cc));
  __PSB200GridCopyoutLocal(f1g, slab);
}
void copyout_physis(REAL *f1_host) { __PSGridCopyout(f1g, f1_host, NULL); }

/* same sweeps forced through the generic per-point kernel (tests / comparison) */
void run_kernel_physis_generic(int count, REAL *f1_host, int nx, int ny, int nz,
                               REAL ce, REAL cw, REAL cn, REAL cs, REAL ct, REAL cb, REAL cc) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  __PSGridCopyin(f1g, f1_host, NULL);
  struct __PSStencil_kernel_physis s0 =
      __PSStencilMap_kernel_physis(dom, f1g, f2g, ce, cw, cn, cs, ct, cb, cc);
  struct __PSStencil_kernel_physis s1 =
      __PSStencilMap_kernel_physis(dom, f2g, f1g, ce, cw, cn, cs, ct, cb, cc);
  __PSB200StencilDesc d[2];
  __PSStencilDescribe_kernel_physis(&s0, &d[0]);
  __PSStencilDescribe_kernel_physis(&s1, &d[1]);
  d[0].kind = d[1].kind = PSB200_KIND_GENERIC;
  __PSB200StencilRun(count/2, 2, d);
  __PSGridCopyout(f1g, f1_host, NULL);
}

}  // extern "C"
