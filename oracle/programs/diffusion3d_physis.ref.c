/*
 * ORACLE program (test infrastructure only).
 *
 * Hand-emitted REFERENCE-target translation of
 *   /root/reference/examples/diffusion-benchmark/diffusion3d_physis.c
 * in the shape `physisc --ref` produces (the translator needs ROSE and cannot
 * be built here).  Shape follows translator/reference_runtime_builder.cc:
 *   BuildGridGet/Emit/Offset :82-100,142-177,220-247  -> ((T*)g->p)[__PSGridGetOffset3D(..)]
 *   BuildStencilMapType/BuildMap :391-541             -> struct __PSStencil_<k>, __PSStencilMap_<k>
 *   BuildRunKernelFuncBody :605-664                   -> z,y,x loops, `<= max-1`
 *   BuildRunFuncBody/LoopBody :837-893                -> __PSStencilRun_0
 *   BuildTypeInfo :959-1065 / TranslateNew reference_translator.cc:189-240
 * The kernel body is the user's, untouched apart from the Get/Emit rewrites
 * (diffusion3d_physis.c:29-58): clamp boundaries by branches, and
 *   cc*c + cw*w + ce*e + cs*s + cn*n + cb*b + ct*t   left to right in fp32.
 * Compile WITHOUT -ffast-math and without -march (no FMA contraction), as
 * examples/diffusion-benchmark/Makefile.cmake:2-3 does.
 */
#define PHYSIS_REF
#include "physis/physis.h"

#define REAL float

static __PSGrid *f1g;
static __PSGrid *f2g;

void initialize_physis(int argc, char **argv, int nx, int ny, int nz) {
  PSInit(&argc, &argv, 3, nx, ny, nz);
}

void initialize_benchmark_physis(int nx, int ny, int nz) {
  {
    PSVectorInt dims = {nx, ny, nz};
    __PSGridTypeInfo type_info = {PS_FLOAT, sizeof(float), 0, NULL};
    f1g = __PSGridNew(&type_info, 3, dims);
  }
  {
    PSVectorInt dims = {nx, ny, nz};
    __PSGridTypeInfo type_info = {PS_FLOAT, sizeof(float), 0, NULL};
    f2g = __PSGridNew(&type_info, 3, dims);
  }
}

void finalize_benchmark_physis(void) {
  PSGridFree(f1g);
  PSGridFree(f2g);
  PSFinalize();
}

static inline void kernel_physis(const int x, const int y, const int z,
                                 __PSGrid *g1, __PSGrid *g2,
                                 REAL ce, REAL cw, REAL cn, REAL cs,
                                 REAL ct, REAL cb, REAL cc) {
  int nx, ny, nz;
  nx = PSGridDim(g1, 0);
  ny = PSGridDim(g1, 1);
  nz = PSGridDim(g1, 2);

  REAL c, w, e, n, s, b, t;
  c = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)];
  if (x == 0)    w = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)]; else w = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x-1, y, z)];
  if (x == nx-1) e = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)]; else e = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x+1, y, z)];
  if (y == 0)    n = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)]; else n = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y-1, z)];
  if (y == ny-1) s = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)]; else s = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y+1, z)];
  if (z == 0)    b = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)]; else b = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z-1)];
  if (z == nz-1) t = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z)]; else t = ((float *)(g1->p))[__PSGridGetOffset3D(g1, x, y, z+1)];
  ((float *)(g2->p))[__PSGridGetOffset3D(g2, x, y, z)] =
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

static void __PSStencilRun_kernel_physis(const struct __PSStencil_kernel_physis *const s) {
  int i3;
  for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {
    int i2;
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {
      int i1;
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) {
        kernel_physis(i1, i2, i3, s->g1, s->g2,
                      s->ce, s->cw, s->cn, s->cs, s->ct, s->cb, s->cc);
      }
    }
  }
}

static float __PSStencilRun_0(int iter, struct __PSStencil_kernel_physis s0,
                              struct __PSStencil_kernel_physis s1) {
  int i;
  for (i = 0; i < iter; i++) {
    __PSStencilRun_kernel_physis(&s0);
    __PSStencilRun_kernel_physis(&s1);
  }
  return 0.0f;
}

void run_kernel_physis(int count, REAL *f1_host,
                       int nx, int ny, int nz,
                       REAL ce, REAL cw, REAL cn, REAL cs,
                       REAL ct, REAL cb, REAL cc) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  PSGridCopyin(f1g, f1_host);

  __PSStencilRun_0(count/2,
                   __PSStencilMap_kernel_physis(dom, f1g, f2g,
                                                ce, cw, cn, cs, ct, cb, cc),
                   __PSStencilMap_kernel_physis(dom, f2g, f1g,
                                                ce, cw, cn, cs, ct, cb, cc));

  PSGridCopyout(f1g, f1_host);
}

/* Timing-only entry for bench.py's cpu_baseline / reference arm: `count`
 * sweeps on data already copied in (no copyin/copyout), count even. */
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
void copyin_physis(const REAL *f1_host) { PSGridCopyin(f1g, f1_host); }
void copyout_physis(REAL *f1_host) { PSGridCopyout(f1g, f1_host); }
