/*
 * ORACLE program (test infrastructure only).
 *
 * Hand-emitted REFERENCE-target translation of
 *   /root/reference/examples/himeno/himenobmtxpa_physis.c
 * (jacobi_kernel :331-361, jacobi :364-393, mat_set :298-312, mat_set_init
 * :314-329, grid creation and initial values :120-151), library-ised so that
 * tests can drive it through ctypes instead of `main`.  Translation shape as
 * in diffusion3d_physis.ref.c (translator/reference_runtime_builder.cc).
 *
 * Added for BASELINE config 3 ("with PSReduce residual"): jacobi_kernel_gosa /
 * himeno_jacobi_gosa, the translation of examples/dsl/himeno_gosa.c in this
 * repo — the same kernel plus `PSGridEmit(gosa_g, ss*ss)` and a trailing
 * `PSReduce(&gosa, PS_SUM, gosa_g)`, which is how the original benchmark's
 * `gosa += ss*ss` (himenobmtxpa_original.c:334) is expressed in the DSL
 * (hinted at himenobmtxpa_physis.c:387-391).
 */
#define PHYSIS_REF
#include "physis/physis.h"

enum { P0, P1, BND, WRK1, A0, A1, A2, A3, B0, B1, B2, C0, C1, C2, GOSA, NGRIDS };
static __PSGrid *G[NGRIDS];
static float omega = 0.8;

static __PSGrid *new_float3d(int nx, int ny, int nz) {
  PSVectorInt dims = {nx, ny, nz};
  __PSGridTypeInfo type_info = {PS_FLOAT, sizeof(float), 0, NULL};
  return __PSGridNew(&type_info, 3, dims);
}

static void mat_set(__PSGrid *mat, float val, float *buf) {
  int i, j, k;
  size_t x = 0;
  for (i = 0; i < PSGridDim(mat, 0); i++)
    for (j = 0; j < PSGridDim(mat, 1); j++)
      for (k = 0; k < PSGridDim(mat, 2); k++) {
        buf[x] = val;
        ++x;
      }
  PSGridCopyin(mat, buf);
}

static void mat_set_init(__PSGrid *Mat, float *buf) {
  int i, j, k;
  int d0 = PSGridDim(Mat, 2);
  size_t x = 0;
  for (k = 0; k < PSGridDim(Mat, 2); k++)
    for (j = 0; j < PSGridDim(Mat, 1); j++)
      for (i = 0; i < PSGridDim(Mat, 0); i++) {
        float v = (float)(k * k) / ((d0 - 1) * (d0 - 1));
        buf[x] = v;
        ++x;
      }
  PSGridCopyin(Mat, buf);
}

void himeno_init(int mimax, int mjmax, int mkmax) {
  int argc = 0;
  char **argv = NULL;
  PSInit(&argc, &argv, 3, mimax, mjmax, mkmax);
  for (int g = 0; g < NGRIDS; ++g) G[g] = new_float3d(mimax, mjmax, mkmax);
  float *host_buf = (float *)malloc((size_t)mimax * mjmax * mkmax * sizeof(float));
  mat_set_init(G[P0], host_buf);
  mat_set_init(G[P1], host_buf);
  mat_set(G[BND], 1.0, host_buf);
  mat_set(G[A0], 1.0, host_buf);
  mat_set(G[A1], 1.0, host_buf);
  mat_set(G[A2], 1.0, host_buf);
  mat_set(G[A3], 1.0 / 6.0, host_buf);
  mat_set(G[B0], 0.0, host_buf);
  mat_set(G[B1], 0.0, host_buf);
  mat_set(G[B2], 0.0, host_buf);
  mat_set(G[C0], 1.0, host_buf);
  mat_set(G[C1], 1.0, host_buf);
  mat_set(G[C2], 1.0, host_buf);
  free(host_buf);
}

/* test hooks: overwrite / read any grid, change omega */
void himeno_set_grid(int which, const float *buf) { PSGridCopyin(G[which], buf); }
void himeno_get_grid(int which, float *buf) { PSGridCopyout(G[which], buf); }
void himeno_set_omega(float w) { omega = w; }

void himeno_finalize(void) {
  for (int g = 0; g < NGRIDS; ++g) PSGridFree(G[g]);
  PSFinalize();
}

#define GET(g, i, j, k) (((float *)((g)->p))[__PSGridGetOffset3D((g), (i), (j), (k))])

static inline void jacobi_kernel(int i, int j, int k,
                                 __PSGrid *p0, __PSGrid *p1,
                                 __PSGrid *a0, __PSGrid *a1, __PSGrid *a2,
                                 __PSGrid *a3, __PSGrid *b0, __PSGrid *b1,
                                 __PSGrid *b2, __PSGrid *c0, __PSGrid *c1,
                                 __PSGrid *c2, __PSGrid *bnd, __PSGrid *wrk1,
                                 float omega) {
  float s0, ss;
  s0 = GET(a0, i, j, k) * GET(p0, i, j, k+1)
      + GET(a1, i, j, k) * GET(p0, i, j+1, k)
      + GET(a2, i, j, k) * GET(p0, i+1, j, k)
      + GET(b0, i, j, k)
      * ( GET(p0, i, j+1, k+1) - GET(p0, i, j-1, k+1)
          - GET(p0, i, j+1, k-1) + GET(p0, i, j-1, k-1) )
      + GET(b1, i, j, k)
      * ( GET(p0, i+1, j+1, k) - GET(p0, i+1, j-1, k)
          - GET(p0, i-1, j+1, k) + GET(p0, i-1, j-1, k) )
      + GET(b2, i, j, k)
      * ( GET(p0, i+1, j, k+1) - GET(p0, i+1, j, k-1)
          - GET(p0, i-1, j, k+1) + GET(p0, i-1, j, k-1) )
      + GET(c0, i, j, k) * GET(p0, i, j, k-1)
      + GET(c1, i, j, k) * GET(p0, i, j-1, k)
      + GET(c2, i, j, k) * GET(p0, i-1, j, k)
      + GET(wrk1, i, j, k);
  ss = (s0 * GET(a3, i, j, k) - GET(p0, i, j, k))
      * GET(bnd, i, j, k);
  float v = GET(p0, i, j, k) + omega * ss;
  GET(p1, i, j, k) = v;
  return;
}

static inline void jacobi_kernel_gosa(int i, int j, int k,
                                      __PSGrid *p0, __PSGrid *p1,
                                      __PSGrid *a0, __PSGrid *a1, __PSGrid *a2,
                                      __PSGrid *a3, __PSGrid *b0, __PSGrid *b1,
                                      __PSGrid *b2, __PSGrid *c0, __PSGrid *c1,
                                      __PSGrid *c2, __PSGrid *bnd, __PSGrid *wrk1,
                                      __PSGrid *gosa_g, float omega) {
  float s0, ss;
  s0 = GET(a0, i, j, k) * GET(p0, i, j, k+1)
      + GET(a1, i, j, k) * GET(p0, i, j+1, k)
      + GET(a2, i, j, k) * GET(p0, i+1, j, k)
      + GET(b0, i, j, k)
      * ( GET(p0, i, j+1, k+1) - GET(p0, i, j-1, k+1)
          - GET(p0, i, j+1, k-1) + GET(p0, i, j-1, k-1) )
      + GET(b1, i, j, k)
      * ( GET(p0, i+1, j+1, k) - GET(p0, i+1, j-1, k)
          - GET(p0, i-1, j+1, k) + GET(p0, i-1, j-1, k) )
      + GET(b2, i, j, k)
      * ( GET(p0, i+1, j, k+1) - GET(p0, i+1, j, k-1)
          - GET(p0, i-1, j, k+1) + GET(p0, i-1, j, k-1) )
      + GET(c0, i, j, k) * GET(p0, i, j, k-1)
      + GET(c1, i, j, k) * GET(p0, i, j-1, k)
      + GET(c2, i, j, k) * GET(p0, i-1, j, k)
      + GET(wrk1, i, j, k);
  ss = (s0 * GET(a3, i, j, k) - GET(p0, i, j, k))
      * GET(bnd, i, j, k);
  float v = GET(p0, i, j, k) + omega * ss;
  GET(p1, i, j, k) = v;
  GET(gosa_g, i, j, k) = ss * ss;
  return;
}

struct __PSStencil_jacobi_kernel {
  PSDomain3D dom;
  __PSGrid *p0; int p0_index;
  __PSGrid *p1; int p1_index;
  __PSGrid *a0; int a0_index;
  __PSGrid *a1; int a1_index;
  __PSGrid *a2; int a2_index;
  __PSGrid *a3; int a3_index;
  __PSGrid *b0; int b0_index;
  __PSGrid *b1; int b1_index;
  __PSGrid *b2; int b2_index;
  __PSGrid *c0; int c0_index;
  __PSGrid *c1; int c1_index;
  __PSGrid *c2; int c2_index;
  __PSGrid *bnd; int bnd_index;
  __PSGrid *wrk1; int wrk1_index;
  float omega;
};

static struct __PSStencil_jacobi_kernel __PSStencilMap_jacobi_kernel(
    PSDomain3D dom, __PSGrid *p0, __PSGrid *p1, __PSGrid *a0, __PSGrid *a1,
    __PSGrid *a2, __PSGrid *a3, __PSGrid *b0, __PSGrid *b1, __PSGrid *b2,
    __PSGrid *c0, __PSGrid *c1, __PSGrid *c2, __PSGrid *bnd, __PSGrid *wrk1,
    float omega) {
  struct __PSStencil_jacobi_kernel stencil = {
      dom, p0, __PSGridGetID(p0), p1, __PSGridGetID(p1), a0, __PSGridGetID(a0),
      a1, __PSGridGetID(a1), a2, __PSGridGetID(a2), a3, __PSGridGetID(a3),
      b0, __PSGridGetID(b0), b1, __PSGridGetID(b1), b2, __PSGridGetID(b2),
      c0, __PSGridGetID(c0), c1, __PSGridGetID(c1), c2, __PSGridGetID(c2),
      bnd, __PSGridGetID(bnd), wrk1, __PSGridGetID(wrk1), omega};
  return stencil;
}

static void __PSStencilRun_jacobi_kernel(const struct __PSStencil_jacobi_kernel *const s) {
  int i3;
  for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {
    int i2;
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {
      int i1;
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) {
        jacobi_kernel(i1, i2, i3, s->p0, s->p1, s->a0, s->a1, s->a2, s->a3,
                      s->b0, s->b1, s->b2, s->c0, s->c1, s->c2, s->bnd, s->wrk1,
                      s->omega);
      }
    }
  }
}

static float __PSStencilRun_0(int iter, struct __PSStencil_jacobi_kernel s0,
                              struct __PSStencil_jacobi_kernel s1) {
  int i;
  for (i = 0; i < iter; i++) {
    __PSStencilRun_jacobi_kernel(&s0);
    __PSStencilRun_jacobi_kernel(&s1);
  }
  return 0.0f;
}

struct __PSStencil_jacobi_kernel_gosa {
  struct __PSStencil_jacobi_kernel base; /* same members, then: */
  __PSGrid *gosa_g; int gosa_g_index;
};

static void __PSStencilRun_jacobi_kernel_gosa(
    const struct __PSStencil_jacobi_kernel_gosa *const sg) {
  const struct __PSStencil_jacobi_kernel *const s = &sg->base;
  int i3;
  for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {
    int i2;
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {
      int i1;
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) {
        jacobi_kernel_gosa(i1, i2, i3, s->p0, s->p1, s->a0, s->a1, s->a2, s->a3,
                           s->b0, s->b1, s->b2, s->c0, s->c1, s->c2, s->bnd,
                           s->wrk1, sg->gosa_g, s->omega);
      }
    }
  }
}

static float __PSStencilRun_1(int iter, struct __PSStencil_jacobi_kernel_gosa s0,
                              struct __PSStencil_jacobi_kernel_gosa s1) {
  int i;
  for (i = 0; i < iter; i++) {
    __PSStencilRun_jacobi_kernel_gosa(&s0);
    __PSStencilRun_jacobi_kernel_gosa(&s1);
  }
  return 0.0f;
}

/* jacobi(): himenobmtxpa_physis.c:364-393 (gosa stays 0 in the Physis version) */
float himeno_jacobi(int nn) {
  float gosa = 0.0f;
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  assert(nn % 2 == 0);
  __PSStencilRun_0(nn / 2,
                   __PSStencilMap_jacobi_kernel(innerDom, p0, p1, G[A0], G[A1], G[A2], G[A3],
                                                G[B0], G[B1], G[B2], G[C0], G[C1], G[C2],
                                                G[BND], G[WRK1], omega),
                   __PSStencilMap_jacobi_kernel(innerDom, p1, p0, G[A0], G[A1], G[A2], G[A3],
                                                G[B0], G[B1], G[B2], G[C0], G[C1], G[C2],
                                                G[BND], G[WRK1], omega));
  return gosa;
}

/* examples/dsl/himeno_gosa.c: residual emitted every sweep, reduced at the end */
float himeno_jacobi_gosa(int nn) {
  float gosa = 0.0f;
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  assert(nn % 2 == 0);
  struct __PSStencil_jacobi_kernel_gosa s0 = {
      __PSStencilMap_jacobi_kernel(innerDom, p0, p1, G[A0], G[A1], G[A2], G[A3],
                                   G[B0], G[B1], G[B2], G[C0], G[C1], G[C2],
                                   G[BND], G[WRK1], omega),
      G[GOSA], __PSGridGetID(G[GOSA])};
  struct __PSStencil_jacobi_kernel_gosa s1 = s0;
  s1.base.p0 = p1;
  s1.base.p1 = p0;
  __PSStencilRun_1(nn / 2, s0, s1);
  __PSReduceGridFloat(&gosa, PS_SUM, G[GOSA]);
  return gosa;
}

/* the original benchmark's structure (himenobmtxpa_original.c:299-346): a residual every
 * iteration -- here every PSStencilRun of the ping-pong pair is followed by its PSReduce */
float himeno_jacobi_gosa_each(int nn) {
  float gosa = 0.0f;
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  assert(nn % 2 == 0);
  struct __PSStencil_jacobi_kernel_gosa s0 = {
      __PSStencilMap_jacobi_kernel(innerDom, p0, p1, G[A0], G[A1], G[A2], G[A3],
                                   G[B0], G[B1], G[B2], G[C0], G[C1], G[C2],
                                   G[BND], G[WRK1], omega),
      G[GOSA], __PSGridGetID(G[GOSA])};
  struct __PSStencil_jacobi_kernel_gosa s1 = s0;
  s1.base.p0 = p1;
  s1.base.p1 = p0;
  int n;
  for (n = 0; n < nn / 2; ++n) {
    __PSStencilRun_1(1, s0, s1);
    __PSReduceGridFloat(&gosa, PS_SUM, G[GOSA]);
  }
  return gosa;
}
