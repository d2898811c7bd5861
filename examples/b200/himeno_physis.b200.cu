/*
 * Hand-emitted `b200`-target translation of
 *   /root/reference/examples/himeno/himenobmtxpa_physis.c   (jacobi_kernel, jacobi)
 * and of this repo's examples/dsl/himeno_gosa.c (residual variant), library-ised
 * with the same entry points as oracle/programs/himeno_physis.ref.c so the parity
 * tests drive both sides identically.  Shape: see diffusion3d_physis.b200.cu.
 */
#define PHYSIS_B200
#include "physis/physis.h"
#include "physis/physis_b200_generic.cuh"

enum { P0, P1, BND, WRK1, A0, A1, A2, A3, B0, B1, B2, C0, C1, C2, GOSA, NGRIDS };
static __PSGrid *G[NGRIDS];
static float omega = 0.8;

static __PSGrid *new_float3d(int nx, int ny, int nz) {
  PSVectorInt dims = {nx, ny, nz};
  __PSGridTypeInfo type_info = {PS_FLOAT, sizeof(float), 0, NULL};
  return __PSGridNew(&type_info, 3, dims, NULL);
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
  __PSGridCopyin(mat, buf, NULL);
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
  __PSGridCopyin(Mat, buf, NULL);
}

#define GETD(g, i, j, k) ((g)->p[__PSGridGetOffset3DDev((g), (i), (j), (k))])

template <bool GOSA_EMIT>
__device__ static inline void jacobi_kernel(int i, int j, int k,
                                            __PSGrid3DFloat_dev *p0, __PSGrid3DFloat_dev *p1,
                                            __PSGrid3DFloat_dev *a0, __PSGrid3DFloat_dev *a1,
                                            __PSGrid3DFloat_dev *a2, __PSGrid3DFloat_dev *a3,
                                            __PSGrid3DFloat_dev *b0, __PSGrid3DFloat_dev *b1,
                                            __PSGrid3DFloat_dev *b2, __PSGrid3DFloat_dev *c0,
                                            __PSGrid3DFloat_dev *c1, __PSGrid3DFloat_dev *c2,
                                            __PSGrid3DFloat_dev *bnd, __PSGrid3DFloat_dev *wrk1,
                                            __PSGrid3DFloat_dev *gosa_g, float omega) {
  float s0, ss;
  s0 = GETD(a0, i, j, k) * GETD(p0, i, j, k+1)
      + GETD(a1, i, j, k) * GETD(p0, i, j+1, k)
      + GETD(a2, i, j, k) * GETD(p0, i+1, j, k)
      + GETD(b0, i, j, k)
      * ( GETD(p0, i, j+1, k+1) - GETD(p0, i, j-1, k+1)
          - GETD(p0, i, j+1, k-1) + GETD(p0, i, j-1, k-1) )
      + GETD(b1, i, j, k)
      * ( GETD(p0, i+1, j+1, k) - GETD(p0, i+1, j-1, k)
          - GETD(p0, i-1, j+1, k) + GETD(p0, i-1, j-1, k) )
      + GETD(b2, i, j, k)
      * ( GETD(p0, i+1, j, k+1) - GETD(p0, i+1, j, k-1)
          - GETD(p0, i-1, j, k+1) + GETD(p0, i-1, j, k-1) )
      + GETD(c0, i, j, k) * GETD(p0, i, j, k-1)
      + GETD(c1, i, j, k) * GETD(p0, i, j-1, k)
      + GETD(c2, i, j, k) * GETD(p0, i-1, j, k)
      + GETD(wrk1, i, j, k);
  ss = (s0 * GETD(a3, i, j, k) - GETD(p0, i, j, k))
      * GETD(bnd, i, j, k);
  float v = GETD(p0, i, j, k) + omega * ss;
  GETD(p1, i, j, k) = v;
  if (GOSA_EMIT) GETD(gosa_g, i, j, k) = ss * ss;
  return;
}

struct __PSStencil_jacobi_kernel {
  PSDomain3D dom;
  __PSGrid *g[15];   /* p0,p1,a0,a1,a2,a3,b0,b1,b2,c0,c1,c2,bnd,wrk1[,gosa_g] */
  int g_index[15];
  float omega;
  int with_gosa;
};

static struct __PSStencil_jacobi_kernel __PSStencilMap_jacobi_kernel(
    PSDomain3D dom, __PSGrid *p0, __PSGrid *p1, __PSGrid *gosa_g, float omega) {
  struct __PSStencil_jacobi_kernel s;
  memset(&s, 0, sizeof(s));
  s.dom = dom;
  __PSGrid *order[15] = {p0, p1, G[A0], G[A1], G[A2], G[A3], G[B0], G[B1], G[B2],
                         G[C0], G[C1], G[C2], G[BND], G[WRK1], gosa_g};
  for (int i = 0; i < 15; ++i) {
    s.g[i] = order[i];
    s.g_index[i] = order[i] ? __PSGridGetID(order[i]) : 0;
  }
  s.omega = omega;
  s.with_gosa = gosa_g != NULL;
  return s;
}

struct __PSJacobiDevArgs { __PSGrid3DFloat_dev g[15]; };

template <bool GOSA_EMIT>
__global__ void __PSStencilRun_jacobi_kernel(__PSDomain dom, int zchunk,
                                             __PSJacobiDevArgs v, float omega) {
  __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)
    jacobi_kernel<GOSA_EMIT>(x, y, z, &v.g[0], &v.g[1], &v.g[2], &v.g[3], &v.g[4], &v.g[5],
                             &v.g[6], &v.g[7], &v.g[8], &v.g[9], &v.g[10], &v.g[11], &v.g[12],
                             &v.g[13], &v.g[14], omega);
  __PSB200_FOREACH_POINT_END
}

static void __PSStencilLaunch_jacobi_kernel(const void *sv, const __PSDomain *dom,
        __PSB200Stream stream) {
  const struct __PSStencil_jacobi_kernel *s = (const struct __PSStencil_jacobi_kernel *)sv;
  __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);
  __PSJacobiDevArgs v;
  for (int i = 0; i < 15; ++i)
    v.g[i] = *((__PSGrid3DFloat_dev *)((s->g[i] ? s->g[i] : s->g[0])->dev));
  if (s->with_gosa)
    __PSStencilRun_jacobi_kernel<true><<<sh.grid, sh.block, 0, (cudaStream_t)stream>>>(
        *dom, sh.zchunk, v, s->omega);
  else
    __PSStencilRun_jacobi_kernel<false><<<sh.grid, sh.block, 0, (cudaStream_t)stream>>>(
        *dom, sh.zchunk, v, s->omega);
}

static void __PSStencilDescribe_jacobi_kernel(const struct __PSStencil_jacobi_kernel *s,
                                              __PSB200StencilDesc *d, int force_generic) {
  memset(d, 0, sizeof(*d));
  d->kind = force_generic ? PSB200_KIND_GENERIC
                          : (s->with_gosa ? PSB200_KIND_HIMENO19_GOSA : PSB200_KIND_HIMENO19);
  d->elm_type = PS_FLOAT;
  d->dom = s->dom;
  d->num_grids = s->with_gosa ? 15 : 14;
  for (int i = 0; i < d->num_grids; ++i) {
    d->grids[i] = s->g[i];
    d->members[i] = -1;
  }
  d->num_scalars = 1;
  d->scalars[0] = s->omega;
  d->stencil = s;
  d->launch = __PSStencilLaunch_jacobi_kernel;
  d->name = s->with_gosa ? "jacobi_kernel_gosa" : "jacobi_kernel";
}

static float __PSStencilRun_0(int iter, struct __PSStencil_jacobi_kernel s0,
                              struct __PSStencil_jacobi_kernel s1, int force_generic) {
  __PSB200StencilDesc d[2];
  __PSStencilDescribe_jacobi_kernel(&s0, &d[0], force_generic);
  __PSStencilDescribe_jacobi_kernel(&s1, &d[1], force_generic);
  return __PSB200StencilRun(iter, 2, d);
}

static int g_force_generic = 0;

extern "C" {

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

/* bench hook for runs that scale: the same initial state, but every rank generates and
 * uploads only its own z-slab (the program above builds the whole array on every rank) */
void himeno_init_local(int mimax, int mjmax, int mkmax) {
  int argc = 0;
  char **argv = NULL;
  PSInit(&argc, &argv, 3, mimax, mjmax, mkmax);
  for (int g = 0; g < NGRIDS; ++g) G[g] = new_float3d(mimax, mjmax, mkmax);
  int z_off = 0, z_len = 0;
  __PSB200GridLocalSize(G[P0], &z_off, &z_len);
  const size_t plane = (size_t)mimax * mjmax;
  float *buf = (float *)malloc(plane * (size_t)z_len * sizeof(float));
  for (int k = 0; k < z_len; ++k) {
    const int kg = z_off + k;
    const float v = (float)(kg * kg) / ((mkmax - 1) * (mkmax - 1));
    for (size_t i = 0; i < plane; ++i) buf[(size_t)k * plane + i] = v;
  }
  __PSB200GridCopyinLocal(G[P0], buf);
  __PSB200GridCopyinLocal(G[P1], buf);
  const struct { int g; float v; } consts[] = {{BND, 1.0f}, {A0, 1.0f}, {A1, 1.0f}, {A2, 1.0f},
      {A3, (float)(1.0 / 6.0)}, {B0, 0.0f}, {B1, 0.0f}, {B2, 0.0f}, {C0, 1.0f}, {C1, 1.0f}, {C2, 1.0f}};
  for (size_t c = 0; c < sizeof(consts) / sizeof(consts[0]); ++c) {
    for (size_t i = 0; i < plane * (size_t)z_len; ++i) buf[i] = consts[c].v;
    __PSB200GridCopyinLocal(G[consts[c].g], buf);
  }
  free(buf);
}

void himeno_set_grid(int which, const float *buf) { __PSGridCopyin(G[which], buf, NULL); }
void himeno_get_grid(int which, float *buf) { __PSGridCopyout(G[which], buf, NULL); }
void himeno_set_omega(float w) { omega = w; }
void himeno_force_generic(int on) { g_force_generic = on; }

void himeno_finalize(void) {
  for (int g = 0; g < NGRIDS; ++g) __PSGridFree(G[g], NULL);
  PSFinalize();
}

float himeno_jacobi(int nn) {
  float gosa = 0.0f;
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  assert(nn % 2 == 0);
  __PSStencilRun_0(nn / 2, __PSStencilMap_jacobi_kernel(innerDom, p0, p1, NULL, omega),
                   __PSStencilMap_jacobi_kernel(innerDom, p1, p0, NULL, omega), g_force_generic);
  return gosa;
}

float himeno_jacobi_gosa(int nn) {
  float gosa = 0.0f;
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  assert(nn % 2 == 0);
  __PSStencilRun_0(nn / 2, __PSStencilMap_jacobi_kernel(innerDom, p0, p1, G[GOSA], omega),
                   __PSStencilMap_jacobi_kernel(innerDom, p1, p0, G[GOSA], omega),
                   g_force_generic);
  __PSReduceGridFloat(&gosa, PS_SUM, G[GOSA]);
  return gosa;
}

/* the original benchmark's structure (himenobmtxpa_original.c:299-346): a residual every
 * iteration -- every PSStencilRun of the ping-pong pair is followed by its PSReduce */
float himeno_jacobi_gosa_each(int nn) {
  float gosa = 0.0f;
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  assert(nn % 2 == 0);
  for (int n = 0; n < nn / 2; ++n) {
    __PSStencilRun_0(1, __PSStencilMap_jacobi_kernel(innerDom, p0, p1, G[GOSA], omega),
                     __PSStencilMap_jacobi_kernel(innerDom, p1, p0, G[GOSA], omega),
                     g_force_generic);
    __PSReduceGridFloat(&gosa, PS_SUM, G[GOSA]);
  }
  return gosa;
}

/* bench hook: sweeps only (no reduction), returns nothing */
void himeno_sweeps_only(int nn, int with_gosa) {
  __PSGrid *p0 = G[P0], *p1 = G[P1];
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0) - 1,
                                      1, PSGridDim(p0, 1) - 1,
                                      1, PSGridDim(p0, 2) - 1);
  __PSGrid *gg = with_gosa ? G[GOSA] : NULL;
  __PSStencilRun_0(nn / 2, __PSStencilMap_jacobi_kernel(innerDom, p0, p1, gg, omega),
                   __PSStencilMap_jacobi_kernel(innerDom, p1, p0, gg, omega), g_force_generic);
}
float himeno_reduce_gosa(void) {
  float gosa = 0.0f;
  __PSReduceGridFloat(&gosa, PS_SUM, G[GOSA]);
  return gosa;
}

}  // extern "C"
