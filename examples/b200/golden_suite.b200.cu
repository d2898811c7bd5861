/*
 * b200-target spelling of the shared golden-suite translations
 * (examples/golden/golden_suite.inc): what a B200RuntimeBuilder emits for the reference's
 * system tests (tests/system_tests/test_cases/test_*.c) -- every kernel as a GENERIC sweep:
 * `__device__` kernel body, `__global__` with the generic iteration shape, launch stub,
 * descriptor, and a run function that calls __PSB200StencilRun.
 * User types are SoA on the device: a member access becomes an access to that member's array.
 */
#define PHYSIS_B200
#include "physis/physis.h"
#include "physis/physis_b200_generic.cuh"

/* by-value device views (layout of Grid::dev_view): dims, then one pointer per member */
struct GV3 { int dim[3]; int slab; void *m[4]; };
struct GV1 { int dim[1]; void *m[1]; };
struct GV2 { int dim[2]; void *m[1]; };
static GV3 MakeGV3(const __PSGrid *g, int nmembers) {
  GV3 v;
  memset(&v, 0, sizeof v);
  memcpy(&v, g->dev, 16 + sizeof(void *) * (size_t)nmembers);
  return v;
}
static GV2 MakeGV2(const __PSGrid *g) {
  GV2 v;
  memcpy(&v, g->dev, sizeof v);
  return v;
}
static GV1 MakeGV1(const __PSGrid *g) {
  GV1 v;
  memcpy(&v, g->dev, sizeof v);
  return v;
}

#define GOLDEN_EXPORT extern "C"
#define GK __device__ static inline
#define KG const GV3 *
#define KG1 const GV1 *
#define KGU const GV3 *
#define OFF3(g, x, y, z) __PSGridGetOffset3DDev(g, x, y, z)
#define OFFP3(g, x, y, z) __PSGridGetOffsetPeriodic3DDev(g, x, y, z)
#define OFF1(g, x) __PSGridGetOffset1DDev(g, x)
#define OFF2(g, x, y) __PSGridGetOffset2DDev(g, x, y)
#define OFFP2(g, x, y) __PSGridGetOffsetPeriodic2DDev(g, x, y)
#define KG2 const GV2 *
#define GET(T, g, off) (((T *)((g)->m[0]))[off])
#define GETM(ST, T, g, m_, mi, ci, off) \
  (((T *)((g)->m[mi]))[(size_t)(ci) * __PSGridMemberStride3DDev(g) + (off)])
#define GRID_NEW(ti, nd, dims) __PSGridNew(ti, nd, dims, NULL)

/* largest |z offset| of a kernel's reads where it exceeds one plane (from the translator's
 * StencilRange): a multi-GPU run checks it against the halo width */
template <typename S> struct ZReachOf { static constexpr int v = 0; };
#define STENCIL_ZREACH(K, R)   \
  struct __PSStencil_##K;      \
  template <> struct ZReachOf<__PSStencil_##K> { static constexpr int v = R; };

#define B200_DESCRIBE(K, NG, ...)                                                               \
  static void __PSStencilDescribe_##K(const struct __PSStencil_##K *s, __PSB200StencilDesc *d) { \
    __PSGrid *gs__[] = {__VA_ARGS__};                                                           \
    memset(d, 0, sizeof(*d));                                                                   \
    d->kind = PSB200_KIND_GENERIC;                                                              \
    d->dom = s->dom;                                                                            \
    d->num_grids = NG;                                                                          \
    for (int i = 0; i < NG; ++i) { d->grids[i] = gs__[i]; d->members[i] = -1; }                 \
    d->stencil = s;                                                                             \
    d->launch = __PSStencilLaunch_##K;                                                          \
    d->name = #K;                                                                               \
    d->z_reach = ZReachOf<__PSStencil_##K>::v;                                                  \
  }

#define DEF_STENCIL_1U(K, NM)                                                                   \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g; int g_index; };                         \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g) {               \
    struct __PSStencil_##K stencil = {dom, g, __PSGridGetID(g)};                                \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g) {                       \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      K(x, y, z, &g);                                                                           \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(*dom, sh.zchunk,             \
                                                                   MakeGV3(s->g, NM));          \
  }                                                                                             \
  B200_DESCRIBE(K, 1, s->g)

#define DEF_STENCIL_2X(K, NM)                                                                   \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2)};       \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g1, GV3 g2) {              \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      K(x, y, z, &g1, &g2);                                                                     \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(                             \
        *dom, sh.zchunk, MakeGV3(s->g1, NM), MakeGV3(s->g2, NM));                               \
  }                                                                                             \
  B200_DESCRIBE(K, 2, s->g1, s->g2)
#define DEF_STENCIL_2(K, ND) DEF_STENCIL_2X(K, 1)
/* two grids and a float scalar (passed to the __global__ by value, as the CUDA target does,
 * translator/cuda_runtime_builder.cc:1615-1641) */
#define DEF_STENCIL_2F(K)                                                                       \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; float c; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2, float c) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), c};    \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g1, GV3 g2, float c) {     \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      K(x, y, z, &g1, &g2, c);                                                                  \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(                             \
        *dom, sh.zchunk, MakeGV3(s->g1, 1), MakeGV3(s->g2, 1), s->c);                           \
  }                                                                                             \
  B200_DESCRIBE(K, 2, s->g1, s->g2)
#define DEF_STENCIL_2U(K, NM) DEF_STENCIL_2X(K, NM)

#define DEF_STENCIL_3G(K, T3, MAKE3)                                                            \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; \
                           __PSGrid *g3; int g3_index; };                                       \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2,  \
                                                   __PSGrid *g3) {                              \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), g3,    \
                                      __PSGridGetID(g3)};                                       \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g1, GV3 g2, T3 g3) {       \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      K(x, y, z, &g1, &g2, &g3);                                                                \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(                             \
        *dom, sh.zchunk, MakeGV3(s->g1, 1), MakeGV3(s->g2, 1), MAKE3);                          \
  }                                                                                             \
  B200_DESCRIBE(K, 3, s->g1, s->g2, s->g3)
#define DEF_STENCIL_3(K, ND) DEF_STENCIL_3G(K, GV3, MakeGV3(s->g3, 1))
#define DEF_STENCIL_2_1D(K) DEF_STENCIL_3G(K, GV1, MakeGV1(s->g3))

#define DEF_STENCIL_1D2(K)                                                                      \
  struct __PSStencil_##K { PSDomain1D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain1D dom, __PSGrid *g1, __PSGrid *g2) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2)};       \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, GV1 g1, GV1 g2) {                          \
    __PSB200_FOREACH_POINT1D_BEGIN(dom, x)                                                      \
      K(x, &g1, &g2);                                                                           \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 1);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(*dom, MakeGV1(s->g1),        \
                                                                   MakeGV1(s->g2));             \
  }                                                                                             \
  B200_DESCRIBE(K, 2, s->g1, s->g2)
#define DEF_STENCIL_2D2(K)                                                                      \
  struct __PSStencil_##K { PSDomain2D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain2D dom, __PSGrid *g1, __PSGrid *g2) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2)};       \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, GV2 g1, GV2 g2) {                          \
    __PSB200_FOREACH_POINT2D_BEGIN(dom, x, y)                                                   \
      K(x, y, &g1, &g2);                                                                        \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 2);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(*dom, MakeGV2(s->g1),        \
                                                                   MakeGV2(s->g2));             \
  }                                                                                             \
  B200_DESCRIBE(K, 2, s->g1, s->g2)
/* red-black: one descriptor per colour, the colour travels in the stencil struct the stub sees */
#define DEF_STENCIL_RB1(K)                                                                      \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g; int g_index; int rb; };                 \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g) {               \
    struct __PSStencil_##K stencil = {dom, g, __PSGridGetID(g), 0};                             \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g, int rb) {               \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      if (__PSB200RedBlackActive(dom, x, y, z, rb)) K(x, y, z, &g);                             \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(*dom, sh.zchunk,             \
                                                                   MakeGV3(s->g, 1), s->rb);    \
  }                                                                                             \
  B200_DESCRIBE(K, 1, s->g)
#define RUN_RB(K, S0)                          \
  do {                                         \
    struct __PSStencil_##K s0__ = S0;          \
    struct __PSStencil_##K s1__ = S0;          \
    s0__.rb = 0;                               \
    s1__.rb = 1;                               \
    __PSB200StencilDesc d__[2];                \
    __PSStencilDescribe_##K(&s0__, &d__[0]);   \
    __PSStencilDescribe_##K(&s1__, &d__[1]);   \
    __PSB200StencilRun(1, 2, d__);             \
  } while (0)
#define DEF_STENCIL_5M(K)                                                                       \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; \
                           __PSGrid *g3; int g3_index; __PSGrid *g4; int g4_index;              \
                           __PSGrid *g5; int g5_index; };                                       \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2,  \
                                                   __PSGrid *g3, __PSGrid *g4, __PSGrid *g5) {  \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), g3,    \
                                      __PSGridGetID(g3), g4, __PSGridGetID(g4), g5,             \
                                      __PSGridGetID(g5)};                                       \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g1, GV3 g2, GV1 g3, GV1 g4, \
                                     GV1 g5) {                                                  \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      K(x, y, z, &g1, &g2, &g3, &g4, &g5);                                                      \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(                             \
        *dom, sh.zchunk, MakeGV3(s->g1, 1), MakeGV3(s->g2, 1), MakeGV1(s->g3), MakeGV1(s->g4),  \
        MakeGV1(s->g5));                                                                        \
  }                                                                                             \
  B200_DESCRIBE(K, 5, s->g1, s->g2, s->g3, s->g4, s->g5)
#define DEF_STENCIL_4M(K)                                                                       \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; \
                           __PSGrid *g3; int g3_index; __PSGrid *g4; int g4_index; };           \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2,  \
                                                   __PSGrid *g3, __PSGrid *g4) {                \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), g3,    \
                                      __PSGridGetID(g3), g4, __PSGridGetID(g4)};                \
    return stencil;                                                                             \
  }                                                                                             \
  __global__ void __PSStencilRun_##K(__PSDomain dom, int zchunk, GV3 g1, GV3 g2, GV1 g3, GV2 g4) { \
    __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                          \
      K(x, y, z, &g1, &g2, &g3, &g4);                                                           \
    __PSB200_FOREACH_POINT_END                                                                  \
  }                                                                                             \
  static void __PSStencilLaunch_##K(const void *sv, const __PSDomain *dom, __PSB200Stream st) { \
    const struct __PSStencil_##K *s = (const struct __PSStencil_##K *)sv;                       \
    __PSB200GenericShape sh = __PSB200GenericShapeFor(dom, 3);                                  \
    __PSStencilRun_##K<<<sh.grid, sh.block, 0, (cudaStream_t)st>>>(                             \
        *dom, sh.zchunk, MakeGV3(s->g1, 1), MakeGV3(s->g2, 1), MakeGV1(s->g3), MakeGV2(s->g4)); \
  }                                                                                             \
  B200_DESCRIBE(K, 4, s->g1, s->g2, s->g3, s->g4)

/* the generated __PSStencilRun_<id>(iter, s0, s1, ...) */
#define RUN1(K, S0)                            \
  do {                                         \
    struct __PSStencil_##K s0__ = S0;          \
    __PSB200StencilDesc d__[1];                \
    __PSStencilDescribe_##K(&s0__, &d__[0]);   \
    __PSB200StencilRun(1, 1, d__);             \
  } while (0)
#define RUN2(K, ITER, S0, S1) RUN2K(K, K, ITER, S0, S1)
#define RUN2K(K0, K1, ITER, S0, S1)            \
  do {                                         \
    struct __PSStencil_##K0 s0__ = S0;         \
    struct __PSStencil_##K1 s1__ = S1;         \
    __PSB200StencilDesc d__[2];                \
    __PSStencilDescribe_##K0(&s0__, &d__[0]);  \
    __PSStencilDescribe_##K1(&s1__, &d__[1]);  \
    __PSB200StencilRun(ITER, 2, d__);          \
  } while (0)

#include "../golden/golden_suite.inc"
