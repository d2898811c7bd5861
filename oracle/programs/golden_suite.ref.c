/*
 * ORACLE (test infrastructure only).  REFERENCE-target spelling of the shared golden-suite
 * translations (examples/golden/golden_suite.inc): what `physisc --ref` emits for the
 * reference's system tests, following translator/reference_runtime_builder.cc
 * (map struct :391-441, map function :443-541, z->y->x run loop :605-664, run function
 * :837-893; Get/Emit rewriting :82-100,142-177).
 */
#define PHYSIS_REF
#include <stdlib.h>
#include <string.h>
#include "physis/physis.h"

#define GOLDEN_EXPORT
#define GK static inline
#define STENCIL_ZREACH(K, R)   /* b200-target hint (halo width); nothing on the REF target */
#define KG __PSGrid *
#define KG1 __PSGrid *
#define KGU __PSGrid *
#define OFF3(g, x, y, z) __PSGridGetOffset3D(g, x, y, z)
#define OFFP3(g, x, y, z) __PSGridGetOffsetPeriodic3D(g, x, y, z)
#define OFF1(g, x) __PSGridGetOffset1D(g, x)
#define OFF2(g, x, y) __PSGridGetOffset2D(g, x, y)
#define OFFP2(g, x, y) __PSGridGetOffsetPeriodic2D(g, x, y)
#define KG2 __PSGrid *
#define GET(T, g, off) (((T *)((g)->p))[off])
#define GETM(ST, T, g, m, mi, ci, off) (((ST *)((g)->p))[off].m)
#define GRID_NEW(ti, nd, dims) __PSGridNew(ti, nd, dims)

#define REF_LOOP(CALL)                                                                 \
  int i3;                                                                              \
  for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {             \
    int i2;                                                                            \
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {           \
      int i1;                                                                          \
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) {         \
        CALL;                                                                          \
      }                                                                                \
    }                                                                                  \
  }

#define DEF_STENCIL_1U(K, NM)                                                          \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g; int g_index; };                \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g) {      \
    struct __PSStencil_##K stencil = {dom, g, __PSGridGetID(g)};                       \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    REF_LOOP(K(i1, i2, i3, s->g))                                                      \
  }
#define DEF_STENCIL_2(K, ND)                                                           \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2)}; \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    REF_LOOP(K(i1, i2, i3, s->g1, s->g2))                                              \
  }
#define DEF_STENCIL_2U(K, NM) DEF_STENCIL_2(K, 3)
/* two grids and a float scalar: scalars are copied into the stencil struct after the grids */
#define DEF_STENCIL_2F(K)                                                              \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; float c; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2, float c) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), c}; \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    REF_LOOP(K(i1, i2, i3, s->g1, s->g2, s->c))                                        \
  }
#define DEF_STENCIL_3(K, ND)                                                           \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; \
                           __PSGrid *g3; int g3_index; };                              \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2, \
                                                   __PSGrid *g3) {                     \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), g3, \
                                      __PSGridGetID(g3)};                              \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    REF_LOOP(K(i1, i2, i3, s->g1, s->g2, s->g3))                                       \
  }
#define DEF_STENCIL_2_1D(K) DEF_STENCIL_3(K, 3)

#define DEF_STENCIL_1D2(K)                                                             \
  struct __PSStencil_##K { PSDomain1D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain1D dom, __PSGrid *g1, __PSGrid *g2) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2)}; \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    int i1;                                                                            \
    for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) K(i1, s->g1, s->g2); \
  }
#define DEF_STENCIL_2D2(K)                                                             \
  struct __PSStencil_##K { PSDomain2D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; }; \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain2D dom, __PSGrid *g1, __PSGrid *g2) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2)}; \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    int i2;                                                                            \
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {           \
      int i1;                                                                          \
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) K(i1, i2, s->g1, s->g2); \
    }                                                                                  \
  }
/* red-black variant: extra colour parameter, x starts at min + ((min & 1) ^ ((c + y + z) % 2)),
 * stride 2 (reference_runtime_builder.cc:585-600,634-646) */
#define DEF_STENCIL_RB1(K)                                                             \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g; int g_index; };                \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g) {      \
    struct __PSStencil_##K stencil = {dom, g, __PSGridGetID(g)};                       \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s, int rb) {      \
    int i3;                                                                            \
    for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {           \
      int i2;                                                                          \
      for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {         \
        int i1;                                                                        \
        for (i1 = s->dom.local_min[0] + ((s->dom.local_min[0] & 1) ^ ((rb + i2 + i3) % 2)); \
             i1 <= s->dom.local_max[0] - 1; i1 += 2)                                   \
          K(i1, i2, i3, s->g);                                                         \
      }                                                                                \
    }                                                                                  \
  }
#define RUN_RB(K, S0)                        \
  do {                                       \
    struct __PSStencil_##K s0__ = S0;        \
    __PSStencilRun_##K(&s0__, 0);            \
    __PSStencilRun_##K(&s0__, 1);            \
  } while (0)
#define DEF_STENCIL_5M(K)                                                              \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; \
                           __PSGrid *g3; int g3_index; __PSGrid *g4; int g4_index;     \
                           __PSGrid *g5; int g5_index; };                              \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2, \
                                                   __PSGrid *g3, __PSGrid *g4, __PSGrid *g5) { \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), g3, \
                                      __PSGridGetID(g3), g4, __PSGridGetID(g4), g5,    \
                                      __PSGridGetID(g5)};                              \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    REF_LOOP(K(i1, i2, i3, s->g1, s->g2, s->g3, s->g4, s->g5))                         \
  }
#define DEF_STENCIL_4M(K)                                                              \
  struct __PSStencil_##K { PSDomain3D dom; __PSGrid *g1; int g1_index; __PSGrid *g2; int g2_index; \
                           __PSGrid *g3; int g3_index; __PSGrid *g4; int g4_index; };  \
  static struct __PSStencil_##K __PSStencilMap_##K(PSDomain3D dom, __PSGrid *g1, __PSGrid *g2, \
                                                   __PSGrid *g3, __PSGrid *g4) {       \
    struct __PSStencil_##K stencil = {dom, g1, __PSGridGetID(g1), g2, __PSGridGetID(g2), g3, \
                                      __PSGridGetID(g3), g4, __PSGridGetID(g4)};       \
    return stencil;                                                                    \
  }                                                                                    \
  static void __PSStencilRun_##K(const struct __PSStencil_##K *const s) {              \
    REF_LOOP(K(i1, i2, i3, s->g1, s->g2, s->g3, s->g4))                                \
  }

/* the generated __PSStencilRun_<id>(iter, s0, s1, ...) */
#define RUN1(K, S0)                          \
  do {                                       \
    struct __PSStencil_##K s0__ = S0;        \
    __PSStencilRun_##K(&s0__);               \
  } while (0)
#define RUN2(K, ITER, S0, S1)                \
  do {                                       \
    struct __PSStencil_##K s0__ = S0;        \
    struct __PSStencil_##K s1__ = S1;        \
    for (int i__ = 0; i__ < (ITER); i__++) { \
      __PSStencilRun_##K(&s0__);             \
      __PSStencilRun_##K(&s1__);             \
    }                                        \
  } while (0)
#define RUN2K(K0, K1, ITER, S0, S1)          \
  do {                                       \
    struct __PSStencil_##K0 s0__ = S0;       \
    struct __PSStencil_##K1 s1__ = S1;       \
    for (int i__ = 0; i__ < (ITER); i__++) { \
      __PSStencilRun_##K0(&s0__);            \
      __PSStencilRun_##K1(&s1__);            \
    }                                        \
  } while (0)

#include "../../examples/golden/golden_suite.inc"
