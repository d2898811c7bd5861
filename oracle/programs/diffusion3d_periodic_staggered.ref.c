/*
 * ORACLE program (test infrastructure only).
 *
 * Hand-emitted REFERENCE-target translation of this repo's
 * examples/dsl/diffusion3d_periodic_staggered.c (BASELINE config 5), in the
 * shape of translator/reference_runtime_builder.cc:
 *   user-type Get   : ((struct Cell *)(g->p))[__PSGridGetOffsetPeriodic3D(g,..)].p   (:82-100,220-247)
 *   PSGridEmitUtype : ((struct Cell *)(g->p))[__PSGridGetOffset3D(g,x,y,z)].q = v    (:142-177)
 *   type_info with member_info[] (BuildTypeInfo :959-1065)
 * Same-grid read .p / write .q as in the reference test
 * tests/system_tests/test_cases/test_user-defined-type-7-pt-periodic.c:19-27.
 * PARITY NOTE: the reference has no golden output for this composition; it is
 * pinned only through the real libphysis_rt_ref build (oracle/_ref) and by the
 * single-feature goldens listed in examples/dsl/diffusion3d_periodic_staggered.c.
 */
#define PHYSIS_REF
#include "physis/physis.h"

struct Cell {
  double p;
  double q;
};

#define UGETP(g, m, x, y, z) (((struct Cell *)((g)->p))[__PSGridGetOffsetPeriodic3D((g), (x), (y), (z))].m)
#define KGET(g, x, y, z) (((double *)((g)->p))[__PSGridGetOffset3D((g), (x), (y), (z))])

static inline void step_pq(const int x, const int y, const int z,
                           __PSGrid *u, __PSGrid *kap) {
  double c = UGETP(u, p, x, y, z);
  double w = UGETP(u, p, x-1, y, z);
  double e = UGETP(u, p, x+1, y, z);
  double n = UGETP(u, p, x, y-1, z);
  double s = UGETP(u, p, x, y+1, z);
  double b = UGETP(u, p, x, y, z-1);
  double t = UGETP(u, p, x, y, z+1);
  double k = 0.125 * (KGET(kap, x, y, z) + KGET(kap, x+1, y, z)
                      + KGET(kap, x, y+1, z) + KGET(kap, x, y, z+1)
                      + KGET(kap, x+1, y+1, z) + KGET(kap, x+1, y, z+1)
                      + KGET(kap, x, y+1, z+1) + KGET(kap, x+1, y+1, z+1));
  ((struct Cell *)(u->p))[__PSGridGetOffset3D(u, x, y, z)].q =
      c + k * (w + e + n + s + b + t - 6.0 * c);
}

static inline void step_qp(const int x, const int y, const int z,
                           __PSGrid *u, __PSGrid *kap) {
  double c = UGETP(u, q, x, y, z);
  double w = UGETP(u, q, x-1, y, z);
  double e = UGETP(u, q, x+1, y, z);
  double n = UGETP(u, q, x, y-1, z);
  double s = UGETP(u, q, x, y+1, z);
  double b = UGETP(u, q, x, y, z-1);
  double t = UGETP(u, q, x, y, z+1);
  double k = 0.125 * (KGET(kap, x, y, z) + KGET(kap, x+1, y, z)
                      + KGET(kap, x, y+1, z) + KGET(kap, x, y, z+1)
                      + KGET(kap, x+1, y+1, z) + KGET(kap, x+1, y, z+1)
                      + KGET(kap, x, y+1, z+1) + KGET(kap, x+1, y+1, z+1));
  ((struct Cell *)(u->p))[__PSGridGetOffset3D(u, x, y, z)].p =
      c + k * (w + e + n + s + b + t - 6.0 * c);
}

struct __PSStencil_step_pq { PSDomain3D dom; __PSGrid *u; int u_index; __PSGrid *kap; int kap_index; };
struct __PSStencil_step_qp { PSDomain3D dom; __PSGrid *u; int u_index; __PSGrid *kap; int kap_index; };

static struct __PSStencil_step_pq __PSStencilMap_step_pq(PSDomain3D dom, __PSGrid *u, __PSGrid *kap) {
  struct __PSStencil_step_pq stencil = {dom, u, __PSGridGetID(u), kap, __PSGridGetID(kap)};
  return stencil;
}
static struct __PSStencil_step_qp __PSStencilMap_step_qp(PSDomain3D dom, __PSGrid *u, __PSGrid *kap) {
  struct __PSStencil_step_qp stencil = {dom, u, __PSGridGetID(u), kap, __PSGridGetID(kap)};
  return stencil;
}

static void __PSStencilRun_step_pq(const struct __PSStencil_step_pq *const s) {
  int i3;
  for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {
    int i2;
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {
      int i1;
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) {
        step_pq(i1, i2, i3, s->u, s->kap);
      }
    }
  }
}
static void __PSStencilRun_step_qp(const struct __PSStencil_step_qp *const s) {
  int i3;
  for (i3 = s->dom.local_min[2]; i3 <= s->dom.local_max[2] - 1; i3 += 1) {
    int i2;
    for (i2 = s->dom.local_min[1]; i2 <= s->dom.local_max[1] - 1; i2 += 1) {
      int i1;
      for (i1 = s->dom.local_min[0]; i1 <= s->dom.local_max[0] - 1; i1 += 1) {
        step_qp(i1, i2, i3, s->u, s->kap);
      }
    }
  }
}

static float __PSStencilRun_0(int iter, struct __PSStencil_step_pq s0, struct __PSStencil_step_qp s1) {
  int i;
  for (i = 0; i < iter; i++) {
    __PSStencilRun_step_pq(&s0);
    __PSStencilRun_step_qp(&s1);
  }
  return 0.0f;
}

static __PSGrid *u;
static __PSGrid *kap;

void pstag_init(int argc, char **argv, int nx, int ny, int nz) {
  PSInit(&argc, &argv, 3, nx + 1, ny + 1, nz + 1);
  {
    PSVectorInt dims = {nx, ny, nz};
    __PSGridTypeMemberInfo member_info[2];
    member_info[0].type = PS_DOUBLE;
    member_info[0].size = sizeof(double);
    member_info[0].rank = 0;
    member_info[1].type = PS_DOUBLE;
    member_info[1].size = sizeof(double);
    member_info[1].rank = 0;
    __PSGridTypeInfo type_info = {PS_USER, sizeof(struct Cell), 2, member_info};
    u = __PSGridNew(&type_info, 3, dims);
  }
  {
    PSVectorInt dims = {nx + 1, ny + 1, nz + 1};
    __PSGridTypeInfo type_info = {PS_DOUBLE, sizeof(double), 0, NULL};
    kap = __PSGridNew(&type_info, 3, dims);
  }
}

void pstag_run(int count, struct Cell *u_host, const double *kap_host,
               int nx, int ny, int nz) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  PSGridCopyin(u, u_host);
  PSGridCopyin(kap, kap_host);
  __PSStencilRun_0(count / 2, __PSStencilMap_step_pq(dom, u, kap),
                   __PSStencilMap_step_qp(dom, u, kap));
  PSGridCopyout(u, u_host);
}

void pstag_finalize(void) {
  PSGridFree(u);
  PSGridFree(kap);
  PSFinalize();
}
