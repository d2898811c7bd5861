/*
 * Physis DSL source for BASELINE.json config 5:
 *   "fp64 periodic-boundary staggered-grid diffusion with user-defined point
 *    type, weak scaling 512^3/GPU".
 * The reference ships no such program; this one composes exactly the features
 * its own tests exercise one at a time:
 *   user type {p,q}, read .p / emit .q of the SAME grid, periodic 7-pt
 *       tests/system_tests/test_cases/test_user-defined-type-7-pt-periodic.c:12-27
 *   fp64 7-pt           tests/system_tests/test_cases/test_7-pt-double-type.c:17-25
 *   staggered (N+1)^3 grid read at (x..x+1, y..y+1, z..z+1) from an N^3 domain
 *       examples/test_staggered_grid.c:6-14,24-27
 * u   : cell-centred state, N^3 of struct Cell (ping-pong between members).
 * kap : vertex-centred diffusion number, (N+1)^3 doubles; a cell uses the mean
 *       of its 8 corner vertices.
 */
#include "physis/physis.h"

struct Cell {
  double p;
  double q;
};
DeclareGrid3D(Cell, struct Cell);

static void step_pq(const int x, const int y, const int z,
                    PSGrid3DCell u, PSGrid3DDouble kap) {
  double c = PSGridGetPeriodic(u, x, y, z).p;
  double w = PSGridGetPeriodic(u, x-1, y, z).p;
  double e = PSGridGetPeriodic(u, x+1, y, z).p;
  double n = PSGridGetPeriodic(u, x, y-1, z).p;
  double s = PSGridGetPeriodic(u, x, y+1, z).p;
  double b = PSGridGetPeriodic(u, x, y, z-1).p;
  double t = PSGridGetPeriodic(u, x, y, z+1).p;
  double k = 0.125 * (PSGridGet(kap, x, y, z) + PSGridGet(kap, x+1, y, z)
                      + PSGridGet(kap, x, y+1, z) + PSGridGet(kap, x, y, z+1)
                      + PSGridGet(kap, x+1, y+1, z) + PSGridGet(kap, x+1, y, z+1)
                      + PSGridGet(kap, x, y+1, z+1) + PSGridGet(kap, x+1, y+1, z+1));
  PSGridEmitUtype(u.q, c + k * (w + e + n + s + b + t - 6.0 * c));
}

static void step_qp(const int x, const int y, const int z,
                    PSGrid3DCell u, PSGrid3DDouble kap) {
  double c = PSGridGetPeriodic(u, x, y, z).q;
  double w = PSGridGetPeriodic(u, x-1, y, z).q;
  double e = PSGridGetPeriodic(u, x+1, y, z).q;
  double n = PSGridGetPeriodic(u, x, y-1, z).q;
  double s = PSGridGetPeriodic(u, x, y+1, z).q;
  double b = PSGridGetPeriodic(u, x, y, z-1).q;
  double t = PSGridGetPeriodic(u, x, y, z+1).q;
  double k = 0.125 * (PSGridGet(kap, x, y, z) + PSGridGet(kap, x+1, y, z)
                      + PSGridGet(kap, x, y+1, z) + PSGridGet(kap, x, y, z+1)
                      + PSGridGet(kap, x+1, y+1, z) + PSGridGet(kap, x+1, y, z+1)
                      + PSGridGet(kap, x, y+1, z+1) + PSGridGet(kap, x+1, y+1, z+1));
  PSGridEmitUtype(u.p, c + k * (w + e + n + s + b + t - 6.0 * c));
}

static PSGrid3DCell u;
static PSGrid3DDouble kap;

void pstag_init(int argc, char **argv, int nx, int ny, int nz) {
  PSInit(&argc, &argv, 3, nx + 1, ny + 1, nz + 1);
  u = PSGrid3DCellNew(nx, ny, nz);
  kap = PSGrid3DDoubleNew(nx + 1, ny + 1, nz + 1);
}

void pstag_run(int count, struct Cell *u_host, const double *kap_host,
               int nx, int ny, int nz) {
  PSDomain3D dom = PSDomain3DNew(0, nx, 0, ny, 0, nz);
  PSGridCopyin(u, u_host);
  PSGridCopyin(kap, kap_host);
  PSStencilRun(PSStencilMap(step_pq, dom, u, kap),
               PSStencilMap(step_qp, dom, u, kap),
               count / 2);
  PSGridCopyout(u, u_host);
}

void pstag_finalize(void) {
  PSGridFree(u);
  PSGridFree(kap);
  PSFinalize();
}
