/*
 * ORACLE (test infrastructure only).
 *
 * Restatement of the diffusion benchmark's problem setup:
 *   initial field  : examples/diffusion-benchmark/diffusion3d.cc:14-39 (Initialize)
 *   coefficients   : examples/diffusion-benchmark/diffusion3d.h:54-66 (Diffusion3D ctor)
 * The reference is C++ with REAL=float and <math.h>: exp()/cos() on float
 * arguments therefore bind to the float overloads (expf/cosf), products with
 * the double literals 1.0 / 0.125-cast / 0.1 are evaluated in double and
 * rounded to float on assignment.  Spelled out explicitly here because this
 * file is C.  Pinned against the reference's own Initialize()/ctor through
 * oracle/_ref (tests/test_oracle.py::test_setup_matches_reference).
 */
#include <math.h>

#define REAL float
#ifndef M_PI
#define M_PI (3.1415926535897932384626)
#endif

/* out[0..6] = ce, cw, cn, cs, ct, cb, cc; out[7..9] = dx,dy,dz; out[10] = dt;
 * out[11] = kappa; out[12..14] = kx,ky,kz */
void oracle_diffusion3d_params(int nx, int ny, int nz, REAL *out) {
  REAL kappa = 0.1;
  REAL l = 1.0;
  REAL dx = l / nx;
  REAL dy = l / ny;
  REAL dz = l / nz;
  REAL kx, ky, kz;
  kx = ky = kz = 2.0 * M_PI;
  REAL dt = 0.1 * dx * dx / kappa;          /* double product, rounded once */
  REAL ce, cw, cn, cs, ct, cb, cc;
  ce = cw = kappa * dt / (dx * dx);         /* all-float arithmetic */
  cn = cs = kappa * dt / (dy * dy);
  ct = cb = kappa * dt / (dz * dz);
  cc = 1.0 - (ce + cw + cn + cs + ct + cb); /* float sum, double subtract */
  out[0] = ce; out[1] = cw; out[2] = cn; out[3] = cs; out[4] = ct; out[5] = cb; out[6] = cc;
  out[7] = dx; out[8] = dy; out[9] = dz; out[10] = dt; out[11] = kappa;
  out[12] = kx; out[13] = ky; out[14] = kz;
}

void oracle_diffusion3d_initialize(REAL *buff, const int nx, const int ny, const int nz,
                                   const REAL kx, const REAL ky, const REAL kz,
                                   const REAL dx, const REAL dy, const REAL dz,
                                   const REAL kappa, const REAL time) {
  REAL ax = expf(-kappa * time * (kx * kx));
  REAL ay = expf(-kappa * time * (ky * ky));
  REAL az = expf(-kappa * time * (kz * kz));
  for (int jz = 0; jz < nz; jz++) {
    for (int jy = 0; jy < ny; jy++) {
      for (int jx = 0; jx < nx; jx++) {
        long j = (long)jz * nx * ny + (long)jy * nx + jx;
        REAL x = dx * ((REAL)(jx + 0.5));
        REAL y = dy * ((REAL)(jy + 0.5));
        REAL z = dz * ((REAL)(jz + 0.5));
        REAL f0 = (REAL)0.125
            * (1.0 - ax * cosf(kx * x))
            * (1.0 - ay * cosf(ky * y))
            * (1.0 - az * cosf(kz * z));
        buff[j] = f0;
      }
    }
  }
}

/* RMS error against the analytic solution, as Baseline::GetAccuracy
 * (examples/diffusion-benchmark/baseline.cc:51-60): float accumulation. */
REAL oracle_diffusion3d_accuracy(const REAL *f, const REAL *ref, long len) {
  REAL err = 0.0;
  for (long i = 0; i < len; i++) {
    REAL diff = ref[i] - f[i];
    err += diff * diff;
  }
  return (REAL)sqrt(err / len);
}
