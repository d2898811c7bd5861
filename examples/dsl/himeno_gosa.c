/*
 * Physis DSL source for BASELINE.json config 3 ("Himeno ... with PSReduce
 * residual").  Identical to /root/reference/examples/himeno/
 * himenobmtxpa_physis.c:331-393 except that the kernel also emits ss*ss into
 * an extra grid and jacobi() reduces it — the DSL form of the original
 * benchmark's `gosa += ss*ss` (himenobmtxpa_original.c:334) that the Physis
 * version left as a comment (himenobmtxpa_physis.c:387-391).  Only the parts
 * that differ are shown; grid setup is the reference program's.
 */
#include "physis/physis.h"

void jacobi_kernel_gosa(int i, int j, int k,
                        PSGrid3DFloat p0, PSGrid3DFloat p1,
                        PSGrid3DFloat a0, PSGrid3DFloat a1, PSGrid3DFloat a2,
                        PSGrid3DFloat a3, PSGrid3DFloat b0, PSGrid3DFloat b1,
                        PSGrid3DFloat b2, PSGrid3DFloat c0, PSGrid3DFloat c1,
                        PSGrid3DFloat c2, PSGrid3DFloat bnd, PSGrid3DFloat wrk1,
                        PSGrid3DFloat gosa_g, float omega) {
  float s0, ss;
  s0 = PSGridGet(a0, i, j, k) * PSGridGet(p0, i, j, k+1)
      + PSGridGet(a1, i, j, k) * PSGridGet(p0, i, j+1, k)
      + PSGridGet(a2, i, j, k) * PSGridGet(p0, i+1, j, k)
      + PSGridGet(b0, i, j, k)
      * ( PSGridGet(p0, i, j+1, k+1) - PSGridGet(p0, i, j-1, k+1)
          - PSGridGet(p0, i, j+1, k-1) + PSGridGet(p0, i, j-1, k-1) )
      + PSGridGet(b1, i, j, k)
      * ( PSGridGet(p0, i+1, j+1, k) - PSGridGet(p0, i+1, j-1, k)
          - PSGridGet(p0, i-1, j+1, k) + PSGridGet(p0, i-1, j-1, k) )
      + PSGridGet(b2, i, j, k)
      * ( PSGridGet(p0, i+1, j, k+1) - PSGridGet(p0, i+1, j, k-1)
          - PSGridGet(p0, i-1, j, k+1) + PSGridGet(p0, i-1, j, k-1) )
      + PSGridGet(c0, i, j, k) * PSGridGet(p0, i, j, k-1)
      + PSGridGet(c1, i, j, k) * PSGridGet(p0, i, j-1, k)
      + PSGridGet(c2, i, j, k) * PSGridGet(p0, i-1, j, k)
      + PSGridGet(wrk1, i, j, k);
  ss = (s0 * PSGridGet(a3, i, j, k) - PSGridGet(p0, i, j, k))
      * PSGridGet(bnd, i, j, k);
  float v = PSGridGet(p0, i, j, k) + omega * ss;
  PSGridEmit(p1, v);
  PSGridEmit(gosa_g, ss * ss);
}

float jacobi_gosa(int nn, PSGrid3DFloat a0, PSGrid3DFloat a1, PSGrid3DFloat a2,
                  PSGrid3DFloat a3, PSGrid3DFloat b0, PSGrid3DFloat b1,
                  PSGrid3DFloat b2, PSGrid3DFloat c0, PSGrid3DFloat c1,
                  PSGrid3DFloat c2, PSGrid3DFloat p0, PSGrid3DFloat p1,
                  PSGrid3DFloat bnd, PSGrid3DFloat wrk1, PSGrid3DFloat gosa_g,
                  float omega) {
  float gosa = 0.0f;
  PSDomain3D innerDom = PSDomain3DNew(1, PSGridDim(p0, 0)-1,
                                      1, PSGridDim(p0, 1)-1,
                                      1, PSGridDim(p0, 2)-1);
  PSStencilRun(PSStencilMap(jacobi_kernel_gosa, innerDom,
                            p0, p1, a0, a1, a2, a3, b0, b1, b2,
                            c0, c1, c2, bnd, wrk1, gosa_g, omega),
               PSStencilMap(jacobi_kernel_gosa, innerDom,
                            p1, p0, a0, a1, a2, a3, b0, b1, b2,
                            c0, c1, c2, bnd, wrk1, gosa_g, omega),
               nn/2);
  PSReduce(&gosa, PS_SUM, gosa_g);
  return gosa;
}
