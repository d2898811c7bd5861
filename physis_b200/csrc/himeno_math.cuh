// Arithmetic of one Himeno 19-point Jacobi update, shared by the single-sweep kernel
// (himeno.cu) and the fused two-sweep kernel (himeno_pair.cu): examples/himeno/
// himenobmtxpa_physis.c:331-361 with every operation separately rounded in C's left-to-right
// order (no FMA), i.e. what the REFERENCE target computes.
#pragma once
#include "sweep_common.cuh"

namespace physis_b200 {
namespace sweep {

// One output point.  pXYZ naming: m = -1, c = 0, p = +1 for (x, y, z).
__device__ __forceinline__ float HimenoJacobi(float a0, float a1, float a2, float a3, float b0,
                                        float b1, float b2, float c0, float c1, float c2,
                                        float bnd, float wrk1, float omega,
                                        float ccc, float ccp, float cpc, float pcc, float cpp,
                                        float cmp, float cpm, float cmm, float ppc, float pmc,
                                        float mpc, float mmc, float pcp, float pcm, float mcp,
                                        float mcm, float ccm, float cmc, float mcc, float *ss_out) {
  float s0 = MulRn(a0, ccp);
  s0 = AddRn(s0, MulRn(a1, cpc));
  s0 = AddRn(s0, MulRn(a2, pcc));
  s0 = AddRn(s0, MulRn(b0, AddRn(SubRn(SubRn(cpp, cmp), cpm), cmm)));
  s0 = AddRn(s0, MulRn(b1, AddRn(SubRn(SubRn(ppc, pmc), mpc), mmc)));
  s0 = AddRn(s0, MulRn(b2, AddRn(SubRn(SubRn(pcp, pcm), mcp), mcm)));
  s0 = AddRn(s0, MulRn(c0, ccm));
  s0 = AddRn(s0, MulRn(c1, cmc));
  s0 = AddRn(s0, MulRn(c2, mcc));
  s0 = AddRn(s0, wrk1);
  const float ss = MulRn(SubRn(MulRn(s0, a3), ccc), bnd);
  *ss_out = ss;
  return AddRn(ccc, MulRn(omega, ss));
}


}  // namespace sweep
}  // namespace physis_b200
