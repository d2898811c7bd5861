// ORACLE build shim (test infrastructure only): the reference's diffusion3d_openmp.cc includes its
// header under a "benchmarks/diffusion3d/" prefix that does not exist in its tree; forward to the
// header where it lies (found through -I<reference>/examples/diffusion-benchmark).
#include <diffusion3d_openmp.h>
