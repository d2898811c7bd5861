/*
 * physis_b200_generic.cuh — launch shape of GENERIC sweeps on the b200 target.
 *
 * A `physisc --b200` translation compiles every user kernel body into the
 * program as a `__device__` function (exactly as `--cuda` does,
 * translator/cuda_runtime_builder.cc:152-186,1099-1141) and wraps it in a
 * `__global__` whose iteration shape comes from here instead of the
 * reference's fixed "one thread per (x,y), loop over all z, block 64x4"
 * (cuda_runtime_builder.cc:12-14,1259-1281,1435-1463).  Sweeps recognised as
 * one of the hand-written families never get here; this path exists so that
 * ANY Physis kernel runs on the GPU (there is no CPU fallback).
 *
 * Shape: block = 32 x 8 threads over (x, y); the z extent is split into chunks
 * across blockIdx.z so small xy domains still fill 148 SMs; consecutive lanes
 * touch consecutive x (coalesced); each thread marches its z chunk so reuse in
 * z comes from L1/L2.  Compile the program with `--fmad=false` so per-point
 * arithmetic rounds like the REFERENCE target's (no FMA contraction).
 */
#ifndef PHYSIS_PHYSIS_B200_GENERIC_CUH_
#define PHYSIS_PHYSIS_B200_GENERIC_CUH_

#include <cuda_runtime.h>
#include "physis/physis_b200.h"

#define __PSB200_GENERIC_BX 32
#define __PSB200_GENERIC_BY 8

struct __PSB200GenericShape {
  dim3 grid;
  dim3 block;
  int zchunk;
};

static inline __PSB200GenericShape __PSB200GenericShapeFor(const __PSDomain *dom, int num_dims) {
  __PSB200GenericShape s;
  int ex = dom->local_max[0] - dom->local_min[0];
  int ey = num_dims > 1 ? dom->local_max[1] - dom->local_min[1] : 1;
  int ez = num_dims > 2 ? dom->local_max[2] - dom->local_min[2] : 1;
  if (ex < 1) ex = 1;
  if (ey < 1) ey = 1;
  if (ez < 1) ez = 1;
  s.block = dim3(__PSB200_GENERIC_BX, num_dims > 1 ? __PSB200_GENERIC_BY : 1, 1);
  unsigned gx = (ex + s.block.x - 1) / s.block.x, gy = (ey + s.block.y - 1) / s.block.y;
  /* enough z chunks for ~8 blocks per SM, at least 4 planes per chunk */
  int want = (int)((148u * 8u + gx * gy - 1) / (gx * gy));
  int maxc = (ez + 3) / 4;
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  s.zchunk = (ez + want - 1) / want;
  s.grid = dim3(gx, gy, (ez + s.zchunk - 1) / s.zchunk);
  return s;
}

/* Body macro for the generated __global__: declares x, y and loops z over this
 * block's chunk, skipping points outside the domain.  `stride`/`xoff` express
 * the red-black variants (translator/reference_runtime_builder.cc:585-600):
 * plain sweeps pass (1, 0). */
#define __PSB200_FOREACH_POINT_BEGIN(dom, zchunk, x, y, z)                                   \
  {                                                                                          \
    const int x = (dom).local_min[0] + (int)(blockIdx.x * blockDim.x + threadIdx.x);         \
    const int y = (dom).local_min[1] + (int)(blockIdx.y * blockDim.y + threadIdx.y);         \
    const int __z0 = (dom).local_min[2] + (int)blockIdx.z * (zchunk);                        \
    const int __z1 = min(__z0 + (zchunk), (int)(dom).local_max[2]);                          \
    if (x < (dom).local_max[0] && y < (dom).local_max[1]) {                                  \
      for (int z = __z0; z < __z1; ++z) {
#define __PSB200_FOREACH_POINT_END \
      }                            \
    }                              \
  }

/* 2-D and 1-D domains (grid z extent is 1) */
#define __PSB200_FOREACH_POINT2D_BEGIN(dom, x, y)                                            \
  {                                                                                          \
    const int x = (dom).local_min[0] + (int)(blockIdx.x * blockDim.x + threadIdx.x);         \
    const int y = (dom).local_min[1] + (int)(blockIdx.y * blockDim.y + threadIdx.y);         \
    if (x < (dom).local_max[0] && y < (dom).local_max[1]) {                                  \
      {
#define __PSB200_FOREACH_POINT1D_BEGIN(dom, x)                                               \
  {                                                                                          \
    const int x = (dom).local_min[0] + (int)(blockIdx.x * blockDim.x + threadIdx.x);         \
    if (x < (dom).local_max[0]) {                                                            \
      {

/* Red/black colouring: a point is visited when ((x + y + z + color) & 1) == 0
 * in the reference's sense: x starts at min + ((min&1) ^ ((y+z+color)%2)) and
 * advances by 2. */
__device__ static inline bool __PSB200RedBlackActive(const __PSDomain &dom, int x, int y, int z,
                                                     int color) {
  const int start = dom.local_min[0] + ((dom.local_min[0] & 1) ^ ((y + z + color) % 2));
  return x >= start && (((x - start) & 1) == 0);
}

#endif /* PHYSIS_PHYSIS_B200_GENERIC_CUH_ */
