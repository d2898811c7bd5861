// Does a 1-D tiled TMA load accept an arbitrary (odd) start element of an fp64 array?
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, double *out) {
  __shared__ __align__(128) double buf[80];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(66 * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                 ::"r"(d), "l"((uint64_t)&m), "r"(b), "r"(c0) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(b) : "memory");
  }
  __syncthreads();
  if (threadIdx.x < 66) out[threadIdx.x] = buf[threadIdx.x];
}
int main() {
  const size_t n = 513ull * 513 * 9;
  double *d, *o;
  cudaMalloc(&d, n * 8); cudaMalloc(&o, 80 * 8);
  double *h = new double[n];
  for (size_t i = 0; i < n; ++i) h[i] = (double)i;
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t gdim[1] = {n}; cuuint64_t gs[1] = {0}; cuuint32_t bd[1] = {66}; cuuint32_t es[1] = {1};
  CUresult r = ((EncodeFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, d, gdim, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  for (int c0 : {0, 64, 513, 1027, (int)n - 10}) {
    k<<<1, 128>>>(m, c0, o);
    cudaError_t e = cudaDeviceSynchronize();
    double ho[66]; cudaMemcpy(ho, o, 66 * 8, cudaMemcpyDeviceToHost);
    printf("c0=%d: %s first=%.0f last=%.0f\n", c0, cudaGetErrorString(e), ho[0], ho[65]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
