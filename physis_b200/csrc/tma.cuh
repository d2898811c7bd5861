// sm_100a async-copy primitives used by the sweep kernels: mbarrier
// producer/consumer handshakes and TMA tiled loads (cp.async.bulk.tensor ->
// SASS UTMALDG).  Inline PTX only; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace physis_b200 {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// Make mbarrier initialisation visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// try_wait suspends in hardware up to a time limit, so this loop is not a hot
// spin; its fast path is two instructions.  A bounded number of retries turns a
// lost arrival into a trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 t;\n"
      "mov.u32 t, 0;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "add.u32 t, t, 1;\n"
      "setp.lt.u32 q, t, 0x4000000;\n"
      "@q bra MBAR_WAIT;\n"
      "trap;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// Ampere-style asynchronous copies (SASS LDGSTS) for data TMA cannot describe.
__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 3-D tiled prefetch global -> L2 of the box at (c0, c1, c2): one instruction of one thread for
// the whole box, no shared memory, no completion to wait for.
__device__ __forceinline__ void prefetch_3d(const CUtensorMap *m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// 3-D tiled load global -> shared; completion is signalled on `bar` as
// transaction bytes.  Out-of-bounds box elements are zero-filled.
__device__ __forceinline__ void load_3d(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                        int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 2-D tiled load.  (A box must start on a 16-byte boundary in global memory: a 1-D or 2-D box
// starting at an odd fp64 element raises an illegal-instruction error, tools/microbench/tma1d_test.cu.)
__device__ __forceinline__ void load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                        int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1)
      : "memory");
}

// Same with an L2 cache-policy operand (createpolicy result).
__device__ __forceinline__ void load_3d_hint(void *smem_dst, const CUtensorMap *m,
                                             uint64_t *bar, int c0, int c1, int c2,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

}  // namespace tma

// Host side: cuTensorMapEncodeTiled through the runtime's driver entry point
// lookup (no link-time dependency on libcuda).
enum class TmaElem { F32, F64 };
// Describes a dense 3-D array (x fastest) of `dim` elements and a box of
// `box` elements; returns false if the shape violates a TMA constraint
// (16-byte strides/alignment, box <= 256 per dim).
bool EncodeTensorMap3D(CUtensorMap *out, TmaElem elem, const void *base, const int dim[3],
                       const int box[3]);
// Dense 2-D array (x fastest) of `dim` elements read in boxes of `box` elements.
bool EncodeTensorMap2D(CUtensorMap *out, TmaElem elem, const void *base, const int dim[2],
                       const int box[2]);

}  // namespace physis_b200
