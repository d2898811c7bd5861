// Device-side AoS <-> SoA transposition for user-defined point types.
//
// The host/REF layout of a user-type grid is an array of structs; on the
// device every member is its own array (array members plane-major:
// [flat array index][grid element], as translator/cuda_runtime_builder.cc:
// 267-305 indexes them).  The reference CUDA target transposes on the HOST,
// element by element, through a pinned temporary generated per type
// (cuda_runtime_builder.cc:737-858).  Here the struct bytes are moved with one
// bulk copy and transposed on the GPU: a block stages a tile of structs in
// shared memory with 16-byte coalesced accesses and then streams each member
// out (or in) with unit stride.
#include "runtime.h"

namespace physis_b200 {

namespace {

constexpr int kMaxMembers = 24;
constexpr int kThreads = 256;

struct MemberTable {
  int n;
  int elm_size;
  int padded;   // the struct has padding bytes (no member covers them)
  void *dev[kMaxMembers];
  int offset[kMaxMembers];
  int size[kMaxMembers];
  int count[kMaxMembers];
};

template <typename W>
__device__ __forceinline__ void MoveScalar(char *dst, const char *src) {
  *reinterpret_cast<W *>(dst) = *reinterpret_cast<const W *>(src);
}

// TO_SOA: aos -> members; otherwise members -> aos.
template <bool TO_SOA>
__global__ void __launch_bounds__(kThreads)
TransposeKernel(char *aos, const __grid_constant__ MemberTable tbl, long num_elms,
                int tile_elems) {
  extern __shared__ __align__(16) char tile[];
  const int es = tbl.elm_size;
  const long ntiles = (num_elms + tile_elems - 1) / tile_elems;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long e0 = t * tile_elems;
    const int ne = (int)min((long)tile_elems, num_elms - e0);
    const long byte0 = e0 * es;
    const int nbytes = ne * es;
    const bool vec_ok = ((byte0 & 15) == 0);
    if (TO_SOA) {
      if (vec_ok) {
        const uint4 *src = reinterpret_cast<const uint4 *>(aos + byte0);
        for (int i = threadIdx.x; i < nbytes / 16; i += kThreads)
          reinterpret_cast<uint4 *>(tile)[i] = src[i];
        for (int i = (nbytes / 16) * 16 + threadIdx.x; i < nbytes; i += kThreads)
          tile[i] = aos[byte0 + i];
      } else {
        for (int i = threadIdx.x; i < nbytes; i += kThreads) tile[i] = aos[byte0 + i];
      }
      __syncthreads();
    } else if (tbl.padded) {
      // padding bytes belong to no member: they leave the device as zeros (what a fresh
      // REFERENCE-target grid holds, runtime/libphysis_rt_ref.cc:53-71), never as stale
      // shared memory
      for (int i = threadIdx.x; i < (nbytes + 3) / 4; i += kThreads) reinterpret_cast<unsigned *>(tile)[i] = 0u;
      __syncthreads();
    }
    for (int m = 0; m < tbl.n; ++m) {
      const int sz = tbl.size[m];
      for (int c = 0; c < tbl.count[m]; ++c) {
        char *plane = (char *)tbl.dev[m] + ((size_t)c * num_elms + e0) * sz;
        const int off = tbl.offset[m] + c * sz;
        for (int i = threadIdx.x; i < ne; i += kThreads) {
          char *s = tile + (size_t)i * es + off;
          char *d = plane + (size_t)i * sz;
          if (TO_SOA) {
            if (sz == 8) MoveScalar<unsigned long long>(d, s);
            else MoveScalar<unsigned int>(d, s);
          } else {
            if (sz == 8) MoveScalar<unsigned long long>(s, d);
            else MoveScalar<unsigned int>(s, d);
          }
        }
      }
    }
    if (!TO_SOA) {
      __syncthreads();
      if (vec_ok) {
        uint4 *dst = reinterpret_cast<uint4 *>(aos + byte0);
        for (int i = threadIdx.x; i < nbytes / 16; i += kThreads)
          dst[i] = reinterpret_cast<const uint4 *>(tile)[i];
        for (int i = (nbytes / 16) * 16 + threadIdx.x; i < nbytes; i += kThreads)
          aos[byte0 + i] = tile[i];
      } else {
        for (int i = threadIdx.x; i < nbytes; i += kThreads) aos[byte0 + i] = tile[i];
      }
    }
    __syncthreads();
  }
}

void Launch(const Grid &g, char *aos, bool to_soa, cudaStream_t s) {
  PSB_CHECK((int)g.members.size() <= kMaxMembers, "too many struct members");
  MemberTable tbl;
  tbl.n = (int)g.members.size();
  tbl.elm_size = g.elm_size;
  int covered = 0;
  for (int m = 0; m < tbl.n; ++m) {
    const MemberLayout &ml = g.members[m];
    PSB_CHECK(ml.size == 4 || ml.size == 8, "struct members must be 4- or 8-byte scalars");
    tbl.dev[m] = ml.dev;
    tbl.offset[m] = ml.aos_offset;
    tbl.size[m] = ml.size;
    tbl.count[m] = ml.count;
    covered += ml.size * ml.count;
  }
  tbl.padded = covered < g.elm_size ? 1 : 0;
  // tile: as many structs as fit 32 KiB, multiple of 16 so tiles stay 16-byte aligned
  int tile_elems = (32 * 1024) / g.elm_size;
  tile_elems = tile_elems / 16 * 16;
  PSB_CHECK(tile_elems >= 16, "user-defined point type too large");
  size_t smem = (size_t)tile_elems * g.elm_size;
  long ntiles = (g.n_alloc + tile_elems - 1) / tile_elems;
  int blocks = (int)std::min<long>(ntiles, 148L * 6);
  if (to_soa)
    TransposeKernel<true><<<blocks, kThreads, smem, s>>>(aos, tbl, (long)g.n_alloc, tile_elems);
  else
    TransposeKernel<false><<<blocks, kThreads, smem, s>>>(aos, tbl, (long)g.n_alloc, tile_elems);
  PSB_CUDA(cudaGetLastError());
}

}  // namespace

void LaunchAosToSoa(const Grid &g, const void *aos_dev, cudaStream_t s) {
  Launch(g, (char *)const_cast<void *>(aos_dev), true, s);
}
void LaunchSoaToAos(const Grid &g, void *aos_dev, cudaStream_t s) {
  Launch(g, (char *)aos_dev, false, s);
}

}  // namespace physis_b200
