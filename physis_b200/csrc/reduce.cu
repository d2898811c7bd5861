// PSReduce on the b200 target: whole-grid MAX/MIN/SUM/PROD of a primitive-type
// grid into one host scalar.
//
// Replaces runtime/reduce_cuda.{h,cu} (thrust::reduce / max_element /
// min_element, reduce_cuda.h:19-40) and, for the single-GPU case, the CUB
// two-stage kernels of runtime/reduce_grid_mpi_cuda_exp.cu:176-328.  Semantics
// follow the REFERENCE target (runtime/libphysis_rt_ref.cc:19-30): every one of
// the num_elms elements takes part, the result has the grid's element type.
// MAX/MIN use true identities (lowest/highest), not the reference CUDA
// runtime's FLT_MIN (runtime/reduce.h:61-63), which is wrong for all-negative
// data; REF seeds with d[0] and is right — we agree with REF.
//
// Shape: stage 1 — persistent grid (a multiple of the SM count), each thread
// folds 16-byte vector loads in a grid-stride loop, warp shuffle tree, one
// shared-memory hop across warps, one partial per block; stage 2 — a single
// block folds the partials in a fixed order.  The combine order depends only on
// (num_elms, launch shape), so results are run-to-run deterministic.  Integer
// results are exact (wrap-around like the CPU's two's-complement adds);
// floating SUM/PROD differ from REF's sequential left fold by reassociation
// only (tests bound it; exactly-representable data is bit-identical).
#include "runtime.h"

#include <type_traits>
#include <vector>

#include <cfloat>
#include <climits>

namespace physis_b200 {

namespace {

constexpr int kThreads = 256;

template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec16<int> { using type = int4; static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec16<long> { using type = longlong2; static constexpr int N = 2; };

template <typename T> __host__ __device__ inline T Lowest();
template <> __host__ __device__ inline float Lowest<float>() { return -INFINITY; }
template <> __host__ __device__ inline double Lowest<double>() { return -INFINITY; }
template <> __host__ __device__ inline int Lowest<int>() { return INT_MIN; }
template <> __host__ __device__ inline long Lowest<long>() { return LONG_MIN; }
template <typename T> __host__ __device__ inline T Highest();
template <> __host__ __device__ inline float Highest<float>() { return INFINITY; }
template <> __host__ __device__ inline double Highest<double>() { return INFINITY; }
template <> __host__ __device__ inline int Highest<int>() { return INT_MAX; }
template <> __host__ __device__ inline long Highest<long>() { return LONG_MAX; }

template <typename T> struct Unsigned { using type = T; };
template <> struct Unsigned<int> { using type = unsigned int; };
template <> struct Unsigned<long> { using type = unsigned long; };

template <typename T, int OP>
struct Op {
  __host__ __device__ static T identity() {
    if (OP == PS_MAX) return Lowest<T>();
    if (OP == PS_MIN) return Highest<T>();
    if (OP == PS_SUM) return (T)0;
    return (T)1;
  }
  __host__ __device__ static T apply(T x, T y) {
    using U = typename Unsigned<T>::type;
    if (OP == PS_MAX) return (x > y) ? x : y;
    if (OP == PS_MIN) return (x < y) ? x : y;
    if (OP == PS_SUM) return (T)((U)x + (U)y);
    return (T)((U)x * (U)y);
  }
};

template <typename T>
__device__ __forceinline__ T ShflDown(T v, int d) {
  return __shfl_down_sync(0xffffffffu, v, d);
}
template <>
__device__ __forceinline__ long ShflDown<long>(long v, int d) {
  return (long)__shfl_down_sync(0xffffffffu, (long long)v, d);
}

template <typename T, int OP>
__device__ __forceinline__ T BlockFold(T v) {
  __shared__ T warp_part[kThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = Op<T, OP>::apply(v, ShflDown(v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < kThreads / 32) ? warp_part[lane] : Op<T, OP>::identity();
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) v = Op<T, OP>::apply(v, ShflDown(v, d));
  }
  return v;  // valid in thread 0
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
ReduceStage1(const T *__restrict__ data, long n, T *__restrict__ partials) {
  using V = typename Vec16<T>::type;
  constexpr int N = Vec16<T>::N;
  T acc = Op<T, OP>::identity();
  // scalar head up to the first 16-byte boundary (a slab's interior starts after its
  // halo planes, which need not be a multiple of 16 bytes)
  const long head = min((long)(((16 - (reinterpret_cast<uintptr_t>(data) & 15)) & 15) / sizeof(T)), n);
  if (blockIdx.x == 0 && threadIdx.x < head) acc = Op<T, OP>::apply(acc, data[threadIdx.x]);
  data += head;
  n -= head;
  const long nvec = n / N;
  const V *vdata = reinterpret_cast<const V *>(data);
  const long stride = (long)gridDim.x * kThreads;
  long i = (long)blockIdx.x * kThreads + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    V v0 = __ldg(vdata + i), v1 = __ldg(vdata + i + stride), v2 = __ldg(vdata + i + 2 * stride),
      v3 = __ldg(vdata + i + 3 * stride);
    const T *e0 = reinterpret_cast<const T *>(&v0), *e1 = reinterpret_cast<const T *>(&v1),
            *e2 = reinterpret_cast<const T *>(&v2), *e3 = reinterpret_cast<const T *>(&v3);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      acc = Op<T, OP>::apply(acc, e0[k]);
      acc = Op<T, OP>::apply(acc, e1[k]);
      acc = Op<T, OP>::apply(acc, e2[k]);
      acc = Op<T, OP>::apply(acc, e3[k]);
    }
  }
  for (; i < nvec; i += stride) {
    V v0 = __ldg(vdata + i);
    const T *e0 = reinterpret_cast<const T *>(&v0);
#pragma unroll
    for (int k = 0; k < N; ++k) acc = Op<T, OP>::apply(acc, e0[k]);
  }
  // scalar tail (n not a multiple of the vector width)
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - nvec * N))
    acc = Op<T, OP>::apply(acc, data[nvec * N + threadIdx.x]);
  acc = BlockFold<T, OP>(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
ReduceStage2(const T *__restrict__ partials, int n, T *__restrict__ out) {
  T acc = Op<T, OP>::identity();
  for (int i = threadIdx.x; i < n; i += kThreads) acc = Op<T, OP>::apply(acc, partials[i]);
  acc = BlockFold<T, OP>(acc);
  if (threadIdx.x == 0) *out = acc;
}

// PSReduce(PS_SUM) right after the sweep that emitted the grid: the sweep left one fp64
// partial per CTA (Grid::SumCache), folded here by one block in a fixed order.
__global__ void __launch_bounds__(kThreads)
FoldPartials(const double *__restrict__ partials, int n, float *__restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) acc += partials[i];
  acc = BlockFold<double, PS_SUM>(acc);
  if (threadIdx.x == 0) *out = (float)acc;
}

template <typename T, int OP>
void Run(Runtime *rt, const Grid &g, void *out_host) {
  // this rank's interior planes are contiguous in the local allocation
  const long n = (long)g.plane_elms * g.nz_loc;
  const T *data = (const T *)g.members[0].dev + (size_t)g.halo * g.plane_elms;
  constexpr int N = Vec16<T>::N;
  long want = (n / N + (long)kThreads * 4 - 1) / ((long)kThreads * 4);
  int blocks = (int)std::max<long>(1, std::min<long>(want, (long)rt->sm_count * 8));
  DeviceBuffer &scr = rt->small_scratch(sizeof(T) * (size_t)(blocks + 1));
  T *partials = (T *)scr.get();
  T *result = partials + blocks;
  bool from_partials = false;
  if constexpr (std::is_same<T, float>::value && OP == PS_SUM) {
    if (g.sum_cache.valid && rt->opt.reduce_fuse) {
      // (a rank without a share of the sweep's domain holds no partials: count 0, sum 0)
      const double *part = g.sum_cache.partials ? (const double *)g.sum_cache.partials->get() : nullptr;
      FoldPartials<<<1, kThreads, 0, rt->stream>>>(part, part ? g.sum_cache.count : 0, result);
      rt->stats.kernel_launches += 1;
      rt->stats.reduces_from_partials += 1;
      from_partials = true;
    }
  }
  if (!from_partials) {
    ReduceStage1<T, OP><<<blocks, kThreads, 0, rt->stream>>>(data, n, partials);
    ReduceStage2<T, OP><<<1, kThreads, 0, rt->stream>>>(partials, blocks, result);
    rt->stats.kernel_launches += 2;
  }
  PSB_CUDA(cudaGetLastError());
  PSB_CUDA(cudaMemcpyAsync(out_host, result, sizeof(T), cudaMemcpyDeviceToHost, rt->stream));
  PSB_CUDA(cudaStreamSynchronize(rt->stream));
  rt->CheckDeviceErrors("PSReduce");
  rt->stats.d2h_bytes += sizeof(T);
  if (g.decomposed) {
    // cross-GPU combine: one scalar per rank, folded in rank order on every rank
    // (the reference: MPI_Reduce of the per-rank scalar, grid_space_mpi_cuda.h:571-598)
    const int W = rt->world();
    std::vector<T> all(W);
    rt->comm->AllGather(out_host, all.data(), sizeof(T));
    T acc = all[0];
    for (int r = 1; r < W; ++r) acc = Op<T, OP>::apply(acc, all[r]);
    *(T *)out_host = acc;
  }
}

template <typename T>
void Dispatch(Runtime *rt, const Grid &g, PSReduceOp op, void *out) {
  switch (op) {
    case PS_MAX: Run<T, PS_MAX>(rt, g, out); break;
    case PS_MIN: Run<T, PS_MIN>(rt, g, out); break;
    case PS_SUM: Run<T, PS_SUM>(rt, g, out); break;
    case PS_PROD: Run<T, PS_PROD>(rt, g, out); break;
    default: PSAbort(1);
  }
}

}  // namespace

void ReduceGrid(Runtime *rt, const Grid &g, PSType type, PSReduceOp op, void *out_host) {
  PSB_CHECK(g.num_elms > 0, "PSReduce on an empty grid");
  switch (type) {
    case PS_FLOAT: Dispatch<float>(rt, g, op, out_host); break;
    case PS_DOUBLE: Dispatch<double>(rt, g, op, out_host); break;
    case PS_INT: Dispatch<int>(rt, g, op, out_host); break;
    case PS_LONG: Dispatch<long>(rt, g, op, out_host); break;
    default: PSB_CHECK(false, "PSReduce: unsupported element type");
  }
}

}  // namespace physis_b200
