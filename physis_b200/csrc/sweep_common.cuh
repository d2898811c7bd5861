// Pieces shared by the TMA-ring sweep kernels (star7.cu, himeno.cu, pstag.cu):
// tile geometry, vector element access, exactly-rounded arithmetic helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace physis_b200 {
namespace sweep {

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; };
template <> struct VecOf<double> { using type = double2; };

template <typename T>
struct Geom {
  static constexpr int VEC = 16 / sizeof(T);   // elements per thread vector
  static constexpr int TXB = 512 / sizeof(T);  // interior box width = one warp of vectors
  static constexpr int HX = VEC;               // x halo, 16 bytes each side
  static constexpr int BW = TXB + 2 * HX;      // box width in elements (544 bytes)
  static constexpr int ROW_BYTES = BW * sizeof(T);
};

__device__ __forceinline__ float MulRn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float AddRn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double MulRn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double AddRn(double a, double b) { return __dadd_rn(a, b); }

__device__ __forceinline__ float SubRn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double SubRn(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ float Elem(const float4 &v, int j) {
  return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w;
}
__device__ __forceinline__ double Elem(const double2 &v, int j) { return j == 0 ? v.x : v.y; }
__device__ __forceinline__ void SetElem(float4 &v, int j, float x) {
  if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x;
}
__device__ __forceinline__ void SetElem(double2 &v, int j, double x) {
  if (j == 0) v.x = x; else v.y = x;
}

__device__ __forceinline__ void StoreVec(float4 *p, const float4 &v, bool streaming) {
  if (streaming) __stcs(p, v); else *p = v;
}
__device__ __forceinline__ void StoreVec(double2 *p, const double2 &v, bool streaming) {
  if (streaming) __stcs(p, v); else *p = v;
}

// Neighbour ordering of a multi-GPU sweep, fused into the sweep kernel itself
// (multigpu.cu): before touching halo planes every CTA waits until both ring
// neighbours have finished the previous sweep; the last CTA to finish publishes this
// sweep's number into the neighbours' flag words (peer memory, NVLink).  All null /
// zero on one GPU.
struct SlabSync {
  const uint32_t *flags;   // this rank's two flag words: [0] written by lo, [1] by hi
  uint32_t *to_lo;         // lo neighbour's flags[1]
  uint32_t *to_hi;         // hi neighbour's flags[0]
  unsigned *done;          // CTAs of this launch that have finished (self-resetting)
  uint32_t wait_epoch;     // neighbours must have completed this many sweeps
  uint32_t signal_epoch;   // the number this sweep publishes
  // a neighbour that stays silent longer than this is reported through `err` (host-mapped
  // words read by Runtime::CheckDeviceErrors) instead of hanging or killing the context
  unsigned long long timeout_ns;
  uint32_t *err;
  // halo-exchange profile (option halo_profile=1; role of the reference's DataCopyProfile,
  // runtime/timing.h:11-18): [0] nanoseconds CTAs spent waiting for the neighbours' flags, summed,
  // [1] the longest single wait, [2] CTAs that waited, [3] launches.  nullptr: not measured.
  unsigned long long *prof;
  // Overlap of the exchange with interior compute: work items are ordered so that the z chunks
  // touching the slab's two ends come first (SlabChunkOrder): the first `nb_lo` and the last
  // `nb_hi` chunks -- every chunk that reads a halo plane or computes a plane the neighbours
  // receive.  Once these `boundary_items` items are done -- the halo planes the neighbours wait
  // for are delivered and this rank's own halo planes are no longer read -- the sweep publishes
  // its number, and the rest of the sweep (the interior chunks) runs while the signal travels.
  // 0: publish when the whole sweep is done.
  int boundary_items;
  int nb_lo, nb_hi;
  // single sweeps: when > 0 the slab's first and last `zbl` planes are chunks of their own
  // (chunk sequence 0 and 1), followed by the interior in chunks of the kernel's zc planes: the
  // boundary work is over -- halo planes delivered, fences paid, number published -- a fraction
  // of a sweep after the launch, and the interior chunks carry none of it
  int zbl;
};

// Host side: which chunks are boundary chunks.  `reach` = planes a sweep reads beyond / delivers
// from each end of the slab (1 for a single sweep, 2 for the fused two-sweep pass); chunks are
// `zc` planes long, the last one `nz - (nzc-1)*zc`.
inline void SlabSyncSetBoundary(SlabSync *s, bool early, int nz, int zc, int nzc, int tiles, int reach) {
  s->boundary_items = 0;
  s->nb_lo = s->nb_hi = 0;
  s->zbl = 0;
  if (!early || !s->done) return;
  const int last = nz - (nzc - 1) * zc;
  int lo = (reach + zc - 1) / zc;                       // chunks covering the first `reach` planes
  int hi = 1 + (reach > last ? (reach - last + zc - 1) / zc : 0);  // ... and the last `reach`
  if (lo + hi >= nzc) {          // every chunk is a boundary chunk: natural order
    s->boundary_items = nzc * tiles;
    return;
  }
  s->nb_lo = lo;
  s->nb_hi = hi;
  s->boundary_items = (lo + hi) * tiles;
}

// Single sweeps on a z-slab: short boundary chunks first.  Rewrites the chunking (*zc, *nzc) of
// a domain of `nz` planes when the slab is thick enough; falls back to SlabSyncSetBoundary.
inline void SlabSyncPlanEnds(SlabSync *s, bool early, int nz, int *zc, int *nzc, int tiles, int reach,
                             int zbl) {
  if (early && s->done && zbl >= reach && nz >= 2 * zbl + 8) {
    const int interior = nz - 2 * zbl;
    const int nint = (interior + *zc - 1) / *zc;
    *zc = (interior + nint - 1) / nint;
    *nzc = 2 + nint;
    s->boundary_items = 2 * tiles;
    s->nb_lo = s->nb_hi = 0;
    s->zbl = zbl;
    return;
  }
  SlabSyncSetBoundary(s, early, nz, *zc, *nzc, tiles, reach);
}

// planes [*zb, *ze) of the work item with chunk sequence number `seq` (single-sweep kernels)
__device__ __forceinline__ void SlabChunkRange(const SlabSync &s, int seq, int nzc, int zc, int dz0, int dz1,
                                               int *zb, int *ze);

// z chunk a work item with chunk sequence number `seq` processes: the nb_lo lowest chunks, the
// nb_hi highest, then the interior ones in order
__host__ __device__ __forceinline__ int SlabChunkOrder(const SlabSync &s, int seq, int nzc) {
  const int nb = s.nb_lo + s.nb_hi;
  if (nb == 0) return seq;
  return seq < s.nb_lo ? seq : (seq < nb ? nzc - nb + seq : seq - s.nb_hi);
}

__device__ __forceinline__ void SlabChunkRange(const SlabSync &s, int seq, int nzc, int zc, int dz0, int dz1,
                                               int *zb, int *ze) {
  if (s.zbl > 0) {
    if (seq < 2) {
      *zb = seq == 0 ? dz0 : dz1 - s.zbl;
      *ze = *zb + s.zbl;
    } else {
      *zb = dz0 + s.zbl + (seq - 2) * zc;
      *ze = min(*zb + zc, dz1 - s.zbl);
    }
  } else {
    *zb = dz0 + SlabChunkOrder(s, seq, nzc) * zc;
    *ze = min(*zb + zc, dz1);
  }
}

__device__ __forceinline__ uint32_t LdAcquireSys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void StReleaseSys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long GlobalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spins until both flag words have reached `epoch`.  A lost signal becomes an error the host
// reports (which neighbour, which sweep, what was seen), not a hang and not a dead context:
// the bound is wall time (option sync_timeout_s), so profiler serialisation or ranks
// time-slicing one GPU do not trip it.
static __device__ __forceinline__ void SlabWaitFlags(const uint32_t *flags, uint32_t epoch,
                                                  unsigned long long timeout_ns, uint32_t *err) {
  for (int i = 0; i < 2; ++i) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    uint32_t seen;
    while ((int32_t)((seen = LdAcquireSys(flags + i)) - epoch) < 0) {
      __nanosleep(64);
      if ((++spins & 4095u) == 0) {
        const unsigned long long now = GlobalTimerNs();
        if (t0 == 0) t0 = now;
        volatile uint32_t *e = err;
        if (e && e[0] != 0) return;  // another CTA has already given up
        if (now - t0 > timeout_ns) {
          if (e) {
            e[1] = epoch; e[2] = seen; e[3] = blockIdx.x;
            __threadfence_system();
            e[0] = 1u + (uint32_t)i;
            __threadfence_system();
          }
          return;
        }
      }
    }
  }
}

// Called by EVERY thread of a CTA before it touches a halo plane, and executed by every thread:
// each thread polls the two flag words itself (a warp's loads of one address are one request, so
// a CTA polls with one request per warp).  Deliberately not "thread 0 waits, the rest wait at the
// barrier": a spin loop under any thread-dependent condition in front of the sweep loop made the
// compiler treat the loop as possibly diverged -- it cloned a step and wrapped the shuffles in
// convergence barriers (4272 SASS instructions against 2648), and the z-slab form of the fused
// pass ran 13 % slower than the single-GPU form on ONE GPU with nothing to exchange.  Only CTAs
// that will process a boundary item need the neighbours: an interior item reads and writes
// nothing but this rank's own interior planes, which stream order already protects.  Items are
// visited in increasing order and the boundary items come first, so a CTA whose first item is
// interior never meets one.
__device__ __forceinline__ void SlabSyncWait(const SlabSync &s) {
  if (!s.flags) return;
  if (s.boundary_items > 0 && (int)blockIdx.x >= s.boundary_items) return;
  if (s.prof) {
    const unsigned long long t0 = GlobalTimerNs();
    SlabWaitFlags(s.flags, s.wait_epoch, s.timeout_ns, s.err);
    const unsigned long long dt = GlobalTimerNs() - t0;
    if (threadIdx.x == 0) {
      atomicAdd(s.prof + 0, dt);
      atomicMax(s.prof + 1, dt);
      atomicAdd(s.prof + 2, 1ull);
      if (blockIdx.x == 0) atomicAdd(s.prof + 3, 1ull);
    }
    return;
  }
  SlabWaitFlags(s.flags, s.wait_epoch, s.timeout_ns, s.err);
}

// Called by every consumer thread of a CTA after it has finished work item `item`: the last
// of the boundary items to finish publishes the sweep's number to both neighbours.  A CTA's
// boundary items are the first of its sequence (item, item + gridDim.x, ...): it reports them all
// at once after the last of them -- one barrier and one fence per CTA, not per item (a fence
// after megabytes of stores costs microseconds).
__device__ __forceinline__ void SlabSyncItemDone(const SlabSync &s, int item, int nthreads, bool leader) {
  if (!s.done || item >= s.boundary_items) return;
  if (item + (int)gridDim.x < s.boundary_items) return;  // more boundary items of this CTA follow
  // every consumer's stores (incl. the peer stores) precede the barrier; the leader's
  // system-scope fence after it then orders all of them before the counter and the flags
  // (fence cumulativity -- the pattern of a grid-wide barrier: bar.sync, one thread fences and
  // signals).  (A GPU-scope fence here with a single system-scope one in the publishing CTA
  // measured the same: profiles/r2_experiments.txt.)
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  if (leader) {
    const unsigned mine = 1u + (unsigned)(item / (int)gridDim.x);  // boundary items this CTA ran
    __threadfence_system();
    const unsigned prev = atomicAdd(s.done, mine);
    if (prev + mine == (unsigned)s.boundary_items) {
      *s.done = 0;           // next launch starts from zero (stream order)
      __threadfence_system();
      StReleaseSys(s.to_lo, s.signal_epoch);
      StReleaseSys(s.to_hi, s.signal_epoch);
    }
  }
}

// Called by every consumer thread after its last store; `nthreads` consumer threads
// take part (a multiple of 32), `leader` is true for exactly one of them.  (With boundary
// items the number has already been published by SlabSyncItemDone.)
__device__ __forceinline__ void SlabSyncSignal(const SlabSync &s, int nthreads, bool leader) {
  if (!s.done || s.boundary_items > 0) return;
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  if (leader) {
    __threadfence_system();  // all consumers' stores (incl. peer stores), ordered by the barrier
    const unsigned prev = atomicAdd(s.done, 1u);
    if (prev == gridDim.x - 1) {
      *s.done = 0;           // next launch starts from zero (stream order)
      __threadfence_system();
      StReleaseSys(s.to_lo, s.signal_epoch);
      StReleaseSys(s.to_hi, s.signal_epoch);
    }
  }
}

constexpr int kBarrierBytes = 128;  // full[] + empty[], up to 8 stages
constexpr int kMaxStages = 8;

template <typename T, int TY>
__host__ __device__ constexpr int BoxStride() {
  return ((TY + 2) * Geom<T>::ROW_BYTES + 127) / 128 * 128;
}

}  // namespace sweep
}  // namespace physis_b200
