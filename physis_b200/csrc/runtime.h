// Internal object model of the b200 runtime (host side, C++17).
//
// Replaces, for the `b200` target, the reference's
//   runtime/buffer.{h,cc}, buffer_cuda.{h,cu}   -> DeviceBuffer / PinnedBuffer
//   runtime/grid.{h,cc} (Grid, GridSpace)        -> Grid / GridSpace
//   runtime/runtime_cuda.h, libphysis_rt_cuda.cc -> Runtime + the C entry points
// It is not a translation of those classes: one device allocation per struct
// member (SoA), pinned double-buffered staging for host<->device traffic, a
// single explicit stream pair instead of the default stream, and no virtual
// Buffer hierarchy.
#pragma once
#include <cstdint>
#include <cstddef>
#include <map>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "physis/physis_b200.h"
#include "common.h"
#include "comm.h"
#include "sweep_common.cuh"

namespace physis_b200 {

// Device allocation with grow-only capacity (role of Buffer::EnsureCapacity,
// runtime/buffer.cc:22-35, except that a repeated request of the same size keeps the
// block: the reference's `>=` would re-allocate on every PSReduce / user-type copy).  Fresh memory is zero-filled like
// BufferCUDADev (runtime/buffer_cuda.cu:106-117).
class DeviceBuffer {
 public:
  DeviceBuffer() = default;
  ~DeviceBuffer() { Free(); }
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  // Returns false on out-of-memory (the caller decides whether that is fatal).
  bool Allocate(size_t bytes, cudaStream_t stream);
  void EnsureCapacity(size_t bytes, cudaStream_t stream);
  void Free();
  void *get() const { return ptr_; }
  size_t size() const { return size_; }
  size_t capacity() const { return capacity_; }

 private:
  void *ptr_ = nullptr;
  size_t size_ = 0;
  size_t capacity_ = 0;
};

// Page-locked host staging (role of BufferCUDAHost, runtime/buffer_cuda.cu:12-52).
class PinnedBuffer {
 public:
  PinnedBuffer() = default;
  ~PinnedBuffer() { Free(); }
  PinnedBuffer(const PinnedBuffer &) = delete;
  PinnedBuffer &operator=(const PinnedBuffer &) = delete;
  void EnsureCapacity(size_t bytes);
  void Free();
  void *get() const { return ptr_; }
  size_t capacity() const { return capacity_; }

 private:
  void *ptr_ = nullptr;
  size_t capacity_ = 0;
};

struct MemberLayout {
  PSType type = PS_FLOAT;
  int size = 0;          // bytes of one scalar of this member
  int count = 1;         // array members: product of dims
  int aos_offset = 0;    // byte offset inside the host struct
  void *dev = nullptr;   // local device array [count][n_alloc] (SoA; arrays plane-major)
  // multi-GPU: the ring neighbours' allocations of the same member, mapped into
  // this process with CUDA IPC (nullptr on one GPU)
  void *peer_lo = nullptr;
  void *peer_hi = nullptr;
};

// One Physis grid.  `handle` is what the generated code holds (its address is
// the `__PSGrid*`); everything else is runtime-private.
class Grid {
 public:
  __PSGrid handle;            // must stay first: Grid* <-> __PSGrid* casts
  int id = 0;
  PSType type = PS_FLOAT;
  int num_dims = 0;
  int dim[PS_MAX_DIM] = {1, 1, 1};   // GLOBAL extents (what PSGridDim reports)
  int64_t num_elms = 0;              // global element count
  int elm_size = 0;
  // z-slab decomposition over the process group (role of GridMPI's local_size /
  // local_offset / halo, runtime/grid_mpi.h:20-276).  The local allocation holds the
  // planes [z_off - halo, z_off + nz_loc + halo) of the last dimension, halo planes
  // in place.  On one GPU: z_off = 0, nz_loc = dim[last], halo = 0, ldim == dim.
  bool decomposed = false;
  int ldim[PS_MAX_DIM] = {1, 1, 1};  // local allocated extents
  int z_off = 0, nz_loc = 0, halo = 0;
  int64_t plane_elms = 0;            // elements of one plane of the decomposed dimension
  int64_t n_alloc = 0;               // local allocated elements
  int lo_nz_loc = 0;                 // interior planes of the lower ring neighbour
  std::vector<MemberLayout> members;   // size 1 for primitive grids
  std::vector<DeviceBuffer *> storage; // one per member
  void *dev_view = nullptr;            // host copy of the by-value device view
  bool external_dev = false;           // allocated by a generated devNew function

  // PSReduce(PS_SUM) fused with the producing sweep: a sweep that emits this grid can leave
  // per-CTA partial sums of what it emitted (himeno.cu, the residual form); a PSReduce that
  // follows folds those few numbers instead of reading the grid back.  The partials stand for
  // the whole grid only while everything outside the sweep's domain is still the zero fill of
  // __PSGridNew, which is what the bookkeeping below establishes; any other write drops them.
  struct SumCache {
    bool valid = false;
    int count = 0;              // partial sums (one per CTA of the producing launch)
    DeviceBuffer *partials = nullptr;  // doubles
  } sum_cache;
  bool contents_unknown = false;       // written by the host or by a kernel the runtime cannot see into
  bool emitted = false;                // some sweep has emitted into this grid ...
  int emit_min[PS_MAX_DIM] = {0, 0, 0}, emit_max[PS_MAX_DIM] = {0, 0, 0};  // ... inside this box (global)
  void NoteUnknownWrite() { sum_cache.valid = false; contents_unknown = true; }
  void NoteEmit(const __PSDomain &dom) {
    sum_cache.valid = false;
    for (int i = 0; i < PS_MAX_DIM; ++i) {
      emit_min[i] = emitted ? (dom.local_min[i] < emit_min[i] ? dom.local_min[i] : emit_min[i]) : dom.local_min[i];
      emit_max[i] = emitted ? (dom.local_max[i] > emit_max[i] ? dom.local_max[i] : emit_max[i]) : dom.local_max[i];
    }
    emitted = true;
  }
  // true when every element outside `dom` still holds the zero fill
  bool ZeroOutside(const __PSDomain &dom) const {
    if (contents_unknown) return false;
    if (!emitted) return true;
    for (int i = 0; i < num_dims; ++i)
      if (emit_min[i] < dom.local_min[i] || emit_max[i] > dom.local_max[i]) return false;
    return true;
  }

  bool is_user_type() const { return type == PS_USER; }
  size_t bytes() const { return (size_t)elm_size * (size_t)num_elms; }
  size_t alloc_bytes() const { return (size_t)elm_size * (size_t)n_alloc; }
  // local plane index of global plane zg of the decomposed dimension, or -1 when
  // this rank does not hold it as an interior plane
  int LocalInterior(int zg) const {
    return (zg >= z_off && zg < z_off + nz_loc) ? zg - z_off + halo : -1;
  }
  static Grid *FromHandle(void *h) { return reinterpret_cast<Grid *>(h); }
};

// id -> grid registry (role of GridSpace, runtime/grid.h:73-119).
class Runtime;
class GridSpace {
 public:
  ~GridSpace();
  Grid *Create(const __PSGridTypeInfo *ti, int num_dims, const int *dim, Runtime *rt);
  void Destroy(Grid *g);
  void Clear();
  Grid *Find(int id) const;
  size_t live() const { return grids_.size(); }

 private:
  std::map<int, Grid *> grids_;
  int next_id_ = 1;
};

struct Options {
  // star-7 sweep: tile shape (index into star7.cu's table, -1 = automatic), ring depth, z chunk,
  // resident CTAs (0 = automatic), streaming stores
  int star7_stages = 0, star7_zc = 0, star7_occ = 0;
  int star7_variant = -1, star7_sthint = 0;
  int star7_fuse = 1;     // 1: a ping-pong pair of whole-grid 7-pt sweeps runs as one fused two-sweep pass
  int star7_pair_zc = 0;  // z chunk of the fused kernel; 0 = automatic
  int star7_pair_zbl = 8;       // multi-GPU: planes of the boundary chunks that run first (0: equal chunks)
  int star7_pair_zbias = 8;     // ... and planes by which their CTAs' interior chunks are shorter
  int star7_pair_xtile = 1;     // 1: rows wider than one fused tile are cut into x tiles
  int star7_pair_variant = -1;  // tile shape of the x-tiled form; -1 = automatic
  int star7_iso = 1;      // 1: equal neighbour coefficients use the shared-product form of the fused kernel
  int himeno_by = 0, himeno_zc = 0, himeno_stages = 0, himeno_occ = 0, himeno_carveout = 0;
  int himeno_sthint = 0;    // bit 0 / bit 1: evict-first stores of p1 / of the residual grid
  int himeno_fuse = 1;      // 1: a ping-pong pair of interior Himeno sweeps runs as fused two-sweep passes
                            //    where that pays, 2: wherever it can, 0: never
  int himeno_pair_zc = 0;   // z chunk of the fused Himeno kernel; 0 = automatic
  int himeno_pair_pf = 1;   // planes ahead the coefficient rows are prefetched into L2
  int himeno_pair_pfmode = 0;  // ... 1: by bulk prefetches per row, 2: by tensor-map prefetch of the tile, 0: automatic
  int pstag_push = 1;     // the config-5 sweep's exchange: 1 in the kernel, number published at the end of the
                          // sweep; 2 in the kernel with boundary chunks first and an early signal; 0 copy-based
  int pstag_variant = 0 /* index into pstag.cu's tile shapes */, pstag_stages = 0, pstag_occ = 0;
  int time_kernels = 0;        // per-family CUDA-event timing (for bench roofline)
  size_t stage_chunk = 64u << 20;  // pinned staging chunk for pageable copies
  int copy_threads = 0;            // host threads that fill / drain a staging chunk (0 = automatic:
                                   // up to 8, shared fairly between the ranks of one box)
  // multi-GPU
  int halo = 2;          // halo planes per side of decomposed grids (single sweeps use the one
                         // next to the interior; the fused two-sweep pass needs two)
  int halo_push = 1;     // 1: specialised sweeps store their boundary planes straight into
                         //    the neighbour's halo (fused); 0: peer copies after the kernel
  int sync_mode = 2;     // 2: waits and signals fused into the sweep kernels, 0: stream
                         //    memory operations, 1: one-thread signal/wait kernels
  int slab_zbl = 8;      // single sweeps on z-slabs: planes of the boundary chunks that run first (0: equal chunks)
  int early_signal = 1;  // with sync_mode 2: publish a sweep's number as soon as its boundary z chunks
                         // are done, so that the neighbours' next sweep overlaps the interior chunks
  int copyout_gather = 1;  // PSGridCopyout fills the whole host array on every rank
  int sync_timeout_s = 120;  // a neighbour silent for longer than this is a reported error
  int plan_cache = 1;    // keep prepared sweep plans across PSStencilRun calls
  int pdl = 1;           // 7-point sweeps and fused passes are launched as programmatic dependents of the kernel before them
  int debug_slab = 0;    // timing experiments (WRONG results): bit 0 no halo stores, bit 1 no neighbour ordering
  int halo_profile = 0;  // 1: sweeps record how long their CTAs wait for the ring neighbours
  int reduce_fuse = 1;   // PSReduce(PS_SUM) folds the partial sums the producing sweep left (himeno.cu)
  int autotune = 0;      // 1: the first long PSStencilRun of a shape times the kernel forms that can run it
                         //    on its own first iterations and keeps the fastest (stencil_run.cu)
};

// "k=v,k=v" onto *o; returns how many entries were not understood
int ParseOptionList(Options *o, const std::string &list, bool warn);

class Runtime {
 public:
  static Runtime *Get();          // aborts if PSInit has not run
  static Runtime *GetOrNull();
  static void Create(int *argc, char ***argv);
  static void Destroy();

  int device = 0;
  int sm_count = 0;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;       // compute + ordered copies
  cudaStream_t copy_stream = nullptr;  // second DMA queue for pipelined staging
  GridSpace gs;
  Options opt;
  __PSB200Stats stats{};

  // kernel timing (opt.time_kernels): accumulated device ms + launches of the
  // family last run through __PSB200StencilRun
  double timed_ms = 0.0;
  uint64_t timed_launches = 0;

  // Host<->device transfers of logically contiguous bytes.  Pageable host
  // memory is pipelined through two pinned chunks so the CPU memcpy overlaps
  // the DMA; pinned/registered host memory goes straight to cudaMemcpyAsync.
  void CopyToDevice(void *dst, const void *src, size_t bytes);
  void CopyToHost(void *dst, const void *src, size_t bytes);
  DeviceBuffer &scratch(size_t bytes);        // device staging for AoS<->SoA transposes
  DeviceBuffer &small_scratch(size_t bytes);  // reduction partials

  cudaEvent_t timer_start = nullptr, timer_stop = nullptr;

  // ---- process group (one process per GPU; see comm.h, multigpu.cu) ----
  Comm *comm = nullptr;
  int domain_dims[PS_MAX_DIM] = {0, 0, 0};  // PSInit's maximum grid extents
  int world() const { return comm ? comm->world() : 1; }
  int rank() const { return comm ? comm->rank() : 0; }
  // neighbour-completion flags: flags[0] is written by the lower neighbour,
  // flags[1] by the upper one, with the number of the sweep they finished
  uint32_t *flags = nullptr;
  uint32_t *flags_of_lo = nullptr;   // IPC mappings of the neighbours' flag words
  uint32_t *flags_of_hi = nullptr;
  uint32_t sweep_epoch = 0;          // sweeps enqueued so far (identical on every rank)
  // a synchronous runtime call has written grids from the host since the last PSStencilRun:
  // the next run starts with a host barrier (every rank takes the same decision: SPMD)
  bool group_dirty = true;
  unsigned *done_counter = nullptr;  // finished-CTA counter of sweeps that signal themselves
  // host-mapped words a kernel writes when a neighbour's signal never arrives:
  // [0] = 1 + neighbour (0 lo, 1 hi), [1] = sweep waited for, [2] = last value seen, [3] = CTA
  volatile uint32_t *dev_err = nullptr;
  unsigned long long *halo_prof = nullptr;  // device counters of the halo-exchange profile (4 words)
  // aborts with a message if a kernel reported a lost neighbour signal (call after a stream sync)
  void CheckDeviceErrors(const char *where);
  // fills the device-side view of the flag words for a kernel that waits / signals itself;
  // false when the run is single-GPU or opt.sync_mode selects stream-ordered flags
  bool FillSlabSync(sweep::SlabSync *s);
  void InitGroup();
  void ShutdownGroup();
  // stream-ordered: wait until both neighbours have finished sweep `epoch`
  void WaitNeighbours(uint32_t epoch);
  // stream-ordered: tell both neighbours this rank has finished sweep `epoch`
  void SignalNeighbours(uint32_t epoch);
  // copies this rank's boundary planes of one member into the neighbours' halos
  void PushHalos(Grid &g, int member);
  void PushAllHalos(Grid &g);
  // maps `mine` (a cudaMalloc base) into the neighbours; returns their pointers
  void ExchangeIpc(void *mine, void **of_lo, void **of_hi);
  void CloseIpc(void *peer);

 private:
  Runtime() = default;
  ~Runtime();
  PinnedBuffer pinned_[2];
  cudaEvent_t pinned_free_[2] = {nullptr, nullptr};
  DeviceBuffer scratch_;
  DeviceBuffer small_scratch_;
};

// ---- kernels (defined in the .cu files) ---------------------------------

// Where a fused sweep must store its first / last interior plane of (g, member) so
// that it lands in the ring neighbours' halo planes: base pointers of those planes
// in the peers' IPC-mapped allocations.  False when the exchange is not fused
// (one GPU or opt.halo_push == 0); with wider halos it is the plane next to the interior.
bool SlabPushTargets(Runtime *rt, const Grid &g, int member, void **to_lo, void **to_hi,
                     size_t elem_size);

// AoS (host struct order) <-> SoA (one array per member) on the device.
void LaunchAosToSoa(const Grid &g, const void *aos_dev, cudaStream_t s);
void LaunchSoaToAos(const Grid &g, void *aos_dev, cudaStream_t s);

// PSReduce: whole-grid reduction of a primitive-type grid to one host scalar.
void ReduceGrid(Runtime *rt, const Grid &g, PSType type, PSReduceOp op, void *out_host);

// Specialised sweeps.  Each returns false if the descriptor does not satisfy the
// kernel's preconditions (then the caller reports a fatal error: there is no
// slower fallback for a specialised kind other than the program supplying a
// GENERIC launch stub).
struct SweepPlan;  // opaque per-descriptor prepared state
SweepPlan *PrepareSweep(Runtime *rt, const __PSB200StencilDesc &d);
void LaunchSweep(Runtime *rt, SweepPlan *plan);
void DestroySweep(SweepPlan *plan);
const char *SweepName(const SweepPlan *plan);
// prepared plans hold device addresses and option-dependent choices: dropped whenever a grid
// is freed or an option changes
void ClearPlanCache();
void ClearTuning();

// Launch of a sweep kernel that begins with `griddepcontrol.wait` (and ONLY of such a kernel):
// as a programmatic dependent of the kernel before it in the stream, so that its CTAs may run
// their prologue on SMs that kernel has already left (option pdl).
inline void LaunchSweepKernel(Runtime *rt, const void *fn, int grid, int block, void **args, size_t smem) {
  if (!rt->opt.pdl) {
    PSB_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(block), args, smem, rt->stream));
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = rt->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PSB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
}

}  // namespace physis_b200
