// b200 runtime: process-wide state, grid/buffer layer and the C entry points
// declared in include/physis/physis_b200.h.
//
// Reference behaviour matched here (files relative to the reference tree):
//   PSInit        runtime/libphysis_rt_cuda.cc:25-36 + runtime/runtime.h:20-30
//                 (consume --physis-trace), runtime/runtime_common.cc:14-35
//   __PSGridNew   runtime/libphysis_rt_cuda.cc:48-99   zero-filled device grid
//   Copyin/out    runtime/libphysis_rt_cuda.cc:114-137 synchronous, whole grid
//   __PSGridSet   runtime/libphysis_rt_cuda.cc:162-177 one element H2D
//   PSDomainNDNew runtime/libphysis_rt_cuda.cc:142-160
// Error handling: print + exit (runtime_common_cuda.h:16-27).
#include "runtime.h"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <string>

FILE *__ps_trace = nullptr;

namespace physis_b200 {

// ---------------------------------------------------------------- buffers

bool DeviceBuffer::Allocate(size_t bytes, cudaStream_t stream) {
  PSB_CHECK(ptr_ == nullptr, "DeviceBuffer::Allocate on a live buffer");
  if (bytes == 0) {
    size_ = capacity_ = 0;
    return true;
  }
  cudaError_t e = cudaMalloc(&ptr_, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear
    ptr_ = nullptr;
    return false;
  }
  PSB_CUDA(cudaMemsetAsync(ptr_, 0, bytes, stream));
  size_ = capacity_ = bytes;
  return true;
}

void DeviceBuffer::EnsureCapacity(size_t bytes, cudaStream_t stream) {
  if (bytes >= capacity_) {
    if (ptr_) {
      // the old block may still be in use by work enqueued on `stream`
      PSB_CUDA(cudaStreamSynchronize(stream));
      Free();
    }
    PSB_CHECK(Allocate(bytes, stream), "device out of memory");
  } else {
    size_ = bytes;
  }
}

void DeviceBuffer::Free() {
  if (ptr_) cudaFree(ptr_);
  ptr_ = nullptr;
  size_ = capacity_ = 0;
}

void PinnedBuffer::EnsureCapacity(size_t bytes) {
  if (bytes <= capacity_) return;
  Free();
  PSB_CUDA(cudaHostAlloc(&ptr_, bytes, cudaHostAllocDefault));
  capacity_ = bytes;
}

void PinnedBuffer::Free() {
  if (ptr_) cudaFreeHost(ptr_);
  ptr_ = nullptr;
  capacity_ = 0;
}

// ------------------------------------------------------------------ grids

static int ScalarSize(PSType t) {
  switch (t) {
    case PS_INT: return (int)sizeof(int);
    case PS_LONG: return (int)sizeof(long);
    case PS_FLOAT: return (int)sizeof(float);
    case PS_DOUBLE: return (int)sizeof(double);
    default: return 0;
  }
}

Grid *GridSpace::Create(const __PSGridTypeInfo *ti, int num_dims, const int *dim,
                        cudaStream_t stream) {
  PSB_CHECK(num_dims >= 1 && num_dims <= PS_MAX_DIM, "unsupported grid dimensionality");
  Grid *g = new Grid();
  g->id = next_id_++;
  g->type = ti->type;
  g->num_dims = num_dims;
  g->elm_size = ti->size;
  g->num_elms = 1;
  for (int i = 0; i < num_dims; ++i) {
    g->dim[i] = dim[i];
    g->num_elms *= dim[i];
  }

  if (ti->type == PS_USER) {
    PSB_CHECK(ti->num_members > 0 && ti->members, "user type without member info");
    int off = 0;
    for (int m = 0; m < ti->num_members; ++m) {
      const __PSGridTypeMemberInfo &mi = ti->members[m];
      MemberLayout ml;
      ml.type = mi.type;
      ml.size = mi.size;
      ml.count = 1;
      for (int r = 0; r < mi.rank; ++r) ml.count *= mi.dim[r];
      // C struct layout: each member aligned to its scalar size
      off = (off + ml.size - 1) / ml.size * ml.size;
      ml.aos_offset = off;
      off += ml.size * ml.count;
      g->members.push_back(ml);
    }
    int max_align = 1;
    for (auto &ml : g->members) max_align = std::max(max_align, ml.size);
    off = (off + max_align - 1) / max_align * max_align;
    PSB_CHECK(off == ti->size, "user type layout does not match sizeof(struct)");
  } else {
    MemberLayout ml;
    ml.type = ti->type;
    ml.size = ti->size;
    PSB_CHECK(ml.size == ScalarSize(ti->type), "primitive type size mismatch");
    g->members.push_back(ml);
  }

  for (auto &ml : g->members) {
    DeviceBuffer *b = new DeviceBuffer();
    size_t bytes = (size_t)ml.size * ml.count * (size_t)g->num_elms;
    if (!b->Allocate(bytes, stream)) {
      delete b;
      for (auto *s : g->storage) delete s;
      delete g;
      return nullptr;  // INVALID_GRID on OOM, as libphysis_rt_cuda.cc:66
    }
    ml.dev = b->get();
    g->storage.push_back(b);
  }

  // by-value device view: int dim[nd] (padded to 8) + one pointer per member
  size_t ptr_off = ((size_t)num_dims * sizeof(int) + 7) / 8 * 8;
  size_t view_bytes = ptr_off + sizeof(void *) * g->members.size();
  g->dev_view = calloc(1, view_bytes);
  for (int i = 0; i < num_dims; ++i) ((int *)g->dev_view)[i] = dim[i];
  for (size_t m = 0; m < g->members.size(); ++m)
    ((void **)((char *)g->dev_view + ptr_off))[m] = g->members[m].dev;

  g->handle.p = g->members[0].dev;
  for (int i = 0; i < PS_MAX_DIM; ++i) g->handle.dim[i] = (i < num_dims) ? dim[i] : 0;
  g->handle.elm_size = g->elm_size;
  g->handle.num_dims = num_dims;
  g->handle.num_elms = g->num_elms;
  g->handle.dev = (__PSGrid_dev *)g->dev_view;
  grids_[g->id] = g;
  return g;
}

void GridSpace::Destroy(Grid *g) {
  grids_.erase(g->id);
  for (auto *s : g->storage) delete s;
  free(g->dev_view);
  delete g;
}

Grid *GridSpace::Find(int id) const {
  auto it = grids_.find(id);
  return it == grids_.end() ? nullptr : it->second;
}

GridSpace::~GridSpace() {
  while (!grids_.empty()) Destroy(grids_.begin()->second);
}

// ---------------------------------------------------------------- runtime

static Runtime *g_rt = nullptr;

Runtime *Runtime::GetOrNull() { return g_rt; }
Runtime *Runtime::Get() {
  PSB_CHECK(g_rt != nullptr, "Physis runtime used before PSInit");
  return g_rt;
}

// Removes `--name`/`-name` (+ nargs following values) from argv; returns true if
// it was present and leaves the values in `vals`.
static bool ConsumeOption(int *argc, char ***argv, const char *name, int nargs,
                          std::vector<std::string> *vals) {
  if (!argc || !argv || !*argv) return false;
  std::string l = std::string("--") + name, s = std::string("-") + name;
  for (int i = 0; i < *argc; ++i) {
    if (l != (*argv)[i] && s != (*argv)[i]) continue;
    int last = std::min(*argc, i + 1 + nargs);
    for (int j = i + 1; j < last; ++j) vals->push_back((*argv)[j]);
    int removed = last - i;
    for (int j = i; j + removed < *argc; ++j) (*argv)[j] = (*argv)[j + removed];
    *argc -= removed;
    return true;
  }
  return false;
}

void Runtime::Create(int *argc, char ***argv) {
  PSB_CHECK(g_rt == nullptr, "PSInit called twice");
  Runtime *rt = new Runtime();
  std::vector<std::string> v;
  __ps_trace = nullptr;
  if (ConsumeOption(argc, argv, "physis-trace", 0, &v)) __ps_trace = stderr;
  // accepted for command-line compatibility with the MPI targets
  // (runtime/runtime_common.cc:37-70, runtime_mpi_cuda.cc:35-46)
  v.clear();
  ConsumeOption(argc, argv, "physis-proc", 1, &v);
  v.clear();
  ConsumeOption(argc, argv, "physis-nlp", 1, &v);
  v.clear();
  int dev = 0;
  if (const char *lr = getenv("LOCAL_RANK")) dev = atoi(lr);
  if (ConsumeOption(argc, argv, "physis-device", 1, &v) && !v.empty()) dev = atoi(v[0].c_str());

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fprintf(stderr,
            "[physis-b200] no CUDA device: %s. The b200 target has no CPU fallback.\n",
            cudaGetErrorString(e));
    exit(1);
  }
  dev %= ndev;
  PSB_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  PSB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) {
    fprintf(stderr, "[physis-b200] device %d is sm_%d%d; this runtime is built for sm_100a only\n",
            dev, prop.major, prop.minor);
    exit(1);
  }
  rt->device = dev;
  rt->sm_count = prop.multiProcessorCount;
  rt->l2_bytes = (size_t)prop.l2CacheSize;
  PSB_CUDA(cudaStreamCreateWithFlags(&rt->stream, cudaStreamNonBlocking));
  PSB_CUDA(cudaStreamCreateWithFlags(&rt->copy_stream, cudaStreamNonBlocking));
  PSB_CUDA(cudaEventCreate(&rt->timer_start));
  PSB_CUDA(cudaEventCreate(&rt->timer_stop));
  for (int i = 0; i < 2; ++i)
    PSB_CUDA(cudaEventCreateWithFlags(&rt->pinned_free_[i], cudaEventDisableTiming));
  g_rt = rt;
}

Runtime::~Runtime() {
  if (stream) cudaStreamSynchronize(stream);
  for (int i = 0; i < 2; ++i)
    if (pinned_free_[i]) cudaEventDestroy(pinned_free_[i]);
  if (timer_start) cudaEventDestroy(timer_start);
  if (timer_stop) cudaEventDestroy(timer_stop);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (stream) cudaStreamDestroy(stream);
}

void Runtime::Destroy() {
  if (!g_rt) return;
  delete g_rt;
  g_rt = nullptr;
}

static bool IsPinned(const void *p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void Runtime::CopyToDevice(void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return;
  stats.h2d_bytes += bytes;
  if (IsPinned(src)) {
    PSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    PSB_CUDA(cudaStreamSynchronize(stream));
    return;
  }
  const size_t chunk = opt.stage_chunk;
  for (int i = 0; i < 2; ++i) pinned_[i].EnsureCapacity(std::min(chunk, bytes));
  int b = 0;
  for (size_t off = 0; off < bytes; off += chunk, b ^= 1) {
    size_t n = std::min(chunk, bytes - off);
    PSB_CUDA(cudaEventSynchronize(pinned_free_[b]));  // previous DMA out of this chunk done
    memcpy(pinned_[b].get(), (const char *)src + off, n);
    PSB_CUDA(cudaMemcpyAsync((char *)dst + off, pinned_[b].get(), n, cudaMemcpyHostToDevice,
                             stream));
    PSB_CUDA(cudaEventRecord(pinned_free_[b], stream));
  }
  PSB_CUDA(cudaStreamSynchronize(stream));
}

void Runtime::CopyToHost(void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return;
  stats.d2h_bytes += bytes;
  if (IsPinned(dst)) {
    PSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    PSB_CUDA(cudaStreamSynchronize(stream));
    return;
  }
  const size_t chunk = opt.stage_chunk;
  for (int i = 0; i < 2; ++i) pinned_[i].EnsureCapacity(std::min(chunk, bytes));
  // software pipeline: DMA of chunk k+1 overlaps the CPU memcpy of chunk k
  size_t nchunks = (bytes + chunk - 1) / chunk;
  auto issue = [&](size_t k) {
    size_t off = k * chunk, n = std::min(chunk, bytes - off);
    int b = (int)(k & 1);
    PSB_CUDA(cudaMemcpyAsync(pinned_[b].get(), (const char *)src + off, n,
                             cudaMemcpyDeviceToHost, stream));
    PSB_CUDA(cudaEventRecord(pinned_free_[b], stream));
  };
  issue(0);
  for (size_t k = 0; k < nchunks; ++k) {
    int b = (int)(k & 1);
    PSB_CUDA(cudaEventSynchronize(pinned_free_[b]));
    if (k + 1 < nchunks) issue(k + 1);
    size_t off = k * chunk, n = std::min(chunk, bytes - off);
    memcpy((char *)dst + off, pinned_[b].get(), n);
  }
}

DeviceBuffer &Runtime::scratch(size_t bytes) {
  scratch_.EnsureCapacity(bytes, stream);
  return scratch_;
}

DeviceBuffer &Runtime::small_scratch(size_t bytes) {
  small_scratch_.EnsureCapacity(bytes, stream);
  return small_scratch_;
}

static int ParseKV(Options *o, const std::string &kv) {
  size_t eq = kv.find('=');
  if (eq == std::string::npos) return -1;
  std::string k = kv.substr(0, eq);
  long val = atol(kv.c_str() + eq + 1);
  if (k == "star7_ty") o->star7_ty = (int)val;
  else if (k == "star7_ry") o->star7_ry = (int)val;
  else if (k == "star7_nbx") o->star7_nbx = (int)val;
  else if (k == "star7_stages") o->star7_stages = (int)val;
  else if (k == "star7_zc") o->star7_zc = (int)val;
  else if (k == "star7_occ") o->star7_occ = (int)val;
  else if (k == "star7_variant") o->star7_variant = (int)val;
  else if (k == "star7_l2hint") o->star7_l2hint = (int)val;
  else if (k == "star7_sthint") o->star7_sthint = (int)val;
  else if (k == "himeno_by") o->himeno_by = (int)val;
  else if (k == "himeno_zc") o->himeno_zc = (int)val;
  else if (k == "himeno_stages") o->himeno_stages = (int)val;
  else if (k == "himeno_occ") o->himeno_occ = (int)val;
  else if (k == "time_kernels") o->time_kernels = (int)val;
  else if (k == "stage_chunk_mb") o->stage_chunk = (size_t)val << 20;
  else return -1;
  return 0;
}

}  // namespace physis_b200

using namespace physis_b200;

// ------------------------------------------------------------ C entry points

extern "C" {

void PSInit(int *argc, char ***argv, int grid_num_dims, ...) {
  (void)grid_num_dims;  // max grid extents follow as varargs; not needed on one GPU
  Runtime::Create(argc, argv);
  if (const char *env = getenv("PHYSIS_B200_OPTIONS")) {
    std::string s(env);
    size_t pos = 0;
    while (pos < s.size()) {
      size_t c = s.find(',', pos);
      if (c == std::string::npos) c = s.size();
      if (ParseKV(&Runtime::Get()->opt, s.substr(pos, c - pos)) != 0)
        fprintf(stderr, "[physis-b200] ignoring unknown option '%s'\n",
                s.substr(pos, c - pos).c_str());
      pos = c + 1;
    }
  }
}

void PSFinalize(void) { Runtime::Destroy(); }

PSDomain1D PSDomain1DNew(PSIndex minx, PSIndex maxx) {
  PSDomain1D d = {{minx}, {maxx}, {minx}, {maxx}};
  return d;
}
PSDomain2D PSDomain2DNew(PSIndex minx, PSIndex maxx, PSIndex miny, PSIndex maxy) {
  PSDomain2D d = {{minx, miny}, {maxx, maxy}, {minx, miny}, {maxx, maxy}};
  return d;
}
PSDomain3D PSDomain3DNew(PSIndex minx, PSIndex maxx, PSIndex miny, PSIndex maxy,
                         PSIndex minz, PSIndex maxz) {
  PSDomain3D d = {{minx, miny, minz}, {maxx, maxy, maxz},
                  {minx, miny, minz}, {maxx, maxy, maxz}};
  return d;
}

int __PSGridGetID(__PSGrid *g) { return Grid::FromHandle(g)->id; }

__PSGrid *__PSGridNew(__PSGridTypeInfo *type_info, int num_dims, PSVectorInt dim,
                      __PSGrid_devNewFunc func) {
  Runtime *rt = Runtime::Get();
  if (func) {
    // `--cuda`-style translation with its own device-struct allocator
    // (libphysis_rt_cuda.cc:60-62): the runtime only keeps the handle.
    Grid *g = new Grid();
    g->external_dev = true;
    g->type = type_info->type;
    g->num_dims = num_dims;
    g->elm_size = type_info->size;
    g->num_elms = 1;
    for (int i = 0; i < num_dims; ++i) {
      g->dim[i] = dim[i];
      g->num_elms *= dim[i];
      g->handle.dim[i] = dim[i];
    }
    g->handle.p = nullptr;
    g->handle.elm_size = g->elm_size;
    g->handle.num_dims = num_dims;
    g->handle.num_elms = g->num_elms;
    g->handle.dev = (__PSGrid_dev *)func(num_dims, dim);
    return &g->handle;
  }
  Grid *g = rt->gs.Create(type_info, num_dims, dim, rt->stream);
  if (!g) return INVALID_GRID;
  return &g->handle;
}

void __PSGridFree(void *gv, __PSGrid_devFreeFunc func) {
  if (!gv) return;
  Runtime *rt = Runtime::Get();
  Grid *g = Grid::FromHandle(gv);
  PSB_CUDA(cudaStreamSynchronize(rt->stream));
  if (g->external_dev) {
    if (func && g->handle.dev) func(g->handle.dev);
    delete g;
    return;
  }
  rt->gs.Destroy(g);
}

void __PSGridCopyin(void *gv, const void *src, __PSGrid_devCopyinFunc func) {
  Runtime *rt = Runtime::Get();
  Grid *g = Grid::FromHandle(gv);
  if (func) {
    func(g->handle.dev, src, (size_t)g->num_elms);
    return;
  }
  PSB_CHECK(!g->external_dev, "copyin of an externally allocated grid needs its helper");
  if (!g->is_user_type()) {
    rt->CopyToDevice(g->members[0].dev, src, g->bytes());
    return;
  }
  // user type: stage the AoS bytes on the device, transpose there
  DeviceBuffer &tmp = rt->scratch(g->bytes());
  rt->CopyToDevice(tmp.get(), src, g->bytes());
  LaunchAosToSoa(*g, tmp.get(), rt->stream);
  rt->stats.kernel_launches++;
  PSB_CUDA(cudaStreamSynchronize(rt->stream));
}

void __PSGridCopyout(void *gv, void *dst, __PSGrid_devCopyoutFunc func) {
  Runtime *rt = Runtime::Get();
  Grid *g = Grid::FromHandle(gv);
  if (func) {
    func(g->handle.dev, dst, (size_t)g->num_elms);
    return;
  }
  PSB_CHECK(!g->external_dev, "copyout of an externally allocated grid needs its helper");
  if (!g->is_user_type()) {
    rt->CopyToHost(dst, g->members[0].dev, g->bytes());
    return;
  }
  DeviceBuffer &tmp = rt->scratch(g->bytes());
  LaunchSoaToAos(*g, tmp.get(), rt->stream);
  rt->stats.kernel_launches++;
  rt->CopyToHost(dst, tmp.get(), g->bytes());
}

void PSGridCopyin(void *g, const void *src) { __PSGridCopyin(g, src, nullptr); }
void PSGridCopyout(void *g, void *dst) { __PSGridCopyout(g, dst, nullptr); }
void PSGridFree(void *g) { __PSGridFree(g, nullptr); }
void __PSGridSwap(__PSGrid *g) { (void)g; }

void __PSGridSet(__PSGrid *gh, void *buf, ...) {
  Runtime *rt = Runtime::Get();
  Grid *g = Grid::FromHandle(gh);
  va_list vl;
  va_start(vl, buf);
  int64_t offset = 0, base = 1;
  for (int i = 0; i < g->num_dims; ++i) {
    PSIndex idx = va_arg(vl, PSIndex);
    offset += idx * base;
    base *= g->dim[i];
  }
  va_end(vl);
  // one element; for user types scatter each member of the struct
  for (auto &ml : g->members) {
    for (int c = 0; c < ml.count; ++c) {
      char *d = (char *)ml.dev + ((size_t)c * g->num_elms + offset) * ml.size;
      const char *s = (const char *)buf + ml.aos_offset + (size_t)c * ml.size;
      PSB_CUDA(cudaMemcpyAsync(d, s, ml.size, cudaMemcpyHostToDevice, rt->stream));
    }
  }
  PSB_CUDA(cudaStreamSynchronize(rt->stream));
  rt->stats.h2d_bytes += g->elm_size;
}

void __PSCheckCudaError(const char *message) {
  cudaError_t error = cudaGetLastError();
  if (error != cudaSuccess) {
    fprintf(stderr, "ERROR: %s: %s\n", message, cudaGetErrorString(error));
    PSAbort(1);
  }
}

static void ReduceEntry(void *buf, enum PSReduceOp op, __PSGrid *gh, PSType t) {
  Runtime *rt = Runtime::Get();
  Grid *g = Grid::FromHandle(gh);
  PSB_CHECK(!g->is_user_type(), "PSReduce is not defined on user-defined point types");
  PSB_CHECK(g->type == t, "PSReduce entry point does not match the grid's element type");
  ReduceGrid(rt, *g, t, op, buf);
}
void __PSReduceGridFloat(void *buf, enum PSReduceOp op, __PSGrid *g) {
  ReduceEntry(buf, op, g, PS_FLOAT);
}
void __PSReduceGridDouble(void *buf, enum PSReduceOp op, __PSGrid *g) {
  ReduceEntry(buf, op, g, PS_DOUBLE);
}
void __PSReduceGridInt(void *buf, enum PSReduceOp op, __PSGrid *g) {
  ReduceEntry(buf, op, g, PS_INT);
}
void __PSReduceGridLong(void *buf, enum PSReduceOp op, __PSGrid *g) {
  ReduceEntry(buf, op, g, PS_LONG);
}

__PSB200Stream __PSB200GetStream(void) { return (__PSB200Stream)Runtime::Get()->stream; }
void __PSB200Synchronize(void) { PSB_CUDA(cudaStreamSynchronize(Runtime::Get()->stream)); }

void __PSB200TimerStart(void) {
  Runtime *rt = Runtime::Get();
  PSB_CUDA(cudaStreamSynchronize(rt->stream));
  PSB_CUDA(cudaEventRecord(rt->timer_start, rt->stream));
}
float __PSB200TimerStopMs(void) {
  Runtime *rt = Runtime::Get();
  PSB_CUDA(cudaEventRecord(rt->timer_stop, rt->stream));
  PSB_CUDA(cudaEventSynchronize(rt->timer_stop));
  float ms = 0.f;
  PSB_CUDA(cudaEventElapsedTime(&ms, rt->timer_start, rt->timer_stop));
  return ms;
}

void __PSB200GetStats(__PSB200Stats *out) {
  Runtime *rt = Runtime::Get();
  *out = rt->stats;
  out->last_kernel_ms = rt->timed_launches ? (float)(rt->timed_ms / rt->timed_launches) : 0.f;
}
void __PSB200ResetStats(void) {
  Runtime *rt = Runtime::Get();
  rt->stats = __PSB200Stats{};
  rt->timed_ms = 0;
  rt->timed_launches = 0;
}
int __PSB200SetOption(const char *kv) { return ParseKV(&Runtime::Get()->opt, kv); }
const char *__PSB200Version(void) { return "physis-b200 0.1 (sm_100a)"; }

void *__PSB200HostAlloc(size_t bytes) {
  void *p = nullptr;
  PSB_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
  return p;
}
void __PSB200HostFree(void *p) { cudaFreeHost(p); }

}  // extern "C"
