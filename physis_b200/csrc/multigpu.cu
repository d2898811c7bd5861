// Multi-GPU layer of the b200 runtime: z-slab halo exchange over NVLink.
//
// Replaces, for the `b200` target, the reference's halo path
//   runtime/grid_mpi_cuda_exp.{h,cc}   CopyoutHalo / CopyinHalo (D2H/H2D + pack kernels)
//   runtime/grid_space_mpi_cuda.h:254-407  SendBoundaries / RecvBoundaries
//     = cudaMemcpy D2H -> MPI_Isend / MPI_Recv -> cudaMemcpy H2D, per grid x direction,
//       serialised with the sweep unless MPI_OVERLAP
// and its per-iteration cudaDeviceSynchronize (mpi_cuda_runtime_builder.cc:334-337,481-483).
//
// B200 design (one process per GPU, all on one NVSwitch box):
//  * every decomposed grid member is one cudaMalloc whose CUDA IPC handle is opened
//    by the two ring neighbours, so a neighbour's halo planes are ordinary device
//    addresses here: z-slab halos are whole contiguous planes, there is nothing to
//    pack, and nothing is staged through the host;
//  * halo planes are WRITTEN BY THE PRODUCER: the specialised sweeps store their
//    first/last interior plane a second time through the peer mapping (st.global
//    over NVLink, fused into the sweep, see star7.cu / himeno.cu / pstag.cu); other
//    writers use PushHalos (peer cudaMemcpyAsync of the same planes);
//  * ordering between neighbours is two 32-bit flag words per rank in device memory,
//    written by the neighbours with stream memory operations after their sweep n and
//    waited on, in stream order, before sweep n+1 — the host never synchronises
//    inside PSStencilRun, and a rank only ever waits for its two neighbours;
//  * z periodicity is the ring wrap of the same exchange (rank 0 <-> rank P-1), as in
//    the reference (runtime/grid_space_mpi.h:269-281).
#include "runtime.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace physis_b200 {

namespace {

using MemOpFn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
MemOpFn g_wait32 = nullptr, g_write32 = nullptr;

void ResolveMemOps() {
  if (g_wait32) return;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_wait32 = reinterpret_cast<MemOpFn>(p);
  p = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_write32 = reinterpret_cast<MemOpFn>(p);
}

// Fallback when stream memory operations are unavailable: one-thread kernels.
__global__ void WaitFlagsKernel(const uint32_t *flags, uint32_t epoch, unsigned long long timeout_ns,
                                uint32_t *err) {
  sweep::SlabWaitFlags(flags, epoch, timeout_ns, err);
  __threadfence_system();
}
__global__ void SignalFlagsKernel(uint32_t *to_lo, uint32_t *to_hi, uint32_t epoch) {
  __threadfence_system();
  *reinterpret_cast<volatile uint32_t *>(to_lo) = epoch;
  *reinterpret_cast<volatile uint32_t *>(to_hi) = epoch;
}

}  // namespace

void Runtime::InitGroup() {
  comm = Comm::Create();
  if (world() == 1) return;
  void *f = nullptr;
  PSB_CUDA(cudaMalloc(&f, 256));
  PSB_CUDA(cudaMemset(f, 0, 256));
  PSB_CUDA(cudaDeviceSynchronize());
  flags = static_cast<uint32_t *>(f);
  void *lo = nullptr, *hi = nullptr;
  ExchangeIpc(flags, &lo, &hi);
  flags_of_lo = static_cast<uint32_t *>(lo);
  flags_of_hi = static_cast<uint32_t *>(hi);
  ResolveMemOps();
  if ((!g_wait32 || !g_write32) && opt.sync_mode == 0) opt.sync_mode = 1;
  sweep_epoch = 0;
  done_counter = reinterpret_cast<unsigned *>(flags + 16);  // same zeroed allocation
  halo_prof = reinterpret_cast<unsigned long long *>(flags + 32);  // same zeroed allocation
  void *e = nullptr;
  PSB_CUDA(cudaHostAlloc(&e, 64, cudaHostAllocMapped));
  memset(e, 0, 64);
  dev_err = static_cast<volatile uint32_t *>(e);
}

bool Runtime::FillSlabSync(sweep::SlabSync *s) {
  *s = sweep::SlabSync{};
  if (world() == 1 || opt.sync_mode != 2) return false;
  s->flags = flags;
  s->to_lo = flags_of_lo + 1;  // this rank is the upper neighbour of `lo`
  s->to_hi = flags_of_hi + 0;
  s->done = done_counter;
  s->timeout_ns = (unsigned long long)std::max(1, opt.sync_timeout_s) * 1000000000ull;
  s->err = const_cast<uint32_t *>(dev_err);
  s->prof = opt.halo_profile ? halo_prof : nullptr;
  return true;
}

void Runtime::ShutdownGroup() {
  if (comm && world() > 1) {
    comm->Barrier();
    CloseIpc(flags_of_lo);
    if (flags_of_hi != flags_of_lo) CloseIpc(flags_of_hi);
    if (flags) cudaFree(flags);
    flags = flags_of_lo = flags_of_hi = nullptr;
    if (dev_err) cudaFreeHost(const_cast<uint32_t *>(dev_err));
    dev_err = nullptr;
  }
  delete comm;
  comm = nullptr;
}

void Runtime::CheckDeviceErrors(const char *where) {
  if (!dev_err || dev_err[0] == 0) return;
  fprintf(stderr,
          "[physis-b200] rank %d (%s): the %s ring neighbour never published sweep %u (last seen %u, "
          "CTA %u, waited %d s; option sync_timeout_s). Results are invalid.\n",
          rank(), where, dev_err[0] == 1 ? "lower" : "upper", dev_err[1], dev_err[2], dev_err[3],
          opt.sync_timeout_s);
  exit(1);
}

void Runtime::ExchangeIpc(void *mine, void **of_lo, void **of_hi) {
  const int W = world();
  cudaIpcMemHandle_t h;
  PSB_CUDA(cudaIpcGetMemHandle(&h, mine));
  std::vector<cudaIpcMemHandle_t> all(W);
  comm->AllGather(&h, all.data(), sizeof h);
  const int lo = comm->lo(), hi = comm->hi();
  PSB_CUDA(cudaIpcOpenMemHandle(of_lo, all[lo], cudaIpcMemLazyEnablePeerAccess));
  if (hi == lo)
    *of_hi = *of_lo;  // two ranks: both neighbours are the same process, one mapping
  else
    PSB_CUDA(cudaIpcOpenMemHandle(of_hi, all[hi], cudaIpcMemLazyEnablePeerAccess));
}

void Runtime::CloseIpc(void *peer) {
  if (peer) cudaIpcCloseMemHandle(peer);
}

void Runtime::WaitNeighbours(uint32_t epoch) {
  if (opt.sync_mode == 0 && g_wait32) {
    // CU_STREAM_WAIT_VALUE_GEQ compares cyclically, so the counter may wrap
    for (int i = 0; i < 2; ++i) {
      CUresult r = g_wait32((CUstream)stream, (CUdeviceptr)(flags + i), epoch, CU_STREAM_WAIT_VALUE_GEQ);
      PSB_CHECK(r == CUDA_SUCCESS, "cuStreamWaitValue32 failed");
    }
  } else {
    WaitFlagsKernel<<<1, 1, 0, stream>>>(flags, epoch,
                                         (unsigned long long)std::max(1, opt.sync_timeout_s) * 1000000000ull,
                                         const_cast<uint32_t *>(dev_err));
    stats.kernel_launches++;
  }
}

void Runtime::SignalNeighbours(uint32_t epoch) {
  // this rank is the UPPER neighbour of `lo` (their flags[1]) and the LOWER one of `hi`
  if (opt.sync_mode == 0 && g_write32) {
    // default flags: all earlier writes of this stream (the sweep's peer stores or the
    // peer copies) are visible before the value lands
    CUresult r = g_write32((CUstream)stream, (CUdeviceptr)(flags_of_lo + 1), epoch, CU_STREAM_WRITE_VALUE_DEFAULT);
    PSB_CHECK(r == CUDA_SUCCESS, "cuStreamWriteValue32 failed");
    r = g_write32((CUstream)stream, (CUdeviceptr)(flags_of_hi + 0), epoch, CU_STREAM_WRITE_VALUE_DEFAULT);
    PSB_CHECK(r == CUDA_SUCCESS, "cuStreamWriteValue32 failed");
  } else {
    SignalFlagsKernel<<<1, 1, 0, stream>>>(flags_of_lo + 1, flags_of_hi + 0, epoch);
    stats.kernel_launches++;
  }
}

void Runtime::PushHalos(Grid &g, int member) {
  if (!g.decomposed || g.halo == 0) return;
  MemberLayout &ml = g.members[member];
  const size_t plane = (size_t)g.plane_elms * ml.size;
  const size_t w = (size_t)g.halo * plane;
  // peers' allocation sizes: their interior thickness may differ by one plane
  int hi_off, hi_nz;
  PartitionGridZ(g.dim[g.num_dims - 1], domain_dims[g.num_dims - 1], world(), comm->hi(), &hi_off, &hi_nz);
  const size_t lo_alloc = (size_t)(g.lo_nz_loc + 2 * g.halo) * plane;
  const size_t hi_alloc = (size_t)(hi_nz + 2 * g.halo) * plane;
  const size_t my_alloc = (size_t)g.n_alloc * ml.size;
  for (int c = 0; c < ml.count; ++c) {
    const char *base = (const char *)ml.dev + c * my_alloc;
    // first interior planes -> lower neighbour's upper halo
    PSB_CUDA(cudaMemcpyAsync((char *)ml.peer_lo + c * lo_alloc + (size_t)(g.halo + g.lo_nz_loc) * plane,
                             base + (size_t)g.halo * plane, w, cudaMemcpyDeviceToDevice, stream));
    // last interior planes -> upper neighbour's lower halo
    PSB_CUDA(cudaMemcpyAsync((char *)ml.peer_hi + c * hi_alloc,
                             base + (size_t)g.nz_loc * plane, w, cudaMemcpyDeviceToDevice, stream));
    stats.halo_bytes += 2 * w;
  }
}

bool SlabPushTargets(Runtime *rt, const Grid &g, int member, void **to_lo, void **to_hi,
                     size_t elem_size) {
  *to_lo = *to_hi = nullptr;
  if (!g.decomposed || g.halo < 1 || !rt->opt.halo_push) return false;
  const MemberLayout &ml = g.members[member];
  if (ml.count != 1 || (size_t)ml.size != elem_size) return false;
  const size_t plane = (size_t)g.plane_elms * ml.size;
  *to_lo = (char *)ml.peer_lo + (size_t)(g.halo + g.lo_nz_loc) * plane;  // their upper halo
  *to_hi = (char *)ml.peer_hi + (size_t)(g.halo - 1) * plane;           // their lower halo
  return true;
}

void Runtime::PushAllHalos(Grid &g) {
  for (size_t m = 0; m < g.members.size(); ++m) PushHalos(g, (int)m);
}

}  // namespace physis_b200
