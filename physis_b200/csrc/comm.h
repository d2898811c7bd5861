// Process group of the b200 runtime: one process per GPU on ONE box.
//
// Replaces, for the `b200` target, the reference's inter-process layer
//   runtime/ipc.h, ipc_mpi.{h,cc}, mpi_wrapper.{h,cc}   (MPI two-sided, host staged)
//   runtime/rpc.h master/worker Request broadcast        (rank 0 runs main, others serve)
// with an SPMD model: every rank runs the same Physis program, and the only
// host-side communication is a tiny rendezvous (barrier + all-gather of a few
// bytes: CUDA IPC handles, reduction partials) through a POSIX shared-memory
// segment.  Bulk data never goes through here: halo planes move GPU-to-GPU over
// NVLink (peer stores / peer copies on IPC-mapped memory, see multigpu.cu).
//
// Ranks come from the launcher's environment (torchrun / torch.distributed.run
// conventions: RANK, WORLD_SIZE, LOCAL_RANK, MASTER_PORT).  This file is pure
// host code with no CUDA dependency so that the decomposition and rendezvous
// logic is testable on a CPU-only machine.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace physis_b200 {

// Block decomposition of `n` planes over `world` ranks: floor(n/world) each, the
// remainder spread one plane each over the LAST ranks (the reference's
// GridSpaceMPI::Partition, runtime/grid_space_mpi.h:741-777).
void PartitionZ(int n, int world, int rank, int *offset, int *length);

// Offsets of a grid whose extent `n` differs from the decomposed domain extent
// `domain_n` (staggered grids, N+1 vs N): interior cuts stay where the domain's
// are, the last rank takes whatever remains.  domain_n <= 0 means "use n".
void PartitionGridZ(int n, int domain_n, int world, int rank, int *offset, int *length);

class Comm {
 public:
  // Builds the group from the environment; world == 1 when WORLD_SIZE is unset.
  // Aborts (print + exit, like every runtime error) if the rendezvous fails.
  static Comm *Create();
  ~Comm();

  int rank() const { return rank_; }
  int world() const { return world_; }
  int local_rank() const { return local_rank_; }
  int lo() const { return (rank_ + world_ - 1) % world_; }  // ring neighbours in z
  int hi() const { return (rank_ + 1) % world_; }

  void Barrier();
  // Every rank contributes `bytes` (<= kSlotBytes per call chunk, any size overall);
  // `out` receives world * bytes, rank-major.
  void AllGather(const void *mine, void *out, size_t bytes);
  // Rank-ordered concatenation of variable-length host segments: rank r owns
  // [offsets[r], offsets[r]+lengths[r]) of `buf` (already filled for r == rank());
  // afterwards every rank holds all segments.  Goes through bounded shm windows.
  void AllGatherV(void *buf, const size_t *offsets, const size_t *lengths);

  static constexpr size_t kSlotBytes = 4096;
  static constexpr size_t kWindowBytes = 8u << 20;

 private:
  Comm() = default;
  struct Shm;
  Shm *shm_ = nullptr;
  size_t shm_bytes_ = 0;
  std::string name_;
  int rank_ = 0, world_ = 1, local_rank_ = 0;
  uint32_t epoch_ = 0;
};

}  // namespace physis_b200
