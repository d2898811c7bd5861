// Shared-memory rendezvous of the per-GPU processes (see comm.h).
#include "comm.h"

#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace physis_b200 {

void PartitionZ(int n, int world, int rank, int *offset, int *length) {
  const int base = n / world, rem = n % world;
  // ranks [world-rem, world) get base+1
  const int first_big = world - rem;
  const int big_before = rank > first_big ? rank - first_big : 0;
  *offset = rank * base + big_before;
  *length = base + (rank >= first_big ? 1 : 0);
}

void PartitionGridZ(int n, int domain_n, int world, int rank, int *offset, int *length) {
  if (domain_n <= 0 || domain_n == n) {
    PartitionZ(n, world, rank, offset, length);
    return;
  }
  int off, len;
  PartitionZ(domain_n, world, rank, &off, &len);
  int end = off + len;
  if (rank == world - 1) end = n;  // the last rank absorbs the difference
  if (off > n) off = n;
  if (end > n) end = n;
  *offset = off;
  *length = end - off;
}

namespace {
constexpr uint32_t kMagic = 0x50534232u;  // "PSB2"
constexpr double kTimeoutSec = 180.0;

double Now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}
int64_t WallSeconds() { return (int64_t)time(nullptr); }

[[noreturn]] void Fail(const char *what, const std::string &name) {
  std::fprintf(stderr, "[physis-b200] FATAL process-group rendezvous: %s (%s, errno %d: %s)\n", what,
               name.c_str(), errno, std::strerror(errno));
  std::exit(1);
}

void Relax(int &spins) {
  if (++spins < 2000) {
    sched_yield();
  } else {
    timespec ts{0, 50000};
    nanosleep(&ts, nullptr);
  }
}
}  // namespace

struct Comm::Shm {
  std::atomic<uint32_t> magic;
  uint32_t world;
  int64_t created;                        // wall-clock seconds, staleness check
  std::atomic<uint32_t> attached;
  alignas(64) std::atomic<uint32_t> barrier_count;
  alignas(64) unsigned char slots[1];     // world * kSlotBytes, then one window
};

Comm *Comm::Create() {
  Comm *c = new Comm();
  const char *ws = getenv("WORLD_SIZE");
  c->world_ = ws ? atoi(ws) : 1;
  if (c->world_ < 1) c->world_ = 1;
  const char *rk = getenv("RANK");
  c->rank_ = rk ? atoi(rk) : 0;
  const char *lr = getenv("LOCAL_RANK");
  c->local_rank_ = lr ? atoi(lr) : c->rank_;
  if (c->world_ == 1) {
    c->rank_ = 0;
    return c;
  }
  if (c->rank_ < 0 || c->rank_ >= c->world_) Fail("RANK outside [0, WORLD_SIZE)", "env");

  if (const char *s = getenv("PHYSIS_B200_SESSION")) {
    c->name_ = std::string("/physis_b200_") + s;
  } else {
    const char *port = getenv("MASTER_PORT");
    char buf[128];
    std::snprintf(buf, sizeof buf, "/physis_b200_%u_%s_%d", (unsigned)getuid(), port ? port : "0",
                  (int)getppid());
    c->name_ = buf;
  }
  const size_t bytes = offsetof(Shm, slots) + (size_t)c->world_ * kSlotBytes + kWindowBytes;
  c->shm_bytes_ = bytes;
  const double t0 = Now();
  int fd = -1;
  if (c->rank_ == 0) {
    shm_unlink(c->name_.c_str());  // a leftover of a crashed run
    fd = shm_open(c->name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) Fail("shm_open(create)", c->name_);
    if (ftruncate(fd, (off_t)bytes) != 0) Fail("ftruncate", c->name_);
  } else {
    int spins = 0;
    for (;;) {
      fd = shm_open(c->name_.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && (size_t)st.st_size == bytes) break;
        close(fd);
        fd = -1;
      }
      if (Now() - t0 > kTimeoutSec) Fail("timed out waiting for rank 0's segment", c->name_);
      Relax(spins);
    }
  }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) Fail("mmap", c->name_);
  c->shm_ = static_cast<Shm *>(p);
  if (c->rank_ == 0) {
    c->shm_->world = (uint32_t)c->world_;
    c->shm_->created = WallSeconds();
    c->shm_->attached.store(1);
    c->shm_->barrier_count.store(0);
    c->shm_->magic.store(kMagic, std::memory_order_release);
  } else {
    int spins = 0;
    for (;;) {
      if (c->shm_->magic.load(std::memory_order_acquire) == kMagic &&
          c->shm_->world == (uint32_t)c->world_ && WallSeconds() - c->shm_->created < 600)
        break;
      if (Now() - t0 > kTimeoutSec) Fail("segment never became valid (stale leftover?)", c->name_);
      Relax(spins);
    }
    c->shm_->attached.fetch_add(1);
  }
  c->Barrier();
  // every rank is attached: the name can go away now (the mapping stays), so a
  // crash later leaves nothing behind
  if (c->rank_ == 0) shm_unlink(c->name_.c_str());
  return c;
}

Comm::~Comm() {
  if (shm_) {
    if (world_ > 1) Barrier();
    if (rank_ == 0) shm_->magic.store(0);
    munmap(shm_, shm_bytes_);
  }
}

void Comm::Barrier() {
  if (world_ == 1) return;
  ++epoch_;
  const uint32_t target = epoch_ * (uint32_t)world_;
  shm_->barrier_count.fetch_add(1, std::memory_order_acq_rel);
  const double t0 = Now();
  int spins = 0;
  while ((int32_t)(shm_->barrier_count.load(std::memory_order_acquire) - target) < 0) {
    if ((spins & 1023) == 1023 && Now() - t0 > kTimeoutSec) Fail("barrier timed out (a rank died?)", name_);
    Relax(spins);
  }
}

void Comm::AllGather(const void *mine, void *out, size_t bytes) {
  if (world_ == 1) {
    std::memcpy(out, mine, bytes);
    return;
  }
  for (size_t done = 0; done < bytes; done += kSlotBytes) {
    const size_t n = bytes - done < kSlotBytes ? bytes - done : kSlotBytes;
    std::memcpy(shm_->slots + (size_t)rank_ * kSlotBytes, (const char *)mine + done, n);
    Barrier();
    for (int r = 0; r < world_; ++r)
      std::memcpy((char *)out + (size_t)r * bytes + done, shm_->slots + (size_t)r * kSlotBytes, n);
    Barrier();
  }
}

void Comm::AllGatherV(void *buf, const size_t *offsets, const size_t *lengths) {
  if (world_ == 1) return;
  unsigned char *window = shm_->slots + (size_t)world_ * kSlotBytes;
  for (int s = 0; s < world_; ++s) {
    for (size_t done = 0; done < lengths[s]; done += kWindowBytes) {
      const size_t n = lengths[s] - done < kWindowBytes ? lengths[s] - done : kWindowBytes;
      char *seg = (char *)buf + offsets[s] + done;
      if (rank_ == s) std::memcpy(window, seg, n);
      Barrier();
      if (rank_ != s) std::memcpy(seg, window, n);
      Barrier();
    }
  }
}

}  // namespace physis_b200
