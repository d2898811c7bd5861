// Shared host-side helpers of the b200 runtime.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

namespace physis_b200 {

// Error policy follows the reference runtimes (runtime/runtime_common_cuda.h:
// 16-27, include/physis/physis_common.h:144-146): no error codes, print and
// exit.  There is deliberately no CPU fallback anywhere.
[[noreturn]] inline void Die(const char *what, const char *file, int line) {
  std::fprintf(stderr, "[physis-b200] FATAL %s (%s:%d)\n", what, file, line);
  std::exit(1);
}

#define PSB_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      std::fprintf(stderr, "[physis-b200] CUDA error %s: %s\n", #call,          \
                   cudaGetErrorString(e__));                                    \
      ::physis_b200::Die("CUDA call failed", __FILE__, __LINE__);               \
    }                                                                           \
  } while (0)

#define PSB_CHECK(cond, msg)                                                    \
  do {                                                                          \
    if (!(cond)) ::physis_b200::Die(msg, __FILE__, __LINE__);                   \
  } while (0)

inline int CeilDiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace physis_b200
