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

constexpr int kBarrierBytes = 128;  // full[] + empty[], up to 8 stages
constexpr int kMaxStages = 8;

template <typename T, int TY>
__host__ __device__ constexpr int BoxStride() {
  return ((TY + 2) * Geom<T>::ROW_BYTES + 127) / 128 * 128;
}

}  // namespace sweep
}  // namespace physis_b200
