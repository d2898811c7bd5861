// Host-side TMA descriptor encoding.  cuTensorMapEncodeTiled is a driver API
// entry; it is resolved at run time through the CUDA runtime so that the
// library has no link-time dependency on libcuda.so (absent on build boxes).
#include "tma.cuh"
#include "common.h"

#include <mutex>

namespace physis_b200 {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                              const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                              const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn ResolveEncode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      Die("cuTensorMapEncodeTiled not available from the driver", __FILE__, __LINE__);
    fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}
}  // namespace

bool EncodeTensorMap3D(CUtensorMap *out, TmaElem elem, const void *base, const int dim[3],
                       const int box[3]) {
  const size_t es = elem == TmaElem::F64 ? 8 : 4;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  if (((size_t)dim[0] * es) % 16 != 0) return false;
  for (int i = 0; i < 3; ++i)
    if (box[i] < 1 || box[i] > 256) return false;
  if (((size_t)box[0] * es) % 16 != 0) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2]};
  cuuint64_t gstride[2] = {(cuuint64_t)dim[0] * es, (cuuint64_t)dim[0] * dim[1] * es};
  cuuint32_t bdim[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2]};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ResolveEncode()(
      out, elem == TmaElem::F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
      3, const_cast<void *>(base), gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool EncodeTensorMap2D(CUtensorMap *out, TmaElem elem, const void *base, const int dim[2],
                       const int box[2]) {
  const size_t es = elem == TmaElem::F64 ? 8 : 4;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  if (((size_t)dim[0] * es) % 16 != 0) return false;
  for (int i = 0; i < 2; ++i)
    if (box[i] < 1 || box[i] > 256 || dim[i] < 1) return false;
  if (((size_t)box[0] * es) % 16 != 0) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)dim[0], (cuuint64_t)dim[1]};
  cuuint64_t gstride[1] = {(cuuint64_t)dim[0] * es};
  cuuint32_t bdim[2] = {(cuuint32_t)box[0], (cuuint32_t)box[1]};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ResolveEncode()(
      out, elem == TmaElem::F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
      2, const_cast<void *>(base), gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace physis_b200
