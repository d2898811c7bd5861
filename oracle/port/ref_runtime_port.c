/*
 * ORACLE (test infrastructure only — nothing under oracle/ is on the product
 * path; the product is the CUDA library built from physis_b200/csrc).
 *
 * Plain-C restatement of the Physis REFERENCE-target runtime, following
 * /root/reference/runtime/libphysis_rt_ref.cc:
 *   PSReduceGridTemplate  :19-30   sequential left fold seeded with d[0], in T
 *   PSInit / PSFinalize   :38-46   (+ runtime/runtime.h:20-30: only strips
 *                                   -physis-trace / --physis-trace from argv)
 *   __PSGridGetID         :48-51   always 0
 *   __PSGridNew           :53-71   calloc'd, num_elms = prod(dim[0..nd))
 *   PSGridFree            :73-79   frees g->p, keeps the handle
 *   PSGridCopyin/Copyout  :81-89   memcpy of elm_size*num_elms
 *   PSDomainNDNew         :91-109  local_{min,max} = {min,max}
 *   __PSGridSet           :111-125 offset = sum idx_i * prod dim_{<i}
 *   __PSReduceGrid{Float,Double,Int,Long} :128-146
 * Reducer semantics follow runtime/reduce.h:14-49 (MAX: x>y?x:y, MIN: x<y?x:y).
 *
 * Pinning: built twice — once here (this file), once against the unmodified
 * reference sources (oracle/_ref) — and compared by tests/test_oracle.py.
 */
#include <stdarg.h>
#include "physis/physis.h"

FILE *__ps_trace;

void PSInit(int *argc, char ***argv, int grid_num_dims, ...) {
  (void)grid_num_dims;
  __ps_trace = NULL;
  if (!argc || !argv) return;
  for (int i = 0; i < *argc; ++i) {
    const char *a = (*argv)[i];
    if (strcmp(a, "-physis-trace") == 0 || strcmp(a, "--physis-trace") == 0) {
      for (int j = i; j + 1 < *argc; ++j) (*argv)[j] = (*argv)[j + 1];
      --*argc;
      __ps_trace = stderr;
      break; /* the reference removes the first occurrence only */
    }
  }
}

void PSFinalize(void) {}

int __PSGridGetID(__PSGrid *g) {
  (void)g;
  return 0;
}

__PSGrid *__PSGridNew(__PSGridTypeInfo *type_info, int num_dims, PSVectorInt dim) {
  __PSGrid *g = (__PSGrid *)malloc(sizeof(__PSGrid));
  g->elm_size = type_info->size;
  g->num_dims = num_dims;
  memcpy(g->dim, dim, sizeof(PSVectorInt));
  g->num_elms = 1;
  for (int i = 0; i < num_dims; ++i) g->num_elms *= dim[i];
  g->p = calloc(g->num_elms, g->elm_size);
  if (!g->p) return INVALID_GRID;
  return g;
}

void PSGridFree(void *p) {
  __PSGrid *g = (__PSGrid *)p;
  if (g->p) free(g->p);
  g->p = NULL;
}

void PSGridCopyin(void *p, const void *src_array) {
  __PSGrid *g = (__PSGrid *)p;
  memcpy(g->p, src_array, (size_t)g->elm_size * g->num_elms);
}

void PSGridCopyout(void *p, void *dst_array) {
  __PSGrid *g = (__PSGrid *)p;
  memcpy(dst_array, g->p, (size_t)g->elm_size * g->num_elms);
}

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

void __PSGridSet(__PSGrid *g, void *buf, ...) {
  va_list vl;
  va_start(vl, buf);
  PSIndex offset = 0, base = 1;
  for (int i = 0; i < g->num_dims; ++i) {
    PSIndex idx = va_arg(vl, PSIndex);
    offset += idx * base;
    base *= g->dim[i];
  }
  va_end(vl);
  memcpy((char *)g->p + (size_t)offset * g->elm_size, buf, g->elm_size);
}

#define DEFINE_REDUCE(NAME, T)                                        \
  void NAME(void *buf, enum PSReduceOp op, __PSGrid *g) {             \
    const T *d = (const T *)g->p;                                     \
    T v = d[0];                                                       \
    for (int64_t i = 1; i < g->num_elms; ++i) {                       \
      T y = d[i];                                                     \
      switch (op) {                                                   \
        case PS_MAX: v = (v > y) ? v : y; break;                      \
        case PS_MIN: v = (v < y) ? v : y; break;                      \
        case PS_SUM: v = v + y; break;                                \
        case PS_PROD: v = v * y; break;                               \
        default: PSAbort(1);                                          \
      }                                                               \
    }                                                                 \
    *(T *)buf = v;                                                    \
  }
DEFINE_REDUCE(__PSReduceGridFloat, float)
DEFINE_REDUCE(__PSReduceGridDouble, double)
DEFINE_REDUCE(__PSReduceGridInt, int)
DEFINE_REDUCE(__PSReduceGridLong, long)
