/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement of the Physis REFERENCE-target ABI that `physisc --ref`
 * output is compiled against.  It lets the hand-emitted `programs/*.ref.c`
 * build on a machine where /root/reference is absent (the GPU box).  The same
 * program sources are also built against the reference's real headers and
 * real libphysis_rt_ref (see oracle/Makefile, target `_ref`), and the two
 * builds are compared bit-for-bit in tests/test_oracle.py — that is what
 * pins this restatement.
 *
 * Layouts and signatures follow (reference, read-only):
 *   include/physis/physis_common.h:46-60,78-103,155-166  PSIndex, PSVectorInt,
 *        __PSDomain, PSDomainNDNew, __PSGridTypeInfo / MemberInfo
 *   include/physis/physis_ref.h:17-23,41,44-74,90-97     __PSGrid, PSGridDim,
 *        __PSGridNew (3 args), offset helpers, __PSReduceGrid*
 *   include/physis/types.h:17-23, reduce.h:16-21, runtime.h:10,16-30
 */
#ifndef ORACLE_PORT_PHYSIS_H_
#define ORACLE_PORT_PHYSIS_H_

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <sys/time.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_MAX_DIM (3)
typedef int32_t PSIndex;
typedef int PSVectorInt[PS_MAX_DIM];

typedef int PSType;
enum PSPrimitiveType { PS_INT = 0, PS_LONG = 1, PS_FLOAT = 2, PS_DOUBLE = 3, PS_USER = 4 };
enum PSReduceOp { PS_MAX, PS_MIN, PS_SUM, PS_PROD };

#define __PS_PERIODIC(x, y) (((x) + (y)) % (y))
#define PSAssert(e) assert(e)
#define INVALID_GRID (NULL)
static inline void PSAbort(int code) { exit(code); }

typedef struct {
  PSIndex min[PS_MAX_DIM];
  PSIndex max[PS_MAX_DIM];
  PSIndex local_min[PS_MAX_DIM];
  PSIndex local_max[PS_MAX_DIM];
} __PSDomain;
typedef __PSDomain PSDomain1D;
typedef __PSDomain PSDomain2D;
typedef __PSDomain PSDomain3D;

#define PS_GRID_USER_TYPE_MAX_ARRAY_RANK (5)
typedef struct {
  PSType type;
  int size;
  int rank;
  int dim[PS_GRID_USER_TYPE_MAX_ARRAY_RANK];
} __PSGridTypeMemberInfo;
typedef struct {
  PSType type;
  int size;
  int num_members;
  __PSGridTypeMemberInfo *members;
} __PSGridTypeInfo;

/* REF grid handle: host memory, logical order x + y*nx + z*nx*ny. */
typedef struct {
  int elm_size;
  int num_dims;
  int64_t num_elms;
  PSVectorInt dim;
  void *p;
} __PSGrid;

#define PSGridDim(g, d) (((__PSGrid *)(g))->dim[(d)])

extern FILE *__ps_trace;
static inline void __PSTraceStencilPre(const char *msg) {
  if (__ps_trace) fprintf(__ps_trace, "Physis: Stencil started (%s)\n", msg);
}
static inline void __PSTraceStencilPost(float time) {
  if (__ps_trace) fprintf(__ps_trace, "Physis: Stencil finished (time: %f)\n", time);
}

void PSInit(int *argc, char ***argv, int grid_num_dims, ...);
void PSFinalize(void);
void PSGridCopyin(void *g, const void *src_array);
void PSGridCopyout(void *g, void *dst_array);
void PSGridFree(void *g);
PSDomain1D PSDomain1DNew(PSIndex minx, PSIndex maxx);
PSDomain2D PSDomain2DNew(PSIndex minx, PSIndex maxx, PSIndex miny, PSIndex maxy);
PSDomain3D PSDomain3DNew(PSIndex minx, PSIndex maxx, PSIndex miny, PSIndex maxy,
                         PSIndex minz, PSIndex maxz);
__PSGrid *__PSGridNew(__PSGridTypeInfo *type_info, int num_dims, PSVectorInt dim);
int __PSGridGetID(__PSGrid *g);
void __PSGridSet(__PSGrid *g, void *buf, ...);
void __PSReduceGridFloat(void *buf, enum PSReduceOp op, __PSGrid *g);
void __PSReduceGridDouble(void *buf, enum PSReduceOp op, __PSGrid *g);
void __PSReduceGridInt(void *buf, enum PSReduceOp op, __PSGrid *g);
void __PSReduceGridLong(void *buf, enum PSReduceOp op, __PSGrid *g);

static inline PSIndex __PSGridGetOffset1D(__PSGrid *g, PSIndex i1) {
  (void)g;
  return i1;
}
static inline PSIndex __PSGridGetOffset2D(__PSGrid *g, PSIndex i1, PSIndex i2) {
  return i1 + i2 * PSGridDim(g, 0);
}
static inline PSIndex __PSGridGetOffset3D(__PSGrid *g, PSIndex i1, PSIndex i2, PSIndex i3) {
  return i1 + i2 * PSGridDim(g, 0) + i3 * PSGridDim(g, 0) * PSGridDim(g, 1);
}
/* Periodic access wraps by ONE period only: (i+n)%n (physis_ref.h:62-74). */
static inline PSIndex __PSGridGetOffsetPeriodic1D(__PSGrid *g, PSIndex i1) {
  return (i1 + PSGridDim(g, 0)) % PSGridDim(g, 0);
}
static inline PSIndex __PSGridGetOffsetPeriodic2D(__PSGrid *g, PSIndex i1, PSIndex i2) {
  return __PSGridGetOffsetPeriodic1D(g, i1) +
         (i2 + PSGridDim(g, 1)) % PSGridDim(g, 1) * PSGridDim(g, 0);
}
static inline PSIndex __PSGridGetOffsetPeriodic3D(__PSGrid *g, PSIndex i1, PSIndex i2,
                                                  PSIndex i3) {
  return __PSGridGetOffsetPeriodic2D(g, i1, i2) +
         (i3 + PSGridDim(g, 2)) % PSGridDim(g, 2) * PSGridDim(g, 0) * PSGridDim(g, 1);
}

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_PORT_PHYSIS_H_ */
