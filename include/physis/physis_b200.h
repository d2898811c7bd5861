/*
 * physis_b200.h — generated-code ABI of the Physis `b200` target.
 *
 * This is the header a `physisc --b200` translation includes (the role
 * include/physis/physis_cuda.h plays for `--cuda` in the reference).  It is a
 * from-scratch declaration of the SAME binary interface so that translator
 * output, and user `main`s written against the Physis C API, link against
 * libphysis_rt_b200.so unchanged.  Every entry is `extern "C"`, takes plain
 * pointers and sizes only.  Reference interface each one replaces (paths
 * relative to the reference tree):
 *
 *   PSInit / PSFinalize                     include/physis/physis_common.h:78-79
 *   PSDomain{1,2,3}DNew, __PSDomain         include/physis/physis_common.h:88-103
 *   PSIndex, PSVectorInt, PS_MAX_DIM        include/physis/physis_common.h:30,46,60
 *   __PSGridTypeInfo / MemberInfo           include/physis/physis_common.h:155-166
 *   PSType enum, PSReduceOp enum            include/physis/types.h:17-23, reduce.h:16-21
 *   __ps_trace, __PSTraceStencilPre/Post    include/physis/runtime.h:16-30
 *   __PSGrid, __PSGrid{1,2,3}D<T>_dev       include/physis/physis_cuda.h:18-106
 *   __PSGridNew/Free/Copyin/Copyout/Set/
 *     GetID/Swap, __PSCheckCudaError        include/physis/physis_cuda.h:131-150
 *   __PSGridGetOffset[Periodic]{1,2,3}D[Dev] include/physis/physis_cuda.h:196-270
 *   __PSReduceGrid{Float,Double,Int,Long}   include/physis/physis_cuda.h:272-282
 *   PSGridCopyin/Copyout/Free (REF-style)   include/physis/physis_common.h:83-86
 *
 * NEW for b200 (no counterpart in the reference, where the sweep is a
 * generated `__global__`): __PSB200StencilRun and its descriptor, which is what
 * a B200RuntimeBuilder would emit in place of
 * translator/cuda_runtime_builder.cc:1465-1583 (BuildRunFuncBody/LoopBody).
 */
#ifndef PHYSIS_PHYSIS_B200_H_
#define PHYSIS_PHYSIS_B200_H_

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- target-neutral surface ------------------------------------------- */

#define PS_MAX_DIM (3)
typedef int32_t PSIndex;
#define PSINDEX_MAX INT32_MAX
#define PSINDEX_MIN INT32_MIN
typedef int PSVectorInt[PS_MAX_DIM];
typedef PSVectorInt PSPoint;

typedef int PSType;
enum PSPrimitiveType { PS_INT = 0, PS_LONG = 1, PS_FLOAT = 2, PS_DOUBLE = 3, PS_USER = 4 };
enum PSReduceOp { PS_MAX, PS_MIN, PS_SUM, PS_PROD };

#define PSAssert(e) assert(e)
#define INVALID_GRID (NULL)
#define __PS_PERIODIC(x, y) (((x) + (y)) % (y))
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

extern FILE *__ps_trace;
static inline void __PSTraceStencilPre(const char *msg) {
  if (__ps_trace) fprintf(__ps_trace, "Physis: Stencil started (%s)\n", msg);
}
static inline void __PSTraceStencilPost(float time) {
  if (__ps_trace) fprintf(__ps_trace, "Physis: Stencil finished (time: %f)\n", time);
}

/* Consumes --physis-trace and the b200 options (--physis-ngpu is accepted for
 * reference-CLI compatibility) from argv, selects the device (LOCAL_RANK when
 * launched one process per GPU), creates the runtime streams. */
void PSInit(int *argc, char ***argv, int grid_num_dims, ...);
void PSFinalize(void);

PSDomain1D PSDomain1DNew(PSIndex minx, PSIndex maxx);
PSDomain2D PSDomain2DNew(PSIndex minx, PSIndex maxx, PSIndex miny, PSIndex maxy);
PSDomain3D PSDomain3DNew(PSIndex minx, PSIndex maxx, PSIndex miny, PSIndex maxy,
                         PSIndex minz, PSIndex maxz);

/* ---- grid handle ------------------------------------------------------- */

/* Device-side view passed BY VALUE to generic (generated) sweep kernels.
 * Primitive grids: dim[] then one pointer.  User types are stored SoA on the
 * device (as the reference CUDA target does, cuda_runtime_builder.cc:351-391):
 * dim[] followed by one pointer per struct member, in declaration order, so a
 * generated `struct __PSGrid3D<Name>_dev { int dim[3]; int slab; T0 *m0; T1 *m1; }`
 * overlays it exactly.
 *
 * 3-D views carry `slab` in the four bytes the CUDA target's struct leaves as padding
 * before its first pointer (same offsets of dim[] and p): 0 when the grid lives whole on
 * this GPU, else the number of z planes of this rank's allocation (interior + halo).  On a
 * z-slab the view's pointers are shifted so that GLOBAL plane indices address the local
 * allocation, and the periodic wrap in z is not a modulo but the halo planes themselves
 * (the ring exchange fills rank 0's lower halo with the last plane and vice versa) -- the
 * role of local_size/local_offset in the MPI-CUDA target's device structs
 * (include/physis/physis_mpi_cuda.h:53-94) and of GridMPI's periodic handling
 * (runtime/grid_mpi.h:213-232). */
typedef struct { int dim[1]; void *p; } __PSGrid1D_dev;
typedef struct { int dim[2]; void *p; } __PSGrid2D_dev;
typedef struct { int dim[3]; int slab; void *p; } __PSGrid3D_dev;
typedef struct { int dim[3]; int slab; void *p; } __PSGrid_dev;
#define __PS_DECL_DEV(N, Name, T) typedef struct { int dim[N]; T *p; } __PSGrid##N##D##Name##_dev;
#define __PS_DECL_DEV3(Name, T) typedef struct { int dim[3]; int slab; T *p; } __PSGrid3D##Name##_dev;
__PS_DECL_DEV(1, Float, float)  __PS_DECL_DEV(2, Float, float)  __PS_DECL_DEV3(Float, float)
__PS_DECL_DEV(1, Double, double) __PS_DECL_DEV(2, Double, double) __PS_DECL_DEV3(Double, double)
__PS_DECL_DEV(1, Int, int)      __PS_DECL_DEV(2, Int, int)      __PS_DECL_DEV3(Int, int)
__PS_DECL_DEV(1, Long, long)    __PS_DECL_DEV(2, Long, long)    __PS_DECL_DEV3(Long, long)

/* Host handle; field order and types as the CUDA target's so `g->dim[d]`,
 * `g->dev` in translated host code keep working. */
typedef struct {
  void *p;            /* device pointer of the data (member 0 for user types) */
  PSVectorInt dim;
  int elm_size;
  int num_dims;
  int64_t num_elms;
  __PSGrid_dev *dev;  /* host-resident device view, see above */
} __PSGrid;

typedef __PSGrid *PSGrid1DFloat;  typedef __PSGrid *PSGrid2DFloat;  typedef __PSGrid *PSGrid3DFloat;
typedef __PSGrid *PSGrid1DDouble; typedef __PSGrid *PSGrid2DDouble; typedef __PSGrid *PSGrid3DDouble;
typedef __PSGrid *PSGrid1DInt;    typedef __PSGrid *PSGrid2DInt;    typedef __PSGrid *PSGrid3DInt;
typedef __PSGrid *PSGrid1DLong;   typedef __PSGrid *PSGrid2DLong;   typedef __PSGrid *PSGrid3DLong;
#define DeclareGrid1D(name, type) typedef __PSGrid *PSGrid1D##name;
#define DeclareGrid2D(name, type) typedef __PSGrid *PSGrid2D##name;
#define DeclareGrid3D(name, type) typedef __PSGrid *PSGrid3D##name;
#define PSGridDim(p, d) ((p)->dim[(d)])
#define __PSGridDimDev(p, d) ((p)->dim[d])

typedef void *(*__PSGrid_devNewFunc)(int num_dims, PSVectorInt dim);
typedef void (*__PSGrid_devFreeFunc)(void *);
typedef void (*__PSGrid_devCopyinFunc)(void *g, const void *src, size_t num_elms);
typedef void (*__PSGrid_devCopyoutFunc)(void *g, void *dst, size_t num_elms);

/* The four func arguments exist for source compatibility with `--cuda`
 * translations (which pass generated per-user-type helpers).  The b200 runtime
 * handles user types itself from type_info->members (device SoA + on-device
 * AoS<->SoA transposition), so translations for b200 pass NULL; a non-NULL
 * function is honoured exactly like the CUDA runtime does. */
__PSGrid *__PSGridNew(__PSGridTypeInfo *type_info, int num_dims, PSVectorInt dim,
                      __PSGrid_devNewFunc func);
void __PSGridFree(void *g, __PSGrid_devFreeFunc func);
void __PSGridCopyin(void *g, const void *src_array, __PSGrid_devCopyinFunc func);
void __PSGridCopyout(void *g, void *dst_array, __PSGrid_devCopyoutFunc func);
void __PSGridSet(__PSGrid *g, void *buf, ...); /* one PSIndex per dimension */
void __PSGridSwap(__PSGrid *g);                /* no-op, as in the reference */
int __PSGridGetID(__PSGrid *g);
void __PSCheckCudaError(const char *message);

/* REF-style three-call surface (physis_common.h:83-86); same as the __ forms
 * with NULL helpers.  Both calls are synchronous: data is valid on return. */
void PSGridCopyin(void *g, const void *src_array);
void PSGridCopyout(void *g, void *dst_array);
void PSGridFree(void *g);

void __PSReduceGridFloat(void *buf, enum PSReduceOp op, __PSGrid *g);
void __PSReduceGridDouble(void *buf, enum PSReduceOp op, __PSGrid *g);
void __PSReduceGridInt(void *buf, enum PSReduceOp op, __PSGrid *g);
void __PSReduceGridLong(void *buf, enum PSReduceOp op, __PSGrid *g);

/* ---- offsets (host + device) ------------------------------------------ */

#if defined(__CUDACC__)
#define PS_FUNCTION_DEVICE __host__ __device__
#else
#define PS_FUNCTION_DEVICE
#endif

static inline PSIndex __PSGridGetOffset1D(__PSGrid *g, PSIndex i1) { (void)g; return i1; }
static inline PSIndex __PSGridGetOffset2D(__PSGrid *g, PSIndex i1, PSIndex i2) {
  return i1 + i2 * PSGridDim(g, 0);
}
static inline PSIndex __PSGridGetOffset3D(__PSGrid *g, PSIndex i1, PSIndex i2, PSIndex i3) {
  return i1 + i2 * PSGridDim(g, 0) + i3 * PSGridDim(g, 0) * PSGridDim(g, 1);
}
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
#define __PS_DEVDIM(g, d) (((const __PSGrid_dev *)(g))->dim[(d)])
PS_FUNCTION_DEVICE static inline PSIndex __PSGridGetOffset1DDev(const void *g, PSIndex i1) {
  (void)g;
  return i1;
}
PS_FUNCTION_DEVICE static inline PSIndex __PSGridGetOffset2DDev(const void *g, PSIndex i1,
                                                                PSIndex i2) {
  return i1 + i2 * __PS_DEVDIM(g, 0);
}
PS_FUNCTION_DEVICE static inline PSIndex __PSGridGetOffset3DDev(const void *g, PSIndex i1,
                                                                PSIndex i2, PSIndex i3) {
  return i1 + i2 * __PS_DEVDIM(g, 0) + i3 * __PS_DEVDIM(g, 0) * __PS_DEVDIM(g, 1);
}
PS_FUNCTION_DEVICE static inline PSIndex __PSGridGetOffsetPeriodic1DDev(const void *g,
                                                                        PSIndex i1) {
  return (i1 + __PS_DEVDIM(g, 0)) % __PS_DEVDIM(g, 0);
}
PS_FUNCTION_DEVICE static inline PSIndex __PSGridGetOffsetPeriodic2DDev(const void *g,
                                                                        PSIndex i1,
                                                                        PSIndex i2) {
  return __PSGridGetOffsetPeriodic1DDev(g, i1) +
         (i2 + __PS_DEVDIM(g, 1)) % __PS_DEVDIM(g, 1) * __PS_DEVDIM(g, 0);
}
#define __PS_DEVSLAB(g) (((const __PSGrid_dev *)(g))->slab)
PS_FUNCTION_DEVICE static inline PSIndex __PSGridGetOffsetPeriodic3DDev(const void *g,
                                                                        PSIndex i1,
                                                                        PSIndex i2,
                                                                        PSIndex i3) {
  /* on a z-slab planes -1 and dim[2] are this rank's halo planes (ring wrap) */
  const PSIndex z = __PS_DEVSLAB(g) ? i3 : (i3 + __PS_DEVDIM(g, 2)) % __PS_DEVDIM(g, 2);
  return __PSGridGetOffsetPeriodic2DDev(g, i1, i2) + z * __PS_DEVDIM(g, 0) * __PS_DEVDIM(g, 1);
}
/* Elements between consecutive components of an array member of a user type (device SoA,
 * components plane-major as translator/cuda_runtime_builder.cc:267-305 indexes them): the
 * number of elements of this rank's allocation. */
PS_FUNCTION_DEVICE static inline size_t __PSGridMemberStride3DDev(const void *g) {
  return (size_t)__PS_DEVDIM(g, 0) * (size_t)__PS_DEVDIM(g, 1) *
         (size_t)(__PS_DEVSLAB(g) ? __PS_DEVSLAB(g) : __PS_DEVDIM(g, 2));
}

/* ---- b200 stencil-run entry (NEW) -------------------------------------- */

/* Sweep families with a hand-written sm_100a kernel.  A translation names the
 * family its kernel body was recognised as; anything else goes GENERIC and
 * carries a launch stub for the per-point kernel compiled into the program. */
enum __PSB200Kind {
  PSB200_KIND_GENERIC = 0,
  /* out = cc*c + cw*w + ce*e + cs*s + cn*n + cb*b + ct*t (left to right, no
   * FMA), faces clamp to the centre value; examples/diffusion-benchmark/
   * diffusion3d_physis.c:29-58.  grids: {in, out}; scalars: ce,cw,cn,cs,ct,cb,cc */
  PSB200_KIND_DIFFUSION7_CLAMP = 1,
  /* Himeno 19-pt Jacobi, examples/himeno/himenobmtxpa_physis.c:331-361.
   * grids: {p0,p1,a0,a1,a2,a3,b0,b1,b2,c0,c1,c2,bnd,wrk1}; scalars: omega */
  PSB200_KIND_HIMENO19 = 2,
  /* same + `PSGridEmit(gosa_g, ss*ss)`; grids: {..., wrk1, gosa_g} */
  PSB200_KIND_HIMENO19_GOSA = 3,
  /* periodic 7-pt on one member of a user type with a vertex-staggered
   * coefficient grid (examples/dsl/diffusion3d_periodic_staggered.c).
   * grids: {u, kap}; members: {read member, write member} */
  PSB200_KIND_PERIODIC7_STAGGERED = 4,
  PSB200_NUM_KINDS
};

#define PSB200_MAX_GRIDS 16
#define PSB200_MAX_SCALARS 8

/* cudaStream_t without dragging cuda_runtime.h into C translation units */
typedef void *__PSB200Stream;
/* Launch stub of a GENERIC sweep: enqueue one sweep of `stencil` over `dom` on
 * `stream`.  `dom` is the stencil's own domain on one GPU; in a multi-GPU run the
 * runtime passes the part of it this rank owns (global coordinates; the device
 * views index globally).  Role of __PSDomainSetLocalSize, mpi_runtime_builder.cc:247-282. */
typedef void (*__PSB200LaunchFunc)(const void *stencil, const __PSDomain *dom,
                                   __PSB200Stream stream);

typedef struct {
  int kind;                          /* enum __PSB200Kind */
  int elm_type;                      /* PS_FLOAT / PS_DOUBLE for the specialised kinds */
  __PSDomain dom;
  int num_grids;
  __PSGrid *grids[PSB200_MAX_GRIDS];
  int members[PSB200_MAX_GRIDS];     /* user-type member index per grid, -1 if primitive */
  int num_scalars;
  double scalars[PSB200_MAX_SCALARS]; /* float scalars widened exactly */
  const void *stencil;               /* GENERIC: the translated __PSStencil_<k> struct */
  __PSB200LaunchFunc launch;         /* GENERIC: enqueues one sweep on `stream` */
  const char *name;                  /* kernel name for --physis-trace */
  unsigned written_mask;             /* GENERIC: bit i set = grids[i] is PSGridEmit'ed (multi-GPU
                                      * halo refresh); 0 = unknown, refresh every grid */
  int z_reach;                       /* GENERIC: largest |z offset| of any PSGridGet in the kernel
                                      * (the translator's StencilRange; the reference sizes its halos
                                      * from it, runtime/grid_space_mpi.h:42-62).  A multi-GPU run
                                      * refuses a sweep that reaches beyond the halo planes (option
                                      * `halo`); 0 = not stated, at most one plane assumed */
} __PSB200StencilDesc;

/* for (i < iter) { sweep descs[0]; sweep descs[1]; ... } enqueued in order on
 * the runtime stream, no host synchronisation (ordering with PSGridCopyout is
 * by stream, as in the reference CUDA target).  Returns elapsed milliseconds
 * when tracing is on (which then synchronises), else 0.0f — the contract of
 * translator/reference_runtime_builder.cc:896-940,1068-1081. */
float __PSB200StencilRun(int iter, int num_stencils, const __PSB200StencilDesc *descs);

/* How many of the `iter` iterations of a fusable ping-pong pair of sweeps run as fused
 * two-sweep passes (the rest run sweep by sweep so that both grids end up exactly as the
 * reference's schedule leaves them, translator/reference_runtime_builder.cc:837-893). */
int __PSB200FusedPassCount(int iter);

__PSB200Stream __PSB200GetStream(void);
void __PSB200Synchronize(void);

/* Device-timed region on the runtime stream (CUDA events), for benchmarks. */
void __PSB200TimerStart(void);
float __PSB200TimerStopMs(void); /* synchronises */

/* Introspection used by tests/bench: kernels launched so far, bytes copied. */
typedef struct {
  uint64_t kernel_launches;
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t halo_bytes;
  float last_kernel_ms;   /* mean device ms per launch of the last timed family, 0 if off */
  uint64_t fused_pairs;   /* of kernel_launches: fused two-sweep passes (two sweeps each) */
  uint64_t fused_pairs_timed; /* with option time_kernels=1: passes covered by fused_pair_ms */
  double fused_pair_ms;       /* ... and their accumulated device time (CUDA events) */
  uint64_t reduces_from_partials; /* PSReduce calls answered from the partial sums the producing
                                   * sweep left behind instead of a pass over the grid */
  uint64_t plan_cache_hits;       /* sweeps whose prepared plan (TMA descriptors, launch shape) was
                                   * reused from an earlier __PSB200StencilRun */
  /* halo-exchange profile of multi-GPU sweeps (option halo_profile=1; cf. the reference's
   * per-grid DataCopyProfile, runtime/timing.h:11-18, grid_space_mpi_cuda.h:556-569): time the
   * sweeps' CTAs spent waiting for the ring neighbours before touching halo planes */
  uint64_t halo_wait_ns_sum;      /* summed over the CTAs that waited */
  uint64_t halo_wait_ns_max;      /* longest single wait */
  uint64_t halo_wait_ctas;        /* CTAs that waited (boundary work items only) */
  uint64_t halo_wait_launches;    /* sweeps that took part */
  /* option autotune=1 (cf. the reference's AUTO_TUNING trial iterations, include/physis/runtime.h:32-52,
   * translator/configuration.cc:27-57): kernel forms timed on a run's own first iterations, and
   * runs that then used a form other than the defaults */
  uint64_t autotune_trials;
  uint64_t autotuned_runs;
} __PSB200Stats;
void __PSB200GetStats(__PSB200Stats *out);
void __PSB200ResetStats(void);
/* Runtime knobs (tile shape etc.) for tuning runs: "key=value". Returns 0 on success. */
int __PSB200SetOption(const char *key_value);
/* What the tuner (option autotune=1) last settled on: "<option overrides | defaults>: <ms> per
 * iteration (defaults <ms>; <n> forms tried)", "" before any tuning.  The string lives until
 * the next tuned run. */
const char *__PSB200LastTuning(void);
const char *__PSB200Version(void);
/* ---- multi-GPU (one process per GPU, SPMD) ------------ */
/* Rank / size of the process group PSInit joined (RANK, WORLD_SIZE, LOCAL_RANK,
 * MASTER_PORT from the launcher's environment; 0 / 1 when launched alone). */
int __PSB200Rank(void);
int __PSB200WorldSize(void);
/* The z-slab of a grid this rank owns (cf. __PSGetLocalOffset / __PSGetLocalSize,
 * include/physis/physis_mpi_cuda.h:412-500). */
void __PSB200GridLocalSize(void *g, int *z_offset, int *z_length);
/* Slab-local transfers: the host buffer holds exactly this rank's z_length planes
 * (PSGridCopyin/Copyout take and return the whole global array on every rank). */
void __PSB200GridCopyinLocal(void *g, const void *src_slab);
void __PSB200GridCopyoutLocal(void *g, void *dst_slab);
/* Pure host helpers (no CUDA): the block decomposition and the shared-memory
 * rendezvous, exposed so that they are testable on a CPU-only machine. */
void __PSB200Partition(int n, int domain_n, int world, int rank, int *offset, int *length);
int __PSB200GroupSelfTest(void);

/* Page-locked host memory for callers that want DMA-speed Copyin/Copyout. */
void *__PSB200HostAlloc(size_t bytes);
void __PSB200HostFree(void *p);

#ifdef __cplusplus
}
#endif
#endif /* PHYSIS_PHYSIS_B200_H_ */
