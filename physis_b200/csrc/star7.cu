// 7-point diffusion sweep with clamped faces — hand-written sm_100a kernel.
//
// Computes, for every point of the domain,
//     out = cc*c + cw*w + ce*e + cs*s + cn*n + cb*b + ct*t
// evaluated left to right with separately rounded multiplies and adds (no FMA
// contraction), a neighbour outside the grid being replaced by the centre
// value — bit-for-bit what the REFERENCE target computes for
// examples/diffusion-benchmark/diffusion3d_physis.c:29-58 (kernel_physis),
// whose CUDA-target form is the generated per-thread z loop of
// translator/cuda_runtime_builder.cc:1259-1281.  fp32 and fp64.
//
// Structure (B200-first, not the reference's launch shape):
//  * persistent CTAs, one work item = (xy tile, z chunk); items are ordered
//    z-chunk-major so concurrently resident CTAs share halo rows through L2;
//  * a producer warp streams haloed xy tiles of successive z planes into a
//    shared-memory ring with TMA (cp.async.bulk.tensor.3d, zero fill outside the
//    grid), completion on mbarriers; consumers release planes on a second set;
//  * 2.5-D blocking: each consumer thread owns a 16-byte vector (4 floats /
//    2 doubles) in RY consecutive rows and keeps bottom/centre/top planes of its
//    own cells in registers while marching along z; y-neighbours of the edge
//    rows come from the shared centre plane, x-neighbours from warp shuffles
//    (lane 0 / 31 read the tile's x-halo column from shared memory);
//  * one 128-bit coalesced store per vector.
// Algorithmic traffic: 1 read + 1 write per point (8 B/LUP fp32, 16 B/LUP fp64;
// examples/diffusion-benchmark/diffusion3d.h:97-100).  HBM-bound; tensor cores
// are deliberately unused (no contraction to feed them).
#include "runtime.h"
#include "tma.cuh"
#include "sweep_common.cuh"
#include "star7_math.cuh"

#include <algorithm>
#include <type_traits>

namespace physis_b200 {

namespace {

using namespace sweep;

template <typename T>
struct Star7Args {
  T *out;
  int nx, ny, nz;
  int dx0, dx1, dy0, dy1, dz0, dz1;  // domain = store mask
  int xbase;                         // x origin of tile column 0 (multiple of VEC)
  T cc, cw, ce, cs, cn, cb, ct;
  int ntx, nty, nzc, zc;             // tiles in x, y; number and length of z chunks
  int nitems;
  int stages;
  int st_hint;   // 1: streaming (evict-first) stores
  // z-slab view (multi-GPU; on one GPU zcl = {0, nz-1} and nothing is pushed):
  // local planes at which the bottom / top neighbour is clamped to the centre,
  int zcl_lo, zcl_hi;
  // and the local planes whose result is also stored, through the CUDA-IPC peer
  // mapping, into the ring neighbour's halo plane (fused halo exchange over NVLink)
  int push_lo_z, push_hi_z;
  T *push_lo, *push_hi;
  SlabSync sync;  // neighbour ordering fused into the kernel
};

// The consumers' instruction count is cut to what the stencil needs (a first form of this
// kernel, since removed, ran at 64 % of issue slots busy with the fp32 work a third of the
// instructions): clamps are taken out of the per-element path (z: register copies on the one
// plane that needs them; y: warp-uniform branch; x: one select on the edge lanes), the z window
// rotates by renaming instead of moving registers, addresses advance by adds, the mbarrier
// wait is two instructions on its fast path, and -- fp32 only -- the six additions per point
// run as packed add.rn.f32x2 (SASS FADD2) on pairs of separately rounded scalar products.
// (ptxas contracts a packed multiply feeding a packed add into FFMA2 even with explicit .rn,
// which would change the rounding, so the multiplies stay scalar.)
//
// TY rows of the CTA tile, RY rows per thread, NBX boxes side by side in x, MINB resident CTAs
// per SM the register budget is sized for; FP: packed adds.


// FR ("full row"): the NBX boxes of a tile cover a whole grid row, so boxes carry no
// x halo (the x neighbours of a box's edge lanes live in the adjacent box of the same
// stage) and a tile's rows are contiguous in memory: no halo over-fetch in x, and
// every plane of a tile is one contiguous DRAM range.
template <typename T, int TY, int RY, int NBX, int MINB, int FP, bool FR>
__global__ void __launch_bounds__((NBX * (TY / RY) + 1) * 32, MINB)
Star7KernelV2(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Star7Args<T> a) {
  using G = Geom<T>;
  using V = typename VecOf<T>::type;
  constexpr int VEC = G::VEC;
  constexpr int NWY = TY / RY;
  constexpr int NW = NBX * NWY;  // consumer warps
  constexpr int HXO = FR ? 0 : G::HX;                        // x halo elements per side
  constexpr int ROWB = (G::TXB + 2 * HXO) * (int)sizeof(T);  // bytes of one box row
  constexpr int BOX_STRIDE = ((TY + 2) * ROWB + 127) / 128 * 128;
  constexpr int STAGE_BYTES = NBX * BOX_STRIDE;
  // where the element left of lane 0's vector / right of lane 31's vector lives,
  // relative to that lane's own position in the row
  constexpr int WEST_OFF = FR ? -BOX_STRIDE + (G::TXB - 1) * (int)sizeof(T) : -(int)sizeof(T);
  constexpr int EAST_OFF = FR ? BOX_STRIDE - 31 * VEC * (int)sizeof(T) : VEC * (int)sizeof(T);
  static_assert(TY % RY == 0, "TY must be a multiple of RY");

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + kMaxStages;
  unsigned char *planes = smem + kBarrierBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = a.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      tma::mbar_init(&full[s], 1);
      tma::mbar_init(&empty[s], NW);
    }
    tma::fence_barrier_init();
  }
  // programmatic dependent launch (LaunchSweepKernel): nothing of the grids is touched before
  // the kernel before this one is complete and visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  SlabSyncWait(a.sync);
  __syncthreads();

  const int tiles_xy = a.ntx * a.nty;

  if (warp == NW) {
    // ------------------------------------------------------------ producer
    if (lane != 0) return;
    tma::prefetch_tensormap(&tmap);
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int zseq = item / tiles_xy;
      const int txy = item - zseq * tiles_xy;
      const int ty = txy / a.ntx;
      const int tx = txy - ty * a.ntx;
      const int x0 = a.xbase + tx * (NBX * G::TXB);
      const int y0 = a.dy0 + ty * TY;
      int zb, ze;
      SlabChunkRange(a.sync, zseq, a.nzc, a.zc, a.dz0, a.dz1, &zb, &ze);
      const int zfirst = zb > 0 ? zb - 1 : zb;
      const int zlast = min(ze, a.nz - 1);
      int nbox = 0;
#pragma unroll
      for (int b = 0; b < NBX; ++b) nbox += (x0 + b * G::TXB < a.nx) ? 1 : 0;
      const uint32_t tx_bytes = (uint32_t)nbox * (uint32_t)((TY + 2) * ROWB);
      for (int z = zfirst; z <= zlast; ++z) {
        tma::mbar_wait(&empty[stage], phase ^ 1u);
        tma::mbar_arrive_expect_tx(&full[stage], tx_bytes);
        unsigned char *dst = planes + stage * STAGE_BYTES;
#pragma unroll
        for (int b = 0; b < NBX; ++b) {
          const int bx0 = x0 + b * G::TXB;
          if (bx0 < a.nx)
            tma::load_3d(dst + b * BOX_STRIDE, &tmap, &full[stage], bx0 - HXO, y0 - 1, z);
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int bx = warp % NBX;
  const int wy = warp / NBX;
  // this thread's first own row (smem row wy*RY + 1) and column inside a stage
  const unsigned char *my = planes + bx * BOX_STRIDE + (wy * RY + 1) * ROWB +
                            (HXO + lane * VEC) * (int)sizeof(T);
  const bool lane_first = (lane == 0), lane_last = (lane == 31);
  // FR: the outermost boxes have no neighbour box (their outer x neighbour is clamped)
  const bool rd_west = lane_first && (!FR || bx > 0);
  const bool rd_east = lane_last && (!FR || bx < NBX - 1);
  const size_t plane_elems = (size_t)a.nx * a.ny;

  int stage = 0;       // ring position of the next plane to consume
  uint32_t phase = 0;

#define S7_ADVANCE() do { if (++stage == S) { stage = 0; phase ^= 1u; } } while (0)
#define S7_RELEASE(st) do { __syncwarp(); if (lane_first) tma::mbar_arrive(&empty[st]); } while (0)
#define S7_LOAD(DST, st) do { \
    const unsigned char *p__ = my + (st) * STAGE_BYTES; \
    _Pragma("unroll") for (int r = 0; r < RY; ++r) DST[r] = *reinterpret_cast<const V *>(p__ + r * ROWB); \
  } while (0)

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int zseq = item / tiles_xy;
    const int txy = item - zseq * tiles_xy;
    const int ty = txy / a.ntx;
    const int tx = txy - ty * a.ntx;
    const int x = a.xbase + tx * (NBX * G::TXB) + bx * G::TXB + lane * VEC;
    const int ybase = a.dy0 + ty * TY + wy * RY;
    int zb, ze;
    SlabChunkRange(a.sync, zseq, a.nzc, a.zc, a.dz0, a.dz1, &zb, &ze);
    const bool x_ok = (x >= a.dx0) && (x + VEC <= a.dx1);
    const bool x_first = (x == 0);
    const bool x_last = (x + VEC == a.nx);
    // rows whose y neighbour leaves the grid take the centre value instead
    // (warp-uniform: every lane of a warp has the same rows)
    const int r_north = -ybase;              // row with y == 0, if in [0, RY)
    const int r_south = a.ny - 1 - ybase;    // row with y == ny-1, if in [0, RY)
    const bool y_edge = (r_north >= 0 && r_north < RY) || (r_south >= 0 && r_south < RY);
    bool st_ok[RY];
    T *outp[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int y = ybase + r;
      st_ok[r] = x_ok && y >= a.dy0 && y < a.dy1;
      outp[r] = a.out + ((size_t)zb * a.ny + y) * a.nx + x;
    }

    V w0[RY], w1[RY], w2[RY];  // the z window; roles rotate, registers do not move

    if (zb > 0) {
      tma::mbar_wait(&full[stage], phase);
      S7_LOAD(w0, stage);
      S7_RELEASE(stage);
      S7_ADVANCE();
    }
    int stage_c = stage;
    tma::mbar_wait(&full[stage], phase);
    S7_LOAD(w1, stage);
    S7_ADVANCE();
    if (zb == a.zcl_lo || zb == 0) {
      // bottom neighbour outside the grid: clamp to the centre
#pragma unroll
      for (int r = 0; r < RY; ++r) w0[r] = w1[r];
    }

    bool has_top = false;
    int z = zb;

    // one plane: BOT / CEN hold planes z-1 / z, TOP receives plane z+1
#define S7_STEP(BOT, CEN, TOP) do { \
      has_top = (z + 1 < a.nz); \
      const int stage_t = stage; \
      if (has_top) { tma::mbar_wait(&full[stage], phase); S7_LOAD(TOP, stage); } \
      if (!has_top || z == a.zcl_hi) { \
        _Pragma("unroll") for (int r = 0; r < RY; ++r) TOP[r] = CEN[r]; \
      } \
      const unsigned char *cb = my + stage_c * STAGE_BYTES; \
      V north = *reinterpret_cast<const V *>(cb - ROWB); \
      V south = *reinterpret_cast<const V *>(cb + RY * ROWB); \
      T *const push0 = (z == a.push_lo_z) ? a.push_lo : nullptr; \
      T *const push1 = (z == a.push_hi_z) ? a.push_hi : nullptr; \
      _Pragma("unroll") for (int r = 0; r < RY; ++r) { \
        const V c = CEN[r]; \
        T wv = __shfl_up_sync(0xffffffffu, v2::Last(c), 1); \
        T ev = __shfl_down_sync(0xffffffffu, v2::First(c), 1); \
        if (rd_west) wv = *reinterpret_cast<const T *>(cb + r * ROWB + WEST_OFF); \
        if (rd_east) ev = *reinterpret_cast<const T *>(cb + r * ROWB + EAST_OFF); \
        if (x_first) wv = v2::First(c); \
        if (x_last) ev = v2::Last(c); \
        V nv = (r == 0) ? north : CEN[r > 0 ? r - 1 : 0]; \
        V sv = (r == RY - 1) ? south : CEN[r < RY - 1 ? r + 1 : r]; \
        if (y_edge) { \
          if (r == r_north) nv = c; \
          if (r == r_south) sv = c; \
        } \
        const V o = v2::Vec7<FP>(a, c, wv, ev, sv, nv, BOT[r], TOP[r]); \
        if (st_ok[r]) { \
          StoreVec(reinterpret_cast<V *>(outp[r]), o, a.st_hint != 0); \
          if (push0) *reinterpret_cast<V *>(push0 + (size_t)(ybase + r) * a.nx + x) = o; \
          if (push1) *reinterpret_cast<V *>(push1 + (size_t)(ybase + r) * a.nx + x) = o; \
        } \
        outp[r] += plane_elems; \
      } \
      S7_RELEASE(stage_c); \
      if (has_top) { stage_c = stage_t; S7_ADVANCE(); } \
      ++z; \
    } while (0)

    for (;;) {
      S7_STEP(w0, w1, w2);
      if (z >= ze) break;
      S7_STEP(w1, w2, w0);
      if (z >= ze) break;
      S7_STEP(w2, w0, w1);
      if (z >= ze) break;
    }
    if (has_top) S7_RELEASE(stage_c);  // plane ze was loaded as the top plane only
    SlabSyncItemDone(a.sync, item, NW * 32, threadIdx.x == 0);
  }
  SlabSyncSignal(a.sync, NW * 32, threadIdx.x == 0);
#undef S7_STEP
#undef S7_LOAD
#undef S7_RELEASE
#undef S7_ADVANCE
}

// ------------------------------------------------------------------ host side

struct VariantInfo {
  int ty, ry, nbx;
  const void *f32;     // packed adds
  const void *f64;
  bool full_row;       // boxes without x halo covering whole rows
};

#define VARIANT(TY, RY, NBX, MINB, FR) \
  { TY, RY, NBX, \
    (const void *)Star7KernelV2<float, TY, RY, NBX, MINB, 1, FR>, \
    (const void *)Star7KernelV2<double, TY, RY, NBX, MINB, 0, FR>, FR }

// Tile shapes (chosen on the B200 by tools/tune_star7.py; shapes that were never selected have
// been removed).  Index = option star7_variant.
const VariantInfo kVariants[] = {
    VARIANT(16, 2, 2, 1, false),  // 0: two haloed boxes x 16 rows: rows wider than 4 boxes, partial x domains
    VARIANT(8, 2, 1, 4, false),   // 1: one haloed box x 8 rows: narrow partial domains
    VARIANT(8, 2, 4, 1, true),    // 2: full rows of 4 boxes (512 floats / 256 doubles)
    VARIANT(16, 2, 2, 1, true),   // 3: full rows of 2 boxes
    VARIANT(8, 2, 1, 4, true),    // 4: full rows of one box
    VARIANT(16, 4, 4, 1, true),   // 5: full rows of 4 boxes, 16-row tiles (tuning alternative)
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kVarHalo2 = 0, kVarHalo1 = 1, kVarFull4 = 2, kVarFull2 = 3, kVarFull1 = 4;

template <typename T>
size_t SmemBytes(const VariantInfo &v, int stages) {
  const size_t rowb = v.full_row ? (size_t)Geom<T>::TXB * sizeof(T) : (size_t)Geom<T>::ROW_BYTES;
  size_t box = ((size_t)(v.ty + 2) * rowb + 127) / 128 * 128;
  return kBarrierBytes + (size_t)stages * v.nbx * box;
}

}  // namespace

struct Star7Plan {
  bool is_double = false;
  int variant = 0;
  int grid = 0, block = 0;
  size_t smem = 0;
  CUtensorMap tmap;
  Star7Args<float> af;
  Star7Args<double> ad;
  const void *fn = nullptr;
  bool pushes = false;  // the kernel itself delivers the halo planes of `out`
  bool syncs = false;   // ... and waits for / signals the neighbours itself
};

template <typename T>
static void FillArgs(Star7Args<T> *a, const __PSB200StencilDesc &d, const Grid *gin, Grid *gout) {
  a->out = (T *)gout->members[0].dev;
  a->nx = gin->ldim[0];
  a->ny = gin->ldim[1];
  a->nz = gin->ldim[2];
  a->zcl_lo = gin->LocalInterior(0);
  a->zcl_hi = gin->LocalInterior(gin->dim[2] - 1);
  a->push_lo_z = a->push_hi_z = -1;
  a->push_lo = a->push_hi = nullptr;
  a->sync = SlabSync{};
  a->dx0 = d.dom.local_min[0]; a->dx1 = d.dom.local_max[0];
  a->dy0 = d.dom.local_min[1]; a->dy1 = d.dom.local_max[1];
  a->dz0 = d.dom.local_min[2]; a->dz1 = d.dom.local_max[2];
  // scalars arrive in the kernel's parameter order: ce, cw, cn, cs, ct, cb, cc
  a->ce = (T)d.scalars[0]; a->cw = (T)d.scalars[1]; a->cn = (T)d.scalars[2];
  a->cs = (T)d.scalars[3]; a->ct = (T)d.scalars[4]; a->cb = (T)d.scalars[5];
  a->cc = (T)d.scalars[6];
}

// Returns nullptr when the descriptor cannot use this kernel.
Star7Plan *PrepareStar7(Runtime *rt, const __PSB200StencilDesc &d, std::string *why) {
  if (d.num_grids != 2 || d.num_scalars != 7) { *why = "expects 2 grids and 7 scalars"; return nullptr; }
  Grid *gin = Grid::FromHandle(d.grids[0]);
  Grid *gout = Grid::FromHandle(d.grids[1]);
  if (gin->num_dims != 3 || gout->num_dims != 3) { *why = "3-D grids only"; return nullptr; }
  if (gin->is_user_type() || gout->is_user_type()) { *why = "primitive element types only"; return nullptr; }
  if (gin->type != gout->type || (gin->type != PS_FLOAT && gin->type != PS_DOUBLE)) {
    *why = "float or double grids of one type"; return nullptr;
  }
  for (int i = 0; i < 3; ++i)
    if (gin->dim[i] != gout->dim[i] || gin->ldim[i] != gout->ldim[i]) { *why = "grids must have equal extents"; return nullptr; }
  if (gin == gout) { *why = "in-place sweep"; return nullptr; }
  const bool dbl = gin->type == PS_DOUBLE;
  const int vec = dbl ? 2 : 4;
  const __PSDomain &dom = d.dom;
  if (gin->ldim[0] % vec != 0 || dom.local_min[0] % vec != 0 || dom.local_max[0] % vec != 0) {
    *why = "x extent and domain x-range must be multiples of 16 bytes"; return nullptr;
  }
  for (int i = 0; i < 3; ++i) {
    if (dom.local_min[i] < 0 || dom.local_max[i] > gin->ldim[i]) { *why = "domain exceeds grid"; return nullptr; }
  }
  if (dom.local_max[0] <= dom.local_min[0] || dom.local_max[1] <= dom.local_min[1] ||
      dom.local_max[2] <= dom.local_min[2]) { *why = "empty domain"; return nullptr; }

  Star7Plan *p = new Star7Plan();
  p->is_double = dbl;
  const Options &o = rt->opt;
  // Tile shape (measured on B200, tools/tune_star7.py, profiles/r1_tune_star7_512.csv): whenever
  // a row fits 4 boxes the tile spans whole rows (no x halo, every plane of a tile one contiguous
  // DRAM range; 8 rows x 4 boxes, 2 rows per thread, 6-deep ring, one CTA per SM reaches 5.8 TB/s
  // at 512^3 fp32).  Wider rows use two haloed boxes x 16 rows.
  int variant = o.star7_variant;
  const size_t row_bytes = (size_t)(dom.local_max[0] - dom.local_min[0]) * (dbl ? 8 : 4);
  const bool autov = variant < 0 || variant >= kNumVariants;
  const int halo_variant = row_bytes >= 1024 ? kVarHalo2 : kVarHalo1;
  if (autov) {
    const int boxes = (int)((gin->ldim[0] * (size_t)(dbl ? 8 : 4) + 511) / 512);
    // (wider full-row tiles would need 4 rows per thread, which spills at 96 registers)
    if (dom.local_min[0] != 0 || boxes > 4) variant = halo_variant;
    else if (boxes <= 1) variant = kVarFull1;
    else if (boxes == 2) variant = kVarFull2;
    else variant = kVarFull4;
  }
  const int txb0 = dbl ? Geom<double>::TXB : Geom<float>::TXB;
  if (kVariants[variant].full_row) {
    // needs a tile that spans the whole row
    const bool fits = dom.local_min[0] == 0 && gin->ldim[0] <= kVariants[variant].nbx * txb0;
    if (!fits) {
      if (!autov) { *why = "full-row tile shape does not span this grid's rows"; delete p; return nullptr; }
      variant = halo_variant;
    }
  }
  const VariantInfo &v = kVariants[variant];
  p->variant = variant;
  p->fn = dbl ? v.f64 : v.f32;
  int stages = o.star7_stages > 0 ? std::min(o.star7_stages, kMaxStages) : (variant == kVarHalo1 ? 5 : 6);
  if (stages < 3) stages = 3;
  // deepest ring that fits the 227 KB of one SM
  while (stages > 3 && (dbl ? SmemBytes<double>(v, stages) : SmemBytes<float>(v, stages)) > 227 * 1024)
    --stages;
  p->smem = dbl ? SmemBytes<double>(v, stages) : SmemBytes<float>(v, stages);
  p->block = (v.nbx * (v.ty / v.ry) + 1) * 32;
  PSB_CHECK(p->block <= 1024, "star7 tile shape needs more than 1024 threads");
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->fn, p->block, p->smem));
  PSB_CHECK(occ > 0, "star7 kernel does not fit on an SM");
  if (o.star7_occ > 0) occ = std::min(occ, o.star7_occ);

  const int txb = dbl ? Geom<double>::TXB : Geom<float>::TXB;
  const int xbase = dom.local_min[0];
  const int ntx = CeilDiv(dom.local_max[0] - xbase, (long)v.nbx * txb);
  const int nty = CeilDiv(dom.local_max[1] - dom.local_min[1], v.ty);
  const int nzd = dom.local_max[2] - dom.local_min[2];
  const int slots = rt->sm_count * occ;
  int zc = o.star7_zc;
  if (zc <= 0) {
    // every z chunk re-reads two planes, and the statically strided items run in waves of
    // `slots`: take the chunk count with the least total plane work per CTA slot (measured:
    // 32 planes at 512^3 with 64 full-row tiles, 128 planes at 1024x1024x512 with 256 tiles)
    const int tiles = ntx * nty;
    long best_cost = -1;
    for (int n = 1; n <= std::max(1, nzd / 8); ++n) {
      const int c = CeilDiv(nzd, n);
      const long waves = CeilDiv((long)tiles * CeilDiv(nzd, c), slots);
      const long cost = waves * (c + 2);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; zc = c; }
    }
    if (zc <= 0) zc = nzd;
  }
  const int nzc = CeilDiv(nzd, zc);
  const int nitems = ntx * nty * nzc;
  p->grid = std::min(nitems, slots);

  int dimv[3] = {gin->ldim[0], gin->ldim[1], gin->ldim[2]};
  int boxv[3] = {v.full_row ? txb0 : (dbl ? Geom<double>::BW : Geom<float>::BW), v.ty + 2, 1};
  if (!EncodeTensorMap3D(&p->tmap, dbl ? TmaElem::F64 : TmaElem::F32, gin->members[0].dev, dimv,
                         boxv)) {
    *why = "grid shape violates a TMA constraint";
    delete p;
    return nullptr;
  }
  auto common = [&](auto *a) {
    a->xbase = xbase;
    a->ntx = ntx; a->nty = nty; a->nzc = nzc; a->zc = zc; a->nitems = nitems;
    a->stages = stages;
    a->st_hint = o.star7_sthint;
    using ET = typename std::remove_pointer<decltype(a->out)>::type;
    if (SlabPushTargets(rt, *gout, 0, (void **)&a->push_lo, (void **)&a->push_hi, sizeof(ET))) {
      a->push_lo_z = gout->halo;
      a->push_hi_z = gout->halo + gout->nz_loc - 1;
      p->pushes = true;
      // the halo stores are in the kernel, so its ordering with the neighbours can be too
      if (rt->FillSlabSync(&a->sync)) {
        p->syncs = true;
        SlabSyncPlanEnds(&a->sync, o.early_signal != 0, nzd, &a->zc, &a->nzc, ntx * nty, 1, o.slab_zbl);
        a->nitems = ntx * nty * a->nzc;
        p->grid = std::min(a->nitems, slots);
      }
    }
  };
  if (dbl) { FillArgs(&p->ad, d, gin, gout); common(&p->ad); }
  else { FillArgs(&p->af, d, gin, gout); common(&p->af); }
  return p;
}

void LaunchStar7(Runtime *rt, Star7Plan *p) {
  if (p->syncs) {
    SlabSync &sy = p->is_double ? p->ad.sync : p->af.sync;
    sy.wait_epoch = rt->sweep_epoch;
    sy.signal_epoch = rt->sweep_epoch + 1;
  }
  void *args[2];
  args[0] = &p->tmap;
  args[1] = p->is_double ? (void *)&p->ad : (void *)&p->af;
  LaunchSweepKernel(rt, p->fn, p->grid, p->block, args, p->smem);
}

void DestroyStar7(Star7Plan *p) { delete p; }
bool Star7Pushes(const Star7Plan *p) { return p->pushes; }
bool Star7Syncs(const Star7Plan *p) { return p->syncs; }

}  // namespace physis_b200
