// Periodic 7-point update of one member of a user-defined point type with a
// vertex-staggered coefficient grid — hand-written sm_100a kernel (fp64).
//
// One sweep of examples/dsl/diffusion3d_periodic_staggered.c (BASELINE config 5):
//   k   = 0.125 * (kap(x,y,z) + kap(x+1,y,z) + kap(x,y+1,z) + kap(x,y,z+1)
//                  + kap(x+1,y+1,z) + kap(x+1,y,z+1) + kap(x,y+1,z+1) + kap(x+1,y+1,z+1))
//   out = c + k * (w + e + n + s + b + t - 6.0*c)
// where c,w,e,n,s,b,t are PSGridGetPeriodic reads of member `rd` of u and `out`
// is PSGridEmitUtype into member `wr` of the same grid; kap is (N+1)^3.  It
// composes what the reference tests one at a time:
//   tests/system_tests/test_cases/test_user-defined-type-7-pt-periodic.c:19-27,
//   test_7-pt-double-type.c:17-25, examples/test_staggered_grid.c:6-14.
// Arithmetic is separately rounded fp64 in source order (bit-exact vs REF).
//
// Device layout: user types are SoA, so the sweep reads one 8-byte member array
// and writes another: 16 B/LUP + 8 B/LUP for the kap stream = 24 B/LUP (AoS
// would move 40).  Structure as star7.cu — persistent CTAs, TMA ring of xy
// tiles along z, register z-window, shuffles for x-neighbours, 128-bit stores —
// plus the periodic wrap done by the producer: the tile's interior rows, its
// two y-halo rows and (for tiles touching x=0 / x=nx-1) a 16-byte-wide wrap
// column are separate TMA boxes, and z wraps in the plane coordinate.  kap's row
// pitch ((N+1)*8 bytes) is not a multiple of 16, so it cannot be a 3-D TMA tensor;
// it is described as a FLAT 1-D tensor instead and every row segment of a tile's
// (TY+1) x (TXB+1) vertex patch is one 1-D TMA box into the same ring stage as the
// p plane.  A box must start on a 16-byte boundary (measured: an odd fp64 start
// element raises an illegal-instruction error), so each row is fetched from the even
// element at or before its first vertex and the consumers add the row's parity when
// they read.  Each thread takes its 3 x (RY+1) values with 8-byte shared-memory loads
// and carries plane z+1 to z in registers.
#include "runtime.h"
#include "tma.cuh"
#include "sweep_common.cuh"

#include <algorithm>
#include <string>

namespace physis_b200 {

namespace {

using namespace sweep;

struct PstagArgs {
  double *out;
  int nx, ny, nz;     // cell grid
  int kx, ky, kz;     // kap extents (nx+1, ny+1, local planes)
  int dx0, dx1, dy0, dy1, dz0, dz1;
  int ntx, nty, nzc, zc, nitems;
  int stages;
  // z-slab view (multi-GPU): with halo planes in place the z wrap is the ring
  // exchange, not a modulo; the written member's boundary planes are also stored
  // into the neighbours' halo planes through the CUDA-IPC mapping
  int zwrap;
  int push_lo_z, push_hi_z;
  double *push_lo, *push_hi;
  SlabSync sync;  // neighbour ordering fused into the kernel
};

__host__ __device__ constexpr int Up128(int v) { return (v + 127) / 128 * 128; }

// Shared-memory layout of one box of one ring stage.  Every TMA destination
// must be 128-byte aligned, so the pieces are separate regions:
//   [MAIN: TY rows][NORTH halo row][SOUTH halo row][WEST wrap col][EAST wrap col]
template <int TY>
struct PstagLayout {
  static constexpr int ROWB = Geom<double>::ROW_BYTES;
  static constexpr int NORTH = Up128(TY * ROWB);
  static constexpr int SOUTH = NORTH + Up128(ROWB);
  static constexpr int WEST = SOUTH + Up128(ROWB);
  static constexpr int EAST = WEST + Up128(TY * 16);
  // kap patch: TY+1 rows of TXB+2 vertices (TXB+1 used; the box must be a multiple of 16 bytes)
  // kap patch: TY+1 rows of TXB+1 vertices in two boxes of the 2-D super-row view (below):
  // the rows with an even flat row number and those with an odd one, KH rows each
  static constexpr int KAP = EAST + Up128(TY * 16);
  static constexpr int KBOX = Geom<double>::TXB + 4;  // TXB+1 vertices, +1 for an odd start, 16-byte multiple
  static constexpr int KROW = KBOX * 8;               // rows of a box are contiguous in shared memory
  static constexpr int KH = TY / 2 + 1;
  static constexpr int KAP_ODD = KAP + Up128(KH * KROW);
  static constexpr int STRIDE = KAP_ODD + Up128(KH * KROW);
};
template <int TY>
constexpr int PstagBoxStride() { return PstagLayout<TY>::STRIDE; }

// NS ring stages (3 or 6).  The consumers' loop is unrolled six times -- the least common
// multiple of the ring depth, the three-plane register window and the two kap planes -- so
// that every ring slot is an immediate and neither window nor kap registers are ever moved;
// every work item restarts at slot 0 and the phase parity of each slot is tracked on its own.
template <int TY, int RY, int NBX, int MINB, int NS>
__global__ void __launch_bounds__((NBX * (TY / RY) + 1) * 32, MINB)
PstagKernel(const __grid_constant__ CUtensorMap map_main, const __grid_constant__ CUtensorMap map_row,
            const __grid_constant__ CUtensorMap map_col, const __grid_constant__ CUtensorMap map_kap,
            const __grid_constant__ PstagArgs a) {
  using G = Geom<double>;
  constexpr int VEC = 2;
  constexpr int NWY = TY / RY;
  constexpr int NW = NBX * NWY;
  constexpr int ROWB = G::ROW_BYTES;            // 544
  using L = PstagLayout<TY>;
  constexpr int BOX_STRIDE = L::STRIDE;
  constexpr int STAGE_BYTES = NBX * BOX_STRIDE;
  constexpr int KROW = L::KROW;
  static_assert(NS == 3 || NS == 6, "ring depth must divide the unroll factor 6");

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + kMaxStages;
  unsigned char *planes = smem + kBarrierBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      tma::mbar_init(&full[s], 1);
      tma::mbar_init(&empty[s], NW);
    }
    tma::fence_barrier_init();
  }
  SlabSyncWait(a.sync);
  __syncthreads();

  const int tiles_xy = a.ntx * a.nty;

  if (warp == NW) {
    // ------------------------------------------------------------ producer
    if (lane != 0) return;
    tma::prefetch_tensormap(&map_main);
    tma::prefetch_tensormap(&map_row);
    tma::prefetch_tensormap(&map_col);
    tma::prefetch_tensormap(&map_kap);
    uint32_t epar = 0;  // bit s: parity of the next "slot s is free" phase to wait for
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int zseq = item / tiles_xy;
      const int txy = item - zseq * tiles_xy;
      const int ty = txy / a.ntx;
      const int tx = txy - ty * a.ntx;
      const int x0 = tx * (NBX * G::TXB);
      const int y0 = a.dy0 + ty * TY;
      int zb, ze;
      SlabChunkRange(a.sync, zseq, a.nzc, a.zc, a.dz0, a.dz1, &zb, &ze);
      const int yn = (y0 - 1 + a.ny) % a.ny;   // periodic halo rows
      const int ys = (y0 + TY) % a.ny;
      uint32_t tx_bytes = 0;
#pragma unroll
      for (int b = 0; b < NBX; ++b) {
        const int bx0 = x0 + b * G::TXB;
        if (bx0 >= a.nx) continue;
        tx_bytes += (uint32_t)((TY + 2) * ROWB) + (uint32_t)(2 * L::KH * L::KROW);
        if (bx0 == 0) tx_bytes += TY * 16;
        if (bx0 + G::TXB >= a.nx) tx_bytes += TY * 16;
      }
      int stage = 0;  // every item starts at slot 0
      for (int zz = zb - 1; zz <= ze; ++zz) {
        const int z = a.zwrap ? (zz + a.nz) % a.nz : zz;
        tma::mbar_wait(&empty[stage], ((epar >> stage) & 1u) ^ 1u);
        epar ^= 1u << stage;
        tma::mbar_arrive_expect_tx(&full[stage], tx_bytes);
        unsigned char *dst = planes + stage * STAGE_BYTES;
        // kap rows of plane zz (kap does not wrap: it has one more plane than u) as super-rows
        // of the 2-D view: rows with an even / odd flat row number are the first / second halves
        const int zk = min(max(zz, 0), a.kz - 1);
        const int r0 = zk * a.ky + y0;          // flat row number of the patch's first row
        const int re = (r0 + 1) >> 1;           // super-row of the first even row (r0 or r0+1)
        const int ro = r0 >> 1;                 // super-row of the first odd row (r0 or r0+1)
#pragma unroll
        for (int b = 0; b < NBX; ++b) {
          const int bx0 = x0 + b * G::TXB;
          if (bx0 >= a.nx) continue;
          unsigned char *bd = dst + b * BOX_STRIDE;
          tma::load_3d(bd, &map_main, &full[stage], bx0 - G::HX, y0, z);
          tma::load_3d(bd + L::NORTH, &map_row, &full[stage], bx0 - G::HX, yn, z);
          tma::load_3d(bd + L::SOUTH, &map_row, &full[stage], bx0 - G::HX, ys, z);
          if (bx0 == 0)
            tma::load_3d(bd + L::WEST, &map_col, &full[stage], a.nx - G::HX, y0, z);
          if (bx0 + G::TXB >= a.nx)
            tma::load_3d(bd + L::EAST, &map_col, &full[stage], 0, y0, z);
          // even rows start at column bx0, odd rows at column kx + bx0 of their super-row,
          // fetched from the even column at or before it (16-byte aligned)
          tma::load_2d(bd + L::KAP, &map_kap, &full[stage], bx0, re);
          tma::load_2d(bd + L::KAP_ODD, &map_kap, &full[stage], (a.kx + bx0) & ~1, ro);
        }
        if (++stage == NS) stage = 0;
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int bx = warp % NBX;
  const int wy = warp / NBX;
  const int row0 = wy * RY;  // first own row inside the MAIN region
  // this thread's first own vector inside a stage, and the rows above / below its rows
  const unsigned char *my = planes + bx * BOX_STRIDE + row0 * ROWB + (G::HX + lane * VEC) * (int)sizeof(double);
  const int north_off = (wy == 0) ? L::NORTH - row0 * ROWB : -ROWB;
  const int south_off = (wy == NWY - 1) ? L::SOUTH - row0 * ROWB : RY * ROWB;
  const unsigned char *my_kap = planes + bx * BOX_STRIDE + L::KAP + lane * 16;
  const bool lane_first = (lane == 0), lane_last = (lane == 31);
  uint32_t fpar = 0;  // bit s: parity of the next "slot s is full" phase to wait for

#define PS_WAIT(SLOT) do { \
    tma::mbar_wait(&full[(SLOT)], (fpar >> (SLOT)) & 1u); \
    fpar ^= 1u << (SLOT); \
  } while (0)
#define PS_RELEASE(SLOT) do { __syncwarp(); if (lane_first) tma::mbar_arrive(&empty[(SLOT)]); } while (0)
#define PS_LOAD(DST, SLOT) do { \
    const unsigned char *p__ = my + (SLOT) * STAGE_BYTES; \
    _Pragma("unroll") for (int r = 0; r < RY; ++r) DST[r] = *reinterpret_cast<const double2 *>(p__ + r * ROWB); \
  } while (0)
  // this thread's 3 x (RY+1) vertices of kap plane KZ, which arrived in ring slot SLOT: patch
  // row q has flat row number R = KZ*ky + ytile + q; the rows with even R are in the first
  // box, those with odd R in the second (whose first wanted column is `kodd` elements in),
  // and in either box row q is row q/2.  (8-byte loads at a 16-byte lane stride are two-way bank
  // conflicts, 20 M per sweep in ncu; fetching pairs with one 16-byte load and taking the third
  // vertex from the next lane by shuffle removed them but added 50 M instructions and 23
  // registers, and the sweep was no faster -- it is bound by fp64 dependency latency at two
  // consumer warps per scheduler, not by shared-memory bandwidth -- so the simple form stays.)
#define PS_LOAD_KAP(K, SLOT, KZ) do { \
    const unsigned char *kp__ = my_kap + (SLOT) * STAGE_BYTES; \
    const int p0__ = ((KZ) * a.ky + ytile) & 1; \
    _Pragma("unroll") for (int r = 0; r <= RY; ++r) { \
      const int q__ = row0 + r; \
      const int odd__ = (p0__ + q__) & 1; \
      const double *src__ = reinterpret_cast<const double *>( \
          kp__ + odd__ * (L::KAP_ODD - L::KAP) + (q__ >> 1) * KROW) + odd__ * kodd; \
      K[r][0] = src__[0]; K[r][1] = src__[1]; K[r][2] = src__[2]; \
    } \
  } while (0)
  // One plane z: the centre plane is in registers (CEN) and in slot CS, the top plane arrives
  // in slot TS together with kap plane z+1.
#define PS_STEP(CS, TS, BOT, CEN, TOP, KLO, KHI) do { \
    PS_WAIT(TS); \
    PS_LOAD(TOP, TS); \
    PS_LOAD_KAP(KHI, TS, z + 1); \
    const unsigned char *cb = my + (CS) * STAGE_BYTES; \
    const double2 north = *reinterpret_cast<const double2 *>(cb + north_off); \
    const double2 south = *reinterpret_cast<const double2 *>(cb + south_off); \
    _Pragma("unroll") for (int r = 0; r < RY; ++r) { \
      const double2 c = CEN[r]; \
      double wv = __shfl_up_sync(0xffffffffu, c.y, 1); \
      double ev = __shfl_down_sync(0xffffffffu, c.x, 1); \
      if (lane_first) wv = *reinterpret_cast<const double *>(cb + w_off + r * w_str); \
      if (need_e) ev = *reinterpret_cast<const double *>(cb + e_off + r * e_str); \
      const double2 nv = (r == 0) ? north : CEN[r > 0 ? r - 1 : 0]; \
      const double2 sv = (r == RY - 1) ? south : CEN[r < RY - 1 ? r + 1 : r]; \
      const double2 bv = BOT[r], tv = TOP[r]; \
      double2 o; \
      _Pragma("unroll") for (int j = 0; j < VEC; ++j) { \
        const double cj = Elem(c, j); \
        const double wj = (j == 0) ? wv : c.x; \
        const double ej = (j == VEC - 1) ? ev : c.y; \
        /* 0.125 * (k000 + k100 + k010 + k001 + k110 + k101 + k011 + k111) */ \
        double ks = AddRn(KLO[r][j], KLO[r][j + 1]); \
        ks = AddRn(ks, KLO[r + 1][j]); \
        ks = AddRn(ks, KHI[r][j]); \
        ks = AddRn(ks, KLO[r + 1][j + 1]); \
        ks = AddRn(ks, KHI[r][j + 1]); \
        ks = AddRn(ks, KHI[r + 1][j]); \
        ks = AddRn(ks, KHI[r + 1][j + 1]); \
        const double k = MulRn(0.125, ks); \
        /* w + e + n + s + b + t - 6.0*c */ \
        double acc = AddRn(wj, ej); \
        acc = AddRn(acc, Elem(nv, j)); \
        acc = AddRn(acc, Elem(sv, j)); \
        acc = AddRn(acc, Elem(bv, j)); \
        acc = AddRn(acc, Elem(tv, j)); \
        acc = SubRn(acc, MulRn(6.0, cj)); \
        SetElem(o, j, AddRn(cj, MulRn(k, acc))); \
      } \
      if (st_ok[r]) { \
        *reinterpret_cast<double2 *>(obase + (size_t)r * a.nx) = o; \
        if (pushing) { \
          if (z == a.push_lo_z) *reinterpret_cast<double2 *>(a.push_lo + prow + (size_t)r * a.nx) = o; \
          if (z == a.push_hi_z) *reinterpret_cast<double2 *>(a.push_hi + prow + (size_t)r * a.nx) = o; \
        } \
      } \
    } \
    obase += plane_elems; \
    PS_RELEASE(CS); \
  } while (0)

  const size_t plane_elems = (size_t)a.nx * a.ny;
  const bool pushing = (a.push_lo_z >= 0) || (a.push_hi_z >= 0);

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int zseq = item / tiles_xy;
    const int txy = item - zseq * tiles_xy;
    const int ty = txy / a.ntx;
    const int tx = txy - ty * a.ntx;
    const int bx0 = tx * (NBX * G::TXB) + bx * G::TXB;
    const int x = bx0 + lane * VEC;
    const int ybase = a.dy0 + ty * TY + wy * RY;
    int zb, ze;
    SlabChunkRange(a.sync, zseq, a.nzc, a.zc, a.dz0, a.dz1, &zb, &ze);
    const bool x_ok = (x >= a.dx0) && (x + VEC <= a.dx1) && (x < a.nx);
    // x neighbours of lane 0 / lane 31: the box's own x halo, or on the periodic faces the
    // wrap column (16 bytes per row: x = nx-2, nx-1 west of the grid, x = 0, 1 east of it)
    const int own = row0 * ROWB + (G::HX + lane * VEC) * (int)sizeof(double);  // of `my` inside a box
    const bool x_first = (x == 0), x_last = (x + VEC == a.nx);
    const int w_off = x_first ? L::WEST + row0 * 16 + 8 - own : -(int)sizeof(double);
    const int w_str = x_first ? 16 : ROWB;
    const int e_off = x_last ? L::EAST + row0 * 16 - own : VEC * (int)sizeof(double);
    const int e_str = x_last ? 16 : ROWB;
    const bool need_e = lane_last || x_last;  // (the last column need not sit in lane 31)
    bool st_ok[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) st_ok[r] = x_ok && (ybase + r) >= a.dy0 && (ybase + r) < a.dy1;
    const size_t prow = (size_t)ybase * a.nx + x;
    double *obase = a.out + (size_t)zb * plane_elems + prow;
    const int ytile = a.dy0 + ty * TY;  // first row of the tile (= of the kap patch)
    const int kodd = (a.kx + bx0) & 1;  // the odd rows' box starts at the even column at or before theirs

    double2 w0[RY], w1[RY], w2[RY];            // z window of own cells; roles rotate
    double ka[RY + 1][3], kb[RY + 1][3];       // kap planes z and z+1; roles alternate

    // plane zb-1 -> bottom (slot 0), plane zb -> centre (slot 1, with kap plane zb)
    PS_WAIT(0);
    PS_LOAD(w0, 0);
    PS_RELEASE(0);
    PS_WAIT(1);
    PS_LOAD(w1, 1);
    PS_LOAD_KAP(ka, 1, zb);

    int z = zb;
    if (NS == 6) {
      for (;;) {
        PS_STEP(1, 2, w0, w1, w2, ka, kb);
        if (++z >= ze) { PS_RELEASE(2); break; }
        PS_STEP(2, 3, w1, w2, w0, kb, ka);
        if (++z >= ze) { PS_RELEASE(3); break; }
        PS_STEP(3, 4, w2, w0, w1, ka, kb);
        if (++z >= ze) { PS_RELEASE(4); break; }
        PS_STEP(4, 5, w0, w1, w2, kb, ka);
        if (++z >= ze) { PS_RELEASE(5); break; }
        PS_STEP(5, 0, w1, w2, w0, ka, kb);
        if (++z >= ze) { PS_RELEASE(0); break; }
        PS_STEP(0, 1, w2, w0, w1, kb, ka);
        if (++z >= ze) { PS_RELEASE(1); break; }
      }
    } else {
      for (;;) {
        PS_STEP(1, 2, w0, w1, w2, ka, kb);
        if (++z >= ze) { PS_RELEASE(2); break; }
        PS_STEP(2, 0, w1, w2, w0, kb, ka);
        if (++z >= ze) { PS_RELEASE(0); break; }
        PS_STEP(0, 1, w2, w0, w1, ka, kb);
        if (++z >= ze) { PS_RELEASE(1); break; }
        PS_STEP(1, 2, w0, w1, w2, kb, ka);
        if (++z >= ze) { PS_RELEASE(2); break; }
        PS_STEP(2, 0, w1, w2, w0, ka, kb);
        if (++z >= ze) { PS_RELEASE(0); break; }
        PS_STEP(0, 1, w2, w0, w1, kb, ka);
        if (++z >= ze) { PS_RELEASE(1); break; }
      }
    }
    SlabSyncItemDone(a.sync, item, NW * 32, threadIdx.x == 0);
  }
  SlabSyncSignal(a.sync, NW * 32, threadIdx.x == 0);
#undef PS_STEP
#undef PS_LOAD_KAP
#undef PS_LOAD
#undef PS_RELEASE
#undef PS_WAIT
}

// tile shapes (rows, rows per thread, boxes side by side); selected by option pstag_variant
struct PstagVariant {
  int ty, ry, nbx;
  int ctas_per_sm;        // the occupancy the shape is meant for (more resident CTAs measured slower)
  const void *fn6, *fn3;  // ring of 6 / 3 stages
  size_t box_stride;
};
#define PSTAG_VARIANT(TY, RY, NBX, MINB) \
  { TY, RY, NBX, MINB, (const void *)PstagKernel<TY, RY, NBX, MINB, 6>, \
    (const void *)PstagKernel<TY, RY, NBX, MINB, 3>, (size_t)PstagBoxStride<TY>() }
const PstagVariant kPstagVariants[] = {
    PSTAG_VARIANT(4, 2, 1, 4),  // 0: 2 consumer warps, four CTAs per SM (measured best, profiles/r1_tune_pstag_512.csv)
    PSTAG_VARIANT(8, 2, 1, 2),  // 1: 4 consumer warps, two CTAs per SM (within 1-3 % of it)
};
constexpr int kNumPstagVariants = sizeof(kPstagVariants) / sizeof(kPstagVariants[0]);

}  // namespace

struct PstagPlan {
  int grid = 0, block = 0;
  size_t smem = 0;
  CUtensorMap map_main, map_row, map_col, map_kap;
  PstagArgs args;
  const void *fn = nullptr;
  bool pushes = false;
  bool syncs = false;
  int wr_member = 0;
};

// kap's rows have a pitch of (N+1)*8 bytes, not a 16-byte multiple, so kap cannot be a 3-D TMA
// tensor.  Two consecutive rows together are ((N+1)*16 bytes), so the flat array is described
// as a 2-D tensor of "super-rows" of two rows each: one box fetches the rows of a tile's
// vertex patch with an even flat row number (first halves of their super-rows), a second box
// those with an odd one (second halves).  The last super-row may extend one row past the
// array; grid allocations carry slack for that (runtime.cu, kAllocSlack).
static bool EncodeKapMap(CUtensorMap *out, const Grid *kap, int ty) {
  const size_t kx = kap->ldim[0];
  const size_t rows = (size_t)kap->ldim[1] * kap->ldim[2];
  const int dim[2] = {(int)(2 * kx), (int)((rows + 1) / 2)};
  const int box[2] = {PstagLayout<8>::KBOX, ty / 2 + 1};
  return EncodeTensorMap2D(out, TmaElem::F64, kap->members[0].dev, dim, box);
}

PstagPlan *PreparePstag(Runtime *rt, const __PSB200StencilDesc &d, std::string *why) {
  if (d.num_grids != 2) { *why = "expects grids {u, kap}"; return nullptr; }
  Grid *u = Grid::FromHandle(d.grids[0]);
  Grid *kap = Grid::FromHandle(d.grids[1]);
  if (u->num_dims != 3 || kap->num_dims != 3) { *why = "3-D grids only"; return nullptr; }
  if (!u->is_user_type() || kap->type != PS_DOUBLE) { *why = "u must be a user type, kap double"; return nullptr; }
  const int rd = d.members[0], wr = d.members[1];
  if (rd < 0 || wr < 0 || rd >= (int)u->members.size() || wr >= (int)u->members.size() || rd == wr) {
    *why = "needs distinct read/write members"; return nullptr;
  }
  const MemberLayout &mr = u->members[rd], &mw = u->members[wr];
  if (mr.type != PS_DOUBLE || mw.type != PS_DOUBLE || mr.count != 1 || mw.count != 1) {
    *why = "members must be scalar doubles"; return nullptr;
  }
  const int nx = u->ldim[0], ny = u->ldim[1], nz = u->ldim[2];  // local allocation
  for (int i = 0; i < 3; ++i)
    if (kap->dim[i] != u->dim[i] + 1) { *why = "kap must be one larger than u per dimension"; return nullptr; }
  const __PSDomain &dom = d.dom;
  if (nx % 2 != 0 || dom.local_min[0] % 2 != 0 || dom.local_max[0] % 2 != 0) {
    *why = "x extent and domain x-range must be even"; return nullptr;
  }
  int vi = rt->opt.pstag_variant;
  if (vi < 0 || vi >= kNumPstagVariants) vi = 0;
  // grids whose y extent is not a multiple of the chosen tile height try 4 rows
  if (ny % kPstagVariants[vi].ty != 0 && ny % 4 == 0) vi = 0;
  const PstagVariant &V = kPstagVariants[vi];
  const int kTY = V.ty, kRY = V.ry, kNBX = V.nbx;
  if (ny % kTY != 0 || (dom.local_min[1] % kTY) != 0) { *why = "y extent must be a multiple of the tile height"; return nullptr; }
  if (nx < Geom<double>::HX || nz < 1) { *why = "grid too small"; return nullptr; }
  for (int i = 0; i < 3; ++i)
    if (dom.local_min[i] < 0 || dom.local_max[i] > u->ldim[i] || dom.local_max[i] <= dom.local_min[i]) {
      *why = "bad domain"; return nullptr;
    }
  if (dom.local_min[0] != 0) { *why = "domain must start at x = 0"; return nullptr; }

  PstagPlan *p = new PstagPlan();
  // ring depth 6, or 3 where six stages do not fit the SM (or option pstag_stages <= 3)
  int stages = (rt->opt.pstag_stages > 0 && rt->opt.pstag_stages <= 3) ? 3 : 6;
  if (kBarrierBytes + (size_t)stages * kNBX * V.box_stride > 227 * 1024) stages = 3;
  p->fn = stages == 6 ? V.fn6 : V.fn3;
  p->smem = kBarrierBytes + (size_t)stages * kNBX * V.box_stride;
  p->block = (kNBX * (kTY / kRY) + 1) * 32;
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  int occ = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->fn, p->block, p->smem));
  PSB_CHECK(occ > 0, "pstag kernel does not fit on an SM");
  occ = std::min(occ, rt->opt.pstag_occ > 0 ? rt->opt.pstag_occ : V.ctas_per_sm);

  PstagArgs &a = p->args;
  p->wr_member = wr;
  a.out = (double *)mw.dev;
  a.nx = nx; a.ny = ny; a.nz = nz;
  a.kx = kap->ldim[0]; a.ky = kap->ldim[1]; a.kz = kap->ldim[2];
  a.dx0 = dom.local_min[0]; a.dx1 = dom.local_max[0];
  a.dy0 = dom.local_min[1]; a.dy1 = dom.local_max[1];
  a.dz0 = dom.local_min[2]; a.dz1 = dom.local_max[2];
  a.ntx = CeilDiv(a.dx1, (long)kNBX * Geom<double>::TXB);
  a.nty = CeilDiv(a.dy1 - a.dy0, kTY);
  const int nzd = a.dz1 - a.dz0;
  const int slots = rt->sm_count * occ;
  const int tiles = a.ntx * a.nty;
  const int want_chunks = std::max(1, CeilDiv(2L * slots, tiles));
  a.zc = std::min(nzd, std::max(8, CeilDiv(nzd, want_chunks)));
  a.nzc = CeilDiv(nzd, a.zc);
  a.nitems = tiles * a.nzc;
  a.stages = stages;
  p->grid = std::min(a.nitems, slots);
  a.zwrap = u->decomposed ? 0 : 1;
  if (u->decomposed && (kap->halo < 1 || u->halo != kap->halo || kap->z_off != u->z_off)) {
    *why = "decomposed run needs equal halo widths and matching z cuts of u and kap";
    delete p;
    return nullptr;
  }
  a.push_lo_z = a.push_hi_z = -1;
  a.sync = SlabSync{};
  // The exchange of this sweep: in the kernel (halo planes stored to the ring neighbours by the
  // CTAs that compute them, ordering by SlabSync), with the sweep's number published when the
  // whole sweep is done (pstag_push=1, the default), the same with the boundary chunks first
  // and an early signal (pstag_push=2), or copy-based (pstag_push=0: peer copies of the two
  // boundary planes and stream-ordered flags after the kernel).  Measured on 2 GPUs at 512^3 per
  // GPU (profiles/r2_experiments.txt): 0.570 / 0.584 / 0.583 ms per sweep -- with 1024 small tiles
  // per slab and four CTAs per SM, boundary-first ordering costs more than the overlap returns.
  if (rt->opt.pstag_push &&
      SlabPushTargets(rt, *u, wr, (void **)&a.push_lo, (void **)&a.push_hi, sizeof(double))) {
    a.push_lo_z = u->halo;
    a.push_hi_z = u->halo + u->nz_loc - 1;
    p->pushes = true;
    if (rt->FillSlabSync(&a.sync)) {
      p->syncs = true;
      SlabSyncPlanEnds(&a.sync, rt->opt.early_signal != 0 && rt->opt.pstag_push == 2, nzd, &a.zc, &a.nzc,
                       a.ntx * a.nty, 1, rt->opt.slab_zbl);
      a.nitems = tiles * a.nzc;
      p->grid = std::min(a.nitems, slots);
    }
    // timing experiments only (results are wrong): what the exchange costs the kernel
    if (rt->opt.debug_slab & 1) a.push_lo_z = a.push_hi_z = -1;
    if (rt->opt.debug_slab & 2) { a.sync.flags = nullptr; a.sync.done = nullptr; a.sync.boundary_items = 0; }
  }

  int dimv[3] = {nx, ny, nz};
  int box_main[3] = {Geom<double>::BW, kTY, 1};
  int box_row[3] = {Geom<double>::BW, 1, 1};
  int box_col[3] = {Geom<double>::HX, kTY, 1};
  if (!EncodeTensorMap3D(&p->map_main, TmaElem::F64, mr.dev, dimv, box_main) ||
      !EncodeTensorMap3D(&p->map_row, TmaElem::F64, mr.dev, dimv, box_row) ||
      !EncodeTensorMap3D(&p->map_col, TmaElem::F64, mr.dev, dimv, box_col) ||
      !EncodeKapMap(&p->map_kap, kap, kTY)) {
    *why = "grid shape violates a TMA constraint";
    delete p;
    return nullptr;
  }
  return p;
}

void LaunchPstag(Runtime *rt, PstagPlan *p) {
  if (p->syncs) {
    p->args.sync.wait_epoch = rt->sweep_epoch;
    p->args.sync.signal_epoch = rt->sweep_epoch + 1;
  }
  void *args[5] = {&p->map_main, &p->map_row, &p->map_col, &p->map_kap, &p->args};
  PSB_CUDA(cudaLaunchKernel(p->fn, dim3(p->grid), dim3(p->block), args, p->smem, rt->stream));
}

void DestroyPstag(PstagPlan *p) { delete p; }
bool PstagPushes(const PstagPlan *p) { return p->pushes; }
bool PstagSyncs(const PstagPlan *p) { return p->syncs; }

}  // namespace physis_b200
