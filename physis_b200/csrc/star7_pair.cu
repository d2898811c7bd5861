// Two 7-point diffusion sweeps in one pass over HBM (temporal blocking).
//
// PSStencilRun(map(kernel, f1 -> f2), map(kernel, f2 -> f1), iter) — the shape of
// examples/diffusion-benchmark/diffusion3d_physis.c:68-72 — applies the same
// clamped 7-point update twice per iteration.  The single-sweep kernel (star7.cu)
// moves 8 B per point per sweep and sits on the HBM roofline; this kernel computes
// sweep n+1 and sweep n+2 of a tile while the tile is on the SM, so a pair of sweeps
// reads the field once and writes it once (the reference's own experiment in that
// direction: examples/diffusion-benchmark/diffusion3d_cuda_temporal_blocking.cu:78-160).
// Every point still sees exactly the reference's arithmetic (star7_math.cuh), so the
// result is bit-identical to two separate sweeps.
//
// Structure
//  * a tile is NBX boxes wide and H = NWY*RY rows high; a CTA marches it along a z chunk.
//    Rows of up to 4 boxes (512 fp32 / 256 fp64) are one tile wide: whole grid rows, no x
//    halo.  Wider rows (BASELINE config 4: 1024 floats) are cut into x tiles that overlap
//    by one 16-byte vector per seam side: first-sweep values ("s1") are computed on every
//    column of a tile (the outermost column of a seam from a garbage neighbour, which no
//    stored value depends on), second-sweep values are stored for the tile's own columns.
//    In y, s1 is computed on all H rows and second-sweep values are stored for the H-2
//    inner rows; tiles overlap by two rows in y and chunks by two planes in z;
//  * input planes (H+2 rows) arrive by TMA (cp.async.bulk.tensor.3d, zero fill
//    outside the grid) in a 3-slot shared-memory ring, completion on mbarriers.
//    There is no producer warp: one CTA-wide barrier per plane already orders the
//    ring, so thread 0 re-arms a slot right after it;
//  * each thread owns one 16-byte vector in RY rows: the z window of the input
//    lives in registers; s1 planes live in a second 3-slot shared-memory ring, from
//    which the second sweep takes its bottom plane, the rows above / below a
//    thread's rows and the columns next to a box; x neighbours inside a box come
//    from warp shuffles;
//  * 128-bit coalesced stores of the second-sweep rows.
// Clamped faces: a neighbour outside the grid is the centre value, in both sweeps;
// s1 values of rows / planes outside the grid are computed from zero fill and never
// selected.
#include "runtime.h"
#include "tma.cuh"
#include "sweep_common.cuh"
#include "star7_math.cuh"

#include <algorithm>

namespace physis_b200 {

namespace {

using namespace sweep;

constexpr int kPairMaxXTiles = 16;

template <typename T>
struct PairArgs {
  T *out;
  int nx, ny, nz;        // extents of the (local) allocation; nz includes halo planes
  T cc, cw, ce, cs, cn, cb, ct;
  int nty, nzc, zc, nitems;
  // z-slab schedule that lets the exchange overlap the interior (zbl > 0; multi-GPU only):
  // chunk sequence 0 / 1 are the slab's first / last zbl planes -- the only chunks that read halo
  // planes or feed the neighbours' -- then `nlong` chunks of `clong` planes and the rest in
  // chunks of `cshort`.  The launch has one CTA per first-wave item, so the CTAs that took the
  // short boundary chunks take the short interior chunks in the second wave and every CTA
  // marches about the same number of planes.  zbl == 0: nzc equal chunks of zc planes.
  int zbl, nlong, clong, cshort;
  // x tiles: tile t loads columns [tx0[t], tx0[t] + NBX boxes) and stores [txs[t], txe[t])
  int ntx;
  int tx0[kPairMaxXTiles], txs[kPairMaxXTiles], txe[kPairMaxXTiles];
  int st_hint;
  // z-slab view (multi-GPU; on one GPU dz = [0, nz), the faces are planes 0 and nz-1 and
  // nothing is pushed).  Local plane indices throughout.
  int dz0, dz1;            // planes this rank computes
  int zface_lo, zface_hi;  // planes holding the global z faces, or -1 when they are elsewhere
  int zld_lo, zld_hi;      // plane range loads are clamped to
  // the first two / last two planes computed here are also stored, through the CUDA-IPC
  // peer mapping, into the ring neighbours' two halo planes (fused exchange over NVLink)
  // (as byte distances from the plane's own address in `out`: uniform over the kernel)
  int push_lo_z, push_hi_z;
  long long push_lo_delta, push_hi_delta;
  SlabSync sync;           // neighbour ordering fused into the kernel
};

constexpr int kPairSlots = 3;

template <typename T, int NBX, int NWY, int RY>
struct PairGeom {
  static constexpr int H = NWY * RY;
  static constexpr int ROWB = Geom<T>::TXB * (int)sizeof(T);  // 512 bytes: one warp of vectors
  static constexpr int IN_BOX = (H + 2) * ROWB;
  static constexpr int IN_STAGE = NBX * IN_BOX;
  static constexpr int S1_BOX = (H + 2) * ROWB;  // one unused row above and below: edge warps read in bounds
  static constexpr int S1_STAGE = NBX * S1_BOX;
  static constexpr int SMEM = kBarrierBytes + kPairSlots * (IN_STAGE + S1_STAGE);
  static constexpr int THREADS = NBX * NWY * 32;
};

// SLAB: the z-slab form (halo planes stored to the ring neighbours, ordering with them)
// ISO: the six neighbour coefficients are equal; the register planes then hold the
// products c6*value instead of the values (star7_math.cuh), 2 multiplies per point
template <typename T, int NBX, int NWY, int RY, int MINB, int FP, bool SLAB, bool ISO>
__global__ void __launch_bounds__(NBX * NWY * 32, MINB)
Star7PairKernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PairArgs<T> a) {
  using G = Geom<T>;
  using PG = PairGeom<T, NBX, NWY, RY>;
  using V = typename VecOf<T>::type;
  constexpr int VEC = G::VEC;
  constexpr int H = PG::H;
  constexpr int ROWB = PG::ROWB;
  constexpr int IN_BOX = PG::IN_BOX, IN_STAGE = PG::IN_STAGE;
  constexpr int S1_BOX = PG::S1_BOX, S1_STAGE = PG::S1_STAGE;
  constexpr int NS = kPairSlots;
  // the element left of lane 0's vector / right of lane 31's vector is in the adjacent box
  constexpr int WEST_EL = (G::TXB - 1) * (int)sizeof(T);
  constexpr int EAST_EL = -31 * VEC * (int)sizeof(T);

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  unsigned char *in_ring = smem + kBarrierBytes;
  unsigned char *s1_ring = in_ring + NS * IN_STAGE;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool issuer = (threadIdx.x == 0);

  if (issuer) {
    for (int s = 0; s < NS; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_barrier_init();
    tma::prefetch_tensormap(&tmap);
  }
  // Programmatic dependent launch (LaunchStar7Pair): the pass after this one may bring its CTAs
  // onto SMs this pass has left and run the prologue above while this pass's last CTAs finish;
  // nothing of the grids is touched before the pass before this one is complete and visible.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (SLAB) SlabSyncWait(a.sync);  // before any halo plane is read or any peer halo written

  const int bx = warp % NBX;
  const int wy = warp / NBX;
  const int j0 = wy * RY;  // tile row of this thread's first row
  const unsigned char *my_in = in_ring + bx * IN_BOX + (j0 + 1) * ROWB + lane * 16;
  // the s1 ring has the input ring's geometry, so one base address serves both
  static_assert(S1_BOX == IN_BOX, "rings share their box geometry");
  unsigned char *my_s1 = const_cast<unsigned char *>(my_in) + NS * IN_STAGE;
  const bool rd_west = (lane == 0) && (bx > 0);
  const bool rd_east = (lane == 31) && (bx < NBX - 1);
  const size_t plane_elems = (size_t)a.nx * a.ny;
  const T c6 = a.cw;
  const bool need_w = (lane == 0);
  const int tiles_xy = a.nty * a.ntx;

  uint32_t par = 0;  // bit s: phase parity of input slot s
  int pstage = 0;                          // issuer: slot the next plane is loaded into

#define SP_LOAD(DST, SLOT) do { \
    const unsigned char *p__ = my_in + (SLOT) * IN_STAGE; \
    _Pragma("unroll") for (int r = 0; r < RY; ++r) DST[r] = *reinterpret_cast<const V *>(p__ + r * ROWB); \
  } while (0)
#define SP_LOAD_SCALED(DST, SLOT) do { \
    const unsigned char *p__ = my_in + (SLOT) * IN_STAGE; \
    _Pragma("unroll") for (int r = 0; r < RY; ++r) \
      DST[r] = v2::Scale(c6, *reinterpret_cast<const V *>(p__ + r * ROWB)); \
  } while (0)
#define SP_WAIT(SLOT) do { \
    tma::mbar_wait(&full[(SLOT)], (par >> (SLOT)) & 1u); \
    par ^= 1u << (SLOT); \
  } while (0)
  // planes beyond the z faces are the face planes themselves: the clamp of the first
  // sweep costs nothing in the loop
#define SP_ISSUE(yin, zpl) do { \
    tma::mbar_arrive_expect_tx(&full[pstage], tx_bytes); \
    unsigned char *dst__ = in_ring + pstage * IN_STAGE; \
    const int zz__ = min(max((zpl), a.zld_lo), a.zld_hi); \
    _Pragma("unroll") for (int b = 0; b < NBX; ++b) \
      if (xt0 + b * G::TXB < a.nx) \
        tma::load_3d(dst__ + b * IN_BOX, &tmap, &full[pstage], xt0 + b * G::TXB, (yin), zz__); \
    if (++pstage == NS) pstage = 0; \
  } while (0)

  // Clamped y faces without selects in the row loop: the one row of a thread's register
  // planes that lies just outside the grid is overwritten with the face row next to it, so
  // the in-thread neighbour of the face row is the face row itself; the rows above / below
  // a thread's rows (read from shared memory) are replaced the same way.  Only warps that
  // touch a y face do anything (y_fix is 0 elsewhere; the loop keeps it a real branch).
#define SP_YFIX(PL) do { \
    for (int e__ = y_fix; e__ > 0; --e__) { \
      const int r_north = -ybase, r_south = a.ny - 1 - ybase; \
      _Pragma("unroll") for (int r = 1; r < RY; ++r) if (r == r_north) PL[r - 1] = PL[r]; \
      _Pragma("unroll") for (int r = 0; r < RY - 1; ++r) if (r == r_south) PL[r + 1] = PL[r]; \
    } \
  } while (0)
#define SP_YFIX_NS(NORTH, SOUTH, PL) do { \
    for (int e__ = y_fix; e__ > 0; --e__) { \
      if (ybase == 0) NORTH = PL[0]; \
      if (ybase + RY == a.ny) SOUTH = PL[RY - 1]; \
    } \
  } while (0)

  // One plane.  PH = iteration number mod 3 fixes every ring slot at compile time:
  // the input plane k+2 arrives in slot (PH+2)%3, plane k+1 (the centre of the first
  // sweep) sits in slot (PH+1)%3, the first-sweep plane k+1 is written to s1 slot PH, the
  // second sweep reads s1 plane k from slot (PH+2)%3 and plane k-1 from slot (PH+1)%3.
  // BOT/CEN/TOP and C1/T1 name register sets whose roles rotate with PH.
#define SP_STEP(PH, BOT, CEN, TOP, C1, T1) do { \
    SP_WAIT(((PH) + 2) % 3); \
    SP_LOAD(TOP, ((PH) + 2) % 3); \
    SP_YFIX(TOP); \
    { \
      const unsigned char *cb = my_in + (((PH) + 1) % 3) * IN_STAGE; \
      V north = *reinterpret_cast<const V *>(cb - ROWB); \
      V south = *reinterpret_cast<const V *>(cb + RY * ROWB); \
      SP_YFIX_NS(north, south, CEN); \
      unsigned char *sp = my_s1 + (PH) * S1_STAGE; \
      _Pragma("unroll") for (int r = 0; r < RY; ++r) { \
        const V c = CEN[r]; \
        const T wv = *reinterpret_cast<const T *>(cb + r * ROWB + w_in); \
        const T ev = *reinterpret_cast<const T *>(cb + r * ROWB + e_in); \
        const V nv = (r == 0) ? north : CEN[r > 0 ? r - 1 : 0]; \
        const V sv = (r == RY - 1) ? south : CEN[r < RY - 1 ? r + 1 : r]; \
        T1[r] = v2::Vec7<FP>(a, c, wv, ev, sv, nv, BOT[r], TOP[r]); \
        *reinterpret_cast<V *>(sp + r * ROWB) = T1[r]; \
      } \
      SP_YFIX(T1); \
    } \
    __syncthreads(); \
    if (issuer) { \
      /* thread 0 owns tile row 0: the box starts one row above it */ \
      if (k == k0 && k + 3 <= ze + 1) SP_ISSUE(ybase - 1, k + 3); \
      if (k + 4 <= ze + 1) SP_ISSUE(ybase - 1, k + 4); \
    } \
    if (k >= zb) { \
      const unsigned char *cb = my_s1 + (((PH) + 2) % 3) * S1_STAGE; \
      /* bottom plane: s1 plane k-1, or (z face) the centre plane itself */ \
      const unsigned char *bb = (k == a.zface_lo) ? cb : my_s1 + (((PH) + 1) % 3) * S1_STAGE; \
      V north = *reinterpret_cast<const V *>(cb - ROWB); \
      V south = *reinterpret_cast<const V *>(cb + RY * ROWB); \
      SP_YFIX_NS(north, south, C1); \
      /* top plane beyond the z face: the centre plane (rare: a loop keeps it a branch) */ \
      for (int e__ = (k == a.zface_hi) ? 1 : 0; e__ > 0; --e__) { \
        _Pragma("unroll") for (int r = 0; r < RY; ++r) T1[r] = C1[r]; \
      } \
      _Pragma("unroll") for (int r = 0; r < RY; ++r) { \
        const V c = C1[r]; \
        const T wv = *reinterpret_cast<const T *>(cb + r * ROWB + w_in); \
        const T ev = *reinterpret_cast<const T *>(cb + r * ROWB + e_in); \
        const V nv = (r == 0) ? north : C1[r > 0 ? r - 1 : 0]; \
        const V sv = (r == RY - 1) ? south : C1[r < RY - 1 ? r + 1 : r]; \
        const V bv = *reinterpret_cast<const V *>(bb + r * ROWB); \
        const V o = v2::Vec7<FP>(a, c, wv, ev, sv, nv, bv, T1[r]); \
        if (st_ok[r]) StoreVec(reinterpret_cast<V *>(obase + (size_t)r * a.nx), o, a.st_hint != 0); \
      } \
      obase += plane_elems; \
    } \
  } while (0)

#define SP_STEP_ISO(PH, BOT, CEN, TOP, C1, T1) do { \
    SP_WAIT(((PH) + 2) % 3); \
    SP_LOAD_SCALED(TOP, ((PH) + 2) % 3); \
    SP_YFIX(TOP); \
    { \
      const unsigned char *cb = my_in + (((PH) + 1) % 3) * IN_STAGE; \
      V north = v2::Scale(c6, *reinterpret_cast<const V *>(cb - ROWB)); \
      V south = v2::Scale(c6, *reinterpret_cast<const V *>(cb + RY * ROWB)); \
      SP_YFIX_NS(north, south, CEN); \
      unsigned char *sp = my_s1 + (PH) * S1_STAGE; \
      _Pragma("unroll") for (int r = 0; r < RY; ++r) { \
        const V pc = CEN[r]; \
        const V c = *reinterpret_cast<const V *>(cb + r * ROWB); \
        const V q = v2::Scale(a.cc, c); \
        T wv = __shfl_up_sync(0xffffffffu, v2::Last(c), 1); \
        T ev = __shfl_down_sync(0xffffffffu, v2::First(c), 1); \
        if (need_w) wv = *reinterpret_cast<const T *>(cb + r * ROWB + w_edge); \
        if (need_e) ev = *reinterpret_cast<const T *>(cb + r * ROWB + e_edge); \
        const T wP = MulRn(c6, wv), eP = MulRn(c6, ev); \
        const V nv = (r == 0) ? north : CEN[r > 0 ? r - 1 : 0]; \
        const V sv = (r == RY - 1) ? south : CEN[r < RY - 1 ? r + 1 : r]; \
        const V s1v = v2::Sum7<FP>(q, wP, eP, pc, sv, nv, BOT[r], TOP[r]); \
        *reinterpret_cast<V *>(sp + r * ROWB) = s1v; \
        T1[r] = v2::Scale(c6, s1v); \
      } \
      SP_YFIX(T1); \
    } \
    __syncthreads(); \
    if (issuer) { \
      /* thread 0 owns tile row 0: the box starts one row above it */ \
      if (k == k0 && k + 3 <= ze + 1) SP_ISSUE(ybase - 1, k + 3); \
      if (k + 4 <= ze + 1) SP_ISSUE(ybase - 1, k + 4); \
    } \
    if (k >= zb) { \
      const unsigned char *cb = my_s1 + (((PH) + 2) % 3) * S1_STAGE; \
      /* bottom plane: s1 plane k-1, or (z face) the centre plane itself */ \
      const unsigned char *bb = (k == a.zface_lo) ? cb : my_s1 + (((PH) + 1) % 3) * S1_STAGE; \
      V north = v2::Scale(c6, *reinterpret_cast<const V *>(cb - ROWB)); \
      V south = v2::Scale(c6, *reinterpret_cast<const V *>(cb + RY * ROWB)); \
      SP_YFIX_NS(north, south, C1); \
      /* top plane beyond the z face: the centre plane (rare: a loop keeps it a branch) */ \
      for (int e__ = (k == a.zface_hi) ? 1 : 0; e__ > 0; --e__) { \
        _Pragma("unroll") for (int r = 0; r < RY; ++r) T1[r] = C1[r]; \
      } \
      _Pragma("unroll") for (int r = 0; r < RY; ++r) { \
        const V pc = C1[r]; \
        const V c = *reinterpret_cast<const V *>(cb + r * ROWB); \
        const V q = v2::Scale(a.cc, c); \
        T wv = __shfl_up_sync(0xffffffffu, v2::Last(c), 1); \
        T ev = __shfl_down_sync(0xffffffffu, v2::First(c), 1); \
        if (need_w) wv = *reinterpret_cast<const T *>(cb + r * ROWB + w_edge); \
        if (need_e) ev = *reinterpret_cast<const T *>(cb + r * ROWB + e_edge); \
        const T wP = MulRn(c6, wv), eP = MulRn(c6, ev); \
        const V nv = (r == 0) ? north : C1[r > 0 ? r - 1 : 0]; \
        const V sv = (r == RY - 1) ? south : C1[r < RY - 1 ? r + 1 : r]; \
        const V bv = v2::Scale(c6, *reinterpret_cast<const V *>(bb + r * ROWB)); \
        const V o = v2::Sum7<FP>(q, wP, eP, pc, sv, nv, bv, T1[r]); \
        if (st_ok[r]) StoreVec(reinterpret_cast<V *>(obase + (size_t)r * a.nx), o, a.st_hint != 0); \
      } \
      obase += plane_elems; \
    } \
  } while (0)

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    // x tiles of one (y tile, z chunk) are neighbours in the item order: they run at the same
    // time and share their seam columns through L2
    const int zseq = item / tiles_xy;
    const int zci = SLAB ? SlabChunkOrder(a.sync, zseq, a.nzc) : zseq;
    const int txy = item - zseq * tiles_xy;
    const int ty = txy / a.ntx;
    const int tx = txy - ty * a.ntx;
    const int xt0 = a.tx0[tx];                      // first column the tile loads
    const int x = xt0 + bx * G::TXB + lane * VEC;   // this thread's vector
    const bool x_ok = (x >= a.txs[tx]) && (x + VEC <= a.txe[tx]);
    const bool x_first = (x == 0);
    const bool x_last = (x + VEC == a.nx);
    int nbox = 0;
#pragma unroll
    for (int b = 0; b < NBX; ++b) nbox += (xt0 + b * G::TXB < a.nx) ? 1 : 0;
    const uint32_t tx_bytes = (uint32_t)nbox * (uint32_t)IN_BOX;
    // x neighbours of a thread's vector come from shared memory (the centre plane of either
    // sweep is there): the element before / after the vector, which for lane 0 / lane 31
    // lives in the adjacent box and on a clamped x face is the vector's own edge element.
    // One scalar load per side, no shuffles, no predicates.  (At a seam between x tiles the
    // outermost lanes read a neighbour that is not theirs -- the row above's last element, or
    // lane 0's by the shuffle: it reaches only the seam column's s1 value, which nothing stored
    // depends on.)
    const int w_in = x_first ? 0 : rd_west ? -IN_BOX + WEST_EL : -(int)sizeof(T);
    const int e_in = x_last ? (VEC - 1) * (int)sizeof(T) : rd_east ? IN_BOX + EAST_EL : VEC * (int)sizeof(T);
    // ISO form: the raw x neighbours travel between lanes by shuffle (its shared-memory
    // pipe is the busier one), only the lanes at a box edge load.  The offsets are made
    // opaque so that they stay in registers instead of being recomputed under a branch.
    const bool need_e = x_last || rd_east;
    int w_edge = w_in, e_edge = e_in;
    if (ISO) asm volatile("" : "+r"(w_edge), "+r"(e_edge));
    int zb = a.dz0 + zci * a.zc;
    int ze = min(zb + a.zc, a.dz1);
    if (SLAB && a.zbl > 0) {
      if (zseq < 2) {
        zb = zseq == 0 ? a.dz0 : a.dz1 - a.zbl;
        ze = zb + a.zbl;
      } else {
        const int i = zseq - 2;
        zb = a.dz0 + a.zbl + (i < a.nlong ? i * a.clong : a.nlong * a.clong + (i - a.nlong) * a.cshort);
        ze = min(zb + (i < a.nlong ? a.clong : a.cshort), a.dz1 - a.zbl);
      }
    }
    const int k0 = (zb == a.zface_lo) ? zb - 1 : zb - 2;  // first plane of the input window
    const int y1 = ty * (H - 2) - 1;      // grid row of tile row 0
    const int ybase = y1 + j0;
    // warps holding a y face row (y == 0 or y == ny-1) fix their planes up (warp-uniform)
    const int y_fix = ((ybase <= 0 && ybase > -RY) || (ybase <= a.ny - 1 && ybase + RY > a.ny - 1)) ? 1 : 0;
    bool st_ok[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int j = j0 + r;
      const int y = ybase + r;
      st_ok[r] = x_ok && j >= 1 && j <= H - 2 && y >= 0 && y < a.ny;
    }
    // second-sweep row 0 of this thread in plane zb (never dereferenced outside the grid)
    T *obase = a.out + (ptrdiff_t)zb * (ptrdiff_t)plane_elems + (ptrdiff_t)ybase * a.nx + x;

    // every thread is done with both rings of the previous item; slots restart at 0
    __syncthreads();
    if (issuer) {
      pstage = 0;
      for (int n = 0; n < NS; ++n) SP_ISSUE(ybase - 1, k0 + n);
    }

    V w0[RY], w1[RY], w2[RY];  // input window (own cells); roles rotate, registers do not move
    V q0[RY], q1[RY], q2[RY];  // first-sweep planes (own cells); two of the three are live
    SP_WAIT(0);
    if (ISO) SP_LOAD_SCALED(w0, 0); else SP_LOAD(w0, 0);
    SP_YFIX(w0);
    SP_WAIT(1);
    if (ISO) SP_LOAD_SCALED(w1, 1); else SP_LOAD(w1, 1);
    SP_YFIX(w1);
#pragma unroll
    for (int r = 0; r < RY; ++r) q2[r] = w1[r];  // defined value; selected by no store

    int k = k0;
    if (ISO) {
      for (;;) {
        SP_STEP_ISO(0, w0, w1, w2, q2, q0);
        if (++k >= ze) break;
        SP_STEP_ISO(1, w1, w2, w0, q0, q1);
        if (++k >= ze) break;
        SP_STEP_ISO(2, w2, w0, w1, q1, q2);
        if (++k >= ze) break;
      }
    } else {
      for (;;) {
        SP_STEP(0, w0, w1, w2, q2, q0);
        if (++k >= ze) break;
        SP_STEP(1, w1, w2, w0, q0, q1);
        if (++k >= ze) break;
        SP_STEP(2, w2, w0, w1, q1, q2);
        if (++k >= ze) break;
      }
    }
    if (SLAB) {
      // The slab's first two / last two planes also go to the ring neighbours' halo planes: once
      // the item is done, every thread forwards the vectors it stored itself for those planes
      // (its own stores: no fence needed to read them back; they come from L2).  Kept out of the
      // plane loop on purpose: any slab-specific code inside it made the compiler emit a larger
      // and slower loop (+13 % per pass on ONE GPU with nothing to exchange).
      for (int s = 0; s < 4; ++s) {
        const int z = (s < 2 ? a.push_lo_z : a.push_hi_z) + (s & 1);
        const long long delta = s < 2 ? a.push_lo_delta : a.push_hi_delta;
        if (z < zb || z >= ze) continue;
        const T *src = a.out + (ptrdiff_t)z * (ptrdiff_t)plane_elems + (ptrdiff_t)ybase * a.nx + x;
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          if (st_ok[r]) {
            const V v = *reinterpret_cast<const V *>(src + (size_t)r * a.nx);
            *reinterpret_cast<V *>(reinterpret_cast<char *>(const_cast<T *>(src + (size_t)r * a.nx)) + delta) = v;
          }
        }
      }
      SlabSyncItemDone(a.sync, item, PG::THREADS, threadIdx.x == 0);
    }
  }
  if (SLAB) SlabSyncSignal(a.sync, PG::THREADS, threadIdx.x == 0);
#undef SP_STEP
#undef SP_STEP_ISO
#undef SP_LOAD_SCALED
#undef SP_YFIX_NS
#undef SP_YFIX
#undef SP_ISSUE
#undef SP_WAIT
#undef SP_LOAD
}

// ------------------------------------------------------------------ host side

struct PairVariant {
  int nbx, nwy, ry, minb;
  const void *f32[2][2];  // [one GPU / z-slab][general / equal neighbour coefficients], packed adds
  const void *f64[2][2];
  int smem_f32, smem_f64, threads;
};

#define PAIR_VARIANT(NBX, NWY, RY, MINB) \
  { NBX, NWY, RY, MINB, \
    {{(const void *)Star7PairKernel<float, NBX, NWY, RY, MINB, 1, false, false>, \
      (const void *)Star7PairKernel<float, NBX, NWY, RY, MINB, 1, false, true>}, \
     {(const void *)Star7PairKernel<float, NBX, NWY, RY, MINB, 1, true, false>, \
      (const void *)Star7PairKernel<float, NBX, NWY, RY, MINB, 1, true, true>}}, \
    {{(const void *)Star7PairKernel<double, NBX, NWY, RY, MINB, 0, false, false>, \
      (const void *)Star7PairKernel<double, NBX, NWY, RY, MINB, 0, false, true>}, \
     {(const void *)Star7PairKernel<double, NBX, NWY, RY, MINB, 0, true, false>, \
      (const void *)Star7PairKernel<double, NBX, NWY, RY, MINB, 0, true, true>}}, \
    PairGeom<float, NBX, NWY, RY>::SMEM, PairGeom<double, NBX, NWY, RY>::SMEM, NBX * NWY * 32 }

const PairVariant kPairVariants[] = {
    PAIR_VARIANT(4, 4, 4, 1),  // 0: rows of up to 4 boxes, 16-row tiles
    PAIR_VARIANT(3, 4, 4, 1),  // 1
    PAIR_VARIANT(2, 4, 4, 2),  // 2
    PAIR_VARIANT(1, 4, 4, 4),  // 3
    PAIR_VARIANT(3, 5, 4, 1),  // 4: x tiles of rows wider than 4 boxes, 20-row tiles (15 warps)
};
constexpr int kPairWideVariant = 4;
constexpr int kNumPairVariants = sizeof(kPairVariants) / sizeof(kPairVariants[0]);

}  // namespace

struct Star7PairPlan {
  bool is_double = false;
  bool iso = false;
  int grid = 0, block = 0;
  size_t smem = 0;
  const void *fn = nullptr;
  CUtensorMap tmap[2];       // direction 0 reads grid A, direction 1 reads grid B
  PairArgs<float> af[2];
  PairArgs<double> ad[2];
};

// d0: A -> B, d1: B -> A, both the clamped 7-point update with the same scalars over
// the whole grid.  Returns nullptr (with a reason) when the pair cannot be fused.
Star7PairPlan *PrepareStar7Pair(Runtime *rt, const __PSB200StencilDesc &d0,
                                const __PSB200StencilDesc &d1, std::string *why) {
  const Options &o = rt->opt;
  if (!o.star7_fuse) { *why = "star7_fuse=0"; return nullptr; }
  if (d0.kind != PSB200_KIND_DIFFUSION7_CLAMP || d1.kind != PSB200_KIND_DIFFUSION7_CLAMP) {
    *why = "not a pair of clamped 7-point sweeps"; return nullptr;
  }
  if (d0.num_grids != 2 || d1.num_grids != 2 || d0.num_scalars != 7 || d1.num_scalars != 7) {
    *why = "expects 2 grids and 7 scalars"; return nullptr;
  }
  if (d0.grids[0] != d1.grids[1] || d0.grids[1] != d1.grids[0] || d0.grids[0] == d0.grids[1]) {
    *why = "the sweeps do not ping-pong between two grids"; return nullptr;
  }
  for (int i = 0; i < 7; ++i)
    if (d0.scalars[i] != d1.scalars[i]) { *why = "the sweeps use different coefficients"; return nullptr; }
  Grid *ga = Grid::FromHandle(d0.grids[0]);
  Grid *gb = Grid::FromHandle(d0.grids[1]);
  if (ga->num_dims != 3 || gb->num_dims != 3 || ga->is_user_type() || gb->is_user_type() ||
      ga->type != gb->type || (ga->type != PS_FLOAT && ga->type != PS_DOUBLE)) {
    *why = "float or double 3-D grids of one type"; return nullptr;
  }
  for (int i = 0; i < 3; ++i) {
    if (ga->dim[i] != gb->dim[i]) { *why = "grids must have equal extents"; return nullptr; }
    if (d0.dom.local_min[i] != 0 || d0.dom.local_max[i] != ga->dim[i] ||
        d1.dom.local_min[i] != 0 || d1.dom.local_max[i] != ga->dim[i]) {
      *why = "both sweeps must cover the whole grid"; return nullptr;
    }
  }
  const bool multi = rt->world() > 1;
  sweep::SlabSync sync{};
  if (multi) {
    // a fused pass consumes two halo planes per side and delivers two; every rank must
    // take the same decision, so only group-wide quantities enter it
    if (!ga->decomposed || !gb->decomposed || ga->halo < 2 || gb->halo != ga->halo ||
        ga->z_off != gb->z_off || ga->nz_loc != gb->nz_loc) {
      *why = "z-slabs need two halo planes (option halo>=2) and identical cuts"; return nullptr;
    }
    if (ga->dim[2] / rt->world() < 4) { *why = "z-slabs thinner than four planes"; return nullptr; }
    if (!o.halo_push || !rt->FillSlabSync(&sync)) {
      *why = "needs the in-kernel halo exchange (halo_push=1, sync_mode=2)"; return nullptr;
    }
  }
  const bool dbl = ga->type == PS_DOUBLE;
  const int vec = dbl ? 2 : 4;
  const int txb = dbl ? Geom<double>::TXB : Geom<float>::TXB;
  const int nx = ga->dim[0], ny = ga->dim[1];
  const int nz = ga->nz_loc;         // planes this rank computes
  const int nz_alloc = ga->ldim[2];  // ... out of this many allocated (halo planes included)
  if (nx % vec != 0) { *why = "x extent must be a multiple of 16 bytes"; return nullptr; }
  const int boxes = CeilDiv(nx, txb);
  int variant = -1;
  for (int v = 0; v < kPairWideVariant; ++v)
    if (kPairVariants[v].nbx == boxes) variant = v;
  // rows wider than 4 boxes: x tiles of 3 boxes that overlap by one vector per seam side, the
  // row cut into as few equal segments as fit (1024 floats: 3 tiles storing 344 + 340 + 340)
  int ntx = 1, seg = nx;
  if (variant < 0) {
    if (!o.star7_pair_xtile) { *why = "rows wider than the fused kernel's tile (star7_pair_xtile=0)"; return nullptr; }
    variant = o.star7_pair_variant >= 0 && o.star7_pair_variant < kNumPairVariants ? o.star7_pair_variant
                                                                                   : kPairWideVariant;
    const int w = kPairVariants[variant].nbx * txb;
    for (ntx = 2; ntx <= kPairMaxXTiles; ++ntx) {
      seg = CeilDiv(CeilDiv(nx, ntx), vec) * vec;
      if (seg + 2 * vec <= w) break;
    }
    if (ntx <= kPairMaxXTiles && o.star7_pair_variant < 0 && o.star7_pair_zc <= 0) {
      // 20-row tiles (15 warps) or 16-row tiles of the same width (12 warps; a plane step takes
      // 0.85 of the time instead of 0.8: measured, profiles/r2_experiments.txt): whichever
      // tiling wastes less of the last wave of work items -- on 1024x1024x128 (an eighth of
      // 1024^3) 222 tiles x 2 chunks fill three waves exactly where 171 tiles x 6 chunks leave
      // 7 % of the seventh idle and pay for three times the chunk overlap
      auto plan_cost = [&](int vi, double step) {
        const int hh = kPairVariants[vi].nwy * kPairVariants[vi].ry;
        const long tiles = (long)CeilDiv(ny, hh - 2) * ntx;
        double best = -1.0;
        for (int nzc = 1; nzc <= std::max(1, nz / 4); ++nzc) {
          const int c = CeilDiv(nz, nzc);
          const long waves = CeilDiv(tiles * CeilDiv(nz, c), (long)rt->sm_count);
          const double cost = (double)waves * (c + 3) * step;
          if (best < 0 || cost < best) best = cost;
        }
        return best;
      };
      static_assert(kPairWideVariant == 4, "the 16-row twin of the wide variant is variant 1");
      if (plan_cost(1, 0.85) < plan_cost(kPairWideVariant, 1.0)) variant = 1;
    }
    if (ntx > kPairMaxXTiles) { *why = "rows wider than 16 x tiles of the fused kernel"; return nullptr; }
  }
  if (nz < 2 || ny < 2) { *why = "grid too thin"; return nullptr; }
  const PairVariant &v = kPairVariants[variant];

  Star7PairPlan *p = new Star7PairPlan();
  p->is_double = dbl;
  // equal neighbour coefficients (the benchmark's isotropic case): 2 multiplies per point
  bool iso = o.star7_iso != 0;
  for (int i = 1; i < 6 && iso; ++i) {
    if (dbl) iso = (d0.scalars[i] == d0.scalars[0]);
    else iso = ((float)d0.scalars[i] == (float)d0.scalars[0]);
  }
  // (debug_slab bit 2: the z-slab form of the kernel on one GPU, with nothing to exchange --
  // a timing experiment that separates the form's code from the exchange itself)
  const int ms = (multi || (o.debug_slab & 4)) ? 1 : 0;
  p->fn = dbl ? v.f64[ms][iso ? 1 : 0] : v.f32[ms][iso ? 1 : 0];
  p->iso = iso;
  p->smem = dbl ? v.smem_f64 : v.smem_f32;
  p->block = v.threads;
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->fn, p->block, p->smem));
  PSB_CHECK(occ > 0, "fused star7 kernel does not fit on an SM");
  const int slots = rt->sm_count * occ;
  const int h = v.nwy * v.ry;
  const int nty = CeilDiv(ny, h - 2);
  // z chunk: every chunk re-reads 4 planes and recomputes 2 first-sweep planes, every
  // wave of items costs a chunk; take the chunk count with the least total plane work
  int zc = o.star7_pair_zc;
  if (zc <= 0) {
    long best_cost = -1;
    for (int nzc = 1; nzc <= std::max(1, nz / 4); ++nzc) {
      const int c = CeilDiv(nz, nzc);
      const long waves = CeilDiv((long)nty * ntx * CeilDiv(nz, c), slots);
      const long cost = waves * (c + 3);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; zc = c; }
    }
  }
  zc = std::max(1, std::min(zc, nz));
  int nzc = CeilDiv(nz, zc);
  int nitems = nty * ntx * nzc;
  p->grid = std::min(nitems, slots);
  // Multi-GPU: when whole groups of tiles fit the CTA slots, run the slab's two ends as short
  // chunks of their own in the first wave, so that the pass number is published a fraction of a
  // pass after the launch and the neighbours' next pass never waits (with equal chunks filling
  // one wave exactly, the "boundary" chunks finish with everything else and nothing overlaps)
  int zbl = 0, nlong = 0, clong = 0, cshort = 0;
  if (multi && o.early_signal && o.star7_pair_zc <= 0 && o.star7_pair_zbl >= 2) {
    const int tiles = nty * ntx;
    const int groups = slots / tiles;
    if (groups >= 2 && tiles * groups * 10 >= slots * 9) {
      const int b = o.star7_pair_zbl;
      // a boundary CTA marches (b + 3) + (cs + 3) planes and also pays for the peer stores and the
      // system-scope fence before it signals (`bias` planes' worth), an interior one cl + 3: make
      // them equal
      const int bias = std::max(0, o.star7_pair_zbias);
      const int cs = (nz - 2 * b - (groups - 2) * (b + 3 + bias)) / groups;
      if (cs >= 4) {
        zbl = b;
        nlong = groups - 2;
        clong = cs + b + 3 + bias;
        cshort = CeilDiv(nz - 2 * b - nlong * clong, 2);
        nzc = groups + 2;
        nitems = tiles * nzc;
        p->grid = tiles * groups;
      }
    }
  }

  Grid *gin[2] = {ga, gb};
  Grid *gout[2] = {gb, ga};
  for (int dir = 0; dir < 2; ++dir) {
    int dimv[3] = {nx, ny, nz_alloc};
    int boxv[3] = {txb, h + 2, 1};
    if (!EncodeTensorMap3D(&p->tmap[dir], dbl ? TmaElem::F64 : TmaElem::F32, gin[dir]->members[0].dev,
                           dimv, boxv)) {
      *why = "grid shape violates a TMA constraint";
      delete p;
      return nullptr;
    }
    auto fill = [&](auto *a) {
      using ET = typename std::remove_pointer<decltype(a->out)>::type;
      a->out = (ET *)gout[dir]->members[0].dev;
      a->nx = nx; a->ny = ny; a->nz = nz_alloc;
      a->dz0 = ga->halo; a->dz1 = ga->halo + nz;
      a->zface_lo = ga->LocalInterior(0);
      a->zface_hi = ga->LocalInterior(ga->dim[2] - 1);
      a->zld_lo = a->zface_lo >= 0 ? a->zface_lo : 0;
      a->zld_hi = a->zface_hi >= 0 ? a->zface_hi : nz_alloc - 1;
      a->push_lo_z = a->push_hi_z = -(1 << 30);
      a->push_lo_delta = a->push_hi_delta = 0;
      a->sync = sync;
      // every chunk that reads a halo plane or computes one of the two planes per side the
      // neighbours receive finishes before the pass number is published
      if (multi) SlabSyncSetBoundary(&a->sync, o.early_signal != 0, nz, zc, nzc, nty * ntx, 2);
      if (zbl > 0) {
        a->sync.boundary_items = 2 * nty * ntx;  // chunk sequence 0 and 1
        a->sync.nb_lo = a->sync.nb_hi = 0;
      }
      a->zbl = zbl; a->nlong = nlong; a->clong = clong; a->cshort = cshort;
      // timing experiments only (results are wrong): what the exchange costs the kernel
      if (o.debug_slab & 1) { a->push_lo_z = a->push_hi_z = -(1 << 30); a->push_lo_delta = a->push_hi_delta = 0; }
      if (o.debug_slab & 2) { a->sync.flags = nullptr; a->sync.done = nullptr; a->sync.boundary_items = 0; }
      if (multi) {
        const Grid *go = gout[dir];
        const MemberLayout &ml = go->members[0];
        const size_t plane = (size_t)go->plane_elms;
        // lower neighbour's two upper halo planes <- my first two planes; upper neighbour's
        // two lower halo planes (the ones next to its interior) <- my last two planes
        a->push_lo_z = a->dz0;
        a->push_hi_z = a->dz1 - 2;
        const ET *to_lo = (ET *)ml.peer_lo + (size_t)(go->halo + go->lo_nz_loc) * plane;
        const ET *to_hi = (ET *)ml.peer_hi + (size_t)(go->halo - 2) * plane;
        a->push_lo_delta = (const char *)to_lo - (const char *)(a->out + (size_t)a->push_lo_z * plane);
        a->push_hi_delta = (const char *)to_hi - (const char *)(a->out + (size_t)a->push_hi_z * plane);
      }
      // scalars arrive in the kernel's parameter order: ce, cw, cn, cs, ct, cb, cc
      a->ce = (ET)d0.scalars[0]; a->cw = (ET)d0.scalars[1]; a->cn = (ET)d0.scalars[2];
      a->cs = (ET)d0.scalars[3]; a->ct = (ET)d0.scalars[4]; a->cb = (ET)d0.scalars[5];
      a->cc = (ET)d0.scalars[6];
      a->nty = nty; a->nzc = nzc; a->zc = zc; a->nitems = nitems;
      a->ntx = ntx;
      for (int t = 0; t < ntx; ++t) {
        a->txs[t] = std::min(t * seg, nx);
        a->txe[t] = std::min((t + 1) * seg, nx);
        // the tile loads one vector beyond its own columns on either side, inside the grid
        a->tx0[t] = ntx == 1 ? 0 : std::max(0, std::min(a->txs[t] - vec, nx - v.nbx * txb));
      }
      a->st_hint = o.star7_sthint;
    };
    if (dbl) fill(&p->ad[dir]); else fill(&p->af[dir]);
  }
  return p;
}

void LaunchStar7Pair(Runtime *rt, Star7PairPlan *p, int dir) {
  if (rt->world() > 1) {
    sweep::SlabSync &sy = p->is_double ? p->ad[dir].sync : p->af[dir].sync;
    sy.wait_epoch = rt->sweep_epoch;
    sy.signal_epoch = rt->sweep_epoch + 1;
  }
  void *args[2];
  args[0] = &p->tmap[dir];
  args[1] = p->is_double ? (void *)&p->ad[dir] : (void *)&p->af[dir];
  LaunchSweepKernel(rt, p->fn, p->grid, p->block, args, p->smem);
}

void DestroyStar7Pair(Star7PairPlan *p) { delete p; }

}  // namespace physis_b200
