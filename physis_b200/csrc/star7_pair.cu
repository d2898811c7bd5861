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
//  * a tile is NBX boxes wide = whole grid rows (no x halo) and H = NWY*RY rows high;
//    a CTA marches it along a z chunk.  First-sweep values ("s1") are computed on all
//    H rows, second-sweep values are stored for the H-2 inner rows; tiles overlap by
//    two rows in y and chunks by two planes in z (the only redundant work);
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

template <typename T>
struct PairArgs {
  T *out;
  int nx, ny, nz;
  T cc, cw, ce, cs, cn, cb, ct;
  int nty, nzc, zc, nitems;
  int st_hint;
};

constexpr int kPairSlots = 3;

template <typename T, int NBX, int NWY, int RY>
struct PairGeom {
  static constexpr int H = NWY * RY;
  static constexpr int ROWB = Geom<T>::TXB * (int)sizeof(T);  // 512 bytes: one warp of vectors
  static constexpr int IN_BOX = (H + 2) * ROWB;
  static constexpr int IN_STAGE = NBX * IN_BOX;
  static constexpr int S1_BOX = H * ROWB;
  static constexpr int S1_STAGE = NBX * S1_BOX;
  static constexpr int SMEM = kBarrierBytes + kPairSlots * (IN_STAGE + S1_STAGE);
  static constexpr int THREADS = NBX * NWY * 32;
};

template <typename T, int NBX, int NWY, int RY, int MINB, int FP>
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

  const int bx = warp % NBX;
  const int wy = warp / NBX;
  const int j0 = wy * RY;  // tile row of this thread's first row
  const unsigned char *my_in = in_ring + bx * IN_BOX + (j0 + 1) * ROWB + lane * 16;
  unsigned char *my_s1 = s1_ring + bx * S1_BOX + j0 * ROWB + lane * 16;
  const bool rd_west = (lane == 0) && (bx > 0);
  const bool rd_east = (lane == 31) && (bx < NBX - 1);
  const int x = bx * G::TXB + lane * VEC;
  const bool x_ok = (x + VEC <= a.nx);
  const bool x_first = (x == 0);
  const bool x_last = (x + VEC == a.nx);
  // rows of the s1 ring above / below this thread's rows (kept inside the tile)
  const int s1_north = (wy == 0) ? 0 : -ROWB;
  const int s1_south = (wy == NWY - 1) ? (RY - 1) * ROWB : RY * ROWB;
  const size_t plane_elems = (size_t)a.nx * a.ny;
  int nbox = 0;
#pragma unroll
  for (int b = 0; b < NBX; ++b) nbox += (b * G::TXB < a.nx) ? 1 : 0;
  const uint32_t tx_bytes = (uint32_t)nbox * (uint32_t)IN_BOX;

  int stage = 0;       // input ring: slot of the next plane to consume
  uint32_t phase = 0;
  int pstage = 0;      // issuer: slot the next plane is loaded into
  int s1s = 0;         // s1 ring: slot the next first-sweep plane is written to

#define SP_ADVANCE() do { if (++stage == NS) { stage = 0; phase ^= 1u; } } while (0)
#define SP_LOAD(DST, st) do { \
    const unsigned char *p__ = my_in + (st) * IN_STAGE; \
    _Pragma("unroll") for (int r = 0; r < RY; ++r) DST[r] = *reinterpret_cast<const V *>(p__ + r * ROWB); \
  } while (0)
#define SP_ISSUE(yin, zpl) do { \
    tma::mbar_arrive_expect_tx(&full[pstage], tx_bytes); \
    unsigned char *dst__ = in_ring + pstage * IN_STAGE; \
    _Pragma("unroll") for (int b = 0; b < NBX; ++b) \
      if (b * G::TXB < a.nx) tma::load_3d(dst__ + b * IN_BOX, &tmap, &full[pstage], b * G::TXB, (yin), (zpl)); \
    if (++pstage == NS) pstage = 0; \
  } while (0)

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int zci = item / a.nty;
    const int ty = item - zci * a.nty;
    const int zb = zci * a.zc;
    const int ze = min(zb + a.zc, a.nz);
    const int k0 = zb > 0 ? zb - 2 : -1;  // first plane of the input window
    const int plast = ze + 1;             // last input plane this item touches
    const int y1 = ty * (H - 2) - 1;      // grid row of tile row 0
    const int ybase = y1 + j0;
    // rows whose y neighbour leaves the grid take the centre value (warp-uniform)
    const int r_north = -ybase;
    const int r_south = a.ny - 1 - ybase;
    const bool y_edge = (r_north >= 0 && r_north < RY) || (r_south >= 0 && r_south < RY);
    bool st_ok[RY];
    T *outp[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int j = j0 + r;
      const int y = ybase + r;
      st_ok[r] = x_ok && j >= 1 && j <= H - 2 && y >= 0 && y < a.ny;
      outp[r] = a.out + ((size_t)zb * a.ny + y) * a.nx + x;
    }

    // every thread is done with both rings of the previous item
    __syncthreads();
    int pnext = k0;
    if (issuer) {
      for (int n = 0; n < NS; ++n) { SP_ISSUE(y1 - 1, pnext); ++pnext; }
    }

    V bot[RY], cen[RY], top[RY];  // input window: planes k, k+1, k+2 (own cells)
    V c1[RY], t1[RY];             // first-sweep planes k and k+1 (own cells)
    tma::mbar_wait(&full[stage], phase);
    SP_LOAD(bot, stage);
    SP_ADVANCE();
    int stage_c = stage;
    tma::mbar_wait(&full[stage], phase);
    SP_LOAD(cen, stage);
    SP_ADVANCE();
#pragma unroll
    for (int r = 0; r < RY; ++r) c1[r] = cen[r];  // defined value; selected by no store

    for (int k = k0; k < ze; ++k) {
      // ---------------- first sweep of plane p = k+1 -> t1, s1 ring slot s1s
      const int p = k + 1;
      const int stage_t = stage;
      tma::mbar_wait(&full[stage], phase);
      SP_LOAD(top, stage);
      SP_ADVANCE();
      if (p == 0) {
#pragma unroll
        for (int r = 0; r < RY; ++r) bot[r] = cen[r];
      }
      if (p == a.nz - 1) {
#pragma unroll
        for (int r = 0; r < RY; ++r) top[r] = cen[r];
      }
      {
        const unsigned char *cb = my_in + stage_c * IN_STAGE;
        const V north = *reinterpret_cast<const V *>(cb - ROWB);
        const V south = *reinterpret_cast<const V *>(cb + RY * ROWB);
        unsigned char *sp = my_s1 + s1s * S1_STAGE;
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          const V c = cen[r];
          T wv = __shfl_up_sync(0xffffffffu, v2::Last(c), 1);
          T ev = __shfl_down_sync(0xffffffffu, v2::First(c), 1);
          if (rd_west) wv = *reinterpret_cast<const T *>(cb + r * ROWB - IN_BOX + WEST_EL);
          if (rd_east) ev = *reinterpret_cast<const T *>(cb + r * ROWB + IN_BOX + EAST_EL);
          if (x_first) wv = v2::First(c);
          if (x_last) ev = v2::Last(c);
          V nv = (r == 0) ? north : cen[r > 0 ? r - 1 : 0];
          V sv = (r == RY - 1) ? south : cen[r < RY - 1 ? r + 1 : r];
          if (y_edge) {
            if (r == r_north) nv = c;
            if (r == r_south) sv = c;
          }
          t1[r] = v2::Vec7<FP>(a, c, wv, ev, sv, nv, bot[r], top[r]);
          *reinterpret_cast<V *>(sp + r * ROWB) = t1[r];
        }
      }
      __syncthreads();
      // planes up to k+1 are dead: re-arm their slots
      if (issuer) {
        if (k == k0 && pnext <= plast) { SP_ISSUE(y1 - 1, pnext); ++pnext; }
        if (pnext <= plast) { SP_ISSUE(y1 - 1, pnext); ++pnext; }
      }
      // ---------------- second sweep of plane k from s1 planes k-1, k, k+1
      if (k >= zb) {
        const int slot_c = (s1s + 2) % NS;   // s1 plane k
        const int slot_b = (s1s + 1) % NS;   // s1 plane k-1
        const unsigned char *cb = my_s1 + slot_c * S1_STAGE;
        const unsigned char *bb = my_s1 + slot_b * S1_STAGE;
        const V north = *reinterpret_cast<const V *>(cb + s1_north);
        const V south = *reinterpret_cast<const V *>(cb + s1_south);
        const bool z_first = (k == 0), z_last = (k == a.nz - 1);
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          const V c = c1[r];
          T wv = __shfl_up_sync(0xffffffffu, v2::Last(c), 1);
          T ev = __shfl_down_sync(0xffffffffu, v2::First(c), 1);
          if (rd_west) wv = *reinterpret_cast<const T *>(cb + r * ROWB - S1_BOX + WEST_EL);
          if (rd_east) ev = *reinterpret_cast<const T *>(cb + r * ROWB + S1_BOX + EAST_EL);
          if (x_first) wv = v2::First(c);
          if (x_last) ev = v2::Last(c);
          V nv = (r == 0) ? north : c1[r > 0 ? r - 1 : 0];
          V sv = (r == RY - 1) ? south : c1[r < RY - 1 ? r + 1 : r];
          if (y_edge) {
            if (r == r_north) nv = c;
            if (r == r_south) sv = c;
          }
          V bv = *reinterpret_cast<const V *>(bb + r * ROWB);
          V tv = t1[r];
          if (z_first) bv = c;
          if (z_last) tv = c;
          const V o = v2::Vec7<FP>(a, c, wv, ev, sv, nv, bv, tv);
          if (st_ok[r]) StoreVec(reinterpret_cast<V *>(outp[r]), o, a.st_hint != 0);
          outp[r] += plane_elems;
        }
      }
      // rotate the windows
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        bot[r] = cen[r];
        cen[r] = top[r];
        c1[r] = t1[r];
      }
      stage_c = stage_t;
      if (++s1s == NS) s1s = 0;
    }
  }
#undef SP_ISSUE
#undef SP_LOAD
#undef SP_ADVANCE
}

// ------------------------------------------------------------------ host side

struct PairVariant {
  int nbx, nwy, ry, minb;
  const void *f32[2];  // scalar / packed-add arithmetic
  const void *f64;
  int smem_f32, smem_f64, threads;
};

#define PAIR_VARIANT(NBX, NWY, RY, MINB) \
  { NBX, NWY, RY, MINB, \
    {(const void *)Star7PairKernel<float, NBX, NWY, RY, MINB, 0>, \
     (const void *)Star7PairKernel<float, NBX, NWY, RY, MINB, 1>}, \
    (const void *)Star7PairKernel<double, NBX, NWY, RY, MINB, 0>, \
    PairGeom<float, NBX, NWY, RY>::SMEM, PairGeom<double, NBX, NWY, RY>::SMEM, NBX * NWY * 32 }

const PairVariant kPairVariants[] = {
    PAIR_VARIANT(4, 4, 4, 1),  // 0: rows of up to 4 boxes, 16-row tiles
    PAIR_VARIANT(3, 4, 4, 1),  // 1
    PAIR_VARIANT(2, 4, 4, 2),  // 2
    PAIR_VARIANT(1, 4, 4, 4),  // 3
};
constexpr int kNumPairVariants = sizeof(kPairVariants) / sizeof(kPairVariants[0]);

}  // namespace

struct Star7PairPlan {
  bool is_double = false;
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
  if (rt->world() > 1) { *why = "multi-GPU runs exchange a one-plane halo per sweep"; return nullptr; }
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
  const bool dbl = ga->type == PS_DOUBLE;
  const int vec = dbl ? 2 : 4;
  const int txb = dbl ? Geom<double>::TXB : Geom<float>::TXB;
  const int nx = ga->dim[0], ny = ga->dim[1], nz = ga->dim[2];
  if (nx % vec != 0) { *why = "x extent must be a multiple of 16 bytes"; return nullptr; }
  const int boxes = CeilDiv(nx, txb);
  int variant = -1;
  for (int v = 0; v < kNumPairVariants; ++v)
    if (kPairVariants[v].nbx == boxes) variant = v;
  if (variant < 0) { *why = "rows wider than the fused kernel's tile"; return nullptr; }
  if (nz < 2 || ny < 2) { *why = "grid too thin"; return nullptr; }
  const PairVariant &v = kPairVariants[variant];

  Star7PairPlan *p = new Star7PairPlan();
  p->is_double = dbl;
  p->fn = dbl ? v.f64 : v.f32[o.star7_impl == 2 ? 1 : 0];
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
      const long waves = CeilDiv((long)nty * CeilDiv(nz, c), slots);
      const long cost = waves * (c + 3);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; zc = c; }
    }
  }
  zc = std::max(1, std::min(zc, nz));
  const int nzc = CeilDiv(nz, zc);
  const int nitems = nty * nzc;
  p->grid = std::min(nitems, slots);

  Grid *gin[2] = {ga, gb};
  Grid *gout[2] = {gb, ga};
  for (int dir = 0; dir < 2; ++dir) {
    int dimv[3] = {nx, ny, nz};
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
      a->nx = nx; a->ny = ny; a->nz = nz;
      // scalars arrive in the kernel's parameter order: ce, cw, cn, cs, ct, cb, cc
      a->ce = (ET)d0.scalars[0]; a->cw = (ET)d0.scalars[1]; a->cn = (ET)d0.scalars[2];
      a->cs = (ET)d0.scalars[3]; a->ct = (ET)d0.scalars[4]; a->cb = (ET)d0.scalars[5];
      a->cc = (ET)d0.scalars[6];
      a->nty = nty; a->nzc = nzc; a->zc = zc; a->nitems = nitems;
      a->st_hint = o.star7_sthint;
    };
    if (dbl) fill(&p->ad[dir]); else fill(&p->af[dir]);
  }
  return p;
}

void LaunchStar7Pair(Runtime *rt, Star7PairPlan *p, int dir) {
  void *args[2];
  args[0] = &p->tmap[dir];
  args[1] = p->is_double ? (void *)&p->ad[dir] : (void *)&p->af[dir];
  PSB_CUDA(cudaLaunchKernel(p->fn, dim3(p->grid), dim3(p->block), args, p->smem, rt->stream));
}

void DestroyStar7Pair(Star7PairPlan *p) { delete p; }

}  // namespace physis_b200
