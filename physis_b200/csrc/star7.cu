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
  int l2_hint;   // 1: loads carry an evict_first policy
  int st_hint;   // 1: streaming (evict-first) stores
  // z-slab view (multi-GPU; on one GPU zcl = {0, nz-1} and nothing is pushed):
  // local planes at which the bottom / top neighbour is clamped to the centre,
  int zcl_lo, zcl_hi;
  // and the local planes whose result is also stored, through the CUDA-IPC peer
  // mapping, into the ring neighbour's halo plane (fused halo exchange over NVLink)
  int push_lo_z, push_hi_z;
  T *push_lo, *push_hi;
};

template <typename T>
__device__ __forceinline__ T Point7(const Star7Args<T> &a, T c, T w, T e, T s, T n, T b, T t) {
  // ((((((cc*c + cw*w) + ce*e) + cs*s) + cn*n) + cb*b) + ct*t)
  T r = MulRn(a.cc, c);
  r = AddRn(r, MulRn(a.cw, w));
  r = AddRn(r, MulRn(a.ce, e));
  r = AddRn(r, MulRn(a.cs, s));
  r = AddRn(r, MulRn(a.cn, n));
  r = AddRn(r, MulRn(a.cb, b));
  r = AddRn(r, MulRn(a.ct, t));
  return r;
}

// TY  rows of the CTA tile, RY rows per thread, NBX boxes side by side in x,
// MINB resident CTAs per SM the register budget is sized for.
template <typename T, int TY, int RY, int NBX, int MINB>
__global__ void __launch_bounds__((NBX * (TY / RY) + 1) * 32, MINB)
Star7Kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Star7Args<T> a) {
  using G = Geom<T>;
  using V = typename VecOf<T>::type;
  constexpr int VEC = G::VEC;
  constexpr int NWY = TY / RY;
  constexpr int NW = NBX * NWY;  // consumer warps
  constexpr int ROWB = G::ROW_BYTES;
  constexpr int BOX_STRIDE = BoxStride<T, TY>();
  constexpr int STAGE_BYTES = NBX * BOX_STRIDE;
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
  __syncthreads();

  const int tiles_xy = a.ntx * a.nty;

  if (warp == NW) {
    // ------------------------------------------------------------ producer
    if (lane != 0) return;
    tma::prefetch_tensormap(&tmap);
    const uint64_t policy = tma::policy_evict_first();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int zci = item / tiles_xy;
      const int txy = item - zci * tiles_xy;
      const int ty = txy / a.ntx;
      const int tx = txy - ty * a.ntx;
      const int x0 = a.xbase + tx * (NBX * G::TXB);
      const int y0 = a.dy0 + ty * TY;
      const int zb = a.dz0 + zci * a.zc;
      const int ze = min(zb + a.zc, a.dz1);
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
          if (bx0 < a.nx) {
            if (a.l2_hint)
              tma::load_3d_hint(dst + b * BOX_STRIDE, &tmap, &full[stage], bx0 - G::HX, y0 - 1, z,
                                policy);
            else
              tma::load_3d(dst + b * BOX_STRIDE, &tmap, &full[stage], bx0 - G::HX, y0 - 1, z);
          }
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int bx = warp % NBX;
  const int wy = warp / NBX;
  // byte offset of this thread's vector inside a box row / of its first row
  const int col_off = (G::HX + lane * VEC) * (int)sizeof(T);
  const int row0 = wy * RY;  // smem row of the north halo of this thread's rows

  int stage = 0;       // ring position of the next plane to consume
  uint32_t phase = 0;
  auto advance = [&]() {
    if (++stage == S) { stage = 0; phase ^= 1u; }
  };
  auto release = [&](int st) {
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(&empty[st]);
  };

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int zci = item / tiles_xy;
    const int txy = item - zci * tiles_xy;
    const int ty = txy / a.ntx;
    const int tx = txy - ty * a.ntx;
    const int x = a.xbase + tx * (NBX * G::TXB) + bx * G::TXB + lane * VEC;
    const int ybase = a.dy0 + ty * TY + wy * RY;
    const int zb = a.dz0 + zci * a.zc;
    const int ze = min(zb + a.zc, a.dz1);
    const bool x_ok = (x >= a.dx0) && (x + VEC <= a.dx1);
    const bool x_first = (x == 0);
    const bool x_last = (x + VEC == a.nx);

    const unsigned char *box = planes + bx * BOX_STRIDE;
    V cen[RY], bot[RY], top[RY];

    if (zb > 0) {
      tma::mbar_wait(&full[stage], phase);
      const unsigned char *p = box + stage * STAGE_BYTES + (row0 + 1) * ROWB + col_off;
#pragma unroll
      for (int r = 0; r < RY; ++r) bot[r] = *reinterpret_cast<const V *>(p + r * ROWB);
      release(stage);
      advance();
    }
    int stage_c = stage;
    tma::mbar_wait(&full[stage], phase);
    {
      const unsigned char *p = box + stage * STAGE_BYTES + (row0 + 1) * ROWB + col_off;
#pragma unroll
      for (int r = 0; r < RY; ++r) cen[r] = *reinterpret_cast<const V *>(p + r * ROWB);
    }
    advance();

    bool has_top = false;
    for (int z = zb; z < ze; ++z) {
      has_top = (z + 1 < a.nz);
      const int stage_t = stage;
      if (has_top) {
        tma::mbar_wait(&full[stage], phase);
        const unsigned char *p = box + stage * STAGE_BYTES + (row0 + 1) * ROWB + col_off;
#pragma unroll
        for (int r = 0; r < RY; ++r) top[r] = *reinterpret_cast<const V *>(p + r * ROWB);
      }
      const unsigned char *cb = box + stage_c * STAGE_BYTES;
      const V north = *reinterpret_cast<const V *>(cb + row0 * ROWB + col_off);
      const V south = *reinterpret_cast<const V *>(cb + (row0 + RY + 1) * ROWB + col_off);
      const bool z_first = (z == a.zcl_lo), z_last = (z == a.zcl_hi);
      T *const push0 = (z == a.push_lo_z) ? a.push_lo : nullptr;
      T *const push1 = (z == a.push_hi_z) ? a.push_hi : nullptr;
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const int y = ybase + r;
        const V c = cen[r];
        T wv = __shfl_up_sync(0xffffffffu, Elem(c, VEC - 1), 1);
        T ev = __shfl_down_sync(0xffffffffu, Elem(c, 0), 1);
        const unsigned char *rowp = cb + (row0 + 1 + r) * ROWB;
        if (lane == 0) wv = *reinterpret_cast<const T *>(rowp + (G::HX - 1) * sizeof(T));
        if (lane == 31) ev = *reinterpret_cast<const T *>(rowp + (G::HX + G::TXB) * sizeof(T));
        V nv = (r == 0) ? north : cen[r - 1];
        V sv = (r == RY - 1) ? south : cen[r + 1];
        V bv = bot[r], tv = top[r];
        if (y == 0) nv = c;
        if (y == a.ny - 1) sv = c;
        if (z_first) bv = c;
        if (z_last) tv = c;
        V o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const T cj = Elem(c, j);
          T wj = (j == 0) ? wv : Elem(c, j - 1);
          T ej = (j == VEC - 1) ? ev : Elem(c, j + 1);
          if (j == 0 && x_first) wj = cj;
          if (j == VEC - 1 && x_last) ej = cj;
          SetElem(o, j, Point7<T>(a, cj, wj, ej, Elem(sv, j), Elem(nv, j), Elem(bv, j), Elem(tv, j)));
        }
        if (x_ok && y >= a.dy0 && y < a.dy1) {
          V *dst = reinterpret_cast<V *>(a.out + ((size_t)z * a.ny + y) * a.nx + x);
          StoreVec(dst, o, a.st_hint != 0);
          if (push0) *reinterpret_cast<V *>(push0 + (size_t)y * a.nx + x) = o;
          if (push1) *reinterpret_cast<V *>(push1 + (size_t)y * a.nx + x) = o;
        }
      }
      release(stage_c);
      if (has_top) {
        stage_c = stage_t;
        advance();
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          bot[r] = cen[r];
          cen[r] = top[r];
        }
      }
    }
    if (has_top) release(stage_c);  // plane ze was loaded as `top` only
  }
}

// ------------------------------------------------------------------ host side

struct VariantInfo {
  int ty, ry, nbx;
  const void *f32;
  const void *f64;
};

#define VARIANT(TY, RY, NBX, MINB) \
  { TY, RY, NBX, (const void *)Star7Kernel<float, TY, RY, NBX, MINB>, \
    (const void *)Star7Kernel<double, TY, RY, NBX, MINB> }

const VariantInfo kVariants[] = {
    VARIANT(32, 4, 1, 2),  // 0: default
    VARIANT(16, 2, 1, 3),  // 1
    VARIANT(16, 2, 1, 4),  // 2
    VARIANT(32, 4, 2, 1),  // 3
    VARIANT(16, 2, 2, 1),  // 4
    VARIANT(16, 4, 2, 2),  // 5
    VARIANT(8, 2, 4, 1),   // 6
    VARIANT(8, 1, 2, 2),   // 7
    VARIANT(64, 8, 1, 1),  // 8
    VARIANT(32, 2, 1, 1),  // 9
    VARIANT(8, 2, 1, 4),   // 10
    VARIANT(32, 8, 1, 2),  // 11
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <typename T>
size_t SmemBytes(const VariantInfo &v, int stages) {
  size_t box = ((size_t)(v.ty + 2) * Geom<T>::ROW_BYTES + 127) / 128 * 128;
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
  // Tile shape: measured on B200 at 512^3 fp32 (tools/tune_star7.py, profiles/):
  // two 512-byte boxes side by side x 16 rows, 2 rows per thread, 6-deep ring, one
  // CTA per SM reaches 5.47 TB/s; rows narrower than two boxes use the one-box
  // 8-row shape at 4 CTAs per SM.
  int variant = o.star7_variant;
  const size_t row_bytes = (size_t)(dom.local_max[0] - dom.local_min[0]) * (dbl ? 8 : 4);
  const bool autov = variant < 0 || variant >= kNumVariants;
  if (autov) variant = row_bytes >= 1024 ? 4 : 10;
  const VariantInfo &v = kVariants[variant];
  p->variant = variant;
  p->fn = dbl ? v.f64 : v.f32;
  int stages = o.star7_stages > 0 ? std::min(o.star7_stages, kMaxStages) : (variant == 4 ? 6 : 5);
  if (stages < 3) stages = 3;
  p->smem = dbl ? SmemBytes<double>(v, stages) : SmemBytes<float>(v, stages);
  p->block = (v.nbx * (v.ty / v.ry) + 1) * 32;
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
    // fewest z chunks (least z-halo re-reads) that still give every resident
    // CTA slot at least ~2 items, chunk count chosen so items divide the slots
    // as evenly as possible
    // short chunks keep the statically strided items balanced across CTAs (the
    // two extra halo planes per chunk mostly hit in L2); 32 planes measured best
    // at 512^3, shrink further only when there would be fewer than ~4 items per slot
    int tiles = ntx * nty;
    int want_chunks = std::max(1, CeilDiv(4L * slots, tiles));
    zc = std::min(32, std::max(8, CeilDiv(nzd, want_chunks)));
    zc = std::min(zc, nzd);
  }
  const int nzc = CeilDiv(nzd, zc);
  const int nitems = ntx * nty * nzc;
  p->grid = std::min(nitems, slots);

  int dimv[3] = {gin->ldim[0], gin->ldim[1], gin->ldim[2]};
  int boxv[3] = {dbl ? Geom<double>::BW : Geom<float>::BW, v.ty + 2, 1};
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
    a->l2_hint = o.star7_l2hint;
    a->st_hint = o.star7_sthint;
    using ET = typename std::remove_pointer<decltype(a->out)>::type;
    if (SlabPushTargets(rt, *gout, 0, (void **)&a->push_lo, (void **)&a->push_hi, sizeof(ET))) {
      a->push_lo_z = gout->halo;
      a->push_hi_z = gout->halo + gout->nz_loc - 1;
      p->pushes = true;
    }
  };
  if (dbl) { FillArgs(&p->ad, d, gin, gout); common(&p->ad); }
  else { FillArgs(&p->af, d, gin, gout); common(&p->af); }
  return p;
}

void LaunchStar7(Runtime *rt, Star7Plan *p) {
  void *args[2];
  args[0] = &p->tmap;
  args[1] = p->is_double ? (void *)&p->ad : (void *)&p->af;
  PSB_CUDA(cudaLaunchKernel(p->fn, dim3(p->grid), dim3(p->block), args, p->smem, rt->stream));
}

void DestroyStar7(Star7Plan *p) { delete p; }
bool Star7Pushes(const Star7Plan *p) { return p->pushes; }

}  // namespace physis_b200
