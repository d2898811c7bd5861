// Two Himeno 19-point Jacobi sweeps in one pass over HBM (temporal blocking).
//
// PSStencilRun(map(jacobi, p0 -> p1), map(jacobi, p1 -> p0), nn/2) -- the shape of
// examples/himeno/himenobmtxpa_physis.c:364-393 -- applies the same update twice per
// iteration, and both applications read the SAME twelve coefficient / source arrays.  A single
// sweep moves 56 B per point (himenobmtxpa_physis.c:418-432) of which 48 are those arrays;
// computing sweep n+1 and n+2 of a tile while it is on the SM moves them once per two updates:
// 12 x 4 + 4 (p read) + 4 (p write) = 56 B per TWO updates.  Every point sees exactly the
// arithmetic of himeno.cu (HimenoJacobi, separately rounded fp32 in source order), so the result
// is bit-identical to two separate sweeps.
//
// What the intermediate field looks like.  A Himeno sweep writes the interior [1, n-1)^3 only;
// the boundary cells of the grid it writes keep what they held.  Pass "X -> Y" therefore needs
// the boundary cells of Y as the boundary of the intermediate field.  The schedule
// (stencil_run.cu) runs fused passes only while the boundary cells of both grids are bit-equal
// (checked on the device, HimenoFacesEqual) -- the benchmark initialises both grids identically
// for this very reason (himenobmtxpa_physis.c:138-139) -- so the intermediate's boundary is read
// from X, which is on the SM anyway.
//
// Structure
//  * a tile is one box of 128 floats (+ 16-byte x halo, as himeno.cu) x H = 12 rows, one row
//    per warp; a CTA marches it along a z chunk.  First-sweep values ("s1") are computed on all H
//    rows and all 128 columns, second-sweep values are stored for the H-2 inner rows and the
//    tile's own columns: tiles overlap by two rows in y and by one 16-byte vector per seam side
//    in x, chunks by two planes in z (the redundant work; the re-read rows / columns hit in L2);
//  * p planes (H+2 rows) arrive by TMA in a 5-slot shared-memory ring, completion on mbarriers;
//    s1 planes live in a 4-slot shared-memory ring of the same geometry (the second sweep reads
//    its neighbours' cells in three s1 planes, so a slot is recycled one step later than in the
//    7-point pair);
//  * one CTA-wide barrier per plane orders both rings; thread 0 re-arms the p slot behind it;
//  * the twelve coefficient vectors a thread loads for the first sweep of plane k+1 stay in its
//    registers for one step and serve the second sweep of the same plane: 96 registers of
//    coefficients per thread, which is why a CTA has 12 warps (170 registers each) and not 16.
//    (A first version re-read them for the second sweep and counted on L2: ncu showed 19.7 GB of
//    DRAM reads per pass instead of 14.9 -- with 148 CTAs streaming 12 arrays the reuse distance
//    of one step is ~40 MB, and half of the re-reads missed.)  Bulk L2 prefetches pull the next
//    plane's coefficient rows towards the SM, and the loads are ordinary (evict-normal) ones so
//    that the rows two neighbouring tiles share are still in L2 for the second of them (with
//    evict-first loads DRAM reads were 18.8 GB per pass, with ordinary ones 15.8);
//  * x neighbours of a thread's vector come from the neighbouring lanes by shuffle, only the
//    edge lanes read the halo column (4-byte shared-memory loads at a 16-byte lane stride are
//    four-way bank conflicts).
#include "runtime.h"
#include "tma.cuh"
#include "sweep_common.cuh"
#include "himeno_math.cuh"

#include <algorithm>
#include <string>
#include <vector>

namespace physis_b200 {

namespace {

using namespace sweep;

constexpr int kHpMaxXTiles = 32;
constexpr int kHpInSlots = 5;
constexpr int kHpS1Slots = 4;

struct HimenoPairArgs {
  const float *coef[12];  // a0 a1 a2 a3 b0 b1 b2 c0 c1 c2 bnd wrk1
  float *out;
  float omega;
  int nx, ny, nz;
  int nty, ntx, nzc, zc, nitems;
  int tx0[kHpMaxXTiles], txs[kHpMaxXTiles], txe[kHpMaxXTiles];
  int dz0, dz1;  // planes written (local indices): the global interior [1, n-1) cut to this rank's slab
  int pf;        // coefficient prefetch distance in planes
  int pf_tensor; // 1: by tensor-map prefetch of the tile's box, 0: by bulk prefetches per row
  // z-slab view (multi-GPU; on one GPU the faces are planes 0 and nz-1 and nothing is pushed): local
  // planes holding the global z faces (-1 when they are elsewhere), whose cells no sweep updates;
  // the slab's first two / last two planes are also stored into the ring neighbours' halo planes
  // (byte distances from the plane's own address in `out`)
  int zface_lo, zface_hi;
  int push_lo_z, push_hi_z;
  long long push_lo_delta, push_hi_delta;
  SlabSync sync;
};

__device__ __forceinline__ void PrefetchL2(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Two ways to pull the coefficient rows of the next plane towards L2 (HimenoPairArgs::pf_tensor):
//  - a bulk prefetch per row and array from 12 lanes of every warp: a serialised loop per lane
//    (a fifth of the kernel's instructions), but every warp fetches its own row exactly when it
//    is about to need it;
//  - a tensor-map prefetch of the tile's box of an array (128 columns x H rows x 1 plane): 12
//    instructions of one thread per plane.
// Measured (profiles/r2_experiments.txt): XL 152.1 / 147.1 GLUP/s (none: 144.7-146.2), L 127.8 /
// 133.4-135.9 (none: 127.0-127.8); the planner picks by plane size (PrepareHimenoPair).
struct alignas(64) HpCoefMaps {
  CUtensorMap m[12];
};

// The update of this thread's four cells from three planes of a ring (pb / pc / pt point at the
// thread's own vector in the planes below / at / above): vectors of the rows above and below
// from shared memory, x neighbours from the neighbouring lanes, the halo column for the edge
// lanes.  Cells with upd[j] false keep `keep`'s value.
__device__ __forceinline__ float4 Update(const float4 (&cf)[12], float omega, const unsigned char *pb,
                                         const unsigned char *pc, const unsigned char *pt, const float4 &c_c,
                                         const float4 &keep, const bool (&upd)[4], int lane) {
  constexpr int ROWB = Geom<float>::ROW_BYTES;
  auto vec = [](const unsigned char *row, int dy) {
    return *reinterpret_cast<const float4 *>(row + dy * ROWB);
  };
  // the element west of a vector is the previous lane's last, east the next lane's first
  auto west = [lane](const float4 &v, const unsigned char *row, int dy) {
    float w = __shfl_up_sync(0xffffffffu, v.w, 1);
    if (lane == 0) w = *reinterpret_cast<const float *>(row + dy * ROWB - 4);
    return w;
  };
  auto east = [lane](const float4 &v, const unsigned char *row, int dy) {
    float e = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (lane == 31) e = *reinterpret_cast<const float *>(row + dy * ROWB + 16);
    return e;
  };
  const float4 c_n = vec(pc, -1), c_s = vec(pc, 1);
  const float c_cw = west(c_c, pc, 0), c_ce = east(c_c, pc, 0);
  const float c_nw = west(c_n, pc, -1), c_ne = east(c_n, pc, -1);
  const float c_sw = west(c_s, pc, 1), c_se = east(c_s, pc, 1);
  const float4 b_c = vec(pb, 0), b_n = vec(pb, -1), b_s = vec(pb, 1);
  const float b_w = west(b_c, pb, 0), b_e = east(b_c, pb, 0);
  const float4 t_c = vec(pt, 0), t_n = vec(pt, -1), t_s = vec(pt, 1);
  const float t_w = west(t_c, pt, 0), t_e = east(t_c, pt, 0);
  float4 o = keep;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float c_xm = (j == 0) ? c_cw : Elem(c_c, j - 1);
    const float c_xp = (j == 3) ? c_ce : Elem(c_c, j + 1);
    const float n_xm = (j == 0) ? c_nw : Elem(c_n, j - 1);
    const float n_xp = (j == 3) ? c_ne : Elem(c_n, j + 1);
    const float s_xm = (j == 0) ? c_sw : Elem(c_s, j - 1);
    const float s_xp = (j == 3) ? c_se : Elem(c_s, j + 1);
    const float b_xm = (j == 0) ? b_w : Elem(b_c, j - 1);
    const float b_xp = (j == 3) ? b_e : Elem(b_c, j + 1);
    const float t_xm = (j == 0) ? t_w : Elem(t_c, j - 1);
    const float t_xp = (j == 3) ? t_e : Elem(t_c, j + 1);
    float ss;
    const float v = HimenoJacobi(
        Elem(cf[0], j), Elem(cf[1], j), Elem(cf[2], j), Elem(cf[3], j), Elem(cf[4], j), Elem(cf[5], j),
        Elem(cf[6], j), Elem(cf[7], j), Elem(cf[8], j), Elem(cf[9], j), Elem(cf[10], j), Elem(cf[11], j),
        omega,
        /*ccc*/ Elem(c_c, j), /*ccp*/ Elem(t_c, j), /*cpc*/ Elem(c_s, j), /*pcc*/ c_xp,
        /*cpp*/ Elem(t_s, j), /*cmp*/ Elem(t_n, j), /*cpm*/ Elem(b_s, j), /*cmm*/ Elem(b_n, j),
        /*ppc*/ s_xp, /*pmc*/ n_xp, /*mpc*/ s_xm, /*mmc*/ n_xm,
        /*pcp*/ t_xp, /*pcm*/ b_xp, /*mcp*/ t_xm, /*mcm*/ b_xm,
        /*ccm*/ Elem(b_c, j), /*cmc*/ Elem(c_n, j), /*mcc*/ c_xm, &ss);
    if (upd[j]) SetElem(o, j, v);
  }
  return o;
}

template <int H>
struct HpGeom {
  static constexpr int ROWB = Geom<float>::ROW_BYTES;                  // 544
  static constexpr int STAGE = ((H + 2) * ROWB + 127) / 128 * 128;
  static constexpr int SMEM = kBarrierBytes + (kHpInSlots + kHpS1Slots) * STAGE;
};

// H rows per tile = consumer warps; no producer warp (thread 0 issues the TMA loads).
// SLAB: the z-slab form (halo planes forwarded to the ring neighbours, ordering with them).
template <int H, bool SLAB>
__global__ void __launch_bounds__(H * 32, 1)
HimenoPairKernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ HpCoefMaps cm,
                 const __grid_constant__ HimenoPairArgs a) {
  using G = Geom<float>;
  using PG = HpGeom<H>;
  constexpr int VEC = 4;
  constexpr int ROWB = PG::ROWB;
  constexpr int STAGE = PG::STAGE;

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  unsigned char *in_ring = smem + kBarrierBytes;
  unsigned char *s1_ring = in_ring + kHpInSlots * STAGE;

  const int warp = threadIdx.x >> 5;  // tile row
  const int lane = threadIdx.x & 31;
  const bool issuer = (threadIdx.x == 0);

  if (issuer) {
    for (int s = 0; s < kHpInSlots; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_barrier_init();
    tma::prefetch_tensormap(&tmap);
  }
  if (SLAB) SlabSyncWait(a.sync);  // before any halo plane is read or any peer halo written

  // this thread's vector inside a stage: row warp+1 (row 0 is the halo row above the tile),
  // column 16 bytes of x halo + lane * 16
  const int my_off = (warp + 1) * ROWB + (G::HX + lane * VEC) * (int)sizeof(float);
  const size_t plane_elems = (size_t)a.nx * a.ny;
  const int tiles_xy = a.nty * a.ntx;
  uint32_t par = 0;  // bit s: phase parity of p slot s

  auto vec = [](const unsigned char *row, int dy) {
    return *reinterpret_cast<const float4 *>(row + dy * ROWB);
  };

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int zseq = item / tiles_xy;
    const int zci = SLAB ? SlabChunkOrder(a.sync, zseq, a.nzc) : zseq;
    const int txy = item - zseq * tiles_xy;
    const int ty = txy / a.ntx;
    const int tx = txy - ty * a.ntx;
    const int xt0 = a.tx0[tx];
    const int x = xt0 + lane * VEC;
    const int y = ty * (H - 2) + warp;  // grid row of this warp's tile row
    const int zb = a.dz0 + zci * a.zc;
    const int ze = min(zb + a.zc, a.dz1);
    const int kfirst = zb - 2;   // first p plane of the window (may be -1: zero fill, unused)
    const int klast = ze + 1;    // last p plane of the window (may be nz: zero fill, unused)
    const bool row_in_grid = (y < a.ny);
    const bool row_bnd = (y == 0) || (y >= a.ny - 1);
    // cells of this vector that the sweeps update (the rest keep the input value)
    bool upd[VEC], st[VEC];
    const bool st_row = (warp >= 1) && (warp <= H - 2) && (y >= 1) && (y < a.ny - 1);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      upd[j] = (x + j >= 1) && (x + j < a.nx - 1);
      st[j] = st_row && upd[j] && (x + j >= a.txs[tx]) && (x + j < a.txe[tx]);
    }
    const bool st_any = st[0] || st[1] || st[2] || st[3];
    const bool st_all = st[0] && st[1] && st[2] && st[3];
    // the row computes first-sweep values (warp-uniform: the shuffles inside Update need every
    // lane); lanes beyond the grid's last column compute on zeros and store nothing
    const bool ld_row = row_in_grid && !row_bnd;
    const bool ld_ok = (x < a.nx);
    // offset of this thread's vector inside a plane of the coefficient arrays / of `out`; lanes
    // beyond the last column load the row's last vector instead (their values reach nothing stored)
    const size_t gp = (size_t)y * a.nx + (ld_ok ? x : a.nx - VEC);

    // every thread is done with both rings of the previous item
    __syncthreads();
    if (issuer) {
      for (int n = 0; n < kHpInSlots; ++n) {
        const int q = kfirst + n;
        if (q <= klast) {
          tma::mbar_arrive_expect_tx(&full[n], (uint32_t)((H + 2) * ROWB));
          tma::load_3d(in_ring + n * STAGE, &tmap, &full[n], xt0 - G::HX, ty * (H - 2) - 1, q);
        }
      }
    }
    // pull the coefficient rows of the first planes towards L2
    if (a.pf_tensor) {
      if (issuer) {
        for (int m = max(zb - 1, 0); m < min(zb - 1 + a.pf, a.nz); ++m)
#pragma unroll
          for (int c = 0; c < 12; ++c) tma::prefetch_3d(&cm.m[c], xt0, ty * (H - 2), m);
      }
    } else if (lane < 12 && ld_row) {
      const uint32_t bytes = (uint32_t)min(G::TXB, a.nx - xt0) * 4u;
      for (int m = max(zb - 1, 0); m < min(zb - 1 + a.pf, a.nz); ++m)
        PrefetchL2(a.coef[lane] + (size_t)m * plane_elems + (size_t)y * a.nx + xt0, bytes);
    }

    int slot_b = 0;  // p slot of plane k (bottom of the first sweep's window)
    // planes kfirst and kfirst+1 are waited for here, plane k+2 inside the loop
    tma::mbar_wait(&full[0], (par >> 0) & 1u);
    par ^= 1u << 0;
    tma::mbar_wait(&full[1], (par >> 1) & 1u);
    par ^= 1u << 1;

    float4 hold[12];  // coefficients of plane k (loaded as plane m of the previous step)
    float4 q[12];     // coefficients of plane m (rows / planes no sweep updates load nothing)
#pragma unroll
    for (int c = 0; c < 12; ++c) hold[c] = q[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = kfirst; k < ze; ++k) {
      const int m = k + 1;  // plane whose first-sweep values this step computes
      const int slot_c = (slot_b + 1 == kHpInSlots) ? 0 : slot_b + 1;
      const int slot_t = (slot_c + 1 == kHpInSlots) ? 0 : slot_c + 1;
      const bool plane_upd = (m != a.zface_lo) && (m != a.zface_hi) && (m >= 0) && (m < a.nz);
      const size_t gm = (size_t)m * plane_elems + gp;
      // ---- first sweep: s1(m) from p(m-1), p(m), p(m+1) ---------------------------------
      const bool compute1 = ld_row && plane_upd;
      if (compute1) {
#pragma unroll
        for (int c = 0; c < 12; ++c) q[c] = __ldg(reinterpret_cast<const float4 *>(a.coef[c] + gm));
      }
      // and the tile's rows of plane m + pf towards L2
      {
        const int mp = m + a.pf;
        if (a.pf > 0 && mp < a.nz && mp <= ze) {
          if (a.pf_tensor) {
            if (issuer) {
#pragma unroll
              for (int c = 0; c < 12; ++c) tma::prefetch_3d(&cm.m[c], xt0, ty * (H - 2), mp);
            }
          } else if (compute1 && lane < 12) {
            PrefetchL2(a.coef[lane] + (size_t)mp * plane_elems + (size_t)y * a.nx + xt0,
                       (uint32_t)min(G::TXB, a.nx - xt0) * 4u);
          }
        }
      }
      tma::mbar_wait(&full[slot_t], (par >> slot_t) & 1u);
      par ^= 1u << slot_t;
      {
        const unsigned char *pb = in_ring + slot_b * STAGE + my_off;
        const unsigned char *pc = in_ring + slot_c * STAGE + my_off;
        const unsigned char *pt = in_ring + slot_t * STAGE + my_off;
        const float4 c_c = vec(pc, 0);
        float4 o = c_c;  // cells the sweep does not update keep the input value
        if (compute1) o = Update(q, a.omega, pb, pc, pt, c_c, o, upd, lane);
        *reinterpret_cast<float4 *>(s1_ring + (m & (kHpS1Slots - 1)) * STAGE + my_off) = o;
      }
      __syncthreads();
      if (issuer) {
        // nobody reads p plane k any more: its slot takes plane k + kHpInSlots
        const int qn = k + kHpInSlots;
        if (qn <= klast) {
          tma::mbar_arrive_expect_tx(&full[slot_b], (uint32_t)((H + 2) * ROWB));
          tma::load_3d(in_ring + slot_b * STAGE, &tmap, &full[slot_b], xt0 - G::HX, ty * (H - 2) - 1, qn);
        }
      }
      // ---- second sweep: out(k) from s1(k-1), s1(k), s1(k+1), coefficients held since the
      //      previous step -----------------------------------------------------------------
      if (k >= zb && st_row) {
        const size_t gk = (size_t)k * plane_elems + gp;
        const unsigned char *pb = s1_ring + ((k - 1) & (kHpS1Slots - 1)) * STAGE + my_off;
        const unsigned char *pc = s1_ring + (k & (kHpS1Slots - 1)) * STAGE + my_off;
        const unsigned char *pt = s1_ring + ((k + 1) & (kHpS1Slots - 1)) * STAGE + my_off;
        const float4 c_c = vec(pc, 0);
        const bool all4[VEC] = {true, true, true, true};
        const float4 o = Update(hold, a.omega, pb, pc, pt, c_c, c_c, all4, lane);
        if (st_all) {
          __stcs(reinterpret_cast<float4 *>(a.out + gk), o);
        } else if (st_any) {
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (st[j]) a.out[gk + j] = Elem(o, j);
        }
      }
      // plane m's coefficients serve the second sweep of the next step
#pragma unroll
      for (int c = 0; c < 12; ++c) hold[c] = q[c];
      slot_b = slot_c;
    }
    if (SLAB) {
      // the slab's first two / last two planes also go to the ring neighbours' halo planes: once
      // the item is done every thread forwards the cells it stored itself (outside the plane loop,
      // as in star7_pair.cu)
      for (int s = 0; s < 4; ++s) {
        const int z = (s < 2 ? a.push_lo_z : a.push_hi_z) + (s & 1);
        const long long delta = s < 2 ? a.push_lo_delta : a.push_hi_delta;
        if (z < zb || z >= ze || !st_any) continue;
        float *src = a.out + (size_t)z * plane_elems + gp;
        if (st_all) {
          *reinterpret_cast<float4 *>(reinterpret_cast<char *>(src) + delta) = *reinterpret_cast<const float4 *>(src);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (st[j]) *reinterpret_cast<float *>(reinterpret_cast<char *>(src + j) + delta) = src[j];
        }
      }
      SlabSyncItemDone(a.sync, item, H * 32, threadIdx.x == 0);
    }
  }
  if (SLAB) SlabSyncSignal(a.sync, H * 32, threadIdx.x == 0);
}

// 1 where a boundary cell of the two grids differs (bitwise), else untouched.  Local view of a
// z-slab: the x and y faces of the interior planes [zlo, zhi), and the z faces this rank holds
// (local planes zf_lo / zf_hi, -1 when they are on another rank).
__global__ void HimenoFacesDifferKernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                        int nx, int ny, int zlo, int zhi, int zf_lo, int zf_hi, int *flag) {
  const long nxy = (long)nx * ny;
  const long ring = 2L * nx + 2L * ny;          // boundary cells of one plane's rim (corners twice)
  const long rim = (long)(zhi - zlo) * ring;
  const long total = rim + 2 * nxy;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    int x, y, z;
    if (i < rim) {
      z = zlo + (int)(i / ring);
      const long r = i % ring;
      if (r < 2L * nx) {
        y = (r >= nx) ? ny - 1 : 0;
        x = (int)(r % nx);
      } else {
        const long q = r - 2L * nx;
        x = (q >= ny) ? nx - 1 : 0;
        y = (int)(q % ny);
      }
    } else {
      const long r = i - rim;
      z = (r >= nxy) ? zf_hi : zf_lo;
      if (z < 0) continue;
      y = (int)((r % nxy) / nx);
      x = (int)(r % nx);
    }
    const size_t off = ((size_t)z * ny + y) * nx + x;
    if (a[off] != b[off]) *flag = 1;
  }
}

}  // namespace

struct HimenoPairPlan {
  int grid = 0, block = 0;
  size_t smem = 0;
  const void *fn = nullptr;
  CUtensorMap tmap[2];  // direction 0 reads the first grid, direction 1 the second
  HpCoefMaps cmaps;     // the coefficient arrays (prefetch boxes)
  HimenoPairArgs args[2];
  Grid *g[2] = {nullptr, nullptr};
};

// d0: X -> Y, d1: Y -> X, both the Himeno sweep over the interior with the same coefficient
// grids and omega.  Returns nullptr (with a reason) when the pair cannot be fused.
HimenoPairPlan *PrepareHimenoPair(Runtime *rt, const __PSB200StencilDesc &d0,
                                  const __PSB200StencilDesc &d1, std::string *why) {
  const Options &o = rt->opt;
  if (!o.himeno_fuse) { *why = "himeno_fuse=0"; return nullptr; }
  const bool k0 = d0.kind == PSB200_KIND_HIMENO19 || d0.kind == PSB200_KIND_HIMENO19_GOSA;
  const bool k1 = d1.kind == PSB200_KIND_HIMENO19 || d1.kind == PSB200_KIND_HIMENO19_GOSA;
  if (!k0 || !k1 || d0.kind != d1.kind) { *why = "not a pair of Himeno sweeps"; return nullptr; }
  const int ng = d0.kind == PSB200_KIND_HIMENO19_GOSA ? 15 : 14;
  if (d0.num_grids != ng || d1.num_grids != ng || d0.num_scalars != 1 || d1.num_scalars != 1) {
    *why = "expects 14(+1) grids and omega"; return nullptr;
  }
  if (d0.grids[0] != d1.grids[1] || d0.grids[1] != d1.grids[0] || d0.grids[0] == d0.grids[1]) {
    *why = "the sweeps do not ping-pong between two grids"; return nullptr;
  }
  for (int i = 2; i < ng; ++i)
    if (d0.grids[i] != d1.grids[i]) { *why = "the sweeps use different coefficient grids"; return nullptr; }
  if (d0.scalars[0] != d1.scalars[0]) { *why = "the sweeps use different omega"; return nullptr; }
  const bool multi = rt->world() > 1;
  Grid *g[15];
  for (int i = 0; i < ng; ++i) {
    g[i] = Grid::FromHandle(d0.grids[i]);
    if (g[i]->num_dims != 3 || g[i]->type != PS_FLOAT) { *why = "3-D float grids only"; return nullptr; }
    for (int k = 0; k < 3; ++k)
      if (g[i]->dim[k] != g[0]->dim[k]) { *why = "grids must have equal extents"; return nullptr; }
  }
  // neither p grid may double as a coefficient (or residual) grid
  for (int i = 2; i < ng; ++i)
    if (g[i] == g[0] || g[i] == g[1]) { *why = "a p grid is also a coefficient grid"; return nullptr; }
  sweep::SlabSync sync{};
  if (multi) {
    // a fused pass consumes two halo planes per side and delivers two; every rank must take the
    // same decision, so only group-wide quantities enter it
    for (int i = 0; i < 14; ++i)
      if (!g[i]->decomposed || g[i]->halo < 2 || g[i]->halo != g[0]->halo || g[i]->z_off != g[0]->z_off ||
          g[i]->nz_loc != g[0]->nz_loc) {
        *why = "z-slabs need two halo planes (option halo>=2) and identical cuts"; return nullptr;
      }
    if (g[0]->dim[2] / rt->world() < 4) { *why = "z-slabs thinner than four planes"; return nullptr; }
    if (!o.halo_push || !rt->FillSlabSync(&sync)) {
      *why = "needs the in-kernel halo exchange (halo_push=1, sync_mode=2)"; return nullptr;
    }
  }
  const int nx = g[0]->dim[0], ny = g[0]->dim[1];
  const int gnz = g[0]->dim[2];       // global planes
  const int nz = g[0]->ldim[2];       // planes of this rank's allocation (halo planes included)
  const int halo = g[0]->halo;
  if (nx % 4 != 0) { *why = "x extent must be a multiple of 4"; return nullptr; }
  if (nx < 8 || ny < 3 || gnz < 3) { *why = "grid too small"; return nullptr; }
  for (int s = 0; s < 2; ++s) {
    const __PSDomain &dom = s ? d1.dom : d0.dom;
    for (int i = 0; i < 3; ++i)
      if (dom.local_min[i] != 1 || dom.local_max[i] != g[0]->dim[i] - 1) {
        *why = "both sweeps must cover exactly the interior"; return nullptr;
      }
  }
  constexpr int H = 12;
  HimenoPairPlan *p = new HimenoPairPlan();
  p->fn = multi ? (const void *)HimenoPairKernel<H, true> : (const void *)HimenoPairKernel<H, false>;
  p->smem = HpGeom<H>::SMEM;
  p->block = H * 32;
  p->g[0] = g[0];
  p->g[1] = g[1];
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  // shared memory: just the rings
  {
    const int pct = (int)(((p->smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(pct, 100)));
  }
  int occ = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->fn, p->block, p->smem));
  PSB_CHECK(occ > 0, "fused himeno kernel does not fit on an SM");
  const int slots = rt->sm_count * occ;

  // x tiles: 128 columns loaded per tile (+ halo), the tile's own columns stored; one vector
  // per seam side is recomputed by the neighbour
  const int w = Geom<float>::TXB;
  int ntx = 1, seg = nx;
  if (nx > w) {
    for (ntx = 2; ntx <= kHpMaxXTiles; ++ntx) {
      seg = CeilDiv(CeilDiv(nx, ntx), 4) * 4;
      if (seg + 8 <= w) break;
    }
    if (ntx > kHpMaxXTiles) { *why = "rows wider than 32 x tiles"; delete p; return nullptr; }
  }
  const int nty = CeilDiv(ny - 2, H - 2);
  // planes this rank writes: the global interior [1, gnz-1) cut to its slab, in local indices
  const int dz0 = std::max(1, g[0]->z_off) - g[0]->z_off + halo;
  const int dz1 = std::min(gnz - 1, g[0]->z_off + g[0]->nz_loc) - g[0]->z_off + halo;
  const int nzd = dz1 - dz0;
  if (nzd < 1) { *why = "a rank without interior planes"; delete p; return nullptr; }
  // every chunk re-reads 4 planes and recomputes 2 first-sweep planes; items run in waves
  auto plan_zc = [&](int planes) {
    int best_zc = planes;
    long best = -1;
    for (int n = 1; n <= std::max(1, planes / 4); ++n) {
      const int c = CeilDiv(planes, n);
      const long waves = CeilDiv((long)nty * ntx * CeilDiv(planes, c), slots);
      const long cost = waves * (c + 3);
      if (best < 0 || cost < best) { best = cost; best_zc = c; }
    }
    return best_zc;
  };
  int zc = o.himeno_pair_zc > 0 ? o.himeno_pair_zc : plan_zc(nzd);
  zc = std::max(1, std::min(zc, nzd));
  const int nzc = CeilDiv(nzd, zc);
  // Shapes the fused form wastes itself on run sweep by sweep (himeno_fuse=2 fuses regardless):
  // measured (profiles/r2_experiments.txt) 256x128x128 -- three x tiles that use 96 of the tile's
  // 128 columns, 117 work items for 148 SMs -- 92.9 GLUP/s fused against 115.5 sweep by sweep,
  // while 128x64x64 (launch-bound: fewer launches win), 512x256x256 and 1024x512x512 are
  // faster fused.  Judged on an even share of the planes, so that every rank of a group decides alike.
  {
    const int share = std::max(1, (gnz - 2) / rt->world());
    const int szc = o.himeno_pair_zc > 0 ? std::min(o.himeno_pair_zc, share) : plan_zc(share);
    const long items = (long)nty * ntx * CeilDiv(share, szc);
    const double lanes = (double)nx / ((double)ntx * w);
    const double fill = (double)items / (double)(CeilDiv(items, (long)slots) * slots);
    const double cells = (double)nx * ny * share;
    if (o.himeno_fuse == 1 && lanes * fill < 0.7 && cells > 1.5e6) {
      *why = "too few or too narrow work items for the fused form";
      delete p;
      return nullptr;
    }
  }
  for (int i = 0; i < 12; ++i) {
    int dimv[3] = {nx, ny, nz};
    int boxv[3] = {std::min(Geom<float>::TXB, nx), std::min(H, ny), 1};
    if (!EncodeTensorMap3D(&p->cmaps.m[i], TmaElem::F32, g[2 + i]->members[0].dev, dimv, boxv)) {
      *why = "grid shape violates a TMA constraint";
      delete p;
      return nullptr;
    }
  }
  for (int dir = 0; dir < 2; ++dir) {
    int dimv[3] = {nx, ny, nz};
    int boxv[3] = {Geom<float>::BW, H + 2, 1};
    if (!EncodeTensorMap3D(&p->tmap[dir], TmaElem::F32, g[dir]->members[0].dev, dimv, boxv)) {
      *why = "grid shape violates a TMA constraint";
      delete p;
      return nullptr;
    }
    HimenoPairArgs &a = p->args[dir];
    // descriptor grid order: p0,p1,a0..a3,b0..b2,c0..c2,bnd,wrk1[,gosa]
    for (int i = 0; i < 12; ++i) a.coef[i] = (const float *)g[2 + i]->members[0].dev;
    Grid *go = g[1 - dir];
    a.out = (float *)go->members[0].dev;
    a.omega = (float)d0.scalars[0];
    a.nx = nx; a.ny = ny; a.nz = nz;
    a.nty = nty; a.ntx = ntx; a.nzc = nzc; a.zc = zc;
    a.nitems = nty * ntx * nzc;
    for (int t = 0; t < ntx; ++t) {
      a.txs[t] = std::min(t * seg, nx);
      a.txe[t] = std::min((t + 1) * seg, nx);
      a.tx0[t] = ntx == 1 ? 0 : std::max(0, std::min(a.txs[t] - 4, nx - w));
    }
    a.dz0 = dz0;
    a.dz1 = dz1;
    a.pf = std::max(0, std::min(o.himeno_pair_pf, 8));
    // automatic: per-row bulk prefetches once a plane of the 12 arrays is 20 MB or more (measured on
    // eight shapes: they win on 1024x512x512 and 2048x256x256, 24 MB, by 3-4 %; the tensor form wins on
    // everything from 128x64x64 to 640x512x256, 16 MB, by 1-10 %)
    const bool big_planes = (size_t)nx * ny * sizeof(float) * 12 >= (size_t)20 << 20;
    a.pf_tensor = o.himeno_pair_pfmode == 2 ? 1 : o.himeno_pair_pfmode == 1 ? 0 : (big_planes ? 0 : 1);
    a.zface_lo = g[0]->LocalInterior(0);
    a.zface_hi = g[0]->LocalInterior(gnz - 1);
    a.push_lo_z = a.push_hi_z = -(1 << 30);
    a.push_lo_delta = a.push_hi_delta = 0;
    a.sync = sync;
    if (multi) {
      sweep::SlabSyncSetBoundary(&a.sync, o.early_signal != 0, nzd, zc, nzc, nty * ntx, 2);
      const MemberLayout &ml = go->members[0];
      const size_t plane = (size_t)go->plane_elms;
      // lower neighbour's two upper halo planes <- my first two interior planes; upper neighbour's
      // two lower halo planes (the ones next to its interior) <- my last two (planes holding a
      // global z face are never written, so never forwarded: the kernel forwards what it stored)
      a.push_lo_z = halo;
      a.push_hi_z = halo + go->nz_loc - 2;
      const float *to_lo = (const float *)ml.peer_lo + (size_t)(go->halo + go->lo_nz_loc) * plane;
      const float *to_hi = (const float *)ml.peer_hi + (size_t)(go->halo - 2) * plane;
      a.push_lo_delta = (const char *)to_lo - (const char *)(a.out + (size_t)a.push_lo_z * plane);
      a.push_hi_delta = (const char *)to_hi - (const char *)(a.out + (size_t)a.push_hi_z * plane);
    }
  }
  p->grid = std::min(p->args[0].nitems, slots);
  return p;
}

// true when the boundary cells of the pair's two grids are bit-equal (a fused pass takes the
// intermediate field's boundary from the grid it reads) -- on every rank's share of them: the
// decision is group-wide.  Synchronises the stream (and, multi-GPU, the ranks).
bool HimenoPairFacesEqual(Runtime *rt, HimenoPairPlan *p) {
  DeviceBuffer &scr = rt->small_scratch(sizeof(int));
  int *flag = (int *)scr.get();
  PSB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), rt->stream));
  const HimenoPairArgs &a = p->args[0];
  const Grid *g = p->g[0];
  HimenoFacesDifferKernel<<<rt->sm_count * 2, 256, 0, rt->stream>>>(
      (const uint32_t *)p->g[0]->members[0].dev, (const uint32_t *)p->g[1]->members[0].dev, a.nx, a.ny,
      g->halo, g->halo + g->nz_loc, a.zface_lo, a.zface_hi, flag);
  PSB_CUDA(cudaGetLastError());
  rt->stats.kernel_launches++;
  int differ = 0;
  PSB_CUDA(cudaMemcpyAsync(&differ, flag, sizeof(int), cudaMemcpyDeviceToHost, rt->stream));
  PSB_CUDA(cudaStreamSynchronize(rt->stream));
  if (rt->world() > 1) {
    std::vector<int> all(rt->world());
    rt->comm->AllGather(&differ, all.data(), sizeof(int));
    for (int v : all) differ |= v;
  }
  return differ == 0;
}

void LaunchHimenoPair(Runtime *rt, HimenoPairPlan *p, int dir) {
  if (rt->world() > 1) {
    p->args[dir].sync.wait_epoch = rt->sweep_epoch;
    p->args[dir].sync.signal_epoch = rt->sweep_epoch + 1;
  }
  void *args[3] = {&p->tmap[dir], &p->cmaps, &p->args[dir]};
  PSB_CUDA(cudaLaunchKernel(p->fn, dim3(p->grid), dim3(p->block), args, p->smem, rt->stream));
}

void DestroyHimenoPair(HimenoPairPlan *p) { delete p; }

}  // namespace physis_b200
