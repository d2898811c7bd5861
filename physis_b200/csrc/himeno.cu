// Himeno 19-point Jacobi sweep — hand-written sm_100a kernel.
//
// One sweep of examples/himeno/himenobmtxpa_physis.c:331-361 (jacobi_kernel):
//   s0 = a0*p(k+1) + a1*p(j+1) + a2*p(i+1)
//      + b0*(p(j+1,k+1) - p(j-1,k+1) - p(j+1,k-1) + p(j-1,k-1))
//      + b1*(p(i+1,j+1) - p(i+1,j-1) - p(i-1,j+1) + p(i-1,j-1))
//      + b2*(p(i+1,k+1) - p(i+1,k-1) - p(i-1,k+1) + p(i-1,k-1))
//      + c0*p(k-1) + c1*p(j-1) + c2*p(i-1) + wrk1
//   ss = (s0*a3 - p)*bnd;   p1 = p + omega*ss          [; gosa_g = ss*ss]
// with i = x (fastest), j = y, k = z, on the interior domain.  All operations
// are separately rounded fp32 in C's left-to-right order (no FMA), so the result
// is bit-identical to the REFERENCE target.  The `_GOSA` form additionally
// emits ss*ss (examples/dsl/himeno_gosa.c), the DSL spelling of the original
// benchmark's residual (himenobmtxpa_original.c:334).
//
// Traffic per point: 12 coefficient/source streams + p read + p1 write = 56 B
// (himenobmtxpa_physis.c:418-432), +4 B with the residual emit.  HBM-bound.
//  * p0: haloed xy tiles of planes z-1, z, z+1 through the same TMA /
//    mbarrier shared-memory ring as star7.cu (zero fill outside the grid; those
//    values only reach points outside the domain, which are never stored);
//  * the 12 coefficient arrays are pure streams: one 128-bit read-only load per
//    thread per array, issued before the thread waits on the p planes, so ~200 B
//    per thread are in flight;
//  * 128-bit stores; vectors straddling the domain edge store element-wise.
#include "runtime.h"
#include "tma.cuh"
#include "sweep_common.cuh"
#include "himeno_math.cuh"

#include <algorithm>
#include <string>

namespace physis_b200 {

namespace {

using namespace sweep;

struct HimenoArgs {
  const float *a0, *a1, *a2, *a3, *b0, *b1, *b2, *c0, *c1, *c2, *bnd, *wrk1;
  float *p1;
  float *gosa;  // nullptr unless the _GOSA form
  // _GOSA form: one partial sum per CTA of everything the launch emitted into `gosa`, so that
  // the PSReduce(PS_SUM) which follows folds gridDim.x numbers instead of re-reading the grid
  // (the original benchmark accumulates gosa inside the sweep: himenobmtxpa_original.c:334)
  double *gosa_partials;
  float omega;
  int nx, ny, nz;
  int dx0, dx1, dy0, dy1, dz0, dz1;
  int xbase;
  int ntx, nty, nzc, zc, nitems;
  int stages;
  int st_hint;  // bit 0: streaming (evict-first) stores of p1, bit 1: of the residual grid
  // z-slab view (multi-GPU): local planes of p1 that are also stored into the ring
  // neighbours' halo planes through the CUDA-IPC mapping; -1 / nullptr on one GPU
  int push_lo_z, push_hi_z;
  float *push_lo, *push_hi;
  SlabSync sync;  // neighbour ordering fused into the kernel
};

__device__ __forceinline__ float4 LdStream(const float *p) {
  return __ldcs(reinterpret_cast<const float4 *>(p));
}

// TY rows per CTA tile, one row per consumer warp, one box (128 floats) wide.
template <int TY, bool GOSA>
__global__ void __launch_bounds__((TY + 1) * 32)
HimenoKernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ HimenoArgs a) {
  using G = Geom<float>;
  constexpr int VEC = 4;
  constexpr int NW = TY;
  constexpr int ROWB = G::ROW_BYTES;
  constexpr int STAGE_BYTES = BoxStride<float, TY>();

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
  SlabSyncWait(a.sync);
  __syncthreads();

  const int tiles_xy = a.ntx * a.nty;

  if (warp == NW) {
    if (lane != 0) return;
    tma::prefetch_tensormap(&tmap);
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int zseq = item / tiles_xy;
      const int txy = item - zseq * tiles_xy;
      const int ty = txy / a.ntx;
      const int tx = txy - ty * a.ntx;
      const int x0 = a.xbase + tx * G::TXB;
      const int y0 = a.dy0 + ty * TY;
      int zb, ze;
      SlabChunkRange(a.sync, zseq, a.nzc, a.zc, a.dz0, a.dz1, &zb, &ze);
      // the domain is interior in z (checked on the host): planes zb-1 .. ze exist
      for (int z = zb - 1; z <= ze; ++z) {
        tma::mbar_wait(&empty[stage], phase ^ 1u);
        tma::mbar_arrive_expect_tx(&full[stage], (uint32_t)((TY + 2) * ROWB));
        tma::load_3d(planes + stage * STAGE_BYTES, &tmap, &full[stage], x0 - G::HX, y0 - 1, z);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  const int col_off = (G::HX + lane * VEC) * (int)sizeof(float);
  int stage = 0;
  uint32_t phase = 0;
  auto next_of = [&](int st) { return (st + 1 == S) ? 0 : st + 1; };
  auto release = [&](int st) {
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(&empty[st]);
  };

  // sum of the ss*ss values this thread emits: fp32 within a vector, fp64 across planes and
  // work items (one DADD per four points is free next to the loads, and keeps the partial
  // within an ulp of the exact sum whatever the order)
  double gacc = 0.0;

  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int zseq = item / tiles_xy;
    const int txy = item - zseq * tiles_xy;
    const int ty = txy / a.ntx;
    const int tx = txy - ty * a.ntx;
    const int x = a.xbase + tx * G::TXB + lane * VEC;
    const int y = a.dy0 + ty * TY + warp;
    int zb, ze;
    SlabChunkRange(a.sync, zseq, a.nzc, a.zc, a.dz0, a.dz1, &zb, &ze);
    const bool row_ok = (y < a.dy1) && (x < a.nx);
    // per-element store mask
    bool ok[VEC];
    bool all_ok = row_ok;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      ok[j] = row_ok && (x + j >= a.dx0) && (x + j < a.dx1);
      all_ok = all_ok && ok[j];
    }
    bool any_ok = false;
#pragma unroll
    for (int j = 0; j < VEC; ++j) any_ok = any_ok || ok[j];

    // ring positions of planes z-1 (sb), z (sc), z+1 (st)
    int sb = stage;
    uint32_t phb = phase;
    tma::mbar_wait(&full[sb], phb);
    int sc = next_of(sb);
    uint32_t phc = (sc == 0) ? (phb ^ 1u) : phb;
    tma::mbar_wait(&full[sc], phc);

    for (int z = zb; z < ze; ++z) {
      const size_t g = ((size_t)z * a.ny + y) * a.nx + x;
      float4 va0, va1, va2, va3, vb0, vb1, vb2, vc0, vc1, vc2, vbnd, vwrk;
      if (any_ok) {
        va0 = LdStream(a.a0 + g); va1 = LdStream(a.a1 + g); va2 = LdStream(a.a2 + g);
        va3 = LdStream(a.a3 + g); vb0 = LdStream(a.b0 + g); vb1 = LdStream(a.b1 + g);
        vb2 = LdStream(a.b2 + g); vc0 = LdStream(a.c0 + g); vc1 = LdStream(a.c1 + g);
        vc2 = LdStream(a.c2 + g); vbnd = LdStream(a.bnd + g); vwrk = LdStream(a.wrk1 + g);
      }
      const int st = next_of(sc);
      const uint32_t pht = (st == 0) ? (phc ^ 1u) : phc;
      tma::mbar_wait(&full[st], pht);

      const unsigned char *pb = planes + sb * STAGE_BYTES + (warp + 1) * ROWB;
      const unsigned char *pc = planes + sc * STAGE_BYTES + (warp + 1) * ROWB;
      const unsigned char *pt = planes + st * STAGE_BYTES + (warp + 1) * ROWB;
      auto vec = [&](const unsigned char *row, int dy) {
        return *reinterpret_cast<const float4 *>(row + dy * ROWB + col_off);
      };
      auto west = [&](const unsigned char *row, int dy) {
        return *reinterpret_cast<const float *>(row + dy * ROWB + col_off - 4);
      };
      auto east = [&](const unsigned char *row, int dy) {
        return *reinterpret_cast<const float *>(row + dy * ROWB + col_off + 16);
      };
      if (any_ok) {
        // plane z
        const float4 c_c = vec(pc, 0), c_n = vec(pc, -1), c_s = vec(pc, 1);
        const float c_cw = west(pc, 0), c_ce = east(pc, 0);
        const float c_nw = west(pc, -1), c_ne = east(pc, -1);
        const float c_sw = west(pc, 1), c_se = east(pc, 1);
        // plane z-1 and z+1: 5-point star
        const float4 b_c = vec(pb, 0), b_n = vec(pb, -1), b_s = vec(pb, 1);
        const float b_w = west(pb, 0), b_e = east(pb, 0);
        const float4 t_c = vec(pt, 0), t_n = vec(pt, -1), t_s = vec(pt, 1);
        const float t_w = west(pt, 0), t_e = east(pt, 0);

        float4 o, q;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          // x-1 / x+1 values in each needed row
          const float c_xm = (j == 0) ? c_cw : Elem(c_c, j - 1);
          const float c_xp = (j == VEC - 1) ? c_ce : Elem(c_c, j + 1);
          const float n_xm = (j == 0) ? c_nw : Elem(c_n, j - 1);
          const float n_xp = (j == VEC - 1) ? c_ne : Elem(c_n, j + 1);
          const float s_xm = (j == 0) ? c_sw : Elem(c_s, j - 1);
          const float s_xp = (j == VEC - 1) ? c_se : Elem(c_s, j + 1);
          const float b_xm = (j == 0) ? b_w : Elem(b_c, j - 1);
          const float b_xp = (j == VEC - 1) ? b_e : Elem(b_c, j + 1);
          const float t_xm = (j == 0) ? t_w : Elem(t_c, j - 1);
          const float t_xp = (j == VEC - 1) ? t_e : Elem(t_c, j + 1);
          float ss;
          const float v = HimenoJacobi(
              Elem(va0, j), Elem(va1, j), Elem(va2, j), Elem(va3, j), Elem(vb0, j), Elem(vb1, j),
              Elem(vb2, j), Elem(vc0, j), Elem(vc1, j), Elem(vc2, j), Elem(vbnd, j), Elem(vwrk, j),
              a.omega,
              /*ccc*/ Elem(c_c, j), /*ccp*/ Elem(t_c, j), /*cpc*/ Elem(c_s, j), /*pcc*/ c_xp,
              /*cpp*/ Elem(t_s, j), /*cmp*/ Elem(t_n, j), /*cpm*/ Elem(b_s, j), /*cmm*/ Elem(b_n, j),
              /*ppc*/ s_xp, /*pmc*/ n_xp, /*mpc*/ s_xm, /*mmc*/ n_xm,
              /*pcp*/ t_xp, /*pcm*/ b_xp, /*mcp*/ t_xm, /*mcm*/ b_xm,
              /*ccm*/ Elem(b_c, j), /*cmc*/ Elem(c_n, j), /*mcc*/ c_xm, &ss);
          SetElem(o, j, v);
          SetElem(q, j, MulRn(ss, ss));
        }
        if (GOSA) {
          const float e0 = ok[0] ? q.x : 0.f, e1 = ok[1] ? q.y : 0.f;
          const float e2 = ok[2] ? q.z : 0.f, e3 = ok[3] ? q.w : 0.f;
          gacc += (double)AddRn(AddRn(e0, e1), AddRn(e2, e3));
        }
        float *const push0 = (z == a.push_lo_z) ? a.push_lo : nullptr;
        float *const push1 = (z == a.push_hi_z) ? a.push_hi : nullptr;
        const size_t gp = (size_t)y * a.nx + x;  // offset inside one plane
        if (all_ok) {
          StoreVec(reinterpret_cast<float4 *>(a.p1 + g), o, (a.st_hint & 1) != 0);
          if (GOSA) StoreVec(reinterpret_cast<float4 *>(a.gosa + g), q, (a.st_hint & 2) != 0);
          if (push0) *reinterpret_cast<float4 *>(push0 + gp) = o;
          if (push1) *reinterpret_cast<float4 *>(push1 + gp) = o;
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            if (ok[j]) {
              a.p1[g + j] = Elem(o, j);
              if (GOSA) a.gosa[g + j] = Elem(q, j);
              if (push0) push0[gp + j] = Elem(o, j);
              if (push1) push1[gp + j] = Elem(o, j);
            }
          }
        }
      }
      // plane z-1 is no longer needed
      release(sb);
      sb = sc; phb = phc;
      sc = st; phc = pht;
    }
    // planes ze-1 (sb) and ze (sc) were loaded for this item and are still held
    release(sb);
    release(sc);
    // the next item's first plane follows `sc` in the ring
    stage = next_of(sc);
    phase = (stage == 0) ? (phc ^ 1u) : phc;
    SlabSyncItemDone(a.sync, item, NW * 32, threadIdx.x == 0);
  }
  if (GOSA) {
    // per-CTA partial in a fixed order: lanes by a shuffle tree, warps one after the other
    __shared__ double warp_part[NW];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) gacc += __shfl_down_sync(0xffffffffu, gacc, d);
    if (lane == 0) warp_part[warp] = gacc;
    asm volatile("bar.sync 2, %0;" ::"r"(NW * 32) : "memory");  // consumers only: the producer has left
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += warp_part[w];
      a.gosa_partials[blockIdx.x] = t;
    }
  }
  SlabSyncSignal(a.sync, NW * 32, threadIdx.x == 0);
}

template <int TY>
size_t SmemBytes(int stages) {
  return kBarrierBytes + (size_t)stages * BoxStride<float, TY>();
}

}  // namespace

struct HimenoPlan {
  int grid = 0, block = 0;
  size_t smem = 0;
  CUtensorMap tmap;
  HimenoArgs args;
  const void *fn = nullptr;
  bool pushes = false;  // the kernel itself delivers the halo planes of p1
  bool syncs = false;   // ... and waits for / signals the neighbours itself
};

HimenoPlan *PrepareHimeno(Runtime *rt, const __PSB200StencilDesc &d, std::string *why) {
  const bool gosa = d.kind == PSB200_KIND_HIMENO19_GOSA;
  const int ng = gosa ? 15 : 14;
  if (d.num_grids != ng || d.num_scalars != 1) { *why = "expects 14(+1) grids and omega"; return nullptr; }
  Grid *g[15];
  for (int i = 0; i < ng; ++i) {
    g[i] = Grid::FromHandle(d.grids[i]);
    if (g[i]->num_dims != 3 || g[i]->type != PS_FLOAT) { *why = "3-D float grids only"; return nullptr; }
    for (int k = 0; k < 3; ++k)
      if (g[i]->dim[k] != g[0]->dim[k] || g[i]->ldim[k] != g[0]->ldim[k]) {
        *why = "grids must have equal extents"; return nullptr;
      }
  }
  if (g[0] == g[1]) { *why = "in-place sweep"; return nullptr; }
  const int nx = g[0]->ldim[0], ny = g[0]->ldim[1], nz = g[0]->ldim[2];  // local allocation
  const __PSDomain &dom = d.dom;
  if (nx % 4 != 0) { *why = "x extent must be a multiple of 4"; return nullptr; }
  // every read p(x±1, y±1, z±1) must stay inside the grid
  if (dom.local_min[0] < 1 || dom.local_max[0] > nx - 1 || dom.local_min[1] < 1 ||
      dom.local_max[1] > ny - 1 || dom.local_min[2] < 1 || dom.local_max[2] > nz - 1) {
    *why = "domain must be interior (19-point reach)";
    return nullptr;
  }
  if (dom.local_max[0] <= dom.local_min[0] || dom.local_max[1] <= dom.local_min[1] ||
      dom.local_max[2] <= dom.local_min[2]) { *why = "empty domain"; return nullptr; }

  HimenoPlan *p = new HimenoPlan();
  // rows per CTA tile = consumer warps; +1 producer warp.  Warps are allocated in
  // groups of four, so 7+1 (two CTAs per SM) and 15+1 fill the register file where
  // 8+1 strands three warps' worth.
  int TY = rt->opt.himeno_by;
  // measured (profiles/r1_tune_himeno_XL.csv, r2_experiments.txt): 15 rows best on XL, 7 rows (two
  // CTAs per SM, twice the work items) on 512x256x256 and smaller
  if (TY != 7 && TY != 15) TY = (double)nx * ny * (dom.local_max[2] - dom.local_min[2]) >= 1.0e8 ? 15 : 7;
  int stages = rt->opt.himeno_stages > 0 ? std::min(rt->opt.himeno_stages, kMaxStages) : 6;
  if (stages < 4) stages = 4;
  if (TY == 7) {
    p->fn = gosa ? (const void *)HimenoKernel<7, true> : (const void *)HimenoKernel<7, false>;
    p->smem = SmemBytes<7>(stages);
  } else {
    p->fn = gosa ? (const void *)HimenoKernel<15, true> : (const void *)HimenoKernel<15, false>;
    p->smem = SmemBytes<15>(stages);
  }
  p->block = (TY + 1) * 32;
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  // The coefficient streams are ordinary global loads: the in-flight lines live in L1,
  // so the shared-memory carve-out is only what the p ring needs and L1 keeps the rest.
  PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->fn, p->block, p->smem));
  PSB_CHECK(occ > 0, "himeno kernel does not fit on an SM");
  if (rt->opt.himeno_occ > 0) occ = std::min(occ, rt->opt.himeno_occ);
  {
    const size_t need = (p->smem + 1024) * (size_t)occ;
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    pct = std::max(rt->opt.himeno_carveout > 0 ? rt->opt.himeno_carveout : pct, pct);
    PSB_CUDA(cudaFuncSetAttribute(p->fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  std::min(pct, 100)));
  }

  HimenoArgs &a = p->args;
  const float **coef[] = {&a.a0, &a.a1, &a.a2, &a.a3, &a.b0, &a.b1, &a.b2, &a.c0, &a.c1, &a.c2,
                          &a.bnd, &a.wrk1};
  // descriptor grid order = kernel parameter order: p0,p1,a0..a3,b0..b2,c0..c2,bnd,wrk1[,gosa]
  for (int i = 0; i < 12; ++i) *coef[i] = (const float *)g[2 + i]->members[0].dev;
  a.p1 = (float *)g[1]->members[0].dev;
  a.gosa = gosa ? (float *)g[14]->members[0].dev : nullptr;
  a.gosa_partials = nullptr;
  a.omega = (float)d.scalars[0];
  a.nx = nx; a.ny = ny; a.nz = nz;
  a.dx0 = dom.local_min[0]; a.dx1 = dom.local_max[0];
  a.dy0 = dom.local_min[1]; a.dy1 = dom.local_max[1];
  a.dz0 = dom.local_min[2]; a.dz1 = dom.local_max[2];
  a.xbase = a.dx0 / 4 * 4;
  a.ntx = CeilDiv(a.dx1 - a.xbase, Geom<float>::TXB);
  a.nty = CeilDiv(a.dy1 - a.dy0, TY);
  const int nzd = a.dz1 - a.dz0;
  const int slots = rt->sm_count * occ;
  int zc = rt->opt.himeno_zc;
  if (zc <= 0) {
    int tiles = a.ntx * a.nty;
    int want_chunks = std::max(1, CeilDiv(4L * slots, tiles));
    zc = std::min(32, std::max(8, CeilDiv(nzd, want_chunks)));
    zc = std::min(zc, nzd);
  }
  a.zc = zc;
  a.nzc = CeilDiv(nzd, zc);
  a.nitems = a.ntx * a.nty * a.nzc;
  a.stages = stages;
  a.st_hint = rt->opt.himeno_sthint;
  p->grid = std::min(a.nitems, slots);
  if (gosa) {
    Grid::SumCache &sc = g[14]->sum_cache;
    if (!sc.partials) sc.partials = new DeviceBuffer();
    sc.partials->EnsureCapacity(sizeof(double) * (size_t)slots, rt->stream);
    a.gosa_partials = (double *)sc.partials->get();
  }
  a.push_lo_z = a.push_hi_z = -1;
  a.sync = SlabSync{};
  if (SlabPushTargets(rt, *g[1], 0, (void **)&a.push_lo, (void **)&a.push_hi, sizeof(float))) {
    a.push_lo_z = g[1]->halo;
    a.push_hi_z = g[1]->halo + g[1]->nz_loc - 1;
    p->pushes = true;
    if (rt->FillSlabSync(&a.sync)) {
      p->syncs = true;
      SlabSyncPlanEnds(&a.sync, rt->opt.early_signal != 0, nzd, &a.zc, &a.nzc, a.ntx * a.nty, 1, rt->opt.slab_zbl);
      a.nitems = a.ntx * a.nty * a.nzc;
      p->grid = std::min(a.nitems, slots);
    }
  }

  int dimv[3] = {nx, ny, nz};
  int boxv[3] = {Geom<float>::BW, TY + 2, 1};
  if (!EncodeTensorMap3D(&p->tmap, TmaElem::F32, g[0]->members[0].dev, dimv, boxv)) {
    *why = "grid shape violates a TMA constraint";
    delete p;
    return nullptr;
  }
  return p;
}

void LaunchHimeno(Runtime *rt, HimenoPlan *p) {
  if (p->syncs) {
    p->args.sync.wait_epoch = rt->sweep_epoch;
    p->args.sync.signal_epoch = rt->sweep_epoch + 1;
  }
  void *args[2] = {&p->tmap, &p->args};
  PSB_CUDA(cudaLaunchKernel(p->fn, dim3(p->grid), dim3(p->block), args, p->smem, rt->stream));
}

void DestroyHimeno(HimenoPlan *p) { delete p; }
bool HimenoPushes(const HimenoPlan *p) { return p->pushes; }
bool HimenoSyncs(const HimenoPlan *p) { return p->syncs; }
int HimenoPartialCount(const HimenoPlan *p) { return p->args.gosa_partials ? p->grid : 0; }

}  // namespace physis_b200
