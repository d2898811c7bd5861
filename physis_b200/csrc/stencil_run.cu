// __PSB200StencilRun: the b200 replacement for the host run function the
// translator generates per PSStencilRun call
// (translator/cuda_runtime_builder.cc:1465-1583: build launch dims once, then
// `for (i < iter) { kernel_0<<<>>>; kernel_1<<<>>>; ... }` on one stream with
// no synchronisation; TRACE_KERNEL wrapper from
// translator/reference_runtime_builder.cc:896-940).
#include "runtime.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace physis_b200 {

struct Star7Plan;
Star7Plan *PrepareStar7(Runtime *rt, const __PSB200StencilDesc &d, std::string *why);
void LaunchStar7(Runtime *rt, Star7Plan *p);
void DestroyStar7(Star7Plan *p);
bool Star7Pushes(const Star7Plan *p);
bool Star7Syncs(const Star7Plan *p);
struct Star7PairPlan;
Star7PairPlan *PrepareStar7Pair(Runtime *rt, const __PSB200StencilDesc &d0,
                                const __PSB200StencilDesc &d1, std::string *why);
void LaunchStar7Pair(Runtime *rt, Star7PairPlan *p, int dir);
void DestroyStar7Pair(Star7PairPlan *p);

struct HimenoPlan;
HimenoPlan *PrepareHimeno(Runtime *rt, const __PSB200StencilDesc &d, std::string *why);
void LaunchHimeno(Runtime *rt, HimenoPlan *p);
void DestroyHimeno(HimenoPlan *p);
bool HimenoPushes(const HimenoPlan *p);
bool HimenoSyncs(const HimenoPlan *p);
int HimenoPartialCount(const HimenoPlan *p);

struct HimenoPairPlan;
HimenoPairPlan *PrepareHimenoPair(Runtime *rt, const __PSB200StencilDesc &d0,
                                  const __PSB200StencilDesc &d1, std::string *why);
bool HimenoPairFacesEqual(Runtime *rt, HimenoPairPlan *p);
void LaunchHimenoPair(Runtime *rt, HimenoPairPlan *p, int dir);
void DestroyHimenoPair(HimenoPairPlan *p);

struct PstagPlan;
PstagPlan *PreparePstag(Runtime *rt, const __PSB200StencilDesc &d, std::string *why);
void LaunchPstag(Runtime *rt, PstagPlan *p);
void DestroyPstag(PstagPlan *p);
bool PstagPushes(const PstagPlan *p);
bool PstagSyncs(const PstagPlan *p);

struct SweepPlan {
  int kind = 0;
  std::string name;
  Star7Plan *star7 = nullptr;
  HimenoPlan *himeno = nullptr;
  PstagPlan *pstag = nullptr;
  __PSB200LaunchFunc launch = nullptr;
  const void *stencil = nullptr;
  __PSDomain dom;              // this rank's part of the domain
  bool empty = false;          // ... which may be nothing
  bool fused_push = false;     // the kernel stores the halo planes of what it writes
  bool fused_sync = false;     // ... and orders itself with the neighbours (SlabSync)
  // (grid, member) pairs whose halo planes must reach the neighbours after the sweep
  std::vector<std::pair<Grid *, int>> written;
  // bookkeeping of what the grids hold (Grid::SumCache): the sweep's domain in global
  // coordinates, every grid it emits into, and the grid whose emitted values the kernel also
  // sums up per CTA (the residual form of the Himeno sweep)
  __PSDomain gdom;
  std::vector<Grid *> outputs;
  Grid *sum_grid = nullptr;
  bool cached = false;         // owned by the plan cache, not by the run that prepared it
};

// ---- plan cache ---------------------------------------------------------------------------
// The reference's generated run function builds its launch configuration once per call
// (translator/cuda_runtime_builder.cc:1465-1510) and that costs nothing; here a plan holds
// encoded TMA descriptors, kernel attributes and an occupancy query, which a program calling
// `PSStencilRun(..., 1)` in a loop would pay per sweep.  Plans of the hand-written families are
// therefore kept, keyed by everything they were derived from (family, domain, the grids'
// identities, members, scalars), until a grid is freed or an option changes.  Generic sweeps
// are not kept: their plan is two pointers, and the stencil struct they point to is the caller's.
namespace {

struct PlanCache {
  std::map<std::string, SweepPlan *> sweeps;
  std::map<std::string, Star7PairPlan *> pairs;
  std::map<std::string, HimenoPairPlan *> himeno_pairs;
  uint64_t hits = 0, misses = 0;
};
PlanCache g_plans;
// the tuner prepares plans under option overrides: they are cached beside the default ones
std::string g_plan_tag;
// ... and asks whether a form can run a shape at all, which must not end the program
bool g_soft_fail = false;

template <typename T>
void KeyAdd(std::string *k, const T &v) { k->append(reinterpret_cast<const char *>(&v), sizeof(T)); }

std::string DescKey(const __PSB200StencilDesc &d) {
  std::string k;
  KeyAdd(&k, d.kind);
  KeyAdd(&k, d.elm_type);
  KeyAdd(&k, d.dom);
  KeyAdd(&k, d.num_grids);
  for (int i = 0; i < d.num_grids; ++i) {
    KeyAdd(&k, d.grids[i]);
    KeyAdd(&k, Grid::FromHandle(d.grids[i])->id);  // ids are never reused, addresses are
    KeyAdd(&k, d.members[i]);
  }
  KeyAdd(&k, d.num_scalars);
  for (int i = 0; i < d.num_scalars; ++i) KeyAdd(&k, d.scalars[i]);
  KeyAdd(&k, d.written_mask);
  KeyAdd(&k, d.z_reach);
  k += g_plan_tag;
  return k;
}

}  // namespace

void ClearPlanCache() {
  for (auto &kv : g_plans.sweeps) {
    kv.second->cached = false;
    DestroySweep(kv.second);
  }
  g_plans.sweeps.clear();
  for (auto &kv : g_plans.pairs)
    if (kv.second) DestroyStar7Pair(kv.second);
  g_plans.pairs.clear();
  for (auto &kv : g_plans.himeno_pairs)
    if (kv.second) DestroyHimenoPair(kv.second);
  g_plans.himeno_pairs.clear();
}

static SweepPlan *GetSweepPlan(Runtime *rt, const __PSB200StencilDesc &d) {
  if (d.kind == PSB200_KIND_GENERIC || !rt->opt.plan_cache || g_soft_fail) return PrepareSweep(rt, d);
  const std::string key = DescKey(d);
  auto it = g_plans.sweeps.find(key);
  if (it != g_plans.sweeps.end()) {
    ++g_plans.hits;
    rt->stats.plan_cache_hits++;
    return it->second;
  }
  ++g_plans.misses;
  SweepPlan *p = PrepareSweep(rt, d);
  if (p->kind != PSB200_KIND_GENERIC) {  // a family kernel took it (not the generic fallback)
    p->cached = true;
    g_plans.sweeps[key] = p;
  }
  return p;
}

// nullptr when the pair cannot be fused (a decision that is cached too)
static Star7PairPlan *GetPairPlan(Runtime *rt, const __PSB200StencilDesc &d0,
                                  const __PSB200StencilDesc &d1, bool *owned) {
  std::string why;
  *owned = true;
  if (!rt->opt.plan_cache) return PrepareStar7Pair(rt, d0, d1, &why);
  const std::string key = DescKey(d0) + DescKey(d1);
  auto it = g_plans.pairs.find(key);
  if (it == g_plans.pairs.end()) {
    ++g_plans.misses;
    it = g_plans.pairs.emplace(key, PrepareStar7Pair(rt, d0, d1, &why)).first;
  } else {
    ++g_plans.hits;
  }
  *owned = false;
  return it->second;
}

static HimenoPairPlan *GetHimenoPairPlan(Runtime *rt, const __PSB200StencilDesc &d0,
                                         const __PSB200StencilDesc &d1, bool *owned) {
  std::string why;
  *owned = true;
  if (!rt->opt.plan_cache) return PrepareHimenoPair(rt, d0, d1, &why);
  const std::string key = DescKey(d0) + DescKey(d1);
  auto it = g_plans.himeno_pairs.find(key);
  if (it == g_plans.himeno_pairs.end()) {
    ++g_plans.misses;
    it = g_plans.himeno_pairs.emplace(key, PrepareHimenoPair(rt, d0, d1, &why)).first;
  } else {
    ++g_plans.hits;
  }
  *owned = false;
  return it->second;
}

// Clips the z-range of a domain to this rank's slab.  Specialised kernels work on
// the local allocation (local plane indices); generic kernels index through the
// shifted device view in global coordinates.
static bool LocaliseDomain(const Grid *g, __PSDomain *dom, bool local_coords) {
  if (!g->decomposed) return dom->local_max[2] > dom->local_min[2];
  int lo = std::max(dom->local_min[2], g->z_off);
  int hi = std::min(dom->local_max[2], g->z_off + g->nz_loc);
  if (hi <= lo) {
    dom->local_min[2] = dom->local_max[2] = 0;
    return false;
  }
  const int shift = local_coords ? g->halo - g->z_off : 0;
  dom->local_min[2] = lo + shift;
  dom->local_max[2] = hi + shift;
  return true;
}

static const char *KindName(int k) {
  switch (k) {
    case PSB200_KIND_GENERIC: return "generic";
    case PSB200_KIND_DIFFUSION7_CLAMP: return "diffusion7_clamp";
    case PSB200_KIND_HIMENO19: return "himeno19";
    case PSB200_KIND_HIMENO19_GOSA: return "himeno19_gosa";
    case PSB200_KIND_PERIODIC7_STAGGERED: return "periodic7_staggered";
    default: return "?";
  }
}

SweepPlan *PrepareSweep(Runtime *rt, const __PSB200StencilDesc &d_in) {
  SweepPlan *p = new SweepPlan();
  p->kind = d_in.kind;
  p->name = d_in.name ? d_in.name : KindName(d_in.kind);
  std::string why;
  Grid *g0 = d_in.num_grids > 0 ? Grid::FromHandle(d_in.grids[0]) : nullptr;
  const bool multi = rt->world() > 1;
  p->gdom = d_in.dom;
  switch (d_in.kind) {
    case PSB200_KIND_DIFFUSION7_CLAMP:
    case PSB200_KIND_HIMENO19:
      if (d_in.num_grids > 1) p->outputs.push_back(Grid::FromHandle(d_in.grids[1]));
      break;
    case PSB200_KIND_HIMENO19_GOSA:
      if (d_in.num_grids > 1) p->outputs.push_back(Grid::FromHandle(d_in.grids[1]));
      if (d_in.num_grids > 14) p->sum_grid = Grid::FromHandle(d_in.grids[14]);
      break;
    case PSB200_KIND_PERIODIC7_STAGGERED:
      if (g0) p->outputs.push_back(g0);
      break;
    default:
      break;
  }
  if (multi && d_in.kind != PSB200_KIND_GENERIC && g0) {
    __PSB200StencilDesc d = d_in;
    p->empty = !LocaliseDomain(g0, &d.dom, true);
    p->dom = d.dom;
    if (p->empty) {
      // nothing of the domain lives here; the rank still takes part in the ordering
      return p;
    }
    switch (d.kind) {
      case PSB200_KIND_DIFFUSION7_CLAMP:
        p->star7 = PrepareStar7(rt, d, &why);
        if (p->star7) {
          p->fused_push = Star7Pushes(p->star7);
          p->fused_sync = Star7Syncs(p->star7);
          p->written.push_back({Grid::FromHandle(d.grids[1]), 0});
          return p;
        }
        break;
      case PSB200_KIND_HIMENO19:
      case PSB200_KIND_HIMENO19_GOSA:
        p->himeno = PrepareHimeno(rt, d, &why);
        if (p->himeno) {
          p->fused_push = HimenoPushes(p->himeno);
          p->fused_sync = HimenoSyncs(p->himeno);
          p->written.push_back({Grid::FromHandle(d.grids[1]), 0});
          return p;
        }
        break;
      case PSB200_KIND_PERIODIC7_STAGGERED:
        p->pstag = PreparePstag(rt, d, &why);
        if (p->pstag) {
          p->fused_push = PstagPushes(p->pstag);
          p->fused_sync = PstagSyncs(p->pstag);
          p->written.push_back({Grid::FromHandle(d.grids[0]), d.members[1]});
          return p;
        }
        break;
      default:
        why = "unknown sweep kind";
    }
  } else if (d_in.kind != PSB200_KIND_GENERIC) {
    const __PSB200StencilDesc &d = d_in;
    p->dom = d.dom;
    switch (d.kind) {
      case PSB200_KIND_DIFFUSION7_CLAMP:
        p->star7 = PrepareStar7(rt, d, &why);
        if (p->star7) return p;
        break;
      case PSB200_KIND_HIMENO19:
      case PSB200_KIND_HIMENO19_GOSA:
        p->himeno = PrepareHimeno(rt, d, &why);
        if (p->himeno) return p;
        break;
      case PSB200_KIND_PERIODIC7_STAGGERED:
        p->pstag = PreparePstag(rt, d, &why);
        if (p->pstag) return p;
        break;
      default:
        why = "unknown sweep kind";
    }
  }
  // Shapes a specialised kernel does not cover run through the program's own
  // generic per-point kernel (still on the GPU); without one this is fatal.
  if (d_in.launch) {
    p->kind = PSB200_KIND_GENERIC;
    p->launch = d_in.launch;
    p->stencil = d_in.stencil;
    p->dom = d_in.dom;
    p->empty = false;
    p->fused_push = false;
    p->fused_sync = false;
    p->written.clear();
    // which grids a generated kernel emits into is only known to the translator (written_mask)
    p->outputs.clear();
    p->sum_grid = nullptr;
    for (int i = 0; i < d_in.num_grids; ++i)
      if (!d_in.written_mask || (d_in.written_mask & (1u << i))) p->outputs.push_back(Grid::FromHandle(d_in.grids[i]));
    if (multi && g0) {
      // the 3-D grid of the sweep that is decomposed decides the cut
      const Grid *cut = nullptr;
      for (int i = 0; i < d_in.num_grids && !cut; ++i)
        if (Grid::FromHandle(d_in.grids[i])->decomposed) cut = Grid::FromHandle(d_in.grids[i]);
      if (cut) p->empty = !LocaliseDomain(cut, &p->dom, false);
      // a generated kernel reads its z neighbours from the halo planes: they must be as wide
      // as its reach (the reference derives the halo width from the same number,
      // runtime/grid_space_mpi.h:42-62,337-340; here it is the option `halo`)
      for (int i = 0; i < d_in.num_grids; ++i) {
        const Grid *g = Grid::FromHandle(d_in.grids[i]);
        if (!g->decomposed) continue;
        if (std::max(d_in.z_reach, 1) > g->halo) {
          fprintf(stderr, "[physis-b200] sweep '%s' reads %d planes beyond its own in z but grid %d has "
                          "%d halo plane(s) per side: run with PHYSIS_B200_OPTIONS=halo=%d (or on one GPU)\n",
                  p->name.c_str(), d_in.z_reach, g->id, g->halo, d_in.z_reach);
          exit(1);
        }
      }
      // which grids a generated kernel writes is only known to the translator
      // (written_mask); without it every grid of the sweep is refreshed
      for (int i = 0; i < d_in.num_grids; ++i) {
        Grid *g = Grid::FromHandle(d_in.grids[i]);
        if (!g->decomposed) continue;
        if (d_in.written_mask && !(d_in.written_mask & (1u << i))) continue;
        for (size_t m = 0; m < g->members.size(); ++m) p->written.push_back({g, (int)m});
      }
    }
    return p;
  }
  if (g_soft_fail) {
    delete p;
    return nullptr;
  }
  fprintf(stderr, "[physis-b200] sweep '%s' (%s) cannot run: %s, and the program carries no "
                  "generic launch stub. There is no CPU fallback.\n",
          p->name.c_str(), KindName(d_in.kind), why.c_str());
  exit(1);
}

void LaunchSweep(Runtime *rt, SweepPlan *p) {
  const bool multi = rt->world() > 1;
  // neighbours must have finished the previous sweep: their halo stores into this
  // rank are complete, and they no longer read the halo planes this sweep overwrites
  const bool self_sync = multi && p->fused_sync && !p->empty;
  if (multi && !self_sync) rt->WaitNeighbours(rt->sweep_epoch);
  // what the grids hold after this sweep
  const bool keep_sum = p->sum_grid && (p->himeno || p->empty) && p->sum_grid->ZeroOutside(p->gdom);
  for (Grid *g : p->outputs) g->NoteEmit(p->gdom);
  if (p->sum_grid) {
    p->sum_grid->NoteEmit(p->gdom);
    if (keep_sum) {
      // everything outside the domain is still the zero fill: the per-CTA partials of this launch
      // add up to the sum of the whole grid (a rank without a share of the domain contributes 0)
      p->sum_grid->sum_cache.valid = true;
      p->sum_grid->sum_cache.count = p->empty ? 0 : HimenoPartialCount(p->himeno);
    }
  }
  if (!p->empty) {
    if (p->star7) LaunchStar7(rt, p->star7);
    else if (p->himeno) LaunchHimeno(rt, p->himeno);
    else if (p->pstag) LaunchPstag(rt, p->pstag);
    else p->launch(p->stencil, &p->dom, (__PSB200Stream)rt->stream);
    rt->stats.kernel_launches++;
  }
  if (multi) {
    if (!p->fused_push)
      for (auto &w : p->written) rt->PushHalos(*w.first, w.second);
    ++rt->sweep_epoch;
    if (!self_sync) rt->SignalNeighbours(rt->sweep_epoch);
  }
}

void DestroySweep(SweepPlan *p) {
  if (p->cached) return;  // the plan cache owns it
  if (p->star7) DestroyStar7(p->star7);
  if (p->himeno) DestroyHimeno(p->himeno);
  if (p->pstag) DestroyPstag(p->pstag);
  delete p;
}

const char *SweepName(const SweepPlan *p) { return p->name.c_str(); }

}  // namespace physis_b200

using namespace physis_b200;

// Schedule of a fusable ping-pong pair run for `iter` iterations: the first
// __PSB200FusedPassCount(iter) iterations run as that many fused two-sweep passes (pass i
// reads the grid pass i-1 wrote), the rest sweep by sweep.  The count is even, so the newest
// field is back in the first grid when the single sweeps start, and at least one iteration
// stays unfused, so the second grid ends up holding the second-newest field.
extern "C" int __PSB200FusedPassCount(int iter) { return iter >= 3 ? ((iter - 1) & ~1) : 0; }

namespace {

struct SegmentTiming {
  bool timed = false;
  cudaEvent_t emid = nullptr;  // recorded after the fused passes of the segment
  int fused = 0;
};

// `iter` iterations of the stencils under the options in force.  `whole_fused` (a tuning trial
// in the middle of a run; iter even): every iteration may run as a fused pass, the iterations
// that follow put the second grid right.
void RunSchedule(Runtime *rt, int iter, int num_stencils, const __PSB200StencilDesc *descs,
                 bool whole_fused, SegmentTiming *tm) {
  if (iter <= 0) return;
  std::vector<SweepPlan *> plans;
  plans.reserve(num_stencils);
  for (int s = 0; s < num_stencils; ++s) plans.push_back(GetSweepPlan(rt, descs[s]));
  const int fusable = whole_fused ? (iter & ~1) : __PSB200FusedPassCount(iter);
  // A ping-pong pair of whole-grid clamped 7-point sweeps (A -> B, B -> A) runs as fused
  // two-sweep passes (star7_pair.cu).  A pass reads one grid and writes the other, so an
  // even number of passes leaves the newest field in A; the last iteration(s) run
  // unfused so that B ends up holding the second-newest field, exactly as the
  // sweep-by-sweep schedule leaves it.
  int first_unfused = 0;
  Star7PairPlan *pair = nullptr;
  bool pair_owned = false;
  if (num_stencils == 2 && fusable > 0 && plans[0]->star7 && plans[1]->star7) {
    pair = GetPairPlan(rt, descs[0], descs[1], &pair_owned);
    if (pair) {
      first_unfused = fusable;
      const bool multi = rt->world() > 1;
      if (multi) {
        // single sweeps keep only the halo plane next to the interior current; a fused
        // pass reads two, so the input grid's halos are refreshed once per run
        rt->WaitNeighbours(rt->sweep_epoch);
        rt->PushAllHalos(*Grid::FromHandle(descs[0].grids[0]));
        ++rt->sweep_epoch;
        rt->SignalNeighbours(rt->sweep_epoch);
      }
      for (int s = 0; s < 2; ++s) Grid::FromHandle(descs[0].grids[s])->NoteEmit(descs[0].dom);
      for (int i = 0; i < first_unfused; ++i) {
        LaunchStar7Pair(rt, pair, i & 1);
        if (multi) ++rt->sweep_epoch;
        rt->stats.kernel_launches++;
        rt->stats.fused_pairs++;
      }
    }
  }
  // The same for a ping-pong pair of Himeno sweeps (himeno_pair.cu): a pass takes the boundary
  // cells of the intermediate field from the grid it reads, so it runs only while the boundary
  // cells of the two grids are bit-equal (one small comparison kernel per run).
  HimenoPairPlan *hpair = nullptr;
  bool hpair_owned = false;
  if (!pair && num_stencils == 2 && fusable > 0 && plans[0]->himeno && plans[1]->himeno) {
    hpair = GetHimenoPairPlan(rt, descs[0], descs[1], &hpair_owned);
    if (hpair && HimenoPairFacesEqual(rt, hpair)) {
      first_unfused = fusable;
      for (int s = 0; s < 2; ++s) Grid::FromHandle(descs[0].grids[s])->NoteEmit(descs[0].dom);
      const bool multi = rt->world() > 1;
      if (multi) {
        // as for the 7-point pair: single sweeps keep only the halo plane next to the interior
        // current, a fused pass reads two
        rt->WaitNeighbours(rt->sweep_epoch);
        rt->PushAllHalos(*Grid::FromHandle(descs[0].grids[0]));
        ++rt->sweep_epoch;
        rt->SignalNeighbours(rt->sweep_epoch);
      }
      for (int i = 0; i < first_unfused; ++i) {
        LaunchHimenoPair(rt, hpair, i & 1);
        if (multi) ++rt->sweep_epoch;
        rt->stats.kernel_launches++;
        rt->stats.fused_pairs++;
      }
    }
  }
  if (tm && tm->timed && first_unfused > 0) {
    PSB_CUDA(cudaEventCreate(&tm->emid));
    PSB_CUDA(cudaEventRecord(tm->emid, rt->stream));
    tm->fused = first_unfused;
  }
  // A residual-emitting Himeno sweep whose residual grid a later sweep of this run overwrites
  // over the same domain (the second sweep of the ping-pong pair, or the next iteration) runs in
  // its plain form: the emission would be dead (4 B per point written for nothing, and the
  // per-CTA partial sums with it).  The grid ends up with the last sweep's values either way.
  std::vector<SweepPlan *> lean(num_stencils, nullptr);
  std::vector<char> dead_in_iter(num_stencils, 0), dead_before_next(num_stencils, 0);
  if (rt->opt.reduce_fuse && first_unfused < iter) {
    for (int s = 0; s < num_stencils; ++s) {
      if (descs[s].kind != PSB200_KIND_HIMENO19_GOSA || !plans[s]->himeno || descs[s].num_grids < 15) continue;
      // ... unless some stencil of the run uses that grid in any other role (it may read it)
      bool other_use = false;
      for (int t = 0; t < num_stencils && !other_use; ++t)
        for (int i = 0; i < descs[t].num_grids && !other_use; ++i)
          other_use = descs[t].grids[i] == descs[s].grids[14] &&
                      !(descs[t].kind == PSB200_KIND_HIMENO19_GOSA && i == 14);
      if (other_use) continue;
      for (int t = 0; t < num_stencils; ++t) {
        const bool same = descs[t].kind == PSB200_KIND_HIMENO19_GOSA && plans[t]->himeno && descs[t].num_grids >= 15 &&
                          descs[t].grids[14] == descs[s].grids[14] &&
                          memcmp(&descs[t].dom, &descs[s].dom, sizeof(descs[s].dom)) == 0;
        if (!same) continue;
        if (t > s) dead_in_iter[s] = 1;
        dead_before_next[s] = 1;  // (t == s: the sweep itself, one iteration later)
      }
      if (!dead_in_iter[s] && !(dead_before_next[s] && iter - first_unfused > 1)) continue;
      __PSB200StencilDesc twin = descs[s];
      twin.kind = PSB200_KIND_HIMENO19;
      twin.num_grids = 14;
      lean[s] = GetSweepPlan(rt, twin);
      if (!lean[s]->himeno) {
        DestroySweep(lean[s]);
        lean[s] = nullptr;
      }
    }
  }
  for (int i = first_unfused; i < iter; ++i)
    for (int s = 0; s < num_stencils; ++s) {
      const bool dead = lean[s] && (dead_in_iter[s] || (dead_before_next[s] && i + 1 < iter));
      LaunchSweep(rt, dead ? lean[s] : plans[s]);
    }
  for (auto *p : lean)
    if (p) DestroySweep(p);
  for (auto *p : plans) DestroySweep(p);
  if (pair && pair_owned) DestroyStar7Pair(pair);
  if (hpair && hpair_owned) DestroyHimenoPair(hpair);
}

// ---- tile-shape / schedule tuner ----------------------------------------------------------
// The reference's auto-tuning compiles every CUDA_BLOCK_SIZE pattern into its own module and
// lets the first iterations of PSStencilRun try them in random order before settling on the
// fastest (translator/configuration.cc:27-57, reference_translator.cc:948,
// cuda_translator.cc:35, include/physis/runtime.h:32-52).  Here the patterns are the forms the
// hand-written kernels come in -- tile shapes, fused passes or single sweeps -- all of which
// compute bit-identical results, so the trials are likewise real iterations of the run: with
// option autotune=1 the first long run of a shape spends 6 iterations per form (2 to warm up,
// 4 timed with CUDA events), keeps the fastest as a set of option overrides for that shape, and
// every later run of the shape starts from there.  On a process group the ranks agree on the
// slowest rank's time per form.
struct TuneEntry {
  std::string best;     // option overrides, "" = the defaults won
  float best_ms = 0.0f, default_ms = 0.0f;  // per iteration
  int forms = 0;
};
std::map<std::string, TuneEntry> g_tuned;
std::string g_last_tuning;

constexpr int kTuneWarm = 2, kTuneTimed = 4;

// the shape of a run, without the identity of its grids
std::string TuneKey(int num_stencils, const __PSB200StencilDesc *descs) {
  std::string k;
  KeyAdd(&k, num_stencils);
  for (int s = 0; s < num_stencils; ++s) {
    const __PSB200StencilDesc &d = descs[s];
    KeyAdd(&k, d.kind);
    KeyAdd(&k, d.elm_type);
    KeyAdd(&k, d.dom);
    KeyAdd(&k, d.num_grids);
    for (int i = 0; i < d.num_grids; ++i) {
      const Grid *g = Grid::FromHandle(d.grids[i]);
      for (int a = 0; a < 3; ++a) KeyAdd(&k, g->dim[a]);
      KeyAdd(&k, d.members[i]);
    }
    KeyAdd(&k, d.num_scalars);
    for (int i = 0; i < d.num_scalars; ++i) KeyAdd(&k, d.scalars[i]);
  }
  return k;
}

std::vector<std::string> TuneForms(int num_stencils, const __PSB200StencilDesc *descs, int world) {
  std::vector<std::string> f;
  f.push_back("");  // the defaults
  bool all7 = true, allh = true, allp = true;
  for (int s = 0; s < num_stencils; ++s) {
    all7 = all7 && descs[s].kind == PSB200_KIND_DIFFUSION7_CLAMP;
    allh = allh && (descs[s].kind == PSB200_KIND_HIMENO19 || descs[s].kind == PSB200_KIND_HIMENO19_GOSA);
    allp = allp && descs[s].kind == PSB200_KIND_PERIODIC7_STAGGERED;
  }
  // z chunks of the fused passes: the planner's cost model (star7_pair.cu) against plain
  // divisions of the planes this rank owns
  auto chunk_forms = [&](const std::string &opt) {
    // (an even share of the planes, not this rank's own count: every rank must build the same list)
    const Grid *g = Grid::FromHandle(descs[0].grids[0]);
    const int nz = g->decomposed ? std::max(1, g->dim[2] / world) : g->dim[2];
    int last = 0;
    for (int d : {2, 4, 8, 16}) {
      const int zc = (nz + d - 1) / d;
      if (zc < 8 || zc == last) continue;
      f.push_back(opt + "=" + std::to_string(zc));
      last = zc;
    }
  };
  if (all7) {
    const std::string single = num_stencils == 2 ? "star7_fuse=0," : "";
    if (num_stencils == 2) {
      chunk_forms("star7_pair_zc");
      // rows cut into x tiles: the 16-row tile (fewer warps, more tiles) against the 20-row one
      const Grid *g = Grid::FromHandle(descs[0].grids[0]);
      if ((size_t)g->dim[0] * (g->type == PS_DOUBLE ? 8 : 4) > 2048) {
        f.push_back("star7_pair_variant=1");
        chunk_forms("star7_pair_variant=1,star7_pair_zc");
      }
      f.push_back("star7_fuse=0");
    }
    for (int v = 0; v < 6; ++v) f.push_back(single + "star7_variant=" + std::to_string(v));
  } else if (allh) {
    const std::string single = num_stencils == 2 ? "himeno_fuse=0," : "";
    if (num_stencils == 2) {
      chunk_forms("himeno_pair_zc");
      f.push_back("himeno_fuse=2");  // fused also where the planner would rather not
      f.push_back("himeno_pair_pfmode=1");  // coefficient prefetch by rows / by tensor-map boxes
      f.push_back("himeno_pair_pfmode=2");
    }
    f.push_back(single + "himeno_by=7");
    f.push_back(single + "himeno_by=15");
  } else if (allp) {
    f.push_back("pstag_variant=1");
  } else {
    f.clear();  // generated kernels: nothing to choose from
  }
  return f;
}

}  // namespace

namespace physis_b200 {
void ClearTuning() {
  g_tuned.clear();
  g_last_tuning.clear();
}
}

namespace {

// Runs the trials on the first iterations of this run; returns how many iterations they used.
int TuneOnRun(Runtime *rt, const std::string &key, int iter, int num_stencils,
              const __PSB200StencilDesc *descs) {
  const std::vector<std::string> forms = TuneForms(num_stencils, descs, rt->world());
  const int per_form = kTuneWarm + kTuneTimed;
  if (forms.size() < 2 || iter < (int)forms.size() * per_form + 1) return 0;
  const Options saved = rt->opt;
  struct Trial { std::string form; cudaEvent_t e0, e1; };
  std::vector<Trial> trials;
  int used = 0;
  for (const std::string &form : forms) {
    rt->opt = saved;
    ParseOptionList(&rt->opt, form, false);
    g_plan_tag = form;
    // a form runs only if every rank's hand-written kernel takes the shape under it
    int ok = 1;
    if (!form.empty()) {
      g_soft_fail = true;
      for (int s = 0; s < num_stencils && ok; ++s) {
        SweepPlan *p = GetSweepPlan(rt, descs[s]);
        ok = (p != nullptr && (p->empty || p->star7 || p->himeno || p->pstag)) ? 1 : 0;
        if (p) DestroySweep(p);
      }
      g_soft_fail = false;
    }
    if (rt->world() > 1) {
      std::vector<int> all(rt->world());
      rt->comm->AllGather(&ok, all.data(), sizeof(int));
      for (int v : all) ok = ok && v;
    }
    if (!ok) continue;
    Trial t{form, nullptr, nullptr};
    PSB_CUDA(cudaEventCreate(&t.e0));
    PSB_CUDA(cudaEventCreate(&t.e1));
    RunSchedule(rt, kTuneWarm, num_stencils, descs, true, nullptr);
    PSB_CUDA(cudaEventRecord(t.e0, rt->stream));
    RunSchedule(rt, kTuneTimed, num_stencils, descs, true, nullptr);
    PSB_CUDA(cudaEventRecord(t.e1, rt->stream));
    trials.push_back(t);
    used += per_form;
    rt->stats.autotune_trials++;
  }
  rt->opt = saved;
  g_plan_tag.clear();
  std::vector<float> ms(trials.size(), 0.0f);
  for (size_t i = 0; i < trials.size(); ++i) {
    PSB_CUDA(cudaEventSynchronize(trials[i].e1));
    PSB_CUDA(cudaEventElapsedTime(&ms[i], trials[i].e0, trials[i].e1));
    ms[i] /= kTuneTimed;
    PSB_CUDA(cudaEventDestroy(trials[i].e0));
    PSB_CUDA(cudaEventDestroy(trials[i].e1));
  }
  if (rt->world() > 1 && !ms.empty()) {
    // every rank ran the same forms; the slowest rank's time counts
    std::vector<float> all(ms.size() * rt->world());
    rt->comm->AllGather(ms.data(), all.data(), ms.size() * sizeof(float));
    for (size_t i = 0; i < ms.size(); ++i)
      for (int r = 0; r < rt->world(); ++r) ms[i] = std::max(ms[i], all[r * ms.size() + i]);
  }
  TuneEntry e;
  e.forms = (int)trials.size();
  size_t best = 0;
  for (size_t i = 1; i < trials.size(); ++i)
    if (ms[i] < ms[best] * 0.99f) best = i;  // the defaults keep ties
  e.best = trials[best].form;
  e.best_ms = ms[best];
  e.default_ms = ms[0];
  g_tuned[key] = e;
  char buf[256];
  snprintf(buf, sizeof buf, "%s%s: %.4f ms per iteration (defaults %.4f; %d forms tried)",
           e.best.empty() ? "defaults" : e.best.c_str(), "", e.best_ms, e.default_ms, e.forms);
  g_last_tuning = buf;
  // the trial plans of the forms that lost are of no further use
  ClearPlanCache();
  return used;
}

}  // namespace

extern "C" const char *__PSB200LastTuning(void) { return g_last_tuning.c_str(); }

extern "C" float __PSB200StencilRun(int iter, int num_stencils, const __PSB200StencilDesc *descs) {
  Runtime *rt = Runtime::Get();
  // every rank has finished its earlier synchronous runtime calls that wrote grids from the
  // host (copyin, PSGridSet, free): a neighbour's sweep must not deliver halo planes into a grid
  // that is still being filled.  Between runs with nothing of the kind in between, the sweeps
  // order themselves on the device, and the host stays out of it.
  if (rt->world() > 1 && rt->group_dirty) {
    rt->comm->Barrier();
    rt->group_dirty = false;
  }
  std::string names;
  for (int s = 0; s < num_stencils; ++s) {
    if (s) names += ", ";
    names += descs[s].name ? descs[s].name : KindName(descs[s].kind);
  }
  const bool trace = (__ps_trace != nullptr);
  SegmentTiming tm;
  tm.timed = trace || rt->opt.time_kernels;
  if (trace) __PSTraceStencilPre(names.c_str());
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (tm.timed) {
    PSB_CUDA(cudaEventCreate(&e0));
    PSB_CUDA(cudaEventCreate(&e1));
    PSB_CUDA(cudaEventRecord(e0, rt->stream));
  }
  int done = 0;
  const TuneEntry *tuned = nullptr;
  if (rt->opt.autotune) {
    const std::string key = TuneKey(num_stencils, descs);
    auto it = g_tuned.find(key);
    if (it == g_tuned.end()) {
      done = TuneOnRun(rt, key, iter, num_stencils, descs);
      it = g_tuned.find(key);
    }
    if (it != g_tuned.end()) tuned = &it->second;
  }
  if (tuned && !tuned->best.empty()) {
    const Options saved = rt->opt;
    ParseOptionList(&rt->opt, tuned->best, false);
    g_plan_tag = tuned->best;
    RunSchedule(rt, iter - done, num_stencils, descs, false, done ? nullptr : &tm);
    rt->opt = saved;
    g_plan_tag.clear();
    rt->stats.autotuned_runs++;
  } else {
    RunSchedule(rt, iter - done, num_stencils, descs, false, done ? nullptr : &tm);
  }
  PSB_CUDA(cudaGetLastError());
  float ms = 0.0f;
  if (tm.timed) {
    PSB_CUDA(cudaEventRecord(e1, rt->stream));
    PSB_CUDA(cudaEventSynchronize(e1));
    PSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (tm.emid) {
      float pms = 0.0f;
      PSB_CUDA(cudaEventElapsedTime(&pms, e0, tm.emid));
      rt->stats.fused_pair_ms += pms;
      rt->stats.fused_pairs_timed += (uint64_t)tm.fused;
      PSB_CUDA(cudaEventDestroy(tm.emid));
    }
    PSB_CUDA(cudaEventDestroy(e0));
    PSB_CUDA(cudaEventDestroy(e1));
    rt->timed_ms += ms;
    rt->timed_launches += (uint64_t)iter * num_stencils;
  }
  if (trace) __PSTraceStencilPost(ms);
  return trace ? ms : 0.0f;
}
