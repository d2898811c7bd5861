// __PSB200StencilRun: the b200 replacement for the host run function the
// translator generates per PSStencilRun call
// (translator/cuda_runtime_builder.cc:1465-1583: build launch dims once, then
// `for (i < iter) { kernel_0<<<>>>; kernel_1<<<>>>; ... }` on one stream with
// no synchronisation; TRACE_KERNEL wrapper from
// translator/reference_runtime_builder.cc:896-940).
#include "runtime.h"

#include <string>
#include <vector>

namespace physis_b200 {

struct Star7Plan;
Star7Plan *PrepareStar7(Runtime *rt, const __PSB200StencilDesc &d, std::string *why);
void LaunchStar7(Runtime *rt, Star7Plan *p);
void DestroyStar7(Star7Plan *p);

struct HimenoPlan;
HimenoPlan *PrepareHimeno(Runtime *rt, const __PSB200StencilDesc &d, std::string *why);
void LaunchHimeno(Runtime *rt, HimenoPlan *p);
void DestroyHimeno(HimenoPlan *p);

struct PstagPlan;
PstagPlan *PreparePstag(Runtime *rt, const __PSB200StencilDesc &d, std::string *why);
void LaunchPstag(Runtime *rt, PstagPlan *p);
void DestroyPstag(PstagPlan *p);

struct SweepPlan {
  int kind = 0;
  std::string name;
  Star7Plan *star7 = nullptr;
  HimenoPlan *himeno = nullptr;
  PstagPlan *pstag = nullptr;
  __PSB200LaunchFunc launch = nullptr;
  const void *stencil = nullptr;
};

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

SweepPlan *PrepareSweep(Runtime *rt, const __PSB200StencilDesc &d) {
  SweepPlan *p = new SweepPlan();
  p->kind = d.kind;
  p->name = d.name ? d.name : KindName(d.kind);
  std::string why;
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
    case PSB200_KIND_GENERIC:
      break;
    default:
      why = "unknown sweep kind";
  }
  // Shapes a specialised kernel does not cover run through the program's own
  // generic per-point kernel (still on the GPU); without one this is fatal.
  if (d.launch) {
    p->kind = PSB200_KIND_GENERIC;
    p->launch = d.launch;
    p->stencil = d.stencil;
    return p;
  }
  fprintf(stderr, "[physis-b200] sweep '%s' (%s) cannot run: %s, and the program carries no "
                  "generic launch stub. There is no CPU fallback.\n",
          p->name.c_str(), KindName(d.kind), why.c_str());
  exit(1);
}

void LaunchSweep(Runtime *rt, SweepPlan *p) {
  if (p->star7) LaunchStar7(rt, p->star7);
  else if (p->himeno) LaunchHimeno(rt, p->himeno);
  else if (p->pstag) LaunchPstag(rt, p->pstag);
  else p->launch(p->stencil, (__PSB200Stream)rt->stream);
  rt->stats.kernel_launches++;
}

void DestroySweep(SweepPlan *p) {
  if (p->star7) DestroyStar7(p->star7);
  if (p->himeno) DestroyHimeno(p->himeno);
  if (p->pstag) DestroyPstag(p->pstag);
  delete p;
}

const char *SweepName(const SweepPlan *p) { return p->name.c_str(); }

}  // namespace physis_b200

using namespace physis_b200;

extern "C" float __PSB200StencilRun(int iter, int num_stencils, const __PSB200StencilDesc *descs) {
  Runtime *rt = Runtime::Get();
  std::vector<SweepPlan *> plans;
  plans.reserve(num_stencils);
  std::string names;
  for (int s = 0; s < num_stencils; ++s) {
    plans.push_back(PrepareSweep(rt, descs[s]));
    if (s) names += ", ";
    names += SweepName(plans.back());
  }
  const bool trace = (__ps_trace != nullptr);
  const bool timed = trace || rt->opt.time_kernels;
  if (trace) __PSTraceStencilPre(names.c_str());
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) {
    PSB_CUDA(cudaEventCreate(&e0));
    PSB_CUDA(cudaEventCreate(&e1));
    PSB_CUDA(cudaEventRecord(e0, rt->stream));
  }
  for (int i = 0; i < iter; ++i)
    for (int s = 0; s < num_stencils; ++s) LaunchSweep(rt, plans[s]);
  PSB_CUDA(cudaGetLastError());
  float ms = 0.0f;
  if (timed) {
    PSB_CUDA(cudaEventRecord(e1, rt->stream));
    PSB_CUDA(cudaEventSynchronize(e1));
    PSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    PSB_CUDA(cudaEventDestroy(e0));
    PSB_CUDA(cudaEventDestroy(e1));
    rt->timed_ms += ms;
    rt->timed_launches += (uint64_t)iter * num_stencils;
  }
  if (trace) __PSTraceStencilPost(ms);
  for (auto *p : plans) DestroySweep(p);
  return trace ? ms : 0.0f;
}
