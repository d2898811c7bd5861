// ORACLE (test infrastructure only).  Thin C entry points over the
// REFERENCE's OWN benchmark classes, compiled together with the unmodified
// /root/reference/examples/diffusion-benchmark/{baseline,diffusion3d}.cc into
// oracle/_ref/ (see oracle/Makefile).  Nothing here restates an algorithm; it
// only exposes the reference's protected members to ctypes.
#include "baseline.h"
#include "diffusion3d_openmp.h"
#include <omp.h>
#include <string.h>

namespace {
class BaselineProbe : public diffusion3d::Baseline {
 public:
  BaselineProbe(int nx, int ny, int nz) : diffusion3d::Baseline(nx, ny, nz) {}
  void Params(float *out) const {
    out[0] = ce_; out[1] = cw_; out[2] = cn_; out[3] = cs_; out[4] = ct_; out[5] = cb_;
    out[6] = cc_; out[7] = dx_; out[8] = dy_; out[9] = dz_; out[10] = dt_; out[11] = kappa_;
    out[12] = kx_; out[13] = ky_; out[14] = kz_;
  }
  // InitializeBenchmark + RunKernel(count) + copy the field out.
  void Run(int count, float *out, float *accuracy) {
    InitializeBenchmark();
    RunKernel(count);
    memcpy(out, f1_, sizeof(float) * (size_t)nx_ * ny_ * nz_);
    if (accuracy) *accuracy = GetAccuracy(count);
    FinalizeBenchmark();
  }
  // RunKernel on a caller-provided field (for timing and arbitrary inputs).
  void RunOn(int count, float *field) {
    InitializeBenchmark();
    memcpy(f1_, field, sizeof(float) * (size_t)nx_ * ny_ * nz_);
    RunKernel(count);
    memcpy(field, f1_, sizeof(float) * (size_t)nx_ * ny_ * nz_);
    FinalizeBenchmark();
  }
  void Analytic(int count, float *out) const {
    float *f = GetCorrectAnswer(count);
    memcpy(out, f, sizeof(float) * (size_t)nx_ * ny_ * nz_);
    free(f);
  }
};
// The reference's OpenMP form of the sweep (diffusion3d_openmp.cc), on a caller-provided field.
class OpenMPProbe : public diffusion3d::Diffusion3DOpenMP {
 public:
  OpenMPProbe(int nx, int ny, int nz) : diffusion3d::Diffusion3DOpenMP(nx, ny, nz) {}
  void Load(const float *field) {
    InitializeBenchmark();
    memcpy(f1_, field, sizeof(float) * (size_t)nx_ * ny_ * nz_);
  }
  void Sweeps(int count) { RunKernel(count); }
  void Store(float *field) {
    memcpy(field, f1_, sizeof(float) * (size_t)nx_ * ny_ * nz_);
    FinalizeBenchmark();
  }
};
OpenMPProbe *g_omp = nullptr;
}  // namespace

extern "C" {
int ref_openmp_threads(void) { return omp_get_max_threads(); }
void ref_openmp_load(int nx, int ny, int nz, const float *field) {
  g_omp = new OpenMPProbe(nx, ny, nz);
  g_omp->Load(field);
}
void ref_openmp_sweeps(int count) { g_omp->Sweeps(count); }
void ref_openmp_store(float *field) {
  g_omp->Store(field);
  delete g_omp;
  g_omp = nullptr;
}
void ref_diffusion3d_params(int nx, int ny, int nz, float *out) {
  BaselineProbe b(nx, ny, nz);
  b.Params(out);
}
void ref_diffusion3d_initialize(float *buff, int nx, int ny, int nz, float kx, float ky,
                                float kz, float dx, float dy, float dz, float kappa,
                                float time) {
  diffusion3d::Initialize(buff, nx, ny, nz, kx, ky, kz, dx, dy, dz, kappa, time);
}
void ref_baseline_run(int nx, int ny, int nz, int count, float *out, float *accuracy) {
  BaselineProbe b(nx, ny, nz);
  b.Run(count, out, accuracy);
}
void ref_baseline_run_on(int nx, int ny, int nz, int count, float *field) {
  BaselineProbe b(nx, ny, nz);
  b.RunOn(count, field);
}
void ref_diffusion3d_analytic(int nx, int ny, int nz, int count, float *out) {
  BaselineProbe b(nx, ny, nz);
  b.Analytic(count, out);
}
}
