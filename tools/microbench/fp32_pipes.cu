// Throughput of the fp32 forms the sweep kernels use, per SM per clock (B200 microbenchmark,
// tuning tool only).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define CHAINS 8
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(512) k(float *out, float s0, float s1, u64 sp) {
  float r[CHAINS];
  u64 p[CHAINS / 2];
  for (int i = 0; i < CHAINS; ++i) r[i] = threadIdx.x * 1e-3f + i;
  for (int i = 0; i < CHAINS / 2; ++i) p[i] = ((u64)__float_as_uint(r[2 * i]) << 32) | __float_as_uint(r[2 * i + 1]);
  float v = s1 + threadIdx.x;  // per-thread (vector register) operand
  u64 vp = sp + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) {
        if (OP == 0) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(v));          // reg x reg
        if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(s0));         // reg x uniform
        if (OP == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(v));
        if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(v), "f"(s0));
        if (OP == 4 && i < CHAINS / 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(vp));
        if (OP == 5 && i < CHAINS / 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(vp));
        if (OP == 6 && i < CHAINS / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(vp), "l"(sp));
        if (OP == 7) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(r[i]));           // immediate
        if (OP == 8) {  // the sweep's mix: 7 scalar multiplies (uniform coefficient) + 3 packed adds
          if (i < 7) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(s0));
          if (i < 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(vp));
        }
        if (OP == 9) {  // 7 scalar multiplies + 6 scalar adds
          if (i < 7) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(s0));
          if (i < 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(v));
        }
      }
    }
  }
  float acc = 0;
  for (int i = 0; i < CHAINS; ++i) acc += r[i];
  for (int i = 0; i < CHAINS / 2; ++i) acc += __uint_as_float((unsigned)p[i]);
  if (acc == 123.456f) out[0] = acc;
}

template <int OP>
void run(const char *name, double insts_per_inner, int sms, float clk_ghz, float *d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 2, threads = 512;
  k<OP><<<blocks, threads>>>(d, 1.0000001f, 0.5f, 0x3f8000003f800000ull);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(d, 1.0000001f, 0.5f, 0x3f8000003f800000ull);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_insts = (double)blocks * (threads / 32) * ITERS * 4 * insts_per_inner;
  const double cycles = ms * 1e-3 * clk_ghz * 1e9;
  printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM  (%.2f per SMSP)\n", name, ms,
         warp_insts / cycles / sms, warp_insts / cycles / sms / 4);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const float ghz = clk_khz * 1e-6f;
  printf("%s, %d SMs, nominal %.3f GHz (rates assume this clock)\n", p.name, p.multiProcessorCount, ghz);
  float *d;
  cudaMalloc(&d, 4);
  run<0>("fmul reg*reg", 8, p.multiProcessorCount, ghz, d);
  run<1>("fmul reg*uniform", 8, p.multiProcessorCount, ghz, d);
  run<7>("fmul reg*imm", 8, p.multiProcessorCount, ghz, d);
  run<2>("fadd reg+reg", 8, p.multiProcessorCount, ghz, d);
  run<3>("ffma", 8, p.multiProcessorCount, ghz, d);
  run<4>("fadd2 (f32x2)", 4, p.multiProcessorCount, ghz, d);
  run<5>("fmul2 (f32x2)", 4, p.multiProcessorCount, ghz, d);
  run<6>("ffma2 (f32x2)", 4, p.multiProcessorCount, ghz, d);
  run<8>("7 fmul(u) + 3 fadd2", 10, p.multiProcessorCount, ghz, d);
  run<9>("7 fmul(u) + 6 fadd", 13, p.multiProcessorCount, ghz, d);
  return 0;
}
