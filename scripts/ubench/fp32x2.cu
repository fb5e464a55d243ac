// Microbenchmark: fp32 scalar vs packed f32x2 issue rates on sm_100a (FADD/FMUL/FFMA).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float x, float y) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) { a[i] = __fmaf_rn(a[i], x, y); a[i + 1] = __fmaf_rn(a[i + 1], x, y); }
      if (MODE == 1) { a[i] = __fadd_rn(a[i], x); a[i + 1] = __fadd_rn(a[i + 1], x); }
      if (MODE == 2) { a[i] = __fmul_rn(a[i], x); a[i + 1] = __fmul_rn(a[i + 1], x); }
      if (MODE == 3) {  // fma.rn.f32x2
        unsigned long long v, xx, yy;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(y));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(xx), "l"(yy));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
      }
      if (MODE == 4) {  // add.rn.f32x2
        unsigned long long v, xx;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(v), "l"(xx));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
      }
      if (MODE == 5) {  // mul.rn.f32x2
        unsigned long long v, xx;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(v), "l"(xx));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
      }
      if (MODE == 6) {  // mix: FMUL + FADD alternating (scalar)
        a[i] = __fmul_rn(a[i], x); a[i + 1] = __fadd_rn(a[i + 1], y);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  int blocks = 148 * 4, threads = warps_per_sm * 32 / 4;
  float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double elems = (double)blocks * threads * ITERS * 16;
  printf("%-22s warps/SM=%2d  %.2f ms  %.2f Tera-elem-ops/s\n", name, warps_per_sm, ms, elems / ms / 1e9);
  cudaFree(out);
}
int main() {
  for (int w : {8, 16, 32, 64}) {
    run<0>("FFMA scalar", w); run<1>("FADD scalar", w); run<2>("FMUL scalar", w);
    run<3>("FFMA2 packed", w); run<4>("FADD2 packed", w); run<5>("FMUL2 packed", w);
    run<6>("FMUL+FADD mix", w);
  }
  return 0;
}
