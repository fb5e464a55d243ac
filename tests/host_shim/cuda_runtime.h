// Host stand-in for <cuda_runtime.h>: lets tests/host_arith.cpp compile the device-side score
// arithmetic of torchkge_b200/csrc/reduce.cuh with g++ (no GPU, no nvcc).  Every rounded
// intrinsic maps to the IEEE operation it names; compile with -ffp-contract=off.
#pragma once
#include <math.h>
#include <stdint.h>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
// warp shuffles appear only in the chain-parallel scorers, which the host harness does not call
static inline float __shfl_sync(unsigned, float v, int) { return v; }
