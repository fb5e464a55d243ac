// Device-side score arithmetic shared by every kernel (dense scan, sparse pair scorer).
// One (query, candidate) score = "element term" per embedding index, reduced by replaying
// a Schedule (schedule.h).  Every operation is an explicitly rounded intrinsic so the
// compiler can neither contract mul+add into fma nor reassociate: bit-equality between the
// dense scan, the true-score pass and the filter pass -- and with ATen's CPU kernels --
// rests on that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "schedule.h"

namespace kge {

// Element kinds: how one term is formed from the query planes q[] and candidate planes c[].
enum ElemKind : int {
  EL_DOT1 = 0,     // q0*c0                      DistMult bilinear.py:235,240 ; RESCAL :109,114
  EL_DOT2 = 1,     // q0*c0 + q1*c1              ComplEx  bilinear.py:514-515, 521-522
  EL_L1_TAIL = 2,  // |q0 - c0|                  interfaces.py:253-254 + dissimilarities.py:16
  EL_L1_HEAD = 3,  // |(c0 + q0) - q1|           interfaces.py:258-260 + dissimilarities.py:16
  EL_L2_TAIL = 4,  // (q0 - c0)^2                interfaces.py:253-254 + dissimilarities.py:25
  EL_L2_HEAD = 5,  // ((c0 + q0) - q1)^2         interfaces.py:258-260 + dissimilarities.py:25
  EL_ROT = 6,      // sqrt((q0-c0)^2+(q1-c1)^2)  oracle RotatE restatement
  EL_COUNT = 7
};

template <int EL> struct ElemTraits;
template <> struct ElemTraits<EL_DOT1> { static constexpr int QW = 1, CW = 1, RED = RED_SUM; };
template <> struct ElemTraits<EL_DOT2> { static constexpr int QW = 2, CW = 2, RED = RED_SUM; };
template <> struct ElemTraits<EL_L1_TAIL> { static constexpr int QW = 1, CW = 1, RED = RED_SEQ; };
template <> struct ElemTraits<EL_L1_HEAD> { static constexpr int QW = 2, CW = 1, RED = RED_SEQ; };
template <> struct ElemTraits<EL_L2_TAIL> { static constexpr int QW = 1, CW = 1, RED = RED_NORM2; };
template <> struct ElemTraits<EL_L2_HEAD> { static constexpr int QW = 2, CW = 1, RED = RED_NORM2; };
template <> struct ElemTraits<EL_ROT> { static constexpr int QW = 2, CW = 2, RED = RED_SUM; };

// For the L2 kinds returns the difference x (the caller squares it, fused or not); for all
// other kinds returns the finished term.
template <int EL>
__device__ __forceinline__ float elem_value(float q0, float q1, float c0, float c1) {
  if constexpr (EL == EL_DOT1) {
    return __fmul_rn(q0, c0);
  } else if constexpr (EL == EL_DOT2) {
    return __fadd_rn(__fmul_rn(q0, c0), __fmul_rn(q1, c1));
  } else if constexpr (EL == EL_L1_TAIL) {
    return fabsf(__fsub_rn(q0, c0));
  } else if constexpr (EL == EL_L1_HEAD) {
    return fabsf(__fsub_rn(__fadd_rn(c0, q0), q1));
  } else if constexpr (EL == EL_L2_TAIL) {
    return __fsub_rn(q0, c0);
  } else if constexpr (EL == EL_L2_HEAD) {
    return __fsub_rn(__fadd_rn(c0, q0), q1);
  } else {  // EL_ROT
    const float dr = __fsub_rn(q0, c0);
    const float di = __fsub_rn(q1, c1);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di)));
  }
}

template <int EL>
__device__ __forceinline__ constexpr bool elem_is_l2() {
  return EL == EL_L2_TAIL || EL == EL_L2_HEAD;
}

// The code byte that needs no combine step and is by far the most frequent one.
template <int EL>
__device__ __forceinline__ constexpr uint8_t fast_code() {
  return ElemTraits<EL>::RED == RED_SEQ ? SC_MODE_T : SC_MODE_A;
}

// Registers of one running (query, candidate) reduction.
struct Acc {
  float a, a1, p, t;
};

__device__ __forceinline__ void acc_reset(Acc& r) { r.a = r.a1 = r.p = r.t = 0.f; }

// Fast path: position whose code is fast_code<EL>().
template <int EL>
__device__ __forceinline__ void acc_step_fast(Acc& r, float q0, float q1, float c0, float c1) {
  float v = elem_value<EL>(q0, q1, c0, c1);
  if constexpr (elem_is_l2<EL>()) v = __fmul_rn(v, v);
  if constexpr (ElemTraits<EL>::RED == RED_SEQ)
    r.t = __fadd_rn(r.t, v);
  else
    r.a = __fadd_rn(r.a, v);
}

// General path for ONE pair (sparse pair scorer).  `code` is uniform across the warp there too.
template <int EL, bool CASC>
__device__ __forceinline__ void acc_step(Acc& r, uint8_t code, float q0, float q1, float c0,
                                         float c1) {
  const float v = elem_value<EL>(q0, q1, c0, c1);
  const uint8_t mode = code & SC_MODE_MASK;
  if constexpr (elem_is_l2<EL>()) {
    if (mode == SC_MODE_A)
      r.a = __fadd_rn(r.a, __fmul_rn(v, v));
    else if (mode == SC_MODE_T)
      r.t = __fadd_rn(r.t, __fmul_rn(v, v));
    else
      r.t = __fmaf_rn(v, v, r.t);
  } else {
    if (mode == SC_MODE_A)
      r.a = __fadd_rn(r.a, v);
    else
      r.t = __fadd_rn(r.t, v);
  }
  if constexpr (CASC) {
    if (code & SC_CASC1) { r.a1 = __fadd_rn(r.a1, r.a); r.a = 0.f; }
    if (code & SC_FOLD1) { r.a = __fadd_rn(r.a, r.a1); r.a1 = 0.f; }
  }
  if (code & SC_P_SET) { r.p = r.a; r.a = 0.f; }
  if (code & SC_P_ADD) { r.p = __fadd_rn(r.p, r.a); r.a = 0.f; }
  if (code & SC_T_ADD_P) { r.t = __fadd_rn(r.t, r.p); }
  if (code & SC_T_ADD_A) { r.t = __fadd_rn(r.t, r.a); r.a = 0.f; }
}

// The same step split for register tiles: the element part for one pair under a given
// (uniform) mode, and the combine part for one pair under a given (uniform) op.  The dense
// scan hoists the uniform tests out of its pair loops, so no per-pair selects are generated.
template <int EL>
__device__ __forceinline__ void acc_elem_mode(Acc& r, uint8_t mode, float q0, float q1, float c0,
                                              float c1) {
  const float v = elem_value<EL>(q0, q1, c0, c1);
  if constexpr (elem_is_l2<EL>()) {
    if (mode == SC_MODE_A)
      r.a = __fadd_rn(r.a, __fmul_rn(v, v));
    else if (mode == SC_MODE_T)
      r.t = __fadd_rn(r.t, __fmul_rn(v, v));
    else
      r.t = __fmaf_rn(v, v, r.t);
  } else {
    if (mode == SC_MODE_A)
      r.a = __fadd_rn(r.a, v);
    else
      r.t = __fadd_rn(r.t, v);
  }
}

__device__ __forceinline__ void acc_casc1(Acc& r) { r.a1 = __fadd_rn(r.a1, r.a); r.a = 0.f; }
__device__ __forceinline__ void acc_fold1(Acc& r) { r.a = __fadd_rn(r.a, r.a1); r.a1 = 0.f; }
__device__ __forceinline__ void acc_p_set(Acc& r) { r.p = r.a; r.a = 0.f; }
__device__ __forceinline__ void acc_p_add(Acc& r) { r.p = __fadd_rn(r.p, r.a); r.a = 0.f; }
__device__ __forceinline__ void acc_t_add_p(Acc& r) { r.t = __fadd_rn(r.t, r.p); }
__device__ __forceinline__ void acc_t_add_a(Acc& r) { r.t = __fadd_rn(r.t, r.a); r.a = 0.f; }

// Reduction result -> score, with the reference's sign and the L2 sqrt-then-square
// (dissimilarities.py:25 computes norm(p=2)**2; interfaces.py:254,260 negate).
template <int EL>
__device__ __forceinline__ float acc_finish(const Acc& r) {
  if constexpr (EL == EL_DOT1 || EL == EL_DOT2) {
    return r.t;
  } else if constexpr (elem_is_l2<EL>()) {
    const float n = __fsqrt_rn(r.t);
    return -__fmul_rn(n, n);
  } else {
    return -r.t;
  }
}

}  // namespace kge
