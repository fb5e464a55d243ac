// Host build of the DEVICE score arithmetic (torchkge_b200/csrc/reduce.cuh) -- test infrastructure.
// tests/test_host_arith.py compiles this with g++ and compares, bit for bit, with ATen on the CPU:
//   replay : the schedule replay every sparse pass uses (acc_step per position, acc_finish)
//   natural: pair_score_natural (natural index order; L2-norm and cascade-sum kinds)
// so the element kinds (including ones not yet run on a GPU) are checked against the reference's
// arithmetic without a GPU.  What this cannot check: the tiling / pipelines of the CUDA kernels.
#include <stdint.h>

#include "../torchkge_b200/csrc/reduce.cuh"

using namespace kge;

namespace {
template <int EL, bool CASC>
float replay(int dim, const float* q0, const float* q1, const float* c0, const float* c1,
             const int32_t* perm, const uint8_t* code, const float* qm, const float* cm) {
  Acc r;
  acc_reset(r);
  for (int pos = 0; pos < dim; ++pos) {
    const int k = perm[pos];
    acc_step<EL, CASC>(r, code[pos], q0[k], q1[k], c0[k], c1[k], qm[k], cm[k]);
  }
  return acc_finish<EL>(r);
}

template <int EL>
void run(int mode, int casc, int dim, int nq, int nc, const float* q, const float* c, const int32_t* perm,
         const uint8_t* code, float* out) {
  constexpr int QW = ElemTraits<EL>::QW, CW = ElemTraits<EL>::CW;
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < nc; ++j) {
      const float* q0 = q + (size_t)i * QW * dim;
      const float* q1 = q0 + (size_t)(QW - 1) * dim;
      const float* c0 = c + (size_t)j * CW * dim;
      const float* c1 = c0 + (size_t)(CW - 1) * dim;
      const float* qm = q0 + (size_t)(QW / 2) * dim;   // middle plane (three-plane kinds)
      const float* cm = c0 + (size_t)(CW / 2) * dim;
      float s;
      if (mode == 0) s = casc ? replay<EL, true>(dim, q0, q1, c0, c1, perm, code, qm, cm)
                              : replay<EL, false>(dim, q0, q1, c0, c1, perm, code, qm, cm);
      else {
        if constexpr (ElemTraits<EL>::RED == RED_SEQ) s = 0.f / 0.f;  // no natural form
        else s = pair_score_natural<EL>(dim, q0, q1, c0, c1, qm, cm);
      }
      out[(size_t)i * nc + j] = s;
    }
}
}  // namespace

// q: [nq][QW][dim], c: [nc][CW][dim] (planes inside a row), out: [nq][nc]
extern "C" int host_scores(int el, int mode, int casc, int dim, int nq, int nc, const float* q, const float* c,
                           const int32_t* perm, const uint8_t* code, float* out) {
  switch (el) {
    case EL_DOT1: run<EL_DOT1>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_DOT2: run<EL_DOT2>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_L1_TAIL: run<EL_L1_TAIL>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_L1_HEAD: run<EL_L1_HEAD>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_L2_TAIL: run<EL_L2_TAIL>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_L2_HEAD: run<EL_L2_HEAD>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_ROT: run<EL_ROT>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_DOT_MID: run<EL_DOT_MID>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_TL1_TAIL: run<EL_TL1_TAIL>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_TL1_HEAD: run<EL_TL1_HEAD>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_TL2_TAIL: run<EL_TL2_TAIL>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_TL2_HEAD: run<EL_TL2_HEAD>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    case EL_DOT3: run<EL_DOT3>(mode, casc, dim, nq, nc, q, c, perm, code, out); break;
    default: return 1;
  }
  return 0;
}

// RESCAL query preparation (reduce.cuh: rescal_query_component): out[i][j] for n facts;
// vec: [n][dim], mats: [n][dim][dim]
extern "C" int host_rescal_prep(int tail, int dim, int n, const float* vec, const float* mats, float* out) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < dim; ++j)
      out[(size_t)i * dim + j] =
          rescal_query_component(tail != 0, dim, j, vec + (size_t)i * dim, mats + (size_t)i * dim * dim);
  return 0;
}
