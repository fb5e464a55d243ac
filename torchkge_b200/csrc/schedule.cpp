// Host-side construction of reduction schedules (see schedule.h).
//
// What is being reproduced (observed on torch 2.11.0 CPU, identical under the AVX2 and
// AVX512 dispatch levels; tests/test_schedule.py re-checks it against the installed ATen):
//
//  * Tensor.norm(p=1, dim=-1)   -- utils/dissimilarities.py:16 -- one fp32 accumulator,
//    terms added in index order.
//  * Tensor.norm(p=2, dim=-1)   -- utils/dissimilarities.py:25 -- eight vector lanes;
//    lane l accumulates x[l], x[l+8], ... as acc = acc + x*x (mul and add rounded
//    separately); lanes are then added in order 0..7; the dim % 8 trailing terms follow,
//    in groups of four as tot = tot + x*x and a last (< 4) remainder as fma(x, x, tot).
//  * Tensor.sum(dim=-1)         -- bilinear.py:114,235,240,514-515,521-522 -- cascade_sum:
//    eight lanes times four interleaved "ilp" rows; each (lane, row) chain adds its terms
//    in order, spilling into a second-level accumulator every 16 terms; the vectors that
//    do not fill a group of four are appended to row 0; rows are folded 0+1+2+3 per lane;
//    the scalar tail (dim % 8 terms) is summed first into the result, then lanes 0..7 are
//    added in order.  For dim < 8 the same scheme runs with one lane.
#include "schedule.h"

#include "../../include/kge_b200.h"

namespace kge {

int reduce_kind_for_model(int model) {
  switch (model) {
    case KGE_TRANSE_L1: return RED_SEQ;
    case KGE_TRANSE_L2: return RED_NORM2;
    case KGE_DISTMULT:
    case KGE_RESCAL:
    case KGE_COMPLEX:
    case KGE_ROTATE:
    case KGE_TORUSE_L1:
    case KGE_TORUSE_L2:
    case KGE_ANALOGY: return RED_SUM;
    default: return -1;
  }
}

namespace {

struct Builder {
  Schedule* s;
  void push(int k, uint8_t mode) {
    s->perm.push_back(k);
    s->code.push_back(mode);
  }
  void mark(uint8_t bits) { s->code.back() |= bits; }
};

void build_seq(int dim, Builder& b) {
  for (int k = 0; k < dim; ++k) b.push(k, SC_MODE_T);
}

void build_norm2(int dim, Builder& b) {
  const int V = 8;
  const int main_len = dim - dim % V;
  if (main_len > 0) {
    for (int l = 0; l < V; ++l) {
      for (int k = l; k < main_len; k += V) b.push(k, SC_MODE_A);
      b.mark(SC_T_ADD_A);
    }
  }
  int k = main_len;
  for (; k + 4 <= dim; k += 4)
    for (int j = 0; j < 4; ++j) b.push(k + j, SC_MODE_T);
  for (; k < dim; ++k) b.push(k, SC_MODE_T_FMA);
}

bool build_sum(int dim, Builder& b) {
  const int V = dim >= 8 ? 8 : 1;
  const int ILP = 4;
  const int LEVEL_STEP = 16;  // level_power = max(4, ceil_log2(rows) / 4) = 4 for rows < 2^20
  const int vec_size = dim / V;
  const int rows = vec_size / ILP;  // "size_ilp"
  if (rows >= LEVEL_STEP * LEVEL_STEP) return false;  // second cascade level not modelled
  const bool casc = rows >= LEVEL_STEP;
  b.s->has_cascade = casc;
  // scalar tail first (vectorized_inner_sum: final_acc accumulates it before the lanes)
  for (int k = vec_size * V; k < dim; ++k) b.push(k, SC_MODE_T);
  for (int l = 0; l < V; ++l) {
    bool lane_has_p = false;
    for (int m = 0; m < ILP; ++m) {
      bool chain_has_elems = false;
      for (int i = 0; i < rows; ++i) {
        b.push((i * ILP + m) * V + l, SC_MODE_A);
        chain_has_elems = true;
        if ((i + 1) % LEVEL_STEP == 0) b.mark(SC_CASC1);
      }
      if (chain_has_elems && casc) b.mark(SC_FOLD1);
      if (m == 0) {
        for (int j = rows * ILP; j < vec_size; ++j) {
          b.push(j * V + l, SC_MODE_A);
          chain_has_elems = true;
        }
      }
      if (chain_has_elems) {
        b.mark(lane_has_p ? SC_P_ADD : SC_P_SET);
        lane_has_p = true;
      }
    }
    if (lane_has_p) b.mark(SC_T_ADD_P);
  }
  return true;
}

}  // namespace

int schedule_depth(const Schedule& s) {
  // replay the micro-ops with depths instead of values; -1 = empty accumulator (exact zero)
  auto add = [](int x, int y) { return x < 0 ? y : (y < 0 ? x : (x > y ? x : y) + 1); };
  int a = -1, a1 = -1, p = -1, t = -1;
  for (uint8_t code : s.code) {
    const uint8_t mode = code & SC_MODE_MASK;
    if (mode == SC_MODE_A) a = add(a, 0); else t = add(t, 0);
    if (code & SC_CASC1) { a1 = add(a1, a); a = -1; }
    if (code & SC_FOLD1) { a = add(a, a1); a1 = -1; }
    if (code & SC_P_SET) { p = a; a = -1; }
    if (code & SC_P_ADD) { p = add(p, a); a = -1; }
    if (code & SC_T_ADD_P) { t = add(t, p); }
    if (code & SC_T_ADD_A) { t = add(t, a); a = -1; }
  }
  return t < 0 ? 0 : t;
}

bool build_schedule(int kind, int dim, Schedule* out) {
  if (dim < 1 || dim > 8191) return false;
  out->kind = kind;
  out->dim = dim;
  out->has_cascade = false;
  out->perm.clear();
  out->code.clear();
  out->perm.reserve(dim);
  out->code.reserve(dim);
  Builder b{out};
  switch (kind) {
    case RED_SEQ: build_seq(dim, b); break;
    case RED_NORM2: build_norm2(dim, b); break;
    case RED_SUM:
      if (!build_sum(dim, b)) return false;
      break;
    default: return false;
  }
  return (int)out->perm.size() == dim;
}

}  // namespace kge
