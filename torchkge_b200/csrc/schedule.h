// Reduction schedules: the exact order in which ATen's CPU kernels (torch 2.11, the
// arithmetic behind torchkge's link-prediction path) add up the per-dimension terms
// of one score.  A schedule is a permutation of the embedding index plus, for every
// schedule position, one byte of micro-ops.  Host builder in schedule.cpp, device
// interpreter in reduce.cuh; tests replay it in numpy against ATen (tests/test_schedule.py).
#pragma once
#include <cstdint>
#include <vector>

namespace kge {

// Bits of the per-position code byte.  Bits 0-1 say where the element of THIS position is
// accumulated, bits 2-7 are combine steps executed (in bit order) AFTER the element.
enum : uint8_t {
  SC_MODE_MASK = 0x03,
  SC_MODE_A = 0x00,      // a = a + e                      (separately rounded)
  SC_MODE_T = 0x01,      // t = t + e
  SC_MODE_T_FMA = 0x02,  // t = fma(x, x, t)               (L2 norm remainder only)
  SC_CASC1 = 0x04,       // a1 = a1 + a ; a = 0            (cascade level 0 -> 1)
  SC_FOLD1 = 0x08,       // a  = a + a1 ; a1 = 0           (end of chain: fold level 1 back)
  SC_P_SET = 0x10,       // p = a ; a = 0
  SC_P_ADD = 0x20,       // p = p + a ; a = 0
  SC_T_ADD_P = 0x40,     // t = t + p
  SC_T_ADD_A = 0x80,     // t = t + a ; a = 0
};

enum ReduceKind : int {
  RED_SEQ = 0,    // norm(p=1): one sequential fp32 chain      (ReduceOpsKernel NormOneOps)
  RED_NORM2 = 1,  // norm(p=2) last-dim fast path: 8 lanes, groups-of-4 tail, fma remainder
  RED_SUM = 2,    // sum(dim=-1): cascade_sum / vectorized_inner_sum, 8 lanes x 4 ilp rows
};

struct Schedule {
  int kind = 0;
  int dim = 0;
  bool has_cascade = false;  // any SC_CASC1 present (dim >= 512 for RED_SUM)
  std::vector<int32_t> perm; // schedule position -> embedding index
  std::vector<uint8_t> code; // micro-ops per position
};

int reduce_kind_for_model(int model);
// Depth of the reduction tree the schedule encodes: the largest number of ROUNDED additions any
// single term passes through on its way into the result (adding to an empty accumulator is exact
// and not counted).  The standard summation bound |fl(sum) - sum| <= depth * u * sum|terms|
// (first order, u = 2^-24) then holds for exactly the order the reference uses; the tensor-core
// path's error bound is built on it (tc.h).
int schedule_depth(const Schedule& s);
// returns false if dim is outside the supported range
bool build_schedule(int kind, int dim, Schedule* out);

}  // namespace kge
