// Tensor-core bound-and-refine path: shared definitions (tc.cu, api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kge {
namespace tc {

constexpr int TC_BM = 128;       // queries per MMA tile (TMEM lanes)
constexpr int TC_BN = 256;       // candidates per MMA tile (TMEM columns, fp32)
// k-block width (bf16 per swizzle span) is a run-time choice: bk() = 32 (64-byte swizzle,
// default) or 64 (128-byte swizzle, KGE_TC_BK=64); n_kblocks(k_total) = ceil(k_total / bk()).
int bk();
int n_kblocks(int k_total);
// Tuning / test hook (kge_tc_configure): bk 32|64, resident 0|1, ct_group (0 = automatic),
// max_ctas (0 = one per SM); negative values keep the current setting.  Operand images packed
// under one bk must be scanned under the same bk.
void configure(int bk, int resident, int ct_group, int max_ctas);

struct TcScanParams {
  const unsigned char* apack;  // [n_qt][n_kb][hi,lo][128 rows x 2*bk() B, swizzled]
  const unsigned char* bpack;  // [n_ct][n_kb][hi,lo][256 rows x 2*bk() B, swizzled]
  const float* s_true;         // [n_q] exact (ATen-order) true scores
  const float* qbound;         // [n_q] >= |a|_2
  const float* qnorm2;         // [n_q] |a|_2^2 (L2 only)
  const float* cbound;         // [n_rows] >= |b|_2
  const float* cnorm2;         // [n_rows]
  int32_t* counts;             // [n_q] +=
  unsigned long long* amb_count;  // [n_qt] fill count of each query tile's region
  int2* amb_pairs;                // [n_qt][amb_cap]
  unsigned long long amb_cap;     // capacity of ONE region
  float* dump;                 // debug: write approximate scores [n_q][n_rows] instead of counting
  float gamma;                 // tc_gamma(k)
  float gamma2;                // tc_gamma2(k) (L2 only)
  int l2;                      // 1: score = -(|a|^2 + |b|^2 - 2 a.b)
  int n_kb;
  int k_total;
  int ct_group;                // candidate tiles a CTA walks per query tile (set by launch_tc_scan)
  long long n_q, n_rows, n_qt, n_ct;
};

// Rigorous bound eps >= |s_tc - s_ATen| used by the threshold test:
//   dot models : eps = gamma  * |a| |b|
//   L2         : eps = gamma  * 2 |a| |b|  +  gamma2 * (|a| + |b|)^2
// gamma collects everything proportional to sum_k |a_k b_k| <= |a| |b|:
//   bf16 splitting x = hi + lo + r, |r| <= 2^-16 |x| (two roundings to 8 significant bits):
//     dropped lo*lo, a*r_b, r_a*b                                   -> 3 * 2^-16       (exact bound)
//   fp32 accumulation inside the tensor core: assumed <= 2 ulp of the running magnitude per
//     MMA instruction, 3 instructions per 16 terms, doubled for safety -> 2 (3 ceil(k/16) + 2) 2^-22
//   dot models only: the reference's own fp32 evaluation -- every product rounded once (ComplEx:
//     two products and their sum), then summed in ATen's cascade order, whose tree depth
//     `ref_depth` (schedule.h: schedule_depth, computed from the very schedule the exact kernels
//     replay) gives |fl(sum) - sum| <= ref_depth * u * sum|terms|   -> (ref_depth + 4) * 2^-24
// gamma2 (L2 only) collects what is proportional to the squared norms, (|a| + |b|)^2 >= |a - b|^2:
//   the reference forms x_k = q_k - c_k (head side: (c_k + r_k) - t_k, two roundings, |x| bounded
//   by |c| + |r| + |t| -- the head-side query bound is |t| + |r| for that reason), squares it,
//   sums in the 8-lane norm order (depth ref_depth), takes an exactly rounded sqrt and squares it:
//     relative (ref_depth + 10) u on sum x^2;
//   this side: |a|^2 and |b|^2 rounded to fp32 (2u), |b|^2/2 carried as three bf16 pieces (u),
//   the <= 6 MMA instructions that accumulate it (6 * 2^-21 * |b|^2 / 2 = 24 u |b|^2), threshold
//   arithmetic is directed-rounded                                    -> (ref_depth + 42) * 2^-24
// tests/test_tc_gpu.py measures the actual error on every model and requires it to stay below
// half of the bound.
inline float tc_gamma(int k_total, int ref_depth, bool l2) {
  const double split = 3.0 * 0x1p-16;
  const double accum = 2.0 * (3.0 * ((k_total + 15) / 16) + 2.0) * 0x1p-22;
  const double ref = l2 ? 0.0 : (ref_depth + 4.0) * 0x1p-24;
  return (float)(split + accum + ref);
}
inline float tc_gamma2(int ref_depth) { return (float)((ref_depth + 42.0) * 0x1p-24); }

size_t a_image_bytes(long long n_q, int n_kb);
size_t b_image_bytes(long long n_rows, int n_kb);
// fold = true (L2 models, k_total = dim + 3): the images carry -|b|^2/2 resp. 1.0 in the three k
// slots after the data, so the accumulator already holds  a.b - |b|^2/2.
cudaError_t launch_pack_b(const float* ent0, const float* ent1, long long n_rows, int dim, int k_total,
                          int n_kb, bool fold, unsigned char* bpack, float* cbound, float* cnorm2,
                          cudaStream_t st);
cudaError_t launch_pack_a(const float* qplain, int qw, long long n_q, int dim, int k_total, int n_kb,
                          int sub_mode, bool fold, unsigned char* apack, float* qbound, float* qnorm2,
                          cudaStream_t st);
cudaError_t launch_tc_scan(const TcScanParams& p, cudaStream_t st);
// The near-tie list is split into one region per QUERY TILE (regions = n_qt): region_counts[regions]
// (zeroed by the caller), pairs[regions][region_cap].  A region's pairs all belong to the same 128
// queries, whose rows therefore stay L1-resident during the exact recheck of that region.
int scan_grid_size(long long n_q, long long n_rows, int n_kb);
cudaError_t launch_recheck(int el, int dim, const unsigned long long* region_counts, int regions,
                           unsigned long long region_cap, const int2* pairs, const float* qplain,
                           const float* ent0, const float* ent1, const float* s_true, int32_t* counts,
                           unsigned long long* stats, cudaStream_t st);

}  // namespace tc
}  // namespace kge
