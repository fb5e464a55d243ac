// Tensor-core bound-and-refine path: shared definitions (tc.cu, api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kge {
namespace tc {

constexpr int TC_BM = 128;       // queries per MMA tile (TMEM lanes)
constexpr int TC_BN = 256;       // candidates per MMA tile (TMEM columns, fp32)
constexpr int TC_BK = 64;        // bf16 per k-block = one 128-byte swizzle span
constexpr int TC_CT_GROUP = 16;  // candidate tiles a CTA walks per query tile before moving on

struct TcScanParams {
  const unsigned char* apack;  // [n_qt][n_kb][hi,lo][128 x 128 B swizzled]
  const unsigned char* bpack;  // [n_ct][n_kb][hi,lo][256 x 128 B swizzled]
  const float* s_true;         // [n_q] exact (ATen-order) true scores
  const float* qbound;         // [n_q] >= |a|_2
  const float* qnorm2;         // [n_q] |a|_2^2 (L2 only)
  const float* cbound;         // [n_rows] >= |b|_2
  const float* cnorm2;         // [n_rows]
  int32_t* counts;             // [n_q] +=
  unsigned long long* amb_count;
  int2* amb_pairs;
  unsigned long long amb_cap;
  float* dump;                 // debug: write approximate scores [n_q][n_rows] instead of counting
  float gamma;                 // relative error bound factor (tc_gamma)
  int l2;                      // 1: score = -(|a|^2 + |b|^2 - 2 a.b), eps = gamma (|a|+|b|)^2
  int n_kb;
  int k_total;
  long long n_q, n_rows, n_qt, n_ct;
};

// Bound on |s_tc - s_ATen| relative to |a|_2 |b|_2 (dot models) or (|a|_2 + |b|_2)^2 (L2):
//   splitting x = hi + lo + r, |r| <= 2^-16 |x| (two bf16 roundings)   -> 3 * 2^-16 incl. lo*lo
//   fp32 accumulation on the tensor core, <= 2 ulp per MMA instruction  -> (3 k/16 + 2) * 2^-22
//   the reference's own fp32 evaluation (products, <= k-term sums)      -> (k + 4) * 2^-24
//   epilogue arithmetic (norm sums, scaling)                            -> 8 * 2^-24
// times a safety factor of 2.  tests/test_tc_gpu.py measures the actual error (it is ~20x
// smaller) and checks it stays below half of this.
inline float tc_gamma(int k_total, bool l2) {
  const double split = 3.0 * 0x1p-16;
  const double accum = (3.0 * ((k_total + 15) / 16) + 2.0) * 0x1p-22;
  const double ref = (k_total + 4.0) * 0x1p-24;
  const double epi = 8.0 * 0x1p-24;
  double g = 2.0 * (split + accum + ref + epi);
  if (l2) g *= 1.5;  // the three-term expansion |a|^2 + |b|^2 - 2ab carries each error once more
  return (float)g;
}

size_t a_image_bytes(long long n_q, int n_kb);
size_t b_image_bytes(long long n_rows, int n_kb);
cudaError_t launch_pack_b(const float* ent0, const float* ent1, long long n_rows, int dim, int k_total,
                          int n_kb, unsigned char* bpack, float* cbound, float* cnorm2, cudaStream_t st);
cudaError_t launch_pack_a(const float* qplain, int qw, long long n_q, int dim, int k_total, int n_kb,
                          int sub_mode, unsigned char* apack, float* qbound, float* qnorm2,
                          cudaStream_t st);
cudaError_t launch_tc_scan(const TcScanParams& p, cudaStream_t st);
cudaError_t launch_recheck(int el, bool cascade, int dim, const unsigned long long* n_pairs_dev,
                           unsigned long long cap, const int2* pairs,
                           const float* qplain, const float* ent0, const float* ent1,
                           const int32_t* perm, const uint8_t* code, const float* s_true,
                           int32_t* counts, cudaStream_t st);

}  // namespace tc
}  // namespace kge
