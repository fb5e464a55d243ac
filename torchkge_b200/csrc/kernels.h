// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kge_b200.h"
#include "reduce.cuh"

namespace kge {

constexpr int TILE_Q = KGE_TILE_Q;  // queries per CTA tile
constexpr int TILE_C = KGE_TILE_C;  // candidates per CTA tile

// Three-plane models (KGE_ANALOGY) hand over planes 0 and 1 of a table; the planes are equally
// spaced in memory (include/kge_b200.h), so plane 2 follows from the two pointers.
__host__ __device__ inline const float* third_plane(const float* p0, const float* p1) {
  return p1 == nullptr ? nullptr : p1 + (p1 - p0);
}
__host__ __device__ inline float* third_plane(float* p0, float* p1) {
  return p1 == nullptr ? nullptr : p1 + (p1 - p0);
}

// model/side -> element kind (-1 if unsupported)
int elem_kind_for(int model, int side);
int elem_qw(int el);
int elem_cw(int el);

constexpr int SCAN_KC = 32;          // schedule positions per pipeline stage of the scan
constexpr int SCAN_MAX_DIM = 8192;   // schedule bytes carried in the kernel parameters

// Passed BY VALUE as the kernel parameter (about 9 KB; CUDA >= 12.1 allows 32 KB): the
// schedule then lives in the constant bank, indexed with uniform registers, so that every
// test on it is a uniform branch (no convergence barriers, no shared-memory latency).
struct ScanParams {
  const float* packed;   // [n_ct][dim][CW][TILE_C]
  const float* qpacked;  // [n_qt][dim][QW][TILE_Q]
  const float* s_true;   // [n_qt*TILE_Q], NaN padded
  const uint8_t* code_host;  // [dim] schedule codes (HOST pointer; copied into `code` at launch)
  int32_t* counts;       // [n_q] (+=) or nullptr
  float* scores;         // [n_q][n_rows] or nullptr
  // bound-and-refine form (RotatE): approximate element arithmetic, pairs whose approximate
  // score is within rel_eps * |score| of s_true go to the near-tie list (regions of 128 queries)
  unsigned long long* amb_count;  // [ceil(n_q / 128)] or nullptr (exact scan)
  int2* amb_pairs;                // [regions][amb_cap]
  unsigned long long amb_cap;
  float rel_eps;
  float abs_eps;   // flushed subnormal terms: dim * 1.1e-19
  // top-k collect form (kge_topk_side): s_true holds a per-query THRESHOLD; every candidate whose
  // exact score is not below it is appended to the query's list as (score bits, global id)
  int2* col_buf;                  // [n_q][col_cap] or nullptr
  unsigned* col_count;            // [n_q] fill counts (unused when col_dense)
  unsigned long long col_cap;
  long long col_id_base;          // global id of candidate row 0 of this launch
  int col_dense;                  // 1: slot = row index (first chunk: everything is collected)
  int dim;
  int64_t n_q;
  int64_t n_rows;
  int64_t n_ct;
  int64_t n_qt;
  uint32_t mask[SCAN_MAX_DIM / SCAN_KC];  // per stage: bit kk set <=> position needs the slow path
  uint8_t code[SCAN_MAX_DIM];
};

// dense scan: counts[q] += #{c < n_rows : score(q,c) >= s_true[q]}  (or writes scores)
// approx = true (EL_ROT only): the bound-and-refine form above; p.amb_* and p.rel_eps must be set
cudaError_t launch_scan(int el, bool cascade, const ScanParams& p, cudaStream_t stream,
                        bool approx = false);

// packed[ct][pos][plane][TILE_C] <- ent_plane[row][perm[pos]]
cudaError_t launch_pack_table(const float* ent0, const float* ent1, int planes, int64_t n_rows,
                              int dim, const int32_t* inv_perm, float* packed,
                              cudaStream_t stream);

cudaError_t launch_gather_rows(const float* ent0, const float* ent1, int planes, int64_t ent_lo,
                               int64_t n_rows, int dim, const int64_t* idx, int64_t n,
                               float* out, cudaStream_t stream);

// qplain[i][plane][dim] from the gathered head/tail rows and the relation tables
cudaError_t launch_prep_queries(int model, int side, int dim, int64_t n, const float* hrows,
                                const float* trows, const float* rel0, const float* rel1,
                                const int64_t* r_idx, float* qplain, cudaStream_t stream);

// qpacked[qt][pos][plane][TILE_Q] <- qplain[i][plane][perm[pos]] (zero padded)
cudaError_t launch_pack_queries(const float* qplain, int qw, int dim, int64_t n,
                                const int32_t* perm, float* qpacked, cudaStream_t stream);

cudaError_t launch_fill_f32(float* dst, float value, int64_t n, cudaStream_t stream);

// s_true[i] = score(query i, rows[i])   (rows = [n][CW][dim])
cudaError_t launch_true_scores(int el, bool cascade, int dim, int64_t n, const float* qplain,
                               const float* rows, const int32_t* perm, const uint8_t* code,
                               float* s_true, cudaStream_t stream);

// filt_sub[q] += [s(q,c) >= s_true[q]] - [s_true[q] == -inf] for every CSR entry c of q that
// lies in [ent_lo, ent_lo + n_rows)
cudaError_t launch_filter(int el, bool cascade, int dim, int64_t n, int64_t n_filt,
                          const float* qplain,
                          const float* ent0, const float* ent1, int64_t ent_lo, int64_t n_rows,
                          const int64_t* offs, const int64_t* ids, const int32_t* qid, const int32_t* perm,
                          const uint8_t* code, const float* s_true, int32_t* filt_sub,
                          cudaStream_t stream);

cudaError_t launch_finalize(const int32_t* raw, const int32_t* sub, int64_t n, int64_t* ranks,
                            int64_t* filt_ranks, cudaStream_t stream);

// ---- top-k selection over collected candidates (topk.cu) ----
// best[q][k] sorted 64-bit keys (score order, then smaller id first; 0 = empty slot).
// Merges the `count[q]` (or `dense_count`) entries of col_buf[q] into best[q], masking ids listed in
// the query's sorted CSR row with -inf (filter_scores with true_idx = None, utils/modeling.py:91-102),
// and writes the new threshold thr[q] (the k-th best score, -inf while fewer than k are held).
cudaError_t launch_topk_merge(unsigned long long* best, int k, const int2* col_buf,
                              const unsigned* col_count, unsigned long long col_cap, long long dense_count,
                              const int64_t* mask_offs, const int64_t* mask_ids, float* thr, int64_t n_q,
                              cudaStream_t stream);
// pred[q][j], scores[q][j] from the keys
cudaError_t launch_topk_finish(const unsigned long long* best, int k, int64_t n_q, int64_t* pred,
                               float* scores, cudaStream_t stream);
constexpr int TOPK_MAX_K = 1024;

// ---- dense side paths (dense.cu) ----
// scores[i][c] = ((h_i^T M_c) * t_i).sum()   RESCAL relation case, bilinear.py:115-121
cudaError_t launch_rescal_rel_scores(const float* hrows, const float* trows, const float* rel_mat, int dim,
                                     int64_t n, int64_t n_rel, float* scores, cudaStream_t stream);
cudaError_t launch_rank_dense(const float* scores, int64_t n, int64_t n_c, const int64_t* true_idx,
                              const float* true_score_in, const int64_t* offs, const int64_t* ids,
                              int32_t* raw_count, int32_t* filt_sub, float* true_score_out, cudaStream_t stream);
cudaError_t launch_dense_to_pairs(const float* scores, int64_t n, int64_t n_c, int2* pairs, cudaStream_t stream);

}  // namespace kge
