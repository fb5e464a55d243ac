// Dense-score side paths: RESCAL's relation case and ranking / top-k over a small dense score matrix.
//
// RESCAL relation prediction (bilinear.py:115-121): the candidates are the relation MATRICES, so
// every (fact, relation) pair has its own vector  hr = h^T M_c  (batched matmul over b * n_rel
// (1 x d)(d x d) products) before the usual  (hr * t).sum(dim=2).  There is no shared candidate
// table to scan; the (n, n_rel) score matrix is small (n_rel relations) and is produced densely,
// in the reference's arithmetic: rescal_query_component (oneMKL order) then the ATen cascade sum.
#include "kernels.h"

namespace kge {

namespace {

constexpr int RR_THREADS = 256;
constexpr int RR_ITILE = 8;   // facts per CTA: one warp finishes one fact's score

// grid (n_rel, ceil(n / RR_ITILE)).  Thread t owns columns j = t, t + 256, ... of hr for the
// CTA's RR_ITILE facts: M_c[k][j] is loaded once per (k, j) and used for all of them.
__global__ void __launch_bounds__(RR_THREADS)
    rescal_rel_scores_kernel(const float* __restrict__ hrows, const float* __restrict__ trows,
                             const float* __restrict__ rel_mat, int dim, long long n, long long n_rel,
                             float* __restrict__ scores) {
  extern __shared__ float sm[];       // [RR_ITILE][dim] h rows, then [RR_ITILE][dim] hr
  float* sh = sm;
  float* shr = sm + (size_t)RR_ITILE * dim;
  const long long c = blockIdx.x;
  const long long i0 = (long long)blockIdx.y * RR_ITILE;
  const float* M = rel_mat + (size_t)c * dim * dim;
  for (int x = threadIdx.x; x < RR_ITILE * dim; x += RR_THREADS) {
    const long long i = i0 + x / dim;
    sh[x] = i < n ? hrows[(size_t)i * dim + x % dim] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < dim; j += RR_THREADS) {
#pragma unroll
    for (int il = 0; il < RR_ITILE; ++il)
      shr[(size_t)il * dim + j] = rescal_query_component(true, dim, j, sh + (size_t)il * dim, M);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = i0 + warp;     // RR_THREADS / 32 == RR_ITILE
  if (i >= n) return;
  const float* hr = shr + (size_t)warp * dim;
  const float* t = trows + (size_t)i * dim;
  float s;
  if (dim < 8) s = pair_score_natural<EL_DOT1>(dim, hr, hr, t, t);
  else s = pair_score_chains<EL_DOT1>(dim, hr, hr, t, t, lane);
  if (lane == 0) scores[(size_t)i * n_rel + c] = s;
}

// One warp per row of a dense (n, n_c) score matrix:
//   raw_count[i] += #{c : s[i][c] >= s_true(i)},
//   filt_sub[i]  += sum over the row's CSR entries of [s[i][c] >= s_true(i)] - [s_true(i) == -inf]
// (get_rank, utils/operations.py:37-61, and filter_scores, utils/modeling.py:91-102, on a matrix that
// is small enough to exist).  s_true(i) = true_score_in[i] if given, else s[i][true_idx[i]].
__global__ void rank_dense_kernel(const float* __restrict__ scores, long long n, long long n_c,
                                  const int64_t* __restrict__ true_idx, const float* __restrict__ true_score_in,
                                  const int64_t* __restrict__ offs, const int64_t* __restrict__ ids,
                                  int32_t* __restrict__ raw_count, int32_t* __restrict__ filt_sub,
                                  float* __restrict__ true_score_out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* row = scores + (size_t)i * n_c;
  const float st = true_score_in ? true_score_in[i] : row[true_idx[i]];
  int cnt = 0, sub = 0;
  for (long long c = lane; c < n_c; c += 32) cnt += row[c] >= st ? 1 : 0;
  if (offs) {
    for (long long e = offs[i] + lane; e < offs[i + 1]; e += 32) {
      const long long c = ids[e];
      if (c >= 0 && c < n_c) sub += (row[c] >= st ? 1 : 0) - (st == -INFINITY ? 1 : 0);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sub += __shfl_xor_sync(0xffffffffu, sub, o);
  }
  if (lane == 0) {
    raw_count[i] += cnt;
    if (filt_sub) filt_sub[i] += sub;
    if (true_score_out) true_score_out[i] = st;
  }
}

__global__ void dense_to_pairs_kernel(const float* __restrict__ scores, long long total, long long n_c,
                                      int2* __restrict__ pairs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  pairs[gid] = make_int2(__float_as_int(scores[gid]), (int)(gid % n_c));
}

}  // namespace

cudaError_t launch_rescal_rel_scores(const float* hrows, const float* trows, const float* rel_mat, int dim,
                                     int64_t n, int64_t n_rel, float* scores, cudaStream_t stream) {
  if (n <= 0 || n_rel <= 0) return cudaSuccess;
  const size_t smem = (size_t)2 * RR_ITILE * dim * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static bool configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || !configured[dev])) {
    e = cudaFuncSetAttribute(rescal_rel_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const long long i_tiles = (n + RR_ITILE - 1) / RR_ITILE;
  for (long long y0 = 0; y0 < i_tiles; y0 += 65535) {     // gridDim.y limit
    const long long ny = i_tiles - y0 < 65535 ? i_tiles - y0 : 65535;
    dim3 grid((unsigned)n_rel, (unsigned)ny);
    const long long i_off = y0 * RR_ITILE;
    rescal_rel_scores_kernel<<<grid, RR_THREADS, smem, stream>>>(hrows + (size_t)i_off * dim, trows + (size_t)i_off * dim,
                                                               rel_mat, dim, n - i_off, n_rel,
                                                               scores + (size_t)i_off * n_rel);
  }
  return cudaGetLastError();
}

cudaError_t launch_rank_dense(const float* scores, int64_t n, int64_t n_c, const int64_t* true_idx,
                              const float* true_score_in, const int64_t* offs, const int64_t* ids,
                              int32_t* raw_count, int32_t* filt_sub, float* true_score_out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  rank_dense_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(scores, n, n_c, true_idx, true_score_in, offs,
                                                                          ids, raw_count, filt_sub, true_score_out);
  return cudaGetLastError();
}

cudaError_t launch_dense_to_pairs(const float* scores, int64_t n, int64_t n_c, int2* pairs, cudaStream_t stream) {
  const long long total = (long long)n * n_c;
  if (total <= 0) return cudaSuccess;
  dense_to_pairs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(scores, total, n_c, pairs);
  return cudaGetLastError();
}

}  // namespace kge
