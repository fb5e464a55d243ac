// Dense rank scan: the kernel that replaces inference_scoring_function + get_rank
// (torchkge models/interfaces.py:240-260, models/bilinear.py:98-121,224-245,501-528,
// utils/operations.py:37-61) for one side of a batch of test triples.
//
// Shape of the work: S[q][c] = reduce_k f(Q[q][k], E[c][k]) over n_q queries x n_rows
// candidates x dim, of which only  #{c : S[q][c] >= s_true[q]}  per query is kept.  With
// tens of thousands of queries sharing every streamed candidate tile the kernel is bound by
// fp32 issue, not by HBM, so it is organised like an SGEMM: a persistent CTA per SM walks
// (query tile, candidate tile) pairs; a producer warp streams schedule-ordered k-chunks of
// both operands into a 4-stage shared-memory ring with 1-D bulk async copies (UBLKCP)
// signalled on mbarriers; 8 consumer warps hold a 4 x 8 register tile of running reductions
// per thread.  Both operands are pre-laid out k-major ("packed") so every shared-memory read
// is a conflict-free 128-bit load and the reduction schedule is walked front to back.
#include <stdio.h>

#include "kernels.h"
#include "ptx.cuh"

namespace kge {

namespace {

constexpr int KC = 16;           // schedule positions per pipeline stage
constexpr int STAGES = 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int TQ = 4;            // queries per thread
constexpr int TC = 8;            // candidates per thread (two groups of 4)
constexpr int CODE_BYTES = 8192;

template <int EL>
struct StageLayout {
  static constexpr int QW = ElemTraits<EL>::QW;
  static constexpr int CW = ElemTraits<EL>::CW;
  static constexpr int C_FLOATS = KC * CW * TILE_C;
  static constexpr int Q_FLOATS = KC * QW * TILE_Q;
  static constexpr int STAGE_FLOATS = C_FLOATS + Q_FLOATS;
  static constexpr size_t SMEM_BYTES =
      (size_t)STAGES * STAGE_FLOATS * sizeof(float) + CODE_BYTES + 2 * STAGES * sizeof(uint64_t);
};

template <int EL, bool CASC>
__global__ void __launch_bounds__(THREADS, 1) scan_kernel(const ScanParams p) {
  using L = StageLayout<EL>;
  constexpr int QW = L::QW, CW = L::CW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint8_t* s_code = smem_raw + (size_t)STAGES * L::STAGE_FLOATS * sizeof(float);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_code + CODE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int dim = p.dim;
  const int n_kc = (dim + KC - 1) / KC;
  const long long total_tiles = (long long)p.n_qt * (long long)p.n_ct;

  for (int i = threadIdx.x; i < dim; i += THREADS) s_code[i] = p.code[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], CONSUMER_WARPS);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ------------------------------ producer warp ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long qt = tile / p.n_ct;
        const long long ct = tile - qt * p.n_ct;
        const float* csrc = p.packed + (size_t)ct * dim * (CW * TILE_C);
        const float* qsrc = p.qpacked + (size_t)qt * dim * (QW * TILE_Q);
        for (int kc = 0; kc < n_kc; ++kc) {
          const int kn = min(KC, dim - kc * KC);
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          float* sc = stage_base + (size_t)stage * L::STAGE_FLOATS;
          float* sq = sc + L::C_FLOATS;
          const uint32_t cbytes = (uint32_t)kn * CW * TILE_C * sizeof(float);
          const uint32_t qbytes = (uint32_t)kn * QW * TILE_Q * sizeof(float);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], cbytes + qbytes);
          ptx::bulk_g2s(sc, csrc + (size_t)kc * KC * (CW * TILE_C), cbytes, &full_bar[stage]);
          ptx::bulk_g2s(sq, qsrc + (size_t)kc * KC * (QW * TILE_Q), qbytes, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // -------------------------------- consumer warps --------------------------------
  const int wq = warp >> 1, wc = warp & 1;  // 4 x 2 warps over the 64 x 128 CTA tile
  const int tq = lane >> 3, tc = lane & 7;  // 4 x 8 threads over the 16 x 64 warp tile
  const int q_off = wq * 16 + tq * TQ;
  const int c_off0 = wc * 64 + tc * 4;
  const int c_off1 = c_off0 + 32;
  constexpr uint8_t FAST = fast_code<EL>();

  int stage = 0;
  uint32_t phase = 0;
  long long cur_qt = -1;
  float st[TQ];
  int cnt[TQ];
#pragma unroll
  for (int i = 0; i < TQ; ++i) { st[i] = 0.f; cnt[i] = 0; }

  auto flush_counts = [&](long long qt) {
    if (p.counts == nullptr || qt < 0) return;
#pragma unroll
    for (int i = 0; i < TQ; ++i) {
      int v = cnt[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      const long long q = qt * TILE_Q + q_off + i;
      if (tc == 0 && v != 0 && q < p.n_q) atomicAdd(&p.counts[q], v);
      cnt[i] = 0;
    }
  };

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const long long qt = tile / p.n_ct;
    const long long ct = tile - qt * p.n_ct;
    if (qt != cur_qt) {
      flush_counts(cur_qt);
      cur_qt = qt;
#pragma unroll
      for (int i = 0; i < TQ; ++i) st[i] = p.s_true[qt * TILE_Q + q_off + i];
    }

    Acc acc[TQ][TC];
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
      for (int j = 0; j < TC; ++j) acc_reset(acc[i][j]);

    for (int kc = 0; kc < n_kc; ++kc) {
      const int kn = min(KC, dim - kc * KC);
      ptx::mbar_wait(&full_bar[stage], phase);
      const float* sc = stage_base + (size_t)stage * L::STAGE_FLOATS;
      const float* sq = sc + L::C_FLOATS;
      const uint8_t* codes = s_code + kc * KC;
#pragma unroll 2
      for (int kk = 0; kk < kn; ++kk) {
        float qv[QW][TQ];
        float cv[CW][TC];
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          const float4 v = *reinterpret_cast<const float4*>(sq + (kk * QW + w) * TILE_Q + q_off);
          qv[w][0] = v.x; qv[w][1] = v.y; qv[w][2] = v.z; qv[w][3] = v.w;
        }
#pragma unroll
        for (int w = 0; w < CW; ++w) {
          const float4 v0 = *reinterpret_cast<const float4*>(sc + (kk * CW + w) * TILE_C + c_off0);
          const float4 v1 = *reinterpret_cast<const float4*>(sc + (kk * CW + w) * TILE_C + c_off1);
          cv[w][0] = v0.x; cv[w][1] = v0.y; cv[w][2] = v0.z; cv[w][3] = v0.w;
          cv[w][4] = v1.x; cv[w][5] = v1.y; cv[w][6] = v1.z; cv[w][7] = v1.w;
        }
        const uint8_t code = codes[kk];
        if (code == FAST) {
#pragma unroll
          for (int i = 0; i < TQ; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j)
              acc_step_fast<EL>(acc[i][j], qv[0][i], qv[QW - 1][i], cv[0][j], cv[CW - 1][j]);
        } else {
#pragma unroll
          for (int i = 0; i < TQ; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j)
              acc_step<EL, CASC>(acc[i][j], code, qv[0][i], qv[QW - 1][i], cv[0][j],
                                 cv[CW - 1][j]);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }

    // ---- epilogue: finish the 32 scores, count or store ----
    const long long c_base = ct * TILE_C;
    if (p.scores != nullptr) {
#pragma unroll
      for (int i = 0; i < TQ; ++i) {
        const long long q = qt * TILE_Q + q_off + i;
        if (q >= p.n_q) continue;
        float* row = p.scores + (size_t)q * p.n_rows;
#pragma unroll
        for (int j = 0; j < TC; ++j) {
          const long long c = c_base + (j < 4 ? c_off0 + j : c_off1 + j - 4);
          if (c < p.n_rows) row[c] = acc_finish<EL>(acc[i][j]);
        }
      }
    } else {
      const bool edge = (c_base + TILE_C > p.n_rows);
#pragma unroll
      for (int j = 0; j < TC; ++j) {
        const long long c = c_base + (j < 4 ? c_off0 + j : c_off1 + j - 4);
        const bool valid = !edge || (c < p.n_rows);
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
          const float s = acc_finish<EL>(acc[i][j]);
          cnt[i] += (valid && s >= st[i]) ? 1 : 0;
        }
      }
    }
  }
  flush_counts(cur_qt);
}

template <int EL, bool CASC>
cudaError_t launch_one(const ScanParams& p, cudaStream_t stream) {
  using L = StageLayout<EL>;
  static bool configured = false;  // benign race: attribute set is idempotent
  auto kern = scan_kernel<EL, CASC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const long long total = (long long)p.n_qt * (long long)p.n_ct;
  if (total <= 0) return cudaSuccess;
  const int grid = (int)(total < sms ? total : sms);
  kern<<<grid, THREADS, L::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_scan(int el, bool cascade, const ScanParams& p, cudaStream_t stream) {
  if (p.dim < 1 || p.dim > CODE_BYTES - 1) return cudaErrorInvalidValue;
  switch (el) {
    case EL_DOT1:
      return cascade ? launch_one<EL_DOT1, true>(p, stream) : launch_one<EL_DOT1, false>(p, stream);
    case EL_DOT2:
      return cascade ? launch_one<EL_DOT2, true>(p, stream) : launch_one<EL_DOT2, false>(p, stream);
    case EL_ROT:
      return cascade ? launch_one<EL_ROT, true>(p, stream) : launch_one<EL_ROT, false>(p, stream);
    case EL_L1_TAIL: return launch_one<EL_L1_TAIL, false>(p, stream);
    case EL_L1_HEAD: return launch_one<EL_L1_HEAD, false>(p, stream);
    case EL_L2_TAIL: return launch_one<EL_L2_TAIL, false>(p, stream);
    case EL_L2_HEAD: return launch_one<EL_L2_HEAD, false>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace kge
