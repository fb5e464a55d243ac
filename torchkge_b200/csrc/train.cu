// Training-side kernels: per-triple scoring (Model.scoring_function), Bernoulli corruption
// (BernoulliNegativeSampler.corrupt_batch), margin loss (MarginLoss) and the fused
// sample + score + hinge step, each with its backward.
//
// Reference bodies replaced: models/translation.py:69-81, models/bilinear.py:60-71, 188-199,
// 460-473, models/interfaces.py:39-82, sampling.py:278-327, utils/losses.py:12-44.
//
// These paths are gather-bound (a few 4d-byte rows per triple, random rows): one warp owns
// one triple (or one positive with all its negatives), lanes stride over the embedding
// index so every row read is a run of coalesced 128-B segments, and the per-triple
// reductions (L2 norms, dot products) are warp-shuffle trees.  Parity with the reference is
// by tolerance here (1e-5 relative on scores / loss), so fused
// multiply-adds and tree reductions are allowed, unlike in the ranking kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <stdlib.h>

#include "../../include/kge_b200.h"
#include "ptx.cuh"
#include "train.h"

namespace kge {

namespace {

constexpr float NORM_EPS = 1e-12f;  // torch.nn.functional.normalize default eps
constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ bool model_normalises(int model) {
  return model == KGE_TRANSE_L1 || model == KGE_TRANSE_L2 || model == KGE_DISTMULT ||
         model == KGE_RESCAL;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (idx, offset), key = seed -------------
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32);
  uint32_t c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// One corrupted triple: Bernoulli(p_r) decides head vs tail, the replacement is uniform on
// [1, n_ent) -- entity 0 is never drawn and true triples are not rejected, as in
// sampling.py:318-325.
__device__ __forceinline__ void corrupt_one(uint64_t seed, uint64_t offset, uint64_t idx, float p,
                                            long long n_ent, long long h, long long t,
                                            long long* nh, long long* nt) {
  const uint4 rnd = philox4x32(seed, offset, idx);
  const float u = (rnd.x >> 8) * (1.0f / 16777216.0f);  // [0, 1)
  const long long span = n_ent - 1;
  long long e = 1;
  if (span > 0) e = 1 + (long long)(((unsigned long long)rnd.y * (unsigned long long)span) >> 32);
  const bool head = u < p;
  *nh = head ? e : h;
  *nt = head ? t : e;
}

// ------------------------------------------------------------------------------------------
// Per-lane view of one triple.  Lane l owns embedding indices l, l+32, ...; `cnt` of them.
// ------------------------------------------------------------------------------------------
struct RowPtrs {
  const float* h0; const float* h1;  // head planes
  const float* t0; const float* t1;  // tail planes
  const float* r0; const float* r1;  // relation planes (RESCAL: r0 = matrix)
  const float* h2; const float* t2; const float* r2;  // third planes (Analogy) or nullptr
};

__device__ __forceinline__ float inv_norm_of(const float* row, int dim, int lane) {
  float s = 0.f;
  for (int k = lane; k < dim; k += 32) { const float v = row[k]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  return 1.0f / fmaxf(sqrtf(s), NORM_EPS);
}

// torch.frac: the fractional part keeps the sign of its argument
__device__ __forceinline__ float frac_of(float v) { return v - truncf(v); }

// score of one triple; all lanes return the same value
__device__ float triple_score(int model, int dim, const RowPtrs& p, int lane, float* inv_h_out,
                              float* inv_t_out) {
  float inv_h = 1.f, inv_t = 1.f;
  if (model_normalises(model)) {
    inv_h = inv_norm_of(p.h0, dim, lane);
    inv_t = inv_norm_of(p.t0, dim, lane);
  }
  if (inv_h_out) *inv_h_out = inv_h;
  if (inv_t_out) *inv_t_out = inv_t;
  float s = 0.f;
  switch (model) {
    case KGE_TRANSE_L1:
      for (int k = lane; k < dim; k += 32) s += fabsf(p.h0[k] * inv_h + p.r0[k] - p.t0[k] * inv_t);
      return -warp_sum(s);
    case KGE_TRANSE_L2:
      for (int k = lane; k < dim; k += 32) {
        const float x = p.h0[k] * inv_h + p.r0[k] - p.t0[k] * inv_t;
        s = fmaf(x, x, s);
      }
      return -warp_sum(s);
    case KGE_DISTMULT:
      for (int k = lane; k < dim; k += 32) s = fmaf(p.h0[k] * inv_h * p.r0[k], p.t0[k] * inv_t, s);
      return warp_sum(s);
    case KGE_RESCAL:
      // s = sum_j (sum_i h_i M_ij) t_j ; lanes stride over j so M rows are read coalesced
      for (int j = lane; j < dim; j += 32) {
        float q = 0.f;
        for (int i = 0; i < dim; ++i) q = fmaf(p.h0[i] * inv_h, p.r0[(size_t)i * dim + j], q);
        s = fmaf(q, p.t0[j] * inv_t, s);
      }
      return warp_sum(s);
    case KGE_COMPLEX:
      for (int k = lane; k < dim; k += 32) {
        const float rh = p.h0[k], ih = p.h1[k], rt = p.t0[k], it = p.t1[k], rr = p.r0[k], ir = p.r1[k];
        s += rh * (rr * rt + ir * it) + ih * (rr * it - ir * rt);
      }
      return warp_sum(s);
    case KGE_ROTATE:
      for (int k = lane; k < dim; k += 32) {
        const float rh = p.h0[k], ih = p.h1[k], rt = p.t0[k], it = p.t1[k], rr = p.r0[k], ir = p.r1[k];
        const float a = rh * rr - ih * ir - rt, b = rh * ir + ih * rr - it;
        s += sqrtf(a * a + b * b);
      }
      return -warp_sum(s);
    case KGE_TORUSE_L1:   // translation.py:706-720: -diss(frac(h) + frac(r), frac(t)), dissimilarities.py:28-43
    case KGE_TORUSE_L2:
      for (int k = lane; k < dim; k += 32) {
        const float x = (frac_of(p.h0[k]) + frac_of(p.r0[k])) - frac_of(p.t0[k]);
        if (model == KGE_TORUSE_L1) { const float ax = fabsf(x); s += 2.f * fminf(ax, 1.f - ax); }
        else { const float x2 = x * x; s += 4.f * fminf(x2, 1.f - x2); }
      }
      return -warp_sum(s);
    case KGE_ANALOGY:   // bilinear.py:634-650: DistMult on the scalar plane + ComplEx on (real, imaginary)
      for (int k = lane; k < dim; k += 32) {
        const float rh = p.h1[k], ih = p.h2[k], rt = p.t1[k], it = p.t2[k], rr = p.r1[k], ir = p.r2[k];
        s += p.h0[k] * p.r0[k] * p.t0[k] + (rh * (rr * rt + ir * it) + ih * (rr * it - ir * rt));
      }
      return warp_sum(s);
    default: return 0.f;
  }
}

__device__ __forceinline__ RowPtrs make_rows(int model, int dim, const TrainTables& tb, long long h,
                                             long long t, long long r) {
  RowPtrs p;
  p.h0 = tb.ent0 + (size_t)h * dim;
  p.t0 = tb.ent0 + (size_t)t * dim;
  p.h1 = tb.ent1 ? tb.ent1 + (size_t)h * dim : nullptr;
  p.t1 = tb.ent1 ? tb.ent1 + (size_t)t * dim : nullptr;
  const size_t rstride = model == KGE_RESCAL ? (size_t)dim * dim : (size_t)dim;
  p.r0 = tb.rel0 + (size_t)r * rstride;
  p.r1 = tb.rel1 ? tb.rel1 + (size_t)r * rstride : nullptr;
  p.h2 = p.t2 = p.r2 = nullptr;
  if (model == KGE_ANALOGY) {   // planes equally spaced in memory (include/kge_b200.h)
    const float* e2 = tb.ent1 + (tb.ent1 - tb.ent0);
    p.h2 = e2 + (size_t)h * dim;
    p.t2 = e2 + (size_t)t * dim;
    p.r2 = tb.rel1 + (tb.rel1 - tb.rel0) + (size_t)r * dim;
  }
  return p;
}

// Gradient of one triple's score, scaled by g, scattered into the dense gradient tables.
// Through F.normalize:  d/dh = (G - h~ (h~ . G)) / max(|h|, eps)  with G = dscore/dh~.
__device__ void triple_backward(int model, int dim, const RowPtrs& p, const TrainGrads& gr,
                                long long h, long long t, long long r, float g, int lane) {
  if (g == 0.f) return;
  float* gh0 = gr.ent0 + (size_t)h * dim;
  float* gt0 = gr.ent0 + (size_t)t * dim;
  const size_t rstride = model == KGE_RESCAL ? (size_t)dim * dim : (size_t)dim;
  float* gr0 = gr.rel0 + (size_t)r * rstride;
  if (model == KGE_COMPLEX || model == KGE_ROTATE) {
    float* gh1 = gr.ent1 + (size_t)h * dim;
    float* gt1 = gr.ent1 + (size_t)t * dim;
    float* gr1 = gr.rel1 + (size_t)r * dim;
    for (int k = lane; k < dim; k += 32) {
      const float rh = p.h0[k], ih = p.h1[k], rt = p.t0[k], it = p.t1[k], rr = p.r0[k], ir = p.r1[k];
      float d_rh, d_ih, d_rt, d_it, d_rr, d_ir;
      if (model == KGE_COMPLEX) {
        d_rh = rr * rt + ir * it; d_ih = rr * it - ir * rt;
        d_rt = rh * rr - ih * ir; d_it = rh * ir + ih * rr;
        d_rr = rh * rt + ih * it; d_ir = rh * it - ih * rt;
      } else {
        const float a = rh * rr - ih * ir - rt, b = rh * ir + ih * rr - it;
        const float m = sqrtf(a * a + b * b);
        const float da = m > 0.f ? -a / m : 0.f, db = m > 0.f ? -b / m : 0.f;
        d_rh = da * rr + db * ir; d_ih = -da * ir + db * rr;
        d_rt = -da; d_it = -db;
        d_rr = da * rh + db * ih; d_ir = -da * ih + db * rh;
      }
      atomicAdd(gh0 + k, g * d_rh); atomicAdd(gh1 + k, g * d_ih);
      atomicAdd(gt0 + k, g * d_rt); atomicAdd(gt1 + k, g * d_it);
      atomicAdd(gr0 + k, g * d_rr); atomicAdd(gr1 + k, g * d_ir);
    }
    return;
  }
  if (model == KGE_ANALOGY) {
    float* gh1 = gr.ent1 + (size_t)h * dim;
    float* gt1 = gr.ent1 + (size_t)t * dim;
    float* gr1 = gr.rel1 + (size_t)r * dim;
    float* ge2 = gr.ent1 + (gr.ent1 - gr.ent0);
    float* gh2 = ge2 + (size_t)h * dim;
    float* gt2 = ge2 + (size_t)t * dim;
    float* gr2 = gr.rel1 + (gr.rel1 - gr.rel0) + (size_t)r * dim;
    for (int k = lane; k < dim; k += 32) {
      const float sh = p.h0[k], st = p.t0[k], sr = p.r0[k];
      atomicAdd(gh0 + k, g * (sr * st)); atomicAdd(gt0 + k, g * (sh * sr)); atomicAdd(gr0 + k, g * (sh * st));
      const float rh = p.h1[k], ih = p.h2[k], rt = p.t1[k], it = p.t2[k], rr = p.r1[k], ir = p.r2[k];
      atomicAdd(gh1 + k, g * (rr * rt + ir * it)); atomicAdd(gh2 + k, g * (rr * it - ir * rt));
      atomicAdd(gt1 + k, g * (rh * rr - ih * ir)); atomicAdd(gt2 + k, g * (rh * ir + ih * rr));
      atomicAdd(gr1 + k, g * (rh * rt + ih * it)); atomicAdd(gr2 + k, g * (rh * it - ih * rt));
    }
    return;
  }
  if (model == KGE_TORUSE_L1 || model == KGE_TORUSE_L2) {
    // x = frac(h) + frac(r) - frac(t) (frac has unit slope); min(u, v) sends the gradient to the
    // smaller argument and, as torch.min does, half to each at an exact tie
    for (int k = lane; k < dim; k += 32) {
      const float x = (frac_of(p.h0[k]) + frac_of(p.r0[k])) - frac_of(p.t0[k]);
      float d;   // d score / d x
      if (model == KGE_TORUSE_L1) {
        const float ax = fabsf(x), sg = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
        d = ax < 1.f - ax ? -2.f * sg : (ax > 1.f - ax ? 2.f * sg : 0.f);
      } else {
        const float x2 = x * x;
        d = x2 < 1.f - x2 ? -8.f * x : (x2 > 1.f - x2 ? 8.f * x : 0.f);
      }
      atomicAdd(gh0 + k, g * d); atomicAdd(gr0 + k, g * d); atomicAdd(gt0 + k, -g * d);
    }
    return;
  }
  // normalising models
  const float inv_h = inv_norm_of(p.h0, dim, lane), inv_t = inv_norm_of(p.t0, dim, lane);
  // pass 1: G . h~ and G . t~
  float dot_h = 0.f, dot_t = 0.f;
  for (int k = lane; k < dim; k += 32) {
    const float hn = p.h0[k] * inv_h, tn = p.t0[k] * inv_t;
    float Gh, Gt;
    if (model == KGE_DISTMULT) {
      Gh = p.r0[k] * tn; Gt = hn * p.r0[k];
    } else if (model == KGE_RESCAL) {
      float a = 0.f, b = 0.f;  // Gh_k = sum_j M_kj t_j ; Gt_k = sum_i h_i M_ik
      for (int j = 0; j < dim; ++j) {
        a = fmaf(p.r0[(size_t)k * dim + j], p.t0[j] * inv_t, a);
        b = fmaf(p.h0[j] * inv_h, p.r0[(size_t)j * dim + k], b);
      }
      Gh = a; Gt = b;
    } else {
      const float x = hn + p.r0[k] - tn;
      const float dx = model == KGE_TRANSE_L2 ? -2.f * x : (x > 0.f ? -1.f : (x < 0.f ? 1.f : 0.f));
      Gh = dx; Gt = -dx;
    }
    dot_h = fmaf(Gh, hn, dot_h); dot_t = fmaf(Gt, tn, dot_t);
  }
  dot_h = warp_sum(dot_h); dot_t = warp_sum(dot_t);
  // pass 2: scatter
  for (int k = lane; k < dim; k += 32) {
    const float hn = p.h0[k] * inv_h, tn = p.t0[k] * inv_t;
    float Gh, Gt;
    if (model == KGE_DISTMULT) {
      Gh = p.r0[k] * tn; Gt = hn * p.r0[k];
      atomicAdd(gr0 + k, g * hn * tn);
    } else if (model == KGE_RESCAL) {
      float a = 0.f, b = 0.f;
      for (int j = 0; j < dim; ++j) {
        a = fmaf(p.r0[(size_t)k * dim + j], p.t0[j] * inv_t, a);
        b = fmaf(p.h0[j] * inv_h, p.r0[(size_t)j * dim + k], b);
        atomicAdd(gr0 + (size_t)k * dim + j, g * hn * (p.t0[j] * inv_t));  // dM_kj = h_k t_j
      }
      Gh = a; Gt = b;
    } else {
      const float x = hn + p.r0[k] - tn;
      const float dx = model == KGE_TRANSE_L2 ? -2.f * x : (x > 0.f ? -1.f : (x < 0.f ? 1.f : 0.f));
      Gh = dx; Gt = -dx;
      atomicAdd(gr0 + k, g * dx);
    }
    atomicAdd(gh0 + k, g * (Gh - hn * dot_h) * inv_h);
    atomicAdd(gt0 + k, g * (Gt - tn * dot_t) * inv_t);
  }
}

// ------------------------------------------------------------------------------------------
__global__ void score_triples_fwd_kernel(int model, int dim, TrainTables tb,
                                         const int64_t* __restrict__ h,
                                         const int64_t* __restrict__ t,
                                         const int64_t* __restrict__ r, long long n,
                                         float* __restrict__ out) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const RowPtrs p = make_rows(model, dim, tb, h[w], t[w], r[w]);
  const float s = triple_score(model, dim, p, lane, nullptr, nullptr);
  if (lane == 0) out[w] = s;
}

__global__ void score_triples_bwd_kernel(int model, int dim, TrainTables tb, TrainGrads gr,
                                         const int64_t* __restrict__ h,
                                         const int64_t* __restrict__ t,
                                         const int64_t* __restrict__ r, long long n,
                                         const float* __restrict__ gout) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long hi = h[w], ti = t[w], ri = r[w];
  const RowPtrs p = make_rows(model, dim, tb, hi, ti, ri);
  triple_backward(model, dim, p, gr, hi, ti, ri, gout[w], lane);
}

__global__ void corrupt_batch_kernel(const int64_t* __restrict__ h, const int64_t* __restrict__ t,
                                     const int64_t* __restrict__ r, long long b, int n_neg,
                                     const float* __restrict__ probs, long long n_ent,
                                     uint64_t seed, uint64_t offset, int64_t* __restrict__ nh,
                                     int64_t* __restrict__ nt) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= b * n_neg) return;
  const long long i = gid % b;  // layout: n_neg blocks of the batch (sampling.py:313-314)
  long long a, c;
  corrupt_one(seed, offset, (uint64_t)gid, probs[r[i]], n_ent, h[i], t[i], &a, &c);
  nh[gid] = a; nt[gid] = c;
}

// Fused step, forward: one warp per positive triple.
//   loss += sum_j max(0, margin - pos_i + neg_ij)         (MarginRankingLoss, target +1, sum)
// Negatives come from (nh, nt) if given, else from Philox; optionally written out.
__global__ void margin_step_fwd_kernel(MarginStepParams a) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= a.b) return;
  const long long hi = a.h[w], ti = a.t[w], ri = a.r[w];
  const RowPtrs pp = make_rows(a.model, a.dim, a.tb, hi, ti, ri);
  const float pos = triple_score(a.model, a.dim, pp, lane, nullptr, nullptr);
  if (lane == 0 && a.pos_out) a.pos_out[w] = pos;
  const float p_head = a.nh ? 0.f : a.probs[ri];
  float loss = 0.f;
  for (int j = 0; j < a.n_neg; ++j) {
    const long long idx = (long long)j * a.b + w;
    long long nh, nt;
    if (a.nh) { nh = a.nh[idx]; nt = a.nt[idx]; }
    else corrupt_one(a.seed, a.offset, (uint64_t)idx, p_head, a.n_ent, hi, ti, &nh, &nt);
    const RowPtrs pn = make_rows(a.model, a.dim, a.tb, nh, nt, ri);
    const float neg = triple_score(a.model, a.dim, pn, lane, nullptr, nullptr);
    if (lane == 0) {
      if (a.neg_out) a.neg_out[idx] = neg;
      if (a.nh_out) { a.nh_out[idx] = nh; a.nt_out[idx] = nt; }
      loss += fmaxf(0.f, a.margin - pos + neg);
    }
  }
  if (lane == 0) atomicAdd(a.loss, loss);
}

// Fused step, backward: recompute the same negatives and scores; every active hinge term
// sends -g to the positive triple and +g to the negative one.
__global__ void margin_step_bwd_kernel(MarginStepParams a, TrainGrads gr, const float* gloss) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= a.b) return;
  const float g = *gloss;
  const long long hi = a.h[w], ti = a.t[w], ri = a.r[w];
  const RowPtrs pp = make_rows(a.model, a.dim, a.tb, hi, ti, ri);
  const float pos = triple_score(a.model, a.dim, pp, lane, nullptr, nullptr);
  const float p_head = a.nh ? 0.f : a.probs[ri];
  int active = 0;
  for (int j = 0; j < a.n_neg; ++j) {
    const long long idx = (long long)j * a.b + w;
    long long nh, nt;
    if (a.nh) { nh = a.nh[idx]; nt = a.nt[idx]; }
    else corrupt_one(a.seed, a.offset, (uint64_t)idx, p_head, a.n_ent, hi, ti, &nh, &nt);
    const RowPtrs pn = make_rows(a.model, a.dim, a.tb, nh, nt, ri);
    const float neg = triple_score(a.model, a.dim, pn, lane, nullptr, nullptr);
    if (a.margin - pos + neg > 0.f) {  // same sub-gradient as torch: zero at the kink
      ++active;
      triple_backward(a.model, a.dim, pn, gr, nh, nt, ri, g, lane);
    }
  }
  if (active) triple_backward(a.model, a.dim, pp, gr, hi, ti, ri, -g * (float)active, lane);
}

// ------------------------------------------------------------------------------------------
// Fast fused step for the single-plane normalising models with a closed form per negative
// (TransE-L1 / TransE-L2 / DistMult), dim % 4 == 0, dim <= 256.
//
// One warp per positive triple.  Lane l owns the 16-byte chunks l and l + 32 of every row, so a
// row is two coalesced LDG.128 per lane (and two float4 atomics on the way back).  The positive's
// h, t, r rows are read ONCE and kept in registers as
//     A  = what a tail-corrupted negative is scored against  (DistMult: hn*r,  TransE: hn + r)
//     Bv = what a head-corrupted negative is scored against  (DistMult: r*tn,  TransE: tn - r)
// so each negative costs exactly one random row: its squared norm and its product with A / Bv
// come out of ONE two-value shuffle tree (TransE-L2 expands |P - en|^2 = |P|^2 - 2 P.en + |en|^2).
// Backward: the corrupted row's gradient is scattered immediately (the only unavoidable
// read-modify-write, 4*dim bytes per negative); the gradients of the intact entity and of the
// relation are linear in  V = sum over active negatives of en  (TransE-L1: of sign(P - en)),
// which stays in registers, so the positive's three rows are written once per positive instead of
// once per negative (the relation rows are shared by thousands of triples: 256x fewer atomics on
// the hottest addresses).  Philox draws are made 32 negatives at a time, one per lane.
// Algorithmic traffic per positive: forward (n_neg + 3) * 4 dim bytes, backward the same rows
// again (recomputed, not stored) plus one RMW of each.
// ------------------------------------------------------------------------------------------
constexpr int FAST_NCH = 2;         // float4 chunks per lane
constexpr int FAST_MAX_DIM = 4 * 32 * FAST_NCH;

struct Vec { float4 c[FAST_NCH]; };

__device__ __forceinline__ Vec vec_load(const float* __restrict__ row, int dim, int lane) {
  Vec v;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i) {
    const int k = 4 * (lane + 32 * i);
    v.c[i] = k < dim ? __ldg(reinterpret_cast<const float4*>(row + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return v;
}
__device__ __forceinline__ void vec_atomic_add(float* __restrict__ row, int dim, int lane, const Vec& v) {
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i) {
    const int k = 4 * (lane + 32 * i);
    if (k < dim) atomicAdd(reinterpret_cast<float4*>(row + k), v.c[i]);
  }
}
template <class F>
__device__ __forceinline__ Vec vec_map(const Vec& a, const Vec& b, F f) {
  Vec o;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i)
    o.c[i] = make_float4(f(a.c[i].x, b.c[i].x), f(a.c[i].y, b.c[i].y), f(a.c[i].z, b.c[i].z),
                         f(a.c[i].w, b.c[i].w));
  return o;
}
template <class F>
__device__ __forceinline__ Vec vec_map3(const Vec& a, const Vec& b, const Vec& c, F f) {
  Vec o;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i)
    o.c[i] = make_float4(f(a.c[i].x, b.c[i].x, c.c[i].x), f(a.c[i].y, b.c[i].y, c.c[i].y),
                         f(a.c[i].z, b.c[i].z, c.c[i].z), f(a.c[i].w, b.c[i].w, c.c[i].w));
  return o;
}
__device__ __forceinline__ Vec vec_scale(const Vec& a, float s) {
  return vec_map(a, a, [s](float x, float) { return x * s; });
}
__device__ __forceinline__ float vec_dot(const Vec& a, const Vec& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i) {
    s = fmaf(a.c[i].x, b.c[i].x, s); s = fmaf(a.c[i].y, b.c[i].y, s);
    s = fmaf(a.c[i].z, b.c[i].z, s); s = fmaf(a.c[i].w, b.c[i].w, s);
  }
  return s;
}
__device__ __forceinline__ float vec_l1_diff(const Vec& a, const Vec& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i)
    s += fabsf(a.c[i].x - b.c[i].x) + fabsf(a.c[i].y - b.c[i].y) + fabsf(a.c[i].z - b.c[i].z) +
         fabsf(a.c[i].w - b.c[i].w);
  return s;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}
__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

template <int MODEL, bool BWD>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 4)
margin_step_fast_kernel(MarginStepParams a, TrainGrads gr, const float* __restrict__ gloss) {
  constexpr int PF = BWD ? 2 : 4;  // negatives whose rows are in flight together, per warp
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= a.b) return;
  const int dim = a.dim;
  const float* __restrict__ ent = a.tb.ent0;
  const long long hi = a.h[w], ti = a.t[w], ri = a.r[w];
  const Vec h = vec_load(ent + (size_t)hi * dim, dim, lane);
  const Vec t = vec_load(ent + (size_t)ti * dim, dim, lane);
  const Vec r = vec_load(a.tb.rel0 + (size_t)ri * dim, dim, lane);
  float sh = vec_dot(h, h), stt = vec_dot(t, t);
  warp_sum2(sh, stt);
  const float inv_h = 1.0f / fmaxf(sqrtf(sh), NORM_EPS), inv_t = 1.0f / fmaxf(sqrtf(stt), NORM_EPS);
  const Vec hn = vec_scale(h, inv_h), tn = vec_scale(t, inv_t);
  Vec A, Bv;
  if constexpr (MODEL == KGE_DISTMULT) {
    A = vec_map(hn, r, [](float x, float y) { return x * y; });
    Bv = vec_map(r, tn, [](float x, float y) { return x * y; });
  } else {
    A = vec_map(hn, r, [](float x, float y) { return x + y; });
    Bv = vec_map(tn, r, [](float x, float y) { return x - y; });
  }
  float sA = 0.f, sB = 0.f, pos;
  if constexpr (MODEL == KGE_DISTMULT) {
    pos = warp_sum(vec_dot(A, tn));
  } else if constexpr (MODEL == KGE_TRANSE_L2) {
    sA = vec_dot(A, A); sB = vec_dot(Bv, Bv);
    warp_sum2(sA, sB);
    const Vec x = vec_map(A, tn, [](float p, float q) { return p - q; });
    pos = -warp_sum(vec_dot(x, x));
  } else {
    pos = -warp_sum(vec_l1_diff(A, tn));
  }
  if (lane == 0 && a.pos_out && !BWD) a.pos_out[w] = pos;
  const float p_head = a.nh ? 0.f : a.probs[ri];
  const float g = BWD ? *gloss : 0.f;
  float loss = 0.f;
  Vec Vt, Vh;  // sums over the active tail- / head-corrupted negatives (BWD only)
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i) Vt.c[i] = Vh.c[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int n_t = 0, n_h = 0;  // active negatives per kind
  for (int j0 = 0; j0 < a.n_neg; j0 += 32) {
    // this lane's draw for negative j0 + lane
    long long my_nh = hi, my_nt = ti;
    if (j0 + lane < a.n_neg) {
      const long long idx = (long long)(j0 + lane) * a.b + w;
      if (a.nh) { my_nh = a.nh[idx]; my_nt = a.nt[idx]; }
      else corrupt_one(a.seed, a.offset, (uint64_t)idx, p_head, a.n_ent, hi, ti, &my_nh, &my_nt);
      if (!BWD && a.nh_out) { a.nh_out[idx] = my_nh; a.nt_out[idx] = my_nt; }
    }
    const int jn = min(32, a.n_neg - j0);
    // PF negatives at a time: their rows are requested together (memory-level parallelism:
    // PF x 4 dim bytes in flight per warp), then reduced one after the other
    for (int jj0 = 0; jj0 < jn; jj0 += PF) {
     long long nhs[PF], nts[PF];
     Vec evs[PF];
#pragma unroll
     for (int u = 0; u < PF; ++u) {
       const int jj = min(jj0 + u, 31);
       nhs[u] = __shfl_sync(0xffffffffu, my_nh, jj);
       nts[u] = __shfl_sync(0xffffffffu, my_nt, jj);
       const bool both = nhs[u] != hi && nts[u] != ti;
       const long long e_ = nhs[u] != hi ? nhs[u] : nts[u];
       if (jj0 + u < jn && !both) evs[u] = vec_load(ent + (size_t)e_ * dim, dim, lane);
     }
#pragma unroll
     for (int u = 0; u < PF; ++u) {
      if (jj0 + u >= jn) break;
      const int jj = jj0 + u;
      const long long nh = nhs[u], nt = nts[u];
      const long long idx = (long long)(j0 + jj) * a.b + w;
      if (nh != hi && nt != ti) {
        // both ends replaced (possible with caller-supplied negatives only): generic path
        const RowPtrs pn = make_rows(MODEL, dim, a.tb, nh, nt, ri);
        const float neg = triple_score(MODEL, dim, pn, lane, nullptr, nullptr);
        const float v = a.margin - pos + neg;
        if (lane == 0) {
          if (!BWD && a.neg_out) a.neg_out[idx] = neg;
          loss += fmaxf(0.f, v);
        }
        if (BWD && v > 0.f) {
          triple_backward(MODEL, dim, pn, gr, nh, nt, ri, g, lane);
          const RowPtrs pp = make_rows(MODEL, dim, a.tb, hi, ti, ri);
          triple_backward(MODEL, dim, pp, gr, hi, ti, ri, -g, lane);
        }
        continue;
      }
      const bool head = nh != hi;            // warp-uniform
      const long long e = head ? nh : nt;
      const Vec ev = evs[u];
      const Vec& P = head ? Bv : A;
      float se = vec_dot(ev, ev), sp = vec_dot(ev, P);
      warp_sum2(se, sp);
      const float inv_e = 1.0f / fmaxf(sqrtf(se), NORM_EPS);
      float neg, en_dot_G = 0.f;  // en . (d neg / d en), needed by the normalisation Jacobian
      Vec en;
      if constexpr (MODEL == KGE_DISTMULT) {
        neg = sp * inv_e;
        en_dot_G = neg;
      } else if constexpr (MODEL == KGE_TRANSE_L2) {
        const float sP = head ? sB : sA;
        const float ee = inv_e * inv_e * se, pe = inv_e * sp;
        neg = -(sP - 2.f * pe + ee);
        en_dot_G = 2.f * (pe - ee);
      } else {
        en = vec_scale(ev, inv_e);
        float l1 = vec_l1_diff(P, en), eg = 0.f;
        if (BWD) {
          const Vec sg = vec_map(P, en, [](float p, float q) { return sgn(p - q); });
          eg = vec_dot(en, sg);
        }
        warp_sum2(l1, eg);
        neg = -l1;
        en_dot_G = eg;
      }
      const float v = a.margin - pos + neg;
      if (lane == 0) {
        if (!BWD && a.neg_out) a.neg_out[idx] = neg;
        loss += fmaxf(0.f, v);
      }
      if (BWD && v > 0.f) {  // same sub-gradient as torch: zero at the kink
        if constexpr (MODEL != KGE_TRANSE_L1) en = vec_scale(ev, inv_e);
        // G = d neg / d en ; d neg / d e = (G - en (en . G)) * inv_e
        Vec ge, V;
        const float c = g * inv_e;
        if constexpr (MODEL == KGE_DISTMULT) {
          ge = vec_map(P, en, [=](float p, float q) { return c * (p - q * en_dot_G); });
          V = en;
        } else if constexpr (MODEL == KGE_TRANSE_L2) {
          ge = vec_map(P, en, [=](float p, float q) { return c * (2.f * (p - q) - q * en_dot_G); });
          V = en;
        } else {
          V = vec_map(P, en, [](float p, float q) { return sgn(p - q); });
          ge = vec_map(V, en, [=](float s_, float q) { return c * (s_ - q * en_dot_G); });
        }
        vec_atomic_add(gr.ent0 + (size_t)e * dim, dim, lane, ge);
        if (head) { Vh = vec_map(Vh, V, [](float x, float y) { return x + y; }); ++n_h; }
        else { Vt = vec_map(Vt, V, [](float x, float y) { return x + y; }); ++n_t; }
      }
     }
    }
  }
  if (!BWD) {
    if (lane == 0) atomicAdd(a.loss, loss);
    return;
  }
  if (n_t + n_h == 0) return;
  // Gradients with respect to hn, tn, r: +g per active negative, -g * (n_t + n_h) for the positive.
  const float fn_t = (float)n_t, fn_h = (float)n_h, fn = (float)(n_t + n_h);
  Vec Gh, Gt, Gr;
  if constexpr (MODEL == KGE_DISTMULT) {
    // neg_t = sum A en, A = hn r ;  neg_h = sum en Bv, Bv = r tn ;  pos = sum hn r tn
    Gh = vec_map3(r, Vt, tn, [=](float rr, float vt, float tt) { return g * rr * (vt - fn * tt); });
    Gt = vec_map3(r, Vh, hn, [=](float rr, float vh, float hh) { return g * rr * (vh - fn * hh); });
    const Vec tmp = vec_map3(hn, Vt, tn, [=](float hh, float vt, float tt) { return hh * (vt - fn * tt); });
    Gr = vec_map3(tmp, tn, Vh, [=](float x, float tt, float vh) { return g * (x + tt * vh); });
  } else if constexpr (MODEL == KGE_TRANSE_L2) {
    // neg_t = -|A - en|^2 ; neg_h = -|en - Bv|^2 ; pos = -|x|^2, x = A - tn
    const Vec x = vec_map(A, tn, [](float p, float q) { return p - q; });
    const Vec dt = vec_map3(A, Vt, x, [=](float aa, float vt, float xx) {  // sum_t (A - en) - n x
      return fn_t * aa - vt - fn * xx; });
    const Vec dh = vec_map(Vh, Bv, [=](float vh, float bb) { return vh - fn_h * bb; });  // sum_h (en - Bv)
    Gh = vec_scale(dt, -2.f * g);
    Gt = vec_map3(dh, x, x, [=](float d, float xx, float) { return 2.f * g * (d - fn * xx); });
    Gr = vec_map(dt, dh, [=](float p, float q) { return -2.f * g * (p + q); });
  } else {
    // neg_t = -|A - en|_1 (V = sign(A - en)) ; neg_h = -|Bv - en|_1 (V = sign(Bv - en)) ; pos = -|x|_1
    const Vec sx = vec_map(A, tn, [](float p, float q) { return sgn(p - q); });
    Gh = vec_map(Vt, sx, [=](float vt, float s_) { return g * (fn * s_ - vt); });
    Gt = vec_map(Vh, sx, [=](float vh, float s_) { return g * (-vh - fn * s_); });
    Gr = vec_map3(Vt, Vh, sx, [=](float vt, float vh, float s_) { return g * (vh - vt + fn * s_); });
  }
  float ph = vec_dot(hn, Gh), pt = vec_dot(tn, Gt);
  warp_sum2(ph, pt);
  const Vec gh = vec_map(Gh, hn, [=](float gg, float q) { return (gg - q * ph) * inv_h; });
  const Vec gt = vec_map(Gt, tn, [=](float gg, float q) { return (gg - q * pt) * inv_t; });
  vec_atomic_add(gr.ent0 + (size_t)hi * dim, dim, lane, gh);
  vec_atomic_add(gr.ent0 + (size_t)ti * dim, dim, lane, gt);
  vec_atomic_add(gr.rel0 + (size_t)ri * dim, dim, lane, Gr);
}

// ------------------------------------------------------------------------------------------
// Ring variant of the fast fused step: the corrupted entities' rows travel through a per-warp
// shared-memory ring filled by 1-D bulk async copies (cp.async.bulk, completion on an mbarrier per
// slot) instead of through registers.  The register form keeps PF = 4 (2 backward) rows in flight
// per warp and stalls on them before every reduction -- measured 0.53 of the HBM copy bandwidth
// with 16 resident warps per SM (ncu: warps active 24 %).  Here a warp keeps RING rows in flight
// at all times (8 x 800 B = 6.4 KB), independently of its register budget, the Philox draws of all
// negatives are made up front into shared memory, and row j + RING is requested the moment row j
// has been consumed.  Arithmetic per negative is exactly that of margin_step_fast_kernel.
// Shared memory per warp: RING * row_bytes + 4 * n_neg (codes) + RING barriers.
// ------------------------------------------------------------------------------------------
constexpr int RING = 8;
constexpr unsigned CODE_BOTH = 0xFFFFFFFFu;   // caller-supplied negative with both ends replaced

__device__ __forceinline__ Vec vec_load_smem(const float* row, int dim, int lane) {
  Vec v;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i) {
    const int k = 4 * (lane + 32 * i);
    v.c[i] = k < dim ? *reinterpret_cast<const float4*>(row + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return v;
}

// MINB: minimum resident CTAs per SM the register allocation is held to (0: the compiler's choice --
// 72 registers forward, 128 backward; 5 holds the backward form to 96 registers, 20 warps per SM)
template <int MODEL, bool BWD, int MINB = 0>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB == 0 ? 1 : MINB)
margin_step_ring_kernel(MarginStepParams a, TrainGrads gr, const float* __restrict__ gloss) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * WARPS_PER_BLOCK + warp_in_block;
  const int dim = a.dim;
  const unsigned row_bytes = (unsigned)dim * 4u;                 // dim % 4 == 0: a multiple of 16
  const int n_codes = (a.n_neg + 3) & ~3;
  const size_t per_warp = (size_t)RING * row_bytes + (size_t)n_codes * 4 + RING * sizeof(uint64_t);
  unsigned char* base = ring_smem + (size_t)warp_in_block * ((per_warp + 127) & ~(size_t)127);
  float* ring = reinterpret_cast<float*>(base);
  unsigned* codes = reinterpret_cast<unsigned*>(base + (size_t)RING * row_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)RING * row_bytes + (size_t)n_codes * 4);
  if (w >= a.b) return;          // whole warps only: no block-wide barrier below
  if (lane == 0) {
#pragma unroll
    for (int s_ = 0; s_ < RING; ++s_) ptx::mbar_init(&bars[s_], 1);
    ptx::fence_mbar_init();
  }
  const float* __restrict__ ent = a.tb.ent0;
  const long long hi = a.h[w], ti = a.t[w], ri = a.r[w];
  const float p_head = a.nh ? 0.f : a.probs[ri];
  // ---- all corruptions of this positive, up front: code = entity | head flag ----
  for (int j0 = 0; j0 < a.n_neg; j0 += 32) {
    const int j = j0 + lane;
    if (j < a.n_neg) {
      const long long idx = (long long)j * a.b + w;
      long long nh = hi, nt = ti;
      if (a.nh) { nh = a.nh[idx]; nt = a.nt[idx]; }
      else corrupt_one(a.seed, a.offset, (uint64_t)idx, p_head, a.n_ent, hi, ti, &nh, &nt);
      if (!BWD && a.nh_out) { a.nh_out[idx] = nh; a.nt_out[idx] = nt; }
      const bool head = nh != hi;
      codes[j] = (nh != hi && nt != ti) ? CODE_BOTH : ((unsigned)(head ? nh : nt) | (head ? 0x80000000u : 0u));
    }
  }
  __syncwarp();
  auto request = [&](int j) {      // lane 0: start the copy of negative j's row into its slot
    const unsigned code = codes[j];
    if (code == CODE_BOTH) return;
    const int slot = j % RING;
    ptx::mbar_arrive_expect_tx(&bars[slot], row_bytes);
    ptx::bulk_g2s(ring + (size_t)slot * dim, ent + (size_t)(code & 0x7FFFFFFFu) * dim, row_bytes, &bars[slot]);
  };
  if (lane == 0) {
    const int first = a.n_neg < RING ? a.n_neg : RING;
    for (int j = 0; j < first; ++j) request(j);
  }
  // ---- the positive (its three rows come straight from global memory, once) ----
  const Vec h = vec_load(ent + (size_t)hi * dim, dim, lane);
  const Vec t = vec_load(ent + (size_t)ti * dim, dim, lane);
  const Vec r = vec_load(a.tb.rel0 + (size_t)ri * dim, dim, lane);
  float sh = vec_dot(h, h), stt = vec_dot(t, t);
  warp_sum2(sh, stt);
  const float inv_h = 1.0f / fmaxf(sqrtf(sh), NORM_EPS), inv_t = 1.0f / fmaxf(sqrtf(stt), NORM_EPS);
  const Vec hn = vec_scale(h, inv_h), tn = vec_scale(t, inv_t);
  Vec A, Bv;
  if constexpr (MODEL == KGE_DISTMULT) {
    A = vec_map(hn, r, [](float x, float y) { return x * y; });
    Bv = vec_map(r, tn, [](float x, float y) { return x * y; });
  } else {
    A = vec_map(hn, r, [](float x, float y) { return x + y; });
    Bv = vec_map(tn, r, [](float x, float y) { return x - y; });
  }
  float sA = 0.f, sB = 0.f, pos;
  if constexpr (MODEL == KGE_DISTMULT) {
    pos = warp_sum(vec_dot(A, tn));
  } else if constexpr (MODEL == KGE_TRANSE_L2) {
    sA = vec_dot(A, A); sB = vec_dot(Bv, Bv);
    warp_sum2(sA, sB);
    const Vec x = vec_map(A, tn, [](float p, float q) { return p - q; });
    pos = -warp_sum(vec_dot(x, x));
  } else {
    pos = -warp_sum(vec_l1_diff(A, tn));
  }
  if (lane == 0 && a.pos_out && !BWD) a.pos_out[w] = pos;
  const float g = BWD ? *gloss : 0.f;
  float loss = 0.f;
  Vec Vt, Vh;
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i) Vt.c[i] = Vh.c[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int n_t = 0, n_h = 0;
  unsigned phases = 0u;   // bit s = parity of the next completion of slot s (a skipped use does not advance it)
  for (int j = 0; j < a.n_neg; ++j) {
    const unsigned code = codes[j];
    const long long idx = (long long)j * a.b + w;
    if (code == CODE_BOTH) {
      // both ends replaced (possible with caller-supplied negatives only): generic path, no ring slot
      const long long nh = a.nh[idx], nt = a.nt[idx];
      const RowPtrs pn = make_rows(MODEL, dim, a.tb, nh, nt, ri);
      const float neg = triple_score(MODEL, dim, pn, lane, nullptr, nullptr);
      const float v = a.margin - pos + neg;
      if (lane == 0) {
        if (!BWD && a.neg_out) a.neg_out[idx] = neg;
        loss += fmaxf(0.f, v);
      }
      if (BWD && v > 0.f) {
        triple_backward(MODEL, dim, pn, gr, nh, nt, ri, g, lane);
        const RowPtrs pp = make_rows(MODEL, dim, a.tb, hi, ti, ri);
        triple_backward(MODEL, dim, pp, gr, hi, ti, ri, -g, lane);
      }
      if (lane == 0 && j + RING < a.n_neg) request(j + RING);
      continue;
    }
    const int slot = j % RING;
    ptx::mbar_wait(&bars[slot], (phases >> slot) & 1u);
    phases ^= 1u << slot;
    const Vec ev = vec_load_smem(ring + (size_t)slot * dim, dim, lane);
    __syncwarp();                        // every lane has its copy: the slot may be refilled
    if (lane == 0 && j + RING < a.n_neg) {
      ptx::fence_proxy_async();          // generic-proxy reads above, async-proxy write below
      request(j + RING);
    }
    const bool head = (code & 0x80000000u) != 0u;   // warp-uniform
    const long long e = (long long)(code & 0x7FFFFFFFu);
    const Vec& P = head ? Bv : A;
    float se = vec_dot(ev, ev), sp = vec_dot(ev, P);
    warp_sum2(se, sp);
    const float inv_e = 1.0f / fmaxf(sqrtf(se), NORM_EPS);
    float neg, en_dot_G = 0.f;
    Vec en;
    if constexpr (MODEL == KGE_DISTMULT) {
      neg = sp * inv_e;
      en_dot_G = neg;
    } else if constexpr (MODEL == KGE_TRANSE_L2) {
      const float sP = head ? sB : sA;
      const float ee = inv_e * inv_e * se, pe = inv_e * sp;
      neg = -(sP - 2.f * pe + ee);
      en_dot_G = 2.f * (pe - ee);
    } else {
      en = vec_scale(ev, inv_e);
      float l1 = vec_l1_diff(P, en), eg = 0.f;
      if (BWD) {
        const Vec sg = vec_map(P, en, [](float p, float q) { return sgn(p - q); });
        eg = vec_dot(en, sg);
      }
      warp_sum2(l1, eg);
      neg = -l1;
      en_dot_G = eg;
    }
    const float v = a.margin - pos + neg;
    if (lane == 0) {
      if (!BWD && a.neg_out) a.neg_out[idx] = neg;
      loss += fmaxf(0.f, v);
    }
    if (BWD && v > 0.f) {
      if constexpr (MODEL != KGE_TRANSE_L1) en = vec_scale(ev, inv_e);
      Vec ge, V;
      const float c = g * inv_e;
      if constexpr (MODEL == KGE_DISTMULT) {
        ge = vec_map(P, en, [=](float p, float q) { return c * (p - q * en_dot_G); });
        V = en;
      } else if constexpr (MODEL == KGE_TRANSE_L2) {
        ge = vec_map(P, en, [=](float p, float q) { return c * (2.f * (p - q) - q * en_dot_G); });
        V = en;
      } else {
        V = vec_map(P, en, [](float p, float q) { return sgn(p - q); });
        ge = vec_map(V, en, [=](float s_, float q) { return c * (s_ - q * en_dot_G); });
      }
      vec_atomic_add(gr.ent0 + (size_t)e * dim, dim, lane, ge);
      if (head) { Vh = vec_map(Vh, V, [](float x, float y) { return x + y; }); ++n_h; }
      else { Vt = vec_map(Vt, V, [](float x, float y) { return x + y; }); ++n_t; }
    }
  }
  if (!BWD) {
    if (lane == 0) atomicAdd(a.loss, loss);
    return;
  }
  if (n_t + n_h == 0) return;
  const float fn_t = (float)n_t, fn_h = (float)n_h, fn = (float)(n_t + n_h);
  Vec Gh, Gt, Gr;
  if constexpr (MODEL == KGE_DISTMULT) {
    Gh = vec_map3(r, Vt, tn, [=](float rr, float vt, float tt) { return g * rr * (vt - fn * tt); });
    Gt = vec_map3(r, Vh, hn, [=](float rr, float vh, float hh) { return g * rr * (vh - fn * hh); });
    const Vec tmp = vec_map3(hn, Vt, tn, [=](float hh, float vt, float tt) { return hh * (vt - fn * tt); });
    Gr = vec_map3(tmp, tn, Vh, [=](float x, float tt, float vh) { return g * (x + tt * vh); });
  } else if constexpr (MODEL == KGE_TRANSE_L2) {
    const Vec x = vec_map(A, tn, [](float p, float q) { return p - q; });
    const Vec dt = vec_map3(A, Vt, x, [=](float aa, float vt, float xx) { return fn_t * aa - vt - fn * xx; });
    const Vec dh = vec_map(Vh, Bv, [=](float vh, float bb) { return vh - fn_h * bb; });
    Gh = vec_scale(dt, -2.f * g);
    Gt = vec_map3(dh, x, x, [=](float d, float xx, float) { return 2.f * g * (d - fn * xx); });
    Gr = vec_map(dt, dh, [=](float p, float q) { return -2.f * g * (p + q); });
  } else {
    const Vec sx = vec_map(A, tn, [](float p, float q) { return sgn(p - q); });
    Gh = vec_map(Vt, sx, [=](float vt, float s_) { return g * (fn * s_ - vt); });
    Gt = vec_map(Vh, sx, [=](float vh, float s_) { return g * (-vh - fn * s_); });
    Gr = vec_map3(Vt, Vh, sx, [=](float vt, float vh, float s_) { return g * (vh - vt + fn * s_); });
  }
  float ph = vec_dot(hn, Gh), pt = vec_dot(tn, Gt);
  warp_sum2(ph, pt);
  const Vec gh = vec_map(Gh, hn, [=](float gg, float q) { return (gg - q * ph) * inv_h; });
  const Vec gt = vec_map(Gt, tn, [=](float gg, float q) { return (gg - q * pt) * inv_t; });
  vec_atomic_add(gr.ent0 + (size_t)hi * dim, dim, lane, gh);
  vec_atomic_add(gr.ent0 + (size_t)ti * dim, dim, lane, gt);
  vec_atomic_add(gr.rel0 + (size_t)ri * dim, dim, lane, Gr);
}

__host__ inline size_t ring_smem_bytes(const MarginStepParams& a) {
  const size_t per_warp = (size_t)RING * a.dim * 4 + (size_t)((a.n_neg + 3) & ~3) * 4 + RING * sizeof(uint64_t);
  return WARPS_PER_BLOCK * ((per_warp + 127) & ~(size_t)127);
}
// KGE_TRAIN_RING=0 selects the register-resident form (margin_step_fast_kernel)
__host__ inline bool ring_step_ok(const MarginStepParams& a) {
  static const bool enabled = [] { const char* v = getenv("KGE_TRAIN_RING"); return !(v && v[0] == '0'); }();
  return enabled && a.n_neg <= 8192 && a.n_ent < 0x7FFFFFFFll && ring_smem_bytes(a) <= 96 * 1024;
}

template <int MODEL, bool BWD, int MINB>
cudaError_t launch_ring_variant(const MarginStepParams& a, const TrainGrads& gr, const float* gloss, cudaStream_t st) {
  const size_t smem = ring_smem_bytes(a);
  static bool configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || !configured[dev])) {
    e = cudaFuncSetAttribute(margin_step_ring_kernel<MODEL, BWD, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             96 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const unsigned blocks = (unsigned)((a.b + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  margin_step_ring_kernel<MODEL, BWD, MINB><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(a, gr, gloss);
  return cudaGetLastError();
}

// KGE_TRAIN_BWD_BLOCKS=5 holds the backward kernel to 96 registers (5 CTAs = 20 warps per SM)
template <int MODEL, bool BWD>
cudaError_t launch_ring(const MarginStepParams& a, const TrainGrads& gr, const float* gloss, cudaStream_t st) {
  if constexpr (BWD) {
    static const bool tight = [] { const char* v = getenv("KGE_TRAIN_BWD_BLOCKS"); return v && v[0] == '5'; }();
    if (tight) return launch_ring_variant<MODEL, true, 5>(a, gr, gloss, st);
  }
  return launch_ring_variant<MODEL, BWD, 0>(a, gr, gloss, st);
}

__host__ inline bool fast_step_ok(const MarginStepParams& a) {
  return (a.model == KGE_TRANSE_L1 || a.model == KGE_TRANSE_L2 || a.model == KGE_DISTMULT) &&
         a.dim % 4 == 0 && a.dim <= FAST_MAX_DIM;
}

__global__ void margin_loss_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ neg,
                                       long long n, float margin, float* __restrict__ loss) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    s += fmaxf(0.f, margin - pos[i] + neg[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(loss, s);
}

__global__ void margin_loss_bwd_kernel(const float* __restrict__ pos, const float* __restrict__ neg,
                                       long long n, float margin, const float* __restrict__ gloss,
                                       float* __restrict__ gpos, float* __restrict__ gneg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = (margin - pos[i] + neg[i] > 0.f) ? *gloss : 0.f;
  gpos[i] = -g;
  gneg[i] = g;
}

// LogisticLoss / BinaryCrossEntropyLoss (utils/losses.py:47-112), sum-reduced.
//   logistic: softplus(-pos) + softplus(neg), softplus(x) = max(x, 0) + log1p(exp(-|x|))
//   bce     : -max(log(sig(pos)), -100) - max(log(1 - sig(neg)), -100), sig in fp32 as torch does
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void pair_loss_fwd_kernel(int kind, const float* __restrict__ pos, const float* __restrict__ neg,
                                     long long n, float* __restrict__ loss) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (kind == KGE_LOSS_LOGISTIC) {
      s += softplus_f(-pos[i]) + softplus_f(neg[i]);
    } else {
      const float pp = sigmoid_f(pos[i]), pn = sigmoid_f(neg[i]);
      s += -fmaxf(logf(pp), -100.f) - fmaxf(logf(1.0f - pn), -100.f);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(loss, s);
}

__global__ void pair_loss_bwd_kernel(int kind, const float* __restrict__ pos, const float* __restrict__ neg,
                                     long long n, const float* __restrict__ gloss,
                                     float* __restrict__ gpos, float* __restrict__ gneg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = *gloss;
  const float pp = sigmoid_f(pos[i]), pn = sigmoid_f(neg[i]);
  if (kind == KGE_LOSS_LOGISTIC) {
    gpos[i] = g * (pp - 1.0f);   // d/dx log(1 + exp(-x)) = -sig(-x) = sig(x) - 1
    gneg[i] = g * pn;            // d/dx log(1 + exp(x))  = sig(x)
  } else {
    // torch's BCELoss backward: (p - y) / max(p (1 - p), 1e-12), chained with dp/dx = p (1 - p)
    gpos[i] = g * (pp - 1.0f) / fmaxf(pp * (1.0f - pp), 1e-12f) * (pp * (1.0f - pp));
    gneg[i] = g * pn / fmaxf(pn * (1.0f - pn), 1e-12f) * (pn * (1.0f - pn));
  }
}

inline unsigned blocks_for_warps(long long warps) {
  return (unsigned)((warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
}

}  // namespace

cudaError_t launch_score_triples_fwd(int model, int dim, const TrainTables& tb, const int64_t* h,
                                     const int64_t* t, const int64_t* r, int64_t n, float* out,
                                     cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  score_triples_fwd_kernel<<<blocks_for_warps(n), WARPS_PER_BLOCK * 32, 0, st>>>(model, dim, tb, h, t,
                                                                                 r, n, out);
  return cudaGetLastError();
}

cudaError_t launch_score_triples_bwd(int model, int dim, const TrainTables& tb, const TrainGrads& gr,
                                     const int64_t* h, const int64_t* t, const int64_t* r, int64_t n,
                                     const float* gout, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  score_triples_bwd_kernel<<<blocks_for_warps(n), WARPS_PER_BLOCK * 32, 0, st>>>(model, dim, tb, gr, h,
                                                                                 t, r, n, gout);
  return cudaGetLastError();
}

cudaError_t launch_corrupt_batch(const int64_t* h, const int64_t* t, const int64_t* r, int64_t b,
                                 int n_neg, const float* probs, int64_t n_ent, uint64_t seed,
                                 uint64_t offset, int64_t* nh, int64_t* nt, cudaStream_t st) {
  const long long n = (long long)b * n_neg;
  if (n <= 0) return cudaSuccess;
  corrupt_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h, t, r, b, n_neg, probs, n_ent,
                                                                     seed, offset, nh, nt);
  return cudaGetLastError();
}

cudaError_t launch_margin_step_fwd(const MarginStepParams& a, cudaStream_t st) {
  if (a.b <= 0) return cudaSuccess;
  if (fast_step_ok(a) && ring_step_ok(a)) {
    const TrainGrads none{nullptr, nullptr, nullptr, nullptr};
    switch (a.model) {
      case KGE_TRANSE_L1: return launch_ring<KGE_TRANSE_L1, false>(a, none, nullptr, st);
      case KGE_TRANSE_L2: return launch_ring<KGE_TRANSE_L2, false>(a, none, nullptr, st);
      default: return launch_ring<KGE_DISTMULT, false>(a, none, nullptr, st);
    }
  }
  if (fast_step_ok(a)) {
    const TrainGrads none{nullptr, nullptr, nullptr, nullptr};
    const unsigned blocks = blocks_for_warps(a.b);
    switch (a.model) {
      case KGE_TRANSE_L1: margin_step_fast_kernel<KGE_TRANSE_L1, false><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(a, none, nullptr); break;
      case KGE_TRANSE_L2: margin_step_fast_kernel<KGE_TRANSE_L2, false><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(a, none, nullptr); break;
      default: margin_step_fast_kernel<KGE_DISTMULT, false><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(a, none, nullptr); break;
    }
    return cudaGetLastError();
  }
  margin_step_fwd_kernel<<<blocks_for_warps(a.b), WARPS_PER_BLOCK * 32, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_margin_step_bwd(const MarginStepParams& a, const TrainGrads& gr, const float* gloss,
                                   cudaStream_t st) {
  if (a.b <= 0) return cudaSuccess;
  if (fast_step_ok(a) && ring_step_ok(a)) {
    switch (a.model) {
      case KGE_TRANSE_L1: return launch_ring<KGE_TRANSE_L1, true>(a, gr, gloss, st);
      case KGE_TRANSE_L2: return launch_ring<KGE_TRANSE_L2, true>(a, gr, gloss, st);
      default: return launch_ring<KGE_DISTMULT, true>(a, gr, gloss, st);
    }
  }
  if (fast_step_ok(a)) {
    const unsigned blocks = blocks_for_warps(a.b);
    switch (a.model) {
      case KGE_TRANSE_L1: margin_step_fast_kernel<KGE_TRANSE_L1, true><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(a, gr, gloss); break;
      case KGE_TRANSE_L2: margin_step_fast_kernel<KGE_TRANSE_L2, true><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(a, gr, gloss); break;
      default: margin_step_fast_kernel<KGE_DISTMULT, true><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(a, gr, gloss); break;
    }
    return cudaGetLastError();
  }
  margin_step_bwd_kernel<<<blocks_for_warps(a.b), WARPS_PER_BLOCK * 32, 0, st>>>(a, gr, gloss);
  return cudaGetLastError();
}

cudaError_t launch_margin_loss_fwd(const float* pos, const float* neg, int64_t n, float margin,
                                   float* loss, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  margin_loss_fwd_kernel<<<blocks, 256, 0, st>>>(pos, neg, n, margin, loss);
  return cudaGetLastError();
}

cudaError_t launch_pair_loss_fwd(int kind, const float* pos, const float* neg, int64_t n, float* loss,
                                 cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  pair_loss_fwd_kernel<<<blocks, 256, 0, st>>>(kind, pos, neg, n, loss);
  return cudaGetLastError();
}

cudaError_t launch_pair_loss_bwd(int kind, const float* pos, const float* neg, int64_t n,
                                 const float* gloss, float* gpos, float* gneg, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  pair_loss_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kind, pos, neg, n, gloss, gpos, gneg);
  return cudaGetLastError();
}

cudaError_t launch_margin_loss_bwd(const float* pos, const float* neg, int64_t n, float margin,
                                   const float* gloss, float* gpos, float* gneg, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  margin_loss_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, neg, n, margin, gloss,
                                                                      gpos, gneg);
  return cudaGetLastError();
}

}  // namespace kge
