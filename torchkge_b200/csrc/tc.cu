// Tensor-core bound-and-refine for the rank scan (tcgen05 / TMEM, sm_100a).
//
// For the models whose score is a dot product or a squared L2 distance (DistMult, RESCAL,
// ComplEx, TransE-L2) the count  #{c : s(q,c) >= s_true(q)}  does not need every score to the
// last bit: it needs every score to be on the right SIDE of s_true.  This kernel computes an
// approximation s~(q,c) on the 5th-generation tensor cores -- fp32 operands split into bf16
// (hi, lo) pairs, three bf16 products per term (hi*hi + lo*hi + hi*lo), fp32 accumulation in
// TMEM -- together with a rigorous bound eps(q,c) >= |s~ - s_ATen| (tc_bound.h).  Pairs with
// |s~ - s_true| > eps are decided from s~; the few others (the "near-tie band", ~0.1 % of the
// pairs) are appended to a list and re-scored exactly, with the ATen-order schedule replay, by
// recheck_kernel.  Ranks therefore stay bit-identical to the reference's while ~99.9 % of the
// arithmetic moves from the fp32 pipes to the tensor cores.
//
// Kernel shape: persistent CTAs, 6 warps.  Warp 0 (one lane): producer -- 1-D bulk async
// copies (UBLKCP) of pre-swizzled operand images (the pack kernels write the exact
// shared-memory image of the 128-byte-swizzled K-major tiles, so no tensor map is needed)
// into a 2-stage ring.  Warp 1 (one lane): issues tcgen05.mma (M=128 queries x N=256
// candidates x K=16, bf16 -> f32) into one of two 256-column TMEM accumulators and commits to
// mbarriers.  Warps 2-5: epilogue -- tcgen05.ld the accumulator (one query row per thread),
// apply norms / bound / threshold, count, append near-ties.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc.h"

namespace kge {
namespace tc {

namespace {

constexpr int BM = TC_BM, BN = TC_BN, BK = TC_BK;
constexpr int STAGES = 2;
constexpr int A_PLANE = BM * 128;  // bytes of one 128-row x 64-bf16 swizzled plane
constexpr int B_PLANE = BN * 128;
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;  // A hi, A lo, B hi, B lo = 96 KB
constexpr int NORM_FLOATS = 2 * BN;                     // cb[256], cn[256] per buffer
constexpr int AMB_BUF = 128;                            // near-tie entries buffered per epilogue warp
constexpr int EPI_WARPS = 8;                            // two per TMEM lane quadrant, half the columns each
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES +
                              2 * NORM_FLOATS * sizeof(float) + 16 * sizeof(uint64_t) + 16 +
                              EPI_WARPS * AMB_BUF * sizeof(int2);

// -------- PTX helpers specific to tcgen05 --------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   ptx::smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> f32, both operands K-major
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO = 1 (ignored for swizzled K-major) | SBO = 1024 B (8 rows x 128 B)
// | version 1 (sm_100) | layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, K-major both, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Work order: a unit = (group of TC_CT_GROUP consecutive candidate tiles, query tile); a CTA
// takes units round-robin and walks the group's candidate tiles for that query tile.  CTAs
// running together work on the same group with different query tiles, so the group's B images
// are served from L2 (B streams from HBM once per launch) and the whole A image (tens of MB)
// stays L2-resident; per-query counters are flushed once per unit.
struct Units {
  long long n_qt, n_ct, n_groups, n_units;
  __device__ Units(long long nq, long long nc)
      : n_qt(nq), n_ct(nc), n_groups((nc + TC_CT_GROUP - 1) / TC_CT_GROUP),
        n_units(nq * ((nc + TC_CT_GROUP - 1) / TC_CT_GROUP)) {}
  __device__ void decode(long long u, long long* qt, long long* ct_lo, long long* ct_hi) const {
    const long long g = u / n_qt;
    *qt = u - g * n_qt;
    *ct_lo = g * TC_CT_GROUP;
    *ct_hi = min(n_ct, *ct_lo + TC_CT_GROUP);
  }
};

template <bool L2, bool DUMP>
__global__ void __launch_bounds__(THREADS, 1) tc_scan_kernel(const __grid_constant__ TcScanParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-B alignment for the 128-byte swizzle atoms
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
  unsigned char* stage_base = smem;
  float* s_norm = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_norm + 2 * NORM_FLOATS);
  uint64_t* full_bar = bars;            // [STAGES]
  uint64_t* empty_bar = bars + 2;       // [STAGES]
  uint64_t* tfull_bar = bars + 4;       // [2]
  uint64_t* tempty_bar = bars + 6;      // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8);
  int2* s_amb = reinterpret_cast<int2*>(bars + 10);  // [EPI_WARPS][AMB_BUF]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Units units(p.n_qt, p.n_ct);
  const int n_kb = p.n_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], EPI_WARPS); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long u = blockIdx.x; u < units.n_units; u += gridDim.x) {
        long long qt, ct_lo, ct_hi; units.decode(u, &qt, &ct_lo, &ct_hi);
        const unsigned char* asrc = p.apack + (size_t)qt * n_kb * (2 * A_PLANE);
        for (long long ct = ct_lo; ct < ct_hi; ++ct) {
          const unsigned char* bsrc = p.bpack + (size_t)ct * n_kb * (2 * B_PLANE);
          for (int kb = 0; kb < n_kb; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            unsigned char* sa = stage_base + (size_t)stage * STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            ptx::bulk_g2s(sa, asrc + (size_t)kb * (2 * A_PLANE), 2 * A_PLANE, &full_bar[stage]);
            ptx::bulk_g2s(sa + 2 * A_PLANE, bsrc + (size_t)kb * (2 * B_PLANE), 2 * B_PLANE,
                          &full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (long long u = blockIdx.x; u < units.n_units; u += gridDim.x) {
       long long qt_, ct_lo, ct_hi; units.decode(u, &qt_, &ct_lo, &ct_hi);
       for (long long ct = ct_lo; ct < ct_hi; ++ct) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          fence_after();
          const uint32_t sa = ptx::smem_u32(stage_base + (size_t)stage * STAGE_BYTES);
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_PLANE);
          const uint64_t b_hi = make_desc(sa + 2 * A_PLANE), b_lo = make_desc(sa + 2 * A_PLANE + B_PLANE);
          const int k16s = min(BK / 16, (p.k_total - kb * BK + 15) / 16);
          for (int k = 0; k < k16s; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 B = 2 x 16-B units
            umma_bf16(d_tmem, a_hi + adv, b_hi + adv, IDESC, (kb | k) ? 1u : 0u);
            umma_bf16(d_tmem, a_lo + adv, b_hi + adv, IDESC, 1u);
            umma_bf16(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
       }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // query row within the tile
    const int epi_tid = (warp - 2) * 32 + lane;
    const int col_half = (warp - 2) >> 2;     // which 128 columns of the tile this warp handles
    int acc = 0; uint32_t acc_phase = 0;
    long long cur_qt = -1;
    float st = 0.f, qb = 0.f, qn = 0.f;
    int cnt = 0;
    int nbuf = 0;
    int2* wbuf = s_amb + (warp - 2) * AMB_BUF;  // this warp's near-tie buffer
    int amb_n = 0;                              // entries in it (warp-uniform)
    for (long long u = blockIdx.x; u < units.n_units; u += gridDim.x) {
     long long qt, ct_lo, ct_hi; units.decode(u, &qt, &ct_lo, &ct_hi);
     for (long long ct = ct_lo; ct < ct_hi; ++ct) {
      const long long q = qt * BM + row;
      if (qt != cur_qt) {
        if (cur_qt >= 0 && cnt != 0) {
          const long long pq = cur_qt * BM + row;
          if (pq < p.n_q) atomicAdd(&p.counts[pq], cnt);
        }
        cnt = 0; cur_qt = qt;
        const bool vq = q < p.n_q;
        st = vq ? p.s_true[q] : INFINITY;
        qb = vq ? p.qbound[q] : 0.f;
        qn = vq ? p.qnorm2[q] : 0.f;
      }
      // candidate-side vectors of this tile -> shared (double buffered)
      float* cbs = s_norm + nbuf * NORM_FLOATS;
      float* cns = cbs + BN;
      for (int j = epi_tid; j < BN; j += EPI_WARPS * 32) {
        const long long c = ct * BN + j;
        cbs[j] = c < p.n_rows ? p.cbound[c] : 0.f;
        cns[j] = c < p.n_rows ? p.cnorm2[c] : 0.f;
      }
      named_bar_sync(1, EPI_WARPS * 32);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      fence_after();
      const int ncols = (int)min((long long)BN, p.n_rows - ct * BN);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      // per-thread constants of the threshold test  u = s~ - s_true  vs  eps
      const float g_qb = p.gamma * qb;        // dot: eps = g_qb * cb[c]
      const float g2_qb = 2.f * g_qb;         // L2 : eps = 2 gamma qb cb + gamma2 (qb + cb)^2
      const float l2_off = -(qn + st);        // L2 : u = 2 dot + l2_off - cn[c]
      const int c_begin = col_half * (BN / 2), c_end = min(ncols, c_begin + BN / 2);
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        const int lim = min(32, ncols - c0);
        // 32 independent threshold tests -> two bit masks per thread (no per-element branches)
        unsigned amb_mask = 0, gt_mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dot = __uint_as_float(v[j]);
          float u, e;
          if constexpr (L2) {
            u = fmaf(2.f, dot, l2_off) - cns[c0 + j];
            const float cbj = cbs[c0 + j];
            const float w = qb + cbj;
            e = fmaf(g2_qb, cbj, p.gamma2 * w * w);
          } else {
            u = dot - st;
            e = g_qb * cbs[c0 + j];
          }
          if constexpr (DUMP) {
            if (j < lim && q < p.n_q) p.dump[(size_t)q * p.n_rows + ct * BN + c0 + j] = L2 ? u + st : dot;
          } else {
            amb_mask |= (fabsf(u) <= e ? 1u : 0u) << j;
            gt_mask |= (u > e ? 1u : 0u) << j;
          }
        }
        if constexpr (!DUMP) {
          if (lim < 32) {  // columns past the table (last tile only)
            const unsigned keep = (1u << lim) - 1u;
            amb_mask &= keep; gt_mask &= keep;
          }
          cnt += __popc(gt_mask);
          if (__any_sync(0xffffffffu, amb_mask != 0)) {
            // near-ties in this 32 x 32 block: warp prefix sum of the per-lane counts, entries
            // into this warp's shared buffer, one global atomic per ~100 entries on flush
            const int mine = __popc(amb_mask);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int up = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += up;
            }
            const int total_new = __shfl_sync(0xffffffffu, incl, 31);
            if (amb_n + total_new > AMB_BUF) {  // flush first (uniform decision)
              __syncwarp();
              unsigned long long base = 0;
              if (lane == 0) base = atomicAdd(p.amb_count + blockIdx.x, (unsigned long long)amb_n);
              base = __shfl_sync(0xffffffffu, base, 0);
              for (int i = lane; i < amb_n; i += 32)
                if (base + i < p.amb_cap) p.amb_pairs[(size_t)blockIdx.x * p.amb_cap + base + i] = wbuf[i];
              __syncwarp();
              amb_n = 0;
            }
            if (total_new <= AMB_BUF) {
              int slot = amb_n + incl - mine;
              unsigned m = amb_mask;
              while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                wbuf[slot++] = make_int2((int)q, (int)(ct * BN + c0 + j));
              }
              amb_n += total_new;
            } else {
              // pathological block (more near-ties than the buffer holds): straight to global
              unsigned long long base = 0;
              if (lane == 0) base = atomicAdd(p.amb_count + blockIdx.x, (unsigned long long)total_new);
              base = __shfl_sync(0xffffffffu, base, 0) + (unsigned long long)(incl - mine);
              unsigned m = amb_mask;
              while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                if (base < p.amb_cap)
                  p.amb_pairs[(size_t)blockIdx.x * p.amb_cap + base] = make_int2((int)q, (int)(ct * BN + c0 + j));
                ++base;
              }
            }
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      nbuf ^= 1;
     }
    }
    if (cur_qt >= 0 && cnt != 0) {
      const long long pq = cur_qt * BM + row;
      if (pq < p.n_q) atomicAdd(&p.counts[pq], cnt);
    }
    if (amb_n > 0) {  // final flush of this warp's near-tie buffer
      __syncwarp();
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(p.amb_count + blockIdx.x, (unsigned long long)amb_n);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < amb_n; i += 32)
        if (base + i < p.amb_cap) p.amb_pairs[(size_t)blockIdx.x * p.amb_cap + base + i] = wbuf[i];
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------
// Operand packing: fp32 rows -> (hi, lo) bf16 planes in the shared-memory image of K-major,
// 128-byte-swizzled tiles.  One thread per (row, 16-byte chunk): 8 consecutive k of one plane.
//   image offset of (row r, chunk j) inside a plane = (r/8)*1024 + (r%8)*128 + ((j ^ (r%8))*16)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float operand_value(const float* __restrict__ p0,
                                               const float* __restrict__ p1, int dim, int k,
                                               int k_total) {
  if (k >= k_total) return 0.f;
  return k < dim ? p0[k] : p1[k - dim];
}

template <int ROWS>
__global__ void pack_operand_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                    long long row_stride, long long plane1_offset, long long n_rows,
                                    int dim, int k_total, int n_kb, int sub_mode,
                                    unsigned char* __restrict__ out) {
  // src0 + row*row_stride = first plane of the row; second plane at +plane1_offset (same row)
  // or in src1 (separate table).  sub_mode = 1: value = plane1[k] - plane0[k]  (t - r, L2 head)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_tiles = (n_rows + ROWS - 1) / ROWS;
  const long long total = n_tiles * n_kb * ROWS * 8;
  if (gid >= total) return;
  const int j = (int)(gid & 7);
  long long rest = gid >> 3;
  const int r = (int)(rest % ROWS);
  rest /= ROWS;
  const int kb = (int)(rest % n_kb);
  const long long tile = rest / n_kb;
  const long long row = tile * ROWS + r;
  __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kb * BK + j * 8 + e;
    float x = 0.f;
    if (row < n_rows) {
      const float* a = src0 + (size_t)row * row_stride;
      const float* b = src1 ? src1 + (size_t)row * row_stride : a + plane1_offset;
      if (sub_mode) x = k < dim ? __fsub_rn(b[k], a[k]) : 0.f;
      else x = operand_value(a, b, dim, k, k_total);
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[e] = h;
    lo[e] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
  const size_t plane = (size_t)ROWS * 128;
  const size_t base = ((size_t)tile * n_kb + kb) * (2 * plane);
  const size_t off = (size_t)(r / 8) * 1024 + (size_t)(r % 8) * 128 + (size_t)((j ^ (r % 8)) * 16);
  *reinterpret_cast<uint4*>(out + base + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(out + base + plane + off) = *reinterpret_cast<const uint4*>(lo);
}

// per-row |x|_2 (rounded up a little) and |x|_2^2 of the operand vector, one warp per row
__global__ void row_norms_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                 long long row_stride, long long plane1_offset, long long n_rows,
                                 int dim, int k_total, int sub_mode, float* __restrict__ bound,
                                 float* __restrict__ norm2) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const float* a = src0 + (size_t)w * row_stride;
  const float* b = src1 ? src1 + (size_t)w * row_stride : a + plane1_offset;
  double s = 0.0, sa = 0.0, sb = 0.0;
  for (int k = lane; k < k_total; k += 32) {
    float x;
    if (sub_mode) {
      x = k < dim ? __fsub_rn(b[k], a[k]) : 0.f;
      if (k < dim) { sa += (double)a[k] * (double)a[k]; sb += (double)b[k] * (double)b[k]; }
    } else {
      x = operand_value(a, b, dim, k, k_total);
    }
    s += (double)x * (double)x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (lane == 0) {
    norm2[w] = (float)s;
    // L2 head side: the reference rounds c + r before subtracting t, an error that scales with
    // |r| and |t| separately, so the bound uses |t| + |r| (>= |t - r|) for this operand
    const double nb = sub_mode ? sqrt(sa) + sqrt(sb) : sqrt(s);
    bound[w] = (float)(nb * (1.0 + 1e-6)) + 1e-30f;
  }
}

template <int EL>
__global__ void recheck_kernel(int dim, const unsigned long long* __restrict__ region_counts,
                               unsigned long long region_cap, const int2* __restrict__ pairs,
                               const float* __restrict__ qplain, const float* __restrict__ ent0,
                               const float* __restrict__ ent1, const float* __restrict__ s_true,
                               int32_t* __restrict__ counts) {
  constexpr int QW = ElemTraits<EL>::QW, CW = ElemTraits<EL>::CW;
  constexpr bool NORM = ElemTraits<EL>::RED == RED_NORM2;
  constexpr int PAIRS_PER_WARP = NORM ? 4 : 1;
  const unsigned long long region = blockIdx.y;
  const unsigned long long n_pairs = min(region_counts[region], region_cap);
  const int lane = threadIdx.x & 31;
  const unsigned long long warp_global = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  const int2* list = pairs + region * region_cap;
  if (!NORM && dim < 8) {  // one-lane schedule: plain per-thread scorer
    for (unsigned long long i = warp_global * 32 + lane; i < n_pairs; i += n_warps * 32) {
      const int2 pr = list[i];
      const float* q0 = qplain + (size_t)pr.x * QW * dim;
      const float* c0 = ent0 + (size_t)pr.y * dim;
      if (pair_score_natural<EL>(dim, q0, q0 + (size_t)(QW - 1) * dim, c0,
                                 (CW == 2 ? ent1 : ent0) + (size_t)pr.y * dim) >= s_true[pr.x])
        atomicAdd(&counts[pr.x], 1);
    }
    return;
  }
  for (unsigned long long base = warp_global * PAIRS_PER_WARP; base < n_pairs; base += n_warps * PAIRS_PER_WARP) {
    const unsigned long long i = base + (NORM ? (lane >> 3) : 0);
    const bool valid = i < n_pairs;
    const int2 pr = list[valid ? i : base];  // idle groups redo the first pair (shuffles stay uniform)
    const float* q0 = qplain + (size_t)pr.x * QW * dim;
    const float* q1 = q0 + (size_t)(QW - 1) * dim;
    const float* c0 = ent0 + (size_t)pr.y * dim;
    const float* c1 = (CW == 2 ? ent1 : ent0) + (size_t)pr.y * dim;
    const float sc = pair_score_chains<EL>(dim, q0, q1, c0, c1, lane);
    const bool leader = NORM ? ((lane & 7) == 0) : (lane == 0);
    if (valid && leader && sc >= s_true[pr.x]) atomicAdd(&counts[pr.x], 1);
  }
}

// stats[0] = near-tie pairs found (or capacity + 1 if any region overflowed), stats[1] = capacity
__global__ void tc_stats_kernel(const unsigned long long* __restrict__ region_counts, int regions,
                                unsigned long long region_cap, unsigned long long* __restrict__ stats) {
  unsigned long long total = 0;
  bool over = false;
  for (int r = 0; r < regions; ++r) {
    total += region_counts[r];
    over |= region_counts[r] > region_cap;
  }
  const unsigned long long cap = region_cap * (unsigned long long)regions;
  stats[0] = over ? cap + 1 : total;
  stats[1] = cap;
}

}  // namespace

size_t a_image_bytes(long long n_q, int n_kb) {
  const long long n_qt = (n_q + BM - 1) / BM;
  return (size_t)n_qt * n_kb * 2 * A_PLANE;
}
size_t b_image_bytes(long long n_rows, int n_kb) {
  const long long n_ct = (n_rows + BN - 1) / BN;
  return (size_t)n_ct * n_kb * 2 * B_PLANE;
}

cudaError_t launch_pack_b(const float* ent0, const float* ent1, long long n_rows, int dim, int k_total,
                          int n_kb, unsigned char* bpack, float* cbound, float* cnorm2,
                          cudaStream_t st) {
  if (n_rows <= 0) return cudaSuccess;
  const long long n_ct = (n_rows + BN - 1) / BN;
  const long long total = n_ct * n_kb * BN * 8;
  pack_operand_kernel<BN><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      ent0, ent1, dim, 0, n_rows, dim, k_total, n_kb, 0, bpack);
  row_norms_kernel<<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, st>>>(
      ent0, ent1, dim, 0, n_rows, dim, k_total, 0, cbound, cnorm2);
  return cudaGetLastError();
}

cudaError_t launch_pack_a(const float* qplain, int qw, long long n_q, int dim, int k_total, int n_kb,
                          int sub_mode, unsigned char* apack, float* qbound, float* qnorm2,
                          cudaStream_t st) {
  if (n_q <= 0) return cudaSuccess;
  const long long n_qt = (n_q + BM - 1) / BM;
  const long long total = n_qt * n_kb * BM * 8;
  // qplain rows are [qw][dim]: plane 1 (if any) follows plane 0 inside the row
  pack_operand_kernel<BM><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      qplain, nullptr, (long long)qw * dim, dim, n_q, dim, k_total, n_kb, sub_mode, apack);
  row_norms_kernel<<<(unsigned)((n_q * 32 + 255) / 256), 256, 0, st>>>(
      qplain, nullptr, (long long)qw * dim, dim, n_q, dim, k_total, sub_mode, qbound, qnorm2);
  return cudaGetLastError();
}

cudaError_t launch_tc_scan(const TcScanParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kern) {
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    };
    set(tc_scan_kernel<false, false>); set(tc_scan_kernel<true, false>);
    set(tc_scan_kernel<false, true>); set(tc_scan_kernel<true, true>);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = scan_grid_size(p.n_q, p.n_rows);
  if (grid <= 0) return cudaSuccess;
  if (p.dump) {
    if (p.l2) tc_scan_kernel<true, true><<<grid, THREADS, SMEM_BYTES, st>>>(p);
    else tc_scan_kernel<false, true><<<grid, THREADS, SMEM_BYTES, st>>>(p);
  } else {
    if (p.l2) tc_scan_kernel<true, false><<<grid, THREADS, SMEM_BYTES, st>>>(p);
    else tc_scan_kernel<false, false><<<grid, THREADS, SMEM_BYTES, st>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_recheck(int el, int dim, const unsigned long long* region_counts, int regions,
                           unsigned long long region_cap, const int2* pairs, const float* qplain,
                           const float* ent0, const float* ent1, const float* s_true, int32_t* counts,
                           unsigned long long* stats, cudaStream_t st) {
  if (regions <= 0 || region_cap == 0) return cudaSuccess;
  dim3 grid(96, (unsigned)regions);
#define CALL_RC(EL) \
  recheck_kernel<EL><<<grid, 128, 0, st>>>(dim, region_counts, region_cap, pairs, qplain, ent0, ent1, s_true, counts)
  switch (el) {
    case EL_DOT1: CALL_RC(EL_DOT1); break;
    case EL_DOT2: CALL_RC(EL_DOT2); break;
    case EL_L2_TAIL: CALL_RC(EL_L2_TAIL); break;
    case EL_L2_HEAD: CALL_RC(EL_L2_HEAD); break;
    default: return cudaErrorInvalidValue;
  }
#undef CALL_RC
  if (stats) tc_stats_kernel<<<1, 1, 0, st>>>(region_counts, regions, region_cap, stats);
  return cudaGetLastError();
}

int scan_grid_size(long long n_q, long long n_rows) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  const long long n_qt = (n_q + BM - 1) / BM, n_ct = (n_rows + BN - 1) / BN;
  const long long units = n_qt * ((n_ct + TC_CT_GROUP - 1) / TC_CT_GROUP);
  return (int)(units < sms ? units : sms);
}

}  // namespace tc
}  // namespace kge
