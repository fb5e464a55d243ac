// Tensor-core bound-and-refine for the rank scan (tcgen05 / TMEM, sm_100a).
//
// For the models whose score is a dot product or a squared L2 distance (DistMult, RESCAL,
// ComplEx, TransE-L2) the count  #{c : s(q,c) >= s_true(q)}  does not need every score to the
// last bit: it needs every score to be on the right SIDE of s_true.  This kernel computes an
// approximation s~(q,c) on the 5th-generation tensor cores -- fp32 operands split into bf16
// (hi, lo) pairs, three bf16 products per term (hi*hi + lo*hi + hi*lo), fp32 accumulation in
// TMEM -- together with a rigorous bound eps(q,c) >= |s~ - s_ATen| (tc.h).  Pairs with
// |s~ - s_true| > eps are decided from s~; the few others (the "near-tie band", 0.06-0.2 % of the
// pairs) are appended to a list and re-scored exactly, with the ATen-order arithmetic, by
// recheck_kernel.  Ranks therefore stay bit-identical to the reference's while ~99.9 % of the
// arithmetic moves from the fp32 pipes to the tensor cores.
//
// Kernel shape: persistent CTAs of 10 warps, one per SM.  Warp 0 (one lane): producer -- 1-D bulk
// async copies (UBLKCP) of pre-swizzled operand images (the pack kernels write the exact
// shared-memory image of K-major swizzled tiles, so no tensor map is needed): k-blocks of 32 bf16
// (64-byte swizzle) with the query tile's image resident for a whole work unit and the candidate
// image streaming through a 3-stage ring, or both streamed (4 x 48 KB; 2 x 96 KB with 128-byte
// swizzle).  Warp 1 (one lane): issues tcgen05.mma (M=128 queries x N=256 candidates x K=16,
// bf16 -> f32) into one of two 256-column TMEM accumulators and commits to mbarriers.
// Warps 2-9: epilogue -- tcgen05.ld the accumulator (one query row per thread, two warps per TMEM
// lane quadrant), compare with two per-thread thresholds, count, append near-ties.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc.h"

namespace kge {
namespace tc {

namespace {

constexpr int BM = TC_BM, BN = TC_BN;
constexpr int AMB_BUF = 64;          // near-tie entries buffered per epilogue warp
constexpr int EPI_WARPS = 8;         // two per TMEM lane quadrant, half the columns each
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 4;
constexpr int BAR_SLOTS = 32;        // full[4] empty[4] tfull[2] tempty[2] a_full[8] a_empty[1] ...
constexpr size_t SMEM_TAIL = BAR_SLOTS * sizeof(uint64_t) + EPI_WARPS * AMB_BUF * sizeof(int2);
constexpr size_t SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA on sm_100

// Geometry of one kernel variant: BKT bf16 per k-block (= one swizzle span of 2*BKT bytes);
// RES = the query tile's whole A image stays resident in shared memory for a unit and only B
// is streamed (possible when the image is <= 7 k-blocks of 32, i.e. k_total <= 224).
template <int BKT, bool RES>
struct Geo {
  static constexpr int ROWB = 2 * BKT;                 // bytes per operand row in a k-block
  static constexpr int A_PLANE = BM * ROWB;            // one 128-row (hi or lo) plane
  static constexpr int B_PLANE = BN * ROWB;
  static constexpr int A_BLOCK = 2 * A_PLANE;          // hi + lo
  static constexpr int B_BLOCK = 2 * B_PLANE;
  static constexpr int STAGE_BYTES = RES ? B_BLOCK : A_BLOCK + B_BLOCK;
  static constexpr int STAGES = RES ? 3 : (BKT == 64 ? 2 : 4);
  static constexpr int K16 = BKT / 16;                 // MMA k-steps per block
  static constexpr uint64_t LAYOUT = BKT == 64 ? 2 : 4;  // SWIZZLE_128B : SWIZZLE_64B
  static constexpr uint32_t SBO = 8 * ROWB;            // byte stride between 8-row groups
};

template <int BKT, bool RES>
size_t smem_bytes(int n_kb) {
  using G = Geo<BKT, RES>;
  return 1024 /*align slack*/ + (RES ? (size_t)n_kb * G::A_BLOCK : 0) + (size_t)G::STAGES * G::STAGE_BYTES +
         SMEM_TAIL;
}

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
// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// >> 4 | LBO = 1 (ignored for swizzled K-major) | SBO = 8 rows x row bytes | version 1 (sm_100)
// | layout type (2 = SWIZZLE_128B for 128-byte rows, 4 = SWIZZLE_64B for 64-byte rows).
template <class G>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(G::SBO >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= G::LAYOUT << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32 (bits [4,6) = 1), A / B format in bits [7,10) / [10,13)
// (0 = fp16, 1 = bf16), K-major both, N = 256 (bits [17,23) = N >> 3), M = 128 (bits [24,29) = M >> 4)
constexpr uint32_t IDESC_BASE = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_BF16 = IDESC_BASE | (1u << 7) | (1u << 10);
constexpr uint32_t IDESC_FP16 = IDESC_BASE;

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


// Work order: a unit = (group of p.ct_group consecutive candidate tiles, query tile); a CTA
// takes units round-robin and walks the group's candidate tiles for that query tile.  CTAs
// running together work on the same group with different query tiles, so the group's B images
// are served from L2 (B streams from HBM once per launch) and the whole A image (tens of MB)
// stays L2-resident; per-query counters are flushed once per unit.
struct Units {
  long long n_qt, n_ct, grp, n_units;
  __device__ Units(long long nq, long long nc, long long g)
      : n_qt(nq), n_ct(nc), grp(g), n_units(nq * ((nc + g - 1) / g)) {}
  __device__ void decode(long long u, long long* qt, long long* ct_lo, long long* ct_hi) const {
    const long long g = u / n_qt;
    *qt = u - g * n_qt;
    *ct_lo = g * grp;
    *ct_hi = min(n_ct, *ct_lo + grp);
  }
};

template <bool L2, bool DUMP, int BKT, bool RES>
__global__ void __launch_bounds__(THREADS, 1) tc_scan_kernel(const __grid_constant__ TcScanParams p) {
  using G = Geo<BKT, RES>;
  constexpr int STAGES = G::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // 1024-B alignment for the swizzle atoms
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
  const int n_kb = p.n_kb;
  unsigned char* a_res = smem;  // [n_kb][hi, lo][128 rows]   (RES only)
  unsigned char* stage_base = smem + (RES ? (size_t)n_kb * G::A_BLOCK : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)STAGES * G::STAGE_BYTES);
  uint64_t* full_bar = bars;             // [MAX_STAGES]
  uint64_t* empty_bar = bars + 4;        // [MAX_STAGES]
  uint64_t* tfull_bar = bars + 8;        // [2]
  uint64_t* tempty_bar = bars + 10;      // [2]
  uint64_t* afull_bar = bars + 12;       // [8]  one per resident A k-block
  uint64_t* aempty_bar = bars + 20;      // [1]  resident A region free again
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 30);
  int2* s_amb = reinterpret_cast<int2*>(bars + BAR_SLOTS);  // [EPI_WARPS][AMB_BUF]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Units units(p.n_qt, p.n_ct, p.ct_group);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], EPI_WARPS); }
    for (int k = 0; k < 8; ++k) ptx::mbar_init(&afull_bar[k], 1);
    ptx::mbar_init(aempty_bar, 1);
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
      uint32_t unit_no = 0;
      for (long long u = blockIdx.x; u < units.n_units; u += gridDim.x, ++unit_no) {
        long long qt, ct_lo, ct_hi; units.decode(u, &qt, &ct_lo, &ct_hi);
        const unsigned char* asrc = p.apack + (size_t)qt * n_kb * G::A_BLOCK;
        if constexpr (RES) {
          // the previous unit's MMAs must have finished reading the resident A image
          if (unit_no > 0) ptx::mbar_wait(aempty_bar, (unit_no - 1) & 1u);
          for (int kb = 0; kb < n_kb; ++kb) {
            ptx::mbar_arrive_expect_tx(&afull_bar[kb], G::A_BLOCK);
            ptx::bulk_g2s(a_res + (size_t)kb * G::A_BLOCK, asrc + (size_t)kb * G::A_BLOCK, G::A_BLOCK,
                          &afull_bar[kb]);
          }
        }
        for (long long ct = ct_lo; ct < ct_hi; ++ct) {
          const unsigned char* bsrc = p.bpack + (size_t)ct * n_kb * G::B_BLOCK;
          for (int kb = 0; kb < n_kb; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            unsigned char* sa = stage_base + (size_t)stage * G::STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], G::STAGE_BYTES);
            if constexpr (!RES) {
              ptx::bulk_g2s(sa, asrc + (size_t)kb * G::A_BLOCK, G::A_BLOCK, &full_bar[stage]);
              sa += G::A_BLOCK;
            }
            ptx::bulk_g2s(sa, bsrc + (size_t)kb * G::B_BLOCK, G::B_BLOCK, &full_bar[stage]);
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
      uint32_t unit_no = 0;
      for (long long u = blockIdx.x; u < units.n_units; u += gridDim.x, ++unit_no) {
       long long qt_, ct_lo, ct_hi; units.decode(u, &qt_, &ct_lo, &ct_hi);
       for (long long ct = ct_lo; ct < ct_hi; ++ct) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < n_kb; ++kb) {
          if constexpr (RES) {
            if (ct == ct_lo) ptx::mbar_wait(&afull_bar[kb], unit_no & 1u);
          }
          ptx::mbar_wait(&full_bar[stage], phase);
          fence_after();
          const uint32_t sb = ptx::smem_u32(stage_base + (size_t)stage * G::STAGE_BYTES);
          const uint32_t sa = RES ? ptx::smem_u32(a_res + (size_t)kb * G::A_BLOCK) : sb;
          const uint32_t sbb = RES ? sb : sb + G::A_BLOCK;
          const uint64_t a_hi = make_desc<G>(sa), a_lo = make_desc<G>(sa + G::A_PLANE);
          const uint64_t b_hi = make_desc<G>(sbb), b_lo = make_desc<G>(sbb + G::B_PLANE);
          const int k16s = min(G::K16, (p.k_total - kb * BKT + 15) / 16);
          for (int k = 0; k < k16s; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 B = 2 x 16-B units
            umma_bf16(d_tmem, a_hi + adv, b_hi + adv, p.idesc, (kb | k) ? 1u : 0u);
            umma_bf16(d_tmem, a_lo + adv, b_hi + adv, p.idesc, 1u);
            umma_bf16(d_tmem, a_hi + adv, b_lo + adv, p.idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
       }
       if constexpr (RES) umma_commit(aempty_bar);  // every MMA of the unit has read A
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    // The accumulator holds  f = a.b  (dot models) or  f = a.b - |b|^2/2  (L2: the candidate's
    // squared norm rides in three spare k slots of the operand images, see pack_operand_kernel),
    // so a pair is decided by comparing f with two per-thread thresholds -- no per-candidate
    // loads, no arithmetic per element:
    //   dot: s~ = f,            u = s~ - s_true = f - st
    //   L2 : s~ = 2 f - |a|^2,  u = 2 f - (qn + st)
    //   eps(q, c) <= E(q, cbmax) with cbmax = max bound of the 32 candidates of the block
    //   (eps is increasing in the candidate's norm bound), hence
    //   f >  T_hi = (st + E)            [L2: (qn + st + E) / 2]  =>  s(q,c) >  s_true : count
    //   f <  T_lo = (st - E)            [L2: (qn + st - E) / 2]  =>  s(q,c) <  s_true : skip
    //   otherwise (including any NaN) near-tie: exact recheck.  T_hi is rounded up, T_lo down.
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // query row within the tile
    const int col_half = (warp - 2) >> 2;     // which 128 columns of the tile this warp handles
    int acc = 0; uint32_t acc_phase = 0;
    long long cur_qt = -1;
    float st = 0.f, qb = 0.f, qn = 0.f, kp = 0.f;
    int cnt = 0;
    int2* wbuf = s_amb + (warp - 2) * AMB_BUF;  // this warp's near-tie buffer
    int amb_n = 0;                              // entries in it (warp-uniform)
    float k1 = 0.f, k0 = 0.f, tbase = 0.f;
    constexpr float INFL = 1.f + 0x1p-19f;      // covers the fp32 rounding of E's evaluation
    // the accumulator holds S * (a.b [- |b|^2/2]), S = scale_a * scale_b (a power of two; NaN when an
    // operand could not be represented: both threshold tests then fail and every pair is rechecked)
    const float S = p.meta_a->acc_scale;
    const float inv_S = 1.f / S;
    const float e_abs = p.meta_a->e_abs;
    const float kappa_a = p.meta_a->kappa, kappa_b = p.meta_b->kappa;
    for (long long u = blockIdx.x; u < units.n_units; u += gridDim.x) {
     long long qt, ct_lo, ct_hi; units.decode(u, &qt, &ct_lo, &ct_hi);
     for (long long ct = ct_lo; ct < ct_hi; ++ct) {
      const long long q = qt * BM + row;
      if (qt != cur_qt) {
        if (cur_qt >= 0 && cnt != 0) {
          const long long pq = cur_qt * BM + row;
          if (pq < p.n_q) atomicAdd(&p.counts[pq], cnt);
        }
        // the near-tie list is kept per QUERY TILE (region = qt) so that the exact recheck of a
        // region touches only that tile's 128 query rows (they stay L1-resident there): flush
        // what this warp buffered for the previous tile before moving on
        if (cur_qt >= 0 && amb_n > 0) {
          __syncwarp();
          unsigned long long fbase = 0;
          if (lane == 0) fbase = atomicAdd(p.amb_count + cur_qt, (unsigned long long)amb_n);
          fbase = __shfl_sync(0xffffffffu, fbase, 0);
          for (int i = lane; i < amb_n; i += 32)
            if (fbase + i < p.amb_cap) p.amb_pairs[(size_t)cur_qt * p.amb_cap + fbase + i] = wbuf[i];
          __syncwarp();
          amb_n = 0;
        }
        cnt = 0; cur_qt = qt;
        const bool vq = q < p.n_q;
        st = vq ? p.s_true[q] : INFINITY;
        qb = vq ? __fadd_ru(p.qbound[q], kappa_a) : 0.f;
        qn = vq ? p.qnorm2[q] : 0.f;
        // accumulation term: gamma_p P(a) P(b) (x2 for L2, whose score is 2 f - |a|^2)
        kp = vq ? (L2 ? 2.f : 1.f) * p.gamma_p * p.qprefix[q] * INFL : 0.f;
        if constexpr (L2) {
          // E = 2 gamma qb cb + gamma2 (qb + cb)^2 = cb (k1 + gamma2 cb) + k0
          k1 = 2.f * (p.gamma + p.gamma2) * qb * INFL;
          k0 = p.gamma2 * qb * qb * INFL;
          tbase = qn + st;
        } else {
          k1 = p.gamma * qb * INFL;   // E = gamma qb cb
          tbase = st;
        }
      }
      // largest candidate norm bound of each of this warp's four 32-column blocks (one coalesced
      // load + one redux each, issued before waiting for the accumulator)
      const int c_begin = col_half * (BN / 2);
      // (maxima over aligned blocks of 32 rows, precomputed with the image: two broadcast loads per block)
      float cbm[4], cpm[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const long long blk = (ct * BN + c_begin) / 32 + b;
        cbm[b] = __fadd_ru(__ldg(p.cbmax32 + blk), kappa_b);
        cpm[b] = __ldg(p.cpmax32 + blk);
      }
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      fence_after();
      const int ncols = (int)min((long long)BN, p.n_rows - ct * BN);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      const int c_end = min(ncols, c_begin + BN / 2);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c0 = c_begin + 32 * b;
        if (c0 >= c_end) break;
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        const int lim = min(32, ncols - c0);
        float t_hi, t_lo;
        {
          const float cb = cbm[b];
          float e;
          if constexpr (L2) e = fmaf(cb, fmaf(p.gamma2 * INFL, cb, k1), k0) * INFL;
          else e = k1 * cb;
          e = __fadd_ru(__fmaf_ru(kp, cpm[b], e), e_abs);
          t_hi = __fadd_ru(tbase, e);
          t_lo = __fadd_rd(tbase, -e);
          if constexpr (L2) { t_hi = __fmul_ru(t_hi, 0.5f); t_lo = __fmul_rd(t_lo, 0.5f); }
          t_hi = __fmul_ru(t_hi, S);   // exact (power of two) unless it overflows, then still on the safe side
          t_lo = __fmul_rd(t_lo, S);
        }
        // 32 independent threshold tests -> two bit masks per thread (no per-element branches)
        // (NaN / inf anywhere -- accumulator, thresholds, norm bounds -- fails both tests and lands
        // in the near-tie list, i.e. is adjudicated by the exact arithmetic)
        unsigned lt_mask = 0, gt_mask = 0, amb_mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float f = __uint_as_float(v[j]);
          if constexpr (DUMP) {
            if (j < lim && q < p.n_q)
              p.dump[(size_t)q * p.n_rows + ct * BN + c0 + j] = L2 ? fmaf(2.f, f * inv_S, -qn) : f * inv_S;
          } else {
            gt_mask |= (f > t_hi ? 1u : 0u) << j;
            lt_mask |= (f < t_lo ? 1u : 0u) << j;
          }
        }
        amb_mask = ~(gt_mask | lt_mask);
        if constexpr (!DUMP) {
          if (lim < 32) {  // columns past the table (last tile only)
            const unsigned keep = (1u << lim) - 1u;
            amb_mask &= keep; gt_mask &= keep;
          }
          // padding rows of the last query tile have no list entries (their thresholds are +inf,
          // but a non-finite candidate norm bound turns them into NaN, which fails both tests)
          if (q >= p.n_q) amb_mask = 0u;
          cnt += __popc(gt_mask);
          if (__any_sync(0xffffffffu, amb_mask != 0)) {
            // near-ties in this 32 x 32 block: warp prefix sum of the per-lane counts, entries
            // into this warp's shared buffer, one global atomic per ~50 entries on flush
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
              if (lane == 0) base = atomicAdd(p.amb_count + cur_qt, (unsigned long long)amb_n);
              base = __shfl_sync(0xffffffffu, base, 0);
              for (int i = lane; i < amb_n; i += 32)
                if (base + i < p.amb_cap) p.amb_pairs[(size_t)cur_qt * p.amb_cap + base + i] = wbuf[i];
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
              if (lane == 0) base = atomicAdd(p.amb_count + cur_qt, (unsigned long long)total_new);
              base = __shfl_sync(0xffffffffu, base, 0) + (unsigned long long)(incl - mine);
              unsigned m = amb_mask;
              while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                if (base < p.amb_cap)
                  p.amb_pairs[(size_t)cur_qt * p.amb_cap + base] = make_int2((int)q, (int)(ct * BN + c0 + j));
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
     }
    }
    if (cur_qt >= 0 && cnt != 0) {
      const long long pq = cur_qt * BM + row;
      if (pq < p.n_q) atomicAdd(&p.counts[pq], cnt);
    }
    if (amb_n > 0) {  // final flush of this warp's near-tie buffer
      __syncwarp();
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(p.amb_count + cur_qt, (unsigned long long)amb_n);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < amb_n; i += 32)
        if (base + i < p.amb_cap) p.amb_pairs[(size_t)cur_qt * p.amb_cap + base + i] = wbuf[i];
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------
// Operand packing: fp32 rows -> (hi, lo) bf16 planes in the shared-memory image of K-major
// swizzled tiles.  One thread per (row, 16-byte chunk): 8 consecutive k of one plane.
//   128-byte rows (BKT = 64, SWIZZLE_128B): offset of (row r, chunk j) inside a plane =
//       (r/8)*1024 + (r%8)*128 + ((j ^ (r%8)) * 16)              j = 0..7
//   64-byte rows  (BKT = 32, SWIZZLE_64B) :
//       (r/8)*512  + (r%8)*64  + ((j ^ ((r>>1)&3)) * 16)         j = 0..3
// (the swizzle XORs address bits [4,7) with bits [7,10), resp. bits [4,6) with bits [7,9)).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float operand_value(const float* __restrict__ p0,
                                               const float* __restrict__ p1, int dim, int k,
                                               int k_total) {
  if (k >= k_total) return 0.f;
  if (k < dim) return p0[k];
  if (k < 2 * dim) return p1[k - dim];
  // third plane of a three-plane row (Analogy: k_total = 3 dim): the planes are equally spaced, both in
  // a query row ([3][dim], p1 = p0 + dim) and across the stacked candidate table (include/kge_b200.h)
  return (p1 + (p1 - p0))[k - 2 * dim];
}

// Half-precision formats of the split: bf16 (8 significant bits) or fp16 (11; operands pre-scaled)
// Optional content guard of a cached candidate image: guard[0..1] = checksum of the table as it is
// now, guard[2..3] = checksum of the table the image was built from (0, 0 before the first build is
// committed; a real checksum has bit 0 of word 1 set).  Equal => every pack kernel returns at once.
__device__ __forceinline__ bool guard_unchanged(const unsigned long long* __restrict__ guard) {
  return guard != nullptr && guard[0] == guard[2] && guard[1] == guard[3];
}

template <bool FP16> struct HalfT;
template <> struct HalfT<false> {
  using T = __nv_bfloat16;
  static __device__ __forceinline__ T from(float x) { return __float2bfloat16_rn(x); }
  static __device__ __forceinline__ float to(T h) { return __bfloat162float(h); }
};
template <> struct HalfT<true> {
  using T = __half;
  static __device__ __forceinline__ T from(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ float to(T h) { return __half2float(h); }
};

template <int ROWS, int BKT, bool FP16>
__global__ void pack_operand_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                    long long row_stride, long long plane1_offset, long long n_rows,
                                    int dim, int k_total, int n_kb, int sub_mode, int fold,
                                    const float* __restrict__ norm2, const TcMeta* __restrict__ meta,
                                    const unsigned long long* __restrict__ guard,
                                    unsigned char* __restrict__ out) {
  if (guard_unchanged(guard)) return;   // the image already holds exactly this table
  // src0 + row*row_stride = first plane of the row; second plane at +plane1_offset (same row)
  // or in src1 (separate table).  sub_mode = 1: value = plane1[k] - plane0[k]  (t - r, L2 head)
  // Every value is multiplied by meta->scale (a power of two: exact) before it is split.
  // fold (L2 only; k_total = dim + 3): the three k slots after the data carry, on the candidate
  // side (fold = 2), -|b|^2/2 * phi split exactly into three half-precision pieces (hi plane; lo
  // plane 0) and, on the query side (fold = 1), alpha = scale_a * scale_b / phi -- so the hi*hi
  // product adds -S |b|^2/2 to every accumulator and the epilogue compares the accumulator with
  // thresholds directly.
  using H = HalfT<FP16>;
  using HT = typename H::T;
  constexpr int CH = BKT / 8;        // 16-byte chunks per row
  constexpr int ROWB = 2 * BKT;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_tiles = (n_rows + ROWS - 1) / ROWS;
  const long long total = n_tiles * n_kb * ROWS * CH;
  if (gid >= total) return;
  const float scale = meta->scale, fold_c = meta->fold;
  const int j = (int)(gid % CH);
  long long rest = gid / CH;
  const int r = (int)(rest % ROWS);
  rest /= ROWS;
  const int kb = (int)(rest % n_kb);
  const long long tile = rest / n_kb;
  const long long row = tile * ROWS + r;
  HT hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kb * BKT + j * 8 + e;
    float x = 0.f;
    if (row < n_rows) {
      if (fold && k >= dim) {
        if (k < dim + 3) {
          if (fold == 1) {
            hi[e] = H::from(fold_c);                 // alpha: a power of two inside the format's range
          } else {
            float rem = -0.5f * norm2[row] * fold_c;  // phi: a power of two (exact)
            HT piece = H::from(rem);
            for (int i = 0; i < k - dim; ++i) {
              rem -= H::to(piece);  // exact: piece holds the leading bits of rem
              piece = H::from(rem);
            }
            hi[e] = piece;
          }
          lo[e] = H::from(0.f);
          continue;
        }
      } else {
        const float* a = src0 + (size_t)row * row_stride;
        const float* b = src1 ? src1 + (size_t)row * row_stride : a + plane1_offset;
        if (sub_mode) x = k < dim ? __fsub_rn(b[k], a[k]) : 0.f;
        else x = operand_value(a, b, dim, k, k_total);
      }
    }
    x *= scale;
    const HT h = H::from(x);
    hi[e] = h;
    lo[e] = H::from(x - H::to(h));
  }
  const size_t plane = (size_t)ROWS * ROWB;
  const size_t base = ((size_t)tile * n_kb + kb) * (2 * plane);
  const int sw = BKT == 64 ? (r & 7) : ((r >> 1) & 3);
  const size_t off = (size_t)(r / 8) * (8 * ROWB) + (size_t)(r % 8) * ROWB + (size_t)((j ^ sw) * 16);
  *reinterpret_cast<uint4*>(out + base + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(out + base + plane + off) = *reinterpret_cast<const uint4*>(lo);
}

// per-row |x|_2 (rounded up a little), |x|_2^2 and the running-magnitude factor
//   P(x) = sqrt( sum_{i=1..ceil(k/16)} |x_{<= 16 i}|^2 )     (tc.h: tc_gamma_p)
// of the operand vector, one warp per row: 32 consecutive k per iteration, lanes 0-15 / 16-31 are the
// two MMA k-steps of the iteration.  The operand's largest |x| and largest |x|_2^2 are folded into
// meta (bit-pattern maxima of non-negative floats: a NaN or inf anywhere wins, which invalidates the
// image -- see tc_meta_kernel).
__global__ void row_norms_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                 long long row_stride, long long plane1_offset, long long n_rows,
                                 int dim, int k_total, int extra_steps, int sub_mode, float* __restrict__ bound,
                                 float* __restrict__ norm2, float* __restrict__ prefix, TcMeta* __restrict__ meta,
                                 const unsigned long long* __restrict__ guard) {
  // extra_steps: MMA k-steps beyond ceil(k_total / 16) (the L2 fold slots spilling into a k-step of
  // their own): their incoming accumulator is bounded by the full |x|^2
  if (guard_unchanged(guard)) return;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const float* a = src0 + (size_t)w * row_stride;
  const float* b = src1 ? src1 + (size_t)w * row_stride : a + plane1_offset;
  double s = 0.0, sa = 0.0, sb = 0.0;   // s: running |x_{<= ...}|^2 (all lanes hold the same value)
  double p2 = 0.0;                      // sum of the squared prefix norms at the 16-element boundaries
  unsigned mx = 0u;
  for (int k0 = 0; k0 < k_total; k0 += 32) {
    const int k = k0 + lane;
    float x = 0.f;
    double xa = 0.0, xb = 0.0;
    if (k < k_total) {
      if (sub_mode) {
        x = k < dim ? __fsub_rn(b[k], a[k]) : 0.f;
        if (k < dim) { xa = (double)a[k] * (double)a[k]; xb = (double)b[k] * (double)b[k]; }
      } else {
        x = operand_value(a, b, dim, k, k_total);
      }
    }
    mx = max(mx, __float_as_uint(fabsf(x)));
    double h = (double)x * (double)x;   // -> sum over this lane's half-warp (one k-step)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      h += __shfl_xor_sync(0xffffffffu, h, o);
      xa += __shfl_xor_sync(0xffffffffu, xa, o);
      xb += __shfl_xor_sync(0xffffffffu, xb, o);
    }
    const double h_lo = __shfl_sync(0xffffffffu, h, 0), h_hi = __shfl_sync(0xffffffffu, h, 16);
    sa += xa + __shfl_xor_sync(0xffffffffu, xa, 16);
    sb += xb + __shfl_xor_sync(0xffffffffu, xb, 16);
    s += h_lo;
    p2 += s;                                           // |x_{<= 16 i}|^2 after k-step i = 2 g + 1
    if (k0 + 16 < k_total) { s += h_hi; p2 += s; }     // ... and after k-step 2 g + 2
  }
  p2 += (double)extra_steps * s;
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) {
    const float n2 = (float)s;
    norm2[w] = n2;
    // L2 head side: the reference rounds c + r before subtracting t, an error that scales with
    // |r| and |t| separately, so the bound uses |t| + |r| (>= |t - r|) for this operand
    const double nb = sub_mode ? sqrt(sa) + sqrt(sb) : sqrt(s);
    bound[w] = (float)(nb * (1.0 + 1e-6)) + 1e-30f;
    // the head-side operand t - r is bounded by |t| + |r| the same way: scale P by that ratio
    const double pr = sqrt(p2) * (s > 0.0 ? nb / sqrt(s) : 1.0);
    prefix[w] = (float)(pr * (1.0 + 1e-6)) + 1e-30f;
    atomicMax(reinterpret_cast<unsigned*>(&meta->max_abs), mx);
    atomicMax(reinterpret_cast<unsigned*>(&meta->max_norm2), __float_as_uint(fabsf(n2)));
  }
}

// Scales of one operand image, decided on the device from the maxima row_norms_kernel gathered.
//   bf16: scale = 1, phi = alpha = 1, S = 1, nothing else.
//   fp16: scale = 2^e with max|x| * scale in [2^8, 2^9): hi never overflows (65504), the lo part of
//         every element >= 2^-12 max|x| is a normal fp16 number (residual 2^-22 |x|); smaller elements
//         are off by <= 2^-25 in scaled units = 2^-33 max|x|, which sum_k |y_k| <= sqrt(K) |y| turns into
//         <= 2^-33 sqrt(K) max|x| |y| per pair: covered by adding kappa = 2^-11 sqrt(K) max|x| / 3 (+1 %)
//         to the row bounds (gamma >= 3 2^-22 multiplies them).
//         phi (candidate side, L2 fold): max_rows(|b|^2/2) * phi in [2^13, 2^14): the three pieces of
//         -|b|^2/2 * phi fit fp16 and are exact to 33 bits for every row within 2^-13 of the largest;
//         beyond, the absolute error 3 * 2^-25 / phi (|b|^2/2 units) goes into e_abs.
//         alpha (query side) = scale_a * scale_b / phi must itself be an fp16 normal power of two.
//   An operand with a NaN / inf entry, or whose alpha falls outside fp16, gets scale = NaN: its image,
//   S and hence both thresholds are NaN, every pair fails both tests and is rechecked exactly.
// maxima of the row bounds / running-magnitude factors over aligned blocks of 32 rows (rows past the
// table count as 0): one warp per block
__global__ void block_max_kernel(const float* __restrict__ bound, const float* __restrict__ prefix, long long n_rows,
                                 long long n_blocks, float* __restrict__ bmax, float* __restrict__ pmax,
                                 const unsigned long long* __restrict__ guard) {
  if (guard_unchanged(guard)) return;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_blocks) return;
  const long long r = w * 32 + lane;
  const float x = r < n_rows ? bound[r] : 0.f, y = r < n_rows ? prefix[r] : 0.f;
  const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(x));   // x, y >= 0 (NaN sorts on top)
  const unsigned my = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
  if (lane == 0) { bmax[w] = __uint_as_float(mx); pmax[w] = __uint_as_float(my); }
}

__global__ void tc_meta_reset_kernel(TcMeta* __restrict__ m, const unsigned long long* __restrict__ guard) {
  if (guard_unchanged(guard)) return;
  if (threadIdx.x == 0) { m->max_abs = 0.f; m->max_norm2 = 0.f; }
}

// after a (re)build: remember which table content the image now holds
__global__ void tc_guard_commit_kernel(unsigned long long* __restrict__ guard) {
  if (threadIdx.x == 0) { guard[2] = guard[0]; guard[3] = guard[1]; }
}

// 128-bit content checksum of a table: two order-independent sums over its 64-bit words w_i,
//   h0 = sum w_i (2 i + 1),   h1 = sum (w_i ^ (w_i >> 31) ^ salt) (2 i + 1)        (mod 2^64)
// The odd, position-dependent multipliers make any single changed word change both sums, and several
// changed words cancel only by a 2^-64 coincidence per sum.  One pass at HBM speed: 16-byte loads,
// four in flight per thread.
__global__ void __launch_bounds__(256) table_checksum_kernel(const uint4* __restrict__ data, long long n_vec,
                                                             const uint32_t* __restrict__ tail_words,
                                                             int n_tail, unsigned long long salt,
                                                             unsigned long long* __restrict__ out) {
  unsigned long long h0 = 0ull, h1 = 0ull;
  auto add = [&](unsigned long long w, unsigned long long i) {
    const unsigned long long m = 2ull * i + 1ull;
    h0 += w * m;
    h1 += (w ^ (w >> 31) ^ salt) * m;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(data + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned long long j = 2ull * (unsigned long long)(i + u * stride);
      add(((unsigned long long)v[u].y << 32) | v[u].x, j);
      add(((unsigned long long)v[u].w << 32) | v[u].z, j + 1ull);
    }
  }
  for (; i < n_vec; i += stride) {
    const uint4 v = __ldg(data + i);
    add(((unsigned long long)v.y << 32) | v.x, 2ull * (unsigned long long)i);
    add(((unsigned long long)v.w << 32) | v.z, 2ull * (unsigned long long)i + 1ull);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail)   // words past the last full 16-byte vector
    add(tail_words[threadIdx.x], 2ull * (unsigned long long)n_vec + threadIdx.x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h0 += __shfl_xor_sync(0xffffffffu, h0, o);
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], h0); atomicAdd(&out[1], h1); }
}
__global__ void table_checksum_finish_kernel(unsigned long long* __restrict__ out) {
  if (threadIdx.x == 0) out[1] |= 1ull;   // never (0, 0): distinguishes "no image yet"
}

__global__ void tc_meta_kernel(TcMeta* __restrict__ m, const TcMeta* __restrict__ other, int k_total,
                               int is_query, int l2, int fp16, const unsigned long long* __restrict__ guard) {
  if (guard_unchanged(guard)) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float mx = m->max_abs, n2 = m->max_norm2;
  const bool finite = mx <= 3.0e38f && n2 <= 3.0e38f;   // false for NaN and inf
  float scale = 1.f, fold = 1.f, kappa = 0.f, acc = 1.f, e_abs = 0.f;
  bool ok = finite;
  if (fp16 && finite) {
    if (mx > 0.f) {
      int ex; frexpf(mx, &ex);                 // mx = f * 2^ex, f in [0.5, 1)
      int e = 9 - ex;
      e = max(-120, min(120, e));
      scale = ldexpf(1.f, e);
      kappa = mx * 0x1p-11f * sqrtf((float)k_total) * (1.01f / 3.f);
    }
    if (!is_query && l2 && n2 > 0.f) {
      int ex; frexpf(0.5f * n2, &ex);
      fold = ldexpf(1.f, max(-120, min(120, 14 - ex)));   // phi
    }
  }
  if (is_query) {
    const float sb = other->scale, phi = other->fold;
    acc = scale * sb;                                      // NaN if the table image is invalid
    if (l2) {
      fold = acc / phi;                                    // alpha
      if (fp16) {
        if (!(fold >= 0x1p-14f && fold <= 0x1p15f)) ok = false;
        e_abs = 0x1p-22f / phi;                            // 8 * 2^-25 / phi: fold pieces, score = 2 f - |a|^2
      }
    }
    if (!(acc >= 0x1p-60f && acc <= 0x1p60f)) ok = false;
  }
  const float nanv = __uint_as_float(0x7fc00000u);
  m->scale = ok ? scale : nanv;
  m->fold = ok ? fold : nanv;
  m->acc_scale = ok ? acc : nanv;
  m->e_abs = e_abs;
  m->kappa = kappa;
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
      const float* c1 = (CW == 3 ? ent1 + (ent1 - ent0) : (CW == 2 ? ent1 : ent0)) + (size_t)pr.y * dim;
      const float* cm = (CW == 3 ? ent1 : ent0) + (size_t)pr.y * dim;   // middle plane (three-plane kinds)
      if (pair_score_natural<EL>(dim, q0, q0 + (size_t)(QW - 1) * dim, c0, c1, q0 + (size_t)(QW / 2) * dim, cm) >=
          s_true[pr.x])
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
    const float* c1 = (CW == 3 ? ent1 + (ent1 - ent0) : (CW == 2 ? ent1 : ent0)) + (size_t)pr.y * dim;
    const float* qm = q0 + (size_t)(QW / 2) * dim;                      // middle plane (three-plane kinds)
    const float* cm = (CW == 3 ? ent1 : ent0) + (size_t)pr.y * dim;
    const float sc = pair_score_chains<EL>(dim, q0, q1, c0, c1, lane, qm, cm);
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

// ------------------------------------------------------------------------------------------
// Geometry choice.  k-block = 32 bf16 (64-byte swizzle) by default: finer stages, 7 % instead of
// 23 % padding at k = 200, and the query tile's A image (<= 7 blocks, k_total <= 224) can stay
// resident in shared memory so that only B streams (1/3 less L2 -> SM traffic, the limiter of
// the streamed form).  KGE_TC_BK=64 selects the 128-byte-swizzle layout (2 stages of 96 KB),
// KGE_TC_RESIDENT=0 disables residency, KGE_TC_GROUP sets candidate tiles per unit.
// ------------------------------------------------------------------------------------------
namespace {
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
struct Config {
  int bk, resident, group, max_ctas, fp16;
  Config()
      : bk(env_int("KGE_TC_BK", 32) == 64 ? 64 : 32), resident(env_int("KGE_TC_RESIDENT", 1) != 0),
        group(env_int("KGE_TC_GROUP", 0)), max_ctas(env_int("KGE_TC_MAX_CTAS", 0)),
        fp16(env_int("KGE_TC_FP16", 1) != 0) {}
};
Config& config() {
  static Config c;
  return c;
}
}  // namespace

void configure(int bk_, int resident_, int group_, int max_ctas_, int fp16_) {
  Config& c = config();
  if (bk_ == 32 || bk_ == 64) c.bk = bk_;
  if (resident_ >= 0) c.resident = resident_ != 0;
  if (group_ >= 0) c.group = group_;
  if (max_ctas_ >= 0) c.max_ctas = max_ctas_;
  if (fp16_ >= 0) c.fp16 = fp16_ != 0;
}
bool fp16() { return config().fp16 != 0; }
int bk() { return config().bk; }
int n_kblocks(int k_total) { return (k_total + bk() - 1) / bk(); }
bool resident(int n_kb) {
  return config().resident && bk() == 32 && n_kb <= 7 && smem_bytes<32, true>(n_kb) <= SMEM_LIMIT;
}
int ct_group(int n_kb) {
  if (config().group > 0) return config().group;
  return resident(n_kb) ? 32 : 16;
}
// With few query tiles (a rank's slice of a query-sharded run) whole groups of 32 candidate tiles
// are too coarse a unit for 148 CTAs: shrink the group until there are >= 32 units per CTA.
int ct_group_for(int n_kb, long long n_qt, long long n_ct, int sms) {
  int g = ct_group(n_kb);
  if (config().group > 0) return g;
  while (g > 4 && n_qt * ((n_ct + g - 1) / g) < 32ll * sms) g >>= 1;
  return g;
}

size_t a_image_bytes(long long n_q, int n_kb) {
  const long long n_qt = (n_q + BM - 1) / BM;
  return (size_t)n_qt * n_kb * 2 * BM * 2 * bk();
}
size_t b_image_bytes(long long n_rows, int n_kb) {
  const long long n_ct = (n_rows + BN - 1) / BN;
  return (size_t)n_ct * n_kb * 2 * BN * 2 * bk();
}

namespace {
template <int ROWS>
void launch_pack_operand(const float* src0, const float* src1, long long row_stride, long long plane1_offset,
                         long long n_rows, int dim, int k_total, int n_kb, int sub_mode, int fold,
                         const float* norm2, const TcMeta* meta, const unsigned long long* guard,
                         unsigned char* out, cudaStream_t st) {
  const long long n_tiles = (n_rows + ROWS - 1) / ROWS;
  const long long total = n_tiles * n_kb * ROWS * (bk() / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
#define KGE_PACK(BKT, F16)                                                                              \
  pack_operand_kernel<ROWS, BKT, F16><<<blocks, 256, 0, st>>>(src0, src1, row_stride, plane1_offset, n_rows, dim, \
                                                              k_total, n_kb, sub_mode, fold, norm2, meta, guard, out)
  if (bk() == 64) { if (fp16()) KGE_PACK(64, true); else KGE_PACK(64, false); }
  else { if (fp16()) KGE_PACK(32, true); else KGE_PACK(32, false); }
#undef KGE_PACK
}
}  // namespace

uint32_t instruction_descriptor() { return fp16() ? IDESC_FP16 : IDESC_BF16; }

cudaError_t launch_pack_b(const float* ent0, const float* ent1, long long n_rows, int dim, int k_total,
                          int n_kb, bool fold, unsigned char* bpack, float* cbound, float* cnorm2, float* cprefix,
                          float* cbmax32, float* cpmax32, TcMeta* meta_b, unsigned long long* guard, cudaStream_t st) {
  if (n_rows <= 0) return cudaSuccess;
  if (guard) {
    // checksum of the table as it is now (one pass over it at HBM speed); if it equals the one the
    // image was built from, every kernel below returns immediately
    cudaError_t e = cudaMemsetAsync(guard, 0, 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long n_words = n_rows * dim;
    auto checksum = [&](const float* tab, unsigned long long salt) {
      // 16-byte vector loads over the aligned middle of the table; the <= 3 words before it and the
      // <= 3 words after it go through the kernel's tail path (positions past the vector part)
      const uint32_t* words = reinterpret_cast<const uint32_t*>(tab);
      const long long head = min(n_words, (long long)(((16u - (reinterpret_cast<uintptr_t>(tab) & 15u)) & 15u) / 4u));
      const long long n_vec = (n_words - head) / 4;
      const int n_tail = (int)(n_words - head - 4 * n_vec);
      const unsigned blocks = (unsigned)max(1ll, min((n_vec + 1023) / 1024, (long long)148 * 8));
      table_checksum_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(words + head), n_vec,
                                                    words + head + 4 * n_vec, n_tail, salt, guard);
      if (head > 0)
        table_checksum_kernel<<<1, 32, 0, st>>>(nullptr, 0, words, (int)head, salt ^ 0xA5A5A5A5A5A5A5A5ull, guard);
    };
    checksum(ent0, 0ull);
    if (ent1) checksum(ent1, 0x5851F42D4C957F2Dull);
    if (ent1 && !fold && k_total == 3 * dim) checksum(ent1 + (ent1 - ent0), 0x2545F4914F6CDD1Dull);   // third plane
    table_checksum_finish_kernel<<<1, 32, 0, st>>>(guard);
  }
  tc_meta_reset_kernel<<<1, 32, 0, st>>>(meta_b, guard);
  // norms first: with fold the image carries -|b|^2/2 (the SAME fp32 value the bound uses)
  row_norms_kernel<<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, st>>>(
      ent0, ent1, dim, 0, n_rows, dim, fold ? dim : k_total, fold ? (k_total + 15) / 16 - (dim + 15) / 16 : 0, 0,
      cbound, cnorm2, cprefix, meta_b, guard);
  {
    const long long n_blocks = ((n_rows + BN - 1) / BN) * (BN / 32);
    block_max_kernel<<<(unsigned)((n_blocks * 32 + 255) / 256), 256, 0, st>>>(cbound, cprefix, n_rows, n_blocks, cbmax32,
                                                                              cpmax32, guard);
  }
  tc_meta_kernel<<<1, 32, 0, st>>>(meta_b, nullptr, k_total, 0, fold ? 1 : 0, fp16() ? 1 : 0, guard);
  launch_pack_operand<BN>(ent0, ent1, dim, 0, n_rows, dim, k_total, n_kb, 0, fold ? 2 : 0, cnorm2, meta_b, guard,
                          bpack, st);
  if (guard) tc_guard_commit_kernel<<<1, 32, 0, st>>>(guard);
  return cudaGetLastError();
}

cudaError_t launch_pack_a(const float* qplain, int qw, long long n_q, int dim, int k_total, int n_kb,
                          int sub_mode, bool fold, unsigned char* apack, float* qbound, float* qnorm2, float* qprefix,
                          TcMeta* meta_a, const TcMeta* meta_b, cudaStream_t st) {
  if (n_q <= 0) return cudaSuccess;
  tc_meta_reset_kernel<<<1, 32, 0, st>>>(meta_a, nullptr);
  // qplain rows are [qw][dim]: plane 1 (if any) follows plane 0 inside the row
  row_norms_kernel<<<(unsigned)((n_q * 32 + 255) / 256), 256, 0, st>>>(
      qplain, nullptr, (long long)qw * dim, dim, n_q, dim, fold ? dim : k_total,
      fold ? (k_total + 15) / 16 - (dim + 15) / 16 : 0, sub_mode, qbound, qnorm2, qprefix, meta_a, nullptr);
  tc_meta_kernel<<<1, 32, 0, st>>>(meta_a, meta_b, k_total, 1, fold ? 1 : 0, fp16() ? 1 : 0, nullptr);
  launch_pack_operand<BM>(qplain, nullptr, (long long)qw * dim, dim, n_q, dim, k_total, n_kb, sub_mode,
                          fold ? 1 : 0, nullptr, meta_a, nullptr, apack, st);
  return cudaGetLastError();
}

namespace {
template <int BKT, bool RES>
cudaError_t launch_variant(const TcScanParams& p, int grid, cudaStream_t st) {
  const size_t smem = smem_bytes<BKT, RES>(p.n_kb);
  if (smem > SMEM_LIMIT) return cudaErrorInvalidValue;
  // per template instantiation AND per device (the attribute is a per-device property)
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cudaErrorInvalidDevice;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kern) {
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    };
    set(tc_scan_kernel<false, false, BKT, RES>); set(tc_scan_kernel<true, false, BKT, RES>);
    set(tc_scan_kernel<false, true, BKT, RES>); set(tc_scan_kernel<true, true, BKT, RES>);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (p.dump) {
    if (p.l2) tc_scan_kernel<true, true, BKT, RES><<<grid, THREADS, smem, st>>>(p);
    else tc_scan_kernel<false, true, BKT, RES><<<grid, THREADS, smem, st>>>(p);
  } else {
    if (p.l2) tc_scan_kernel<true, false, BKT, RES><<<grid, THREADS, smem, st>>>(p);
    else tc_scan_kernel<false, false, BKT, RES><<<grid, THREADS, smem, st>>>(p);
  }
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_tc_scan(const TcScanParams& p_in, cudaStream_t st) {
  TcScanParams p = p_in;
  p.idesc = instruction_descriptor();
  const int grid = scan_grid_size(p.n_q, p.n_rows, p.n_kb, &p.ct_group);
  if (grid <= 0) return cudaSuccess;
  if (bk() == 64) return launch_variant<64, false>(p, grid, st);
  if (resident(p.n_kb)) return launch_variant<32, true>(p, grid, st);
  return launch_variant<32, false>(p, grid, st);
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
    case EL_DOT3: CALL_RC(EL_DOT3); break;
    case EL_L2_TAIL: CALL_RC(EL_L2_TAIL); break;
    case EL_L2_HEAD: CALL_RC(EL_L2_HEAD); break;
    case EL_ROT: CALL_RC(EL_ROT); break;
    default: return cudaErrorInvalidValue;
  }
#undef CALL_RC
  if (stats) tc_stats_kernel<<<1, 1, 0, st>>>(region_counts, regions, region_cap, stats);
  return cudaGetLastError();
}

int scan_grid_size(long long n_q, long long n_rows, int n_kb, int* group_out) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  const long long n_qt = (n_q + BM - 1) / BM, n_ct = (n_rows + BN - 1) / BN;
  if (config().max_ctas > 0 && config().max_ctas < sms) sms = config().max_ctas;
  const long long g = ct_group_for(n_kb, n_qt, n_ct, sms);
  if (group_out) *group_out = (int)g;
  const long long units = n_qt * ((n_ct + g - 1) / g);
  return (int)(units < sms ? units : sms);
}

}  // namespace tc
}  // namespace kge
