// K5 on the 5th-gen tensor cores: causal GQA flash-attention FORWARD for head_dim 64 (CSM-1B backbone).
//   warp 0     : TMA producer — Q tile once, then K_j / V_j tiles (64 keys x 64) into a 3-stage ring
//   warp 1     : tcgen05.mma issuer — S_j = Q K_j^T (128x64, fp32 in TMEM, double buffered) and
//                PV_j = P_j V_j (128x64): P is read as the A operand straight out of TMEM, V is consumed MN-major from
//                its row-major smem tile
//   warps 2..5 : softmax — one query row per thread (row == TMEM lane, so no shuffles): the S row is read with
//                tcgen05.ld (max, then exp2/sum), P_j goes back over the same TMEM columns as bf16 pairs
//                (tcgen05.st) — it never touches shared memory; the running output is kept in registers and updated
//                from the PV_j tile one block later, so the tensor pipe computes S_{j+1} and PV_j while the softmax of
//                the next block runs.
// The kernel is MUFU-bound (one ex2 per score, 16 per clock per SM), so what matters is keeping all four XU pipes
// fed: 64-key blocks keep a CTA at 65 KB of shared memory and 256 TMEM columns, so TWO CTAs are resident per SM and
// one CTA's softmax overlaps the other's TMEM round trips and mbarrier hand-offs.
#include <type_traits>

#include "tc_common.cuh"

namespace csm {

using namespace tc;

// -DCSM_ATTN_PROF (tools/ubench/attn_prof.sh only; never in the product build): per-role cycle accounting of the backward
// kernels — CTA 0's MMA warp and two compute warps add up the cycles they spend in each wait.
#ifdef CSM_ATTN_PROF
__device__ long long g_attn_prof[128];
#define PROF_WAIT(acc, ...) { const long long _t = clock64(); __VA_ARGS__; acc += clock64() - _t; }
#define PROF_ONLY(...) __VA_ARGS__
#else
#define PROF_WAIT(acc, ...) { __VA_ARGS__; }
#define PROF_ONLY(...)
#endif

namespace {

constexpr int TQ = 128, TK = 128, THD = 64;
constexpr int FK = 64;   // keys per block of the FORWARD kernel
constexpr int FST = 3;   // K/V ring depth of the forward kernel (a stage is held until its PV MMA has completed, so
                         // two stages would put the ~1 us TMA latency of block j+1 on the softmax's critical path)
constexpr int kAttnThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + TQ * THD * 2;                 // 16 KB
constexpr int SM_V = SM_K + FST * FK * THD * 2;           // + 24 KB
constexpr int SM_BAR = SM_V + FST * FK * THD * 2;         // + 24 KB (P never touches smem: it goes back into TMEM)
constexpr int kAttnSmem = SM_BAR + 256 + 1024;            // 65 KB; two CTAs per SM (256 TMEM columns each)

__device__ __forceinline__ float ex2(float x) {   // one MUFU; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA pipe (Cody-Waite split + degree-3 polynomial, relative error ~1e-4 — far below the bf16 rounding P gets
// next): floor(x) through the 1.5 * 2^23 magic add in round-down mode, 2^frac by Horner, the integer part added straight
// into the exponent field.  The softmax is MUFU-bound (one ex2 per score, 4 lanes per clock per scheduler); evaluating a
// share of the scores here runs both pipes side by side.  Inputs <= ~8 (lazy maximum); -inf (masked) -> 2^-127 ~ 0.
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -127.f);
  float t;
  asm("add.rm.f32 %0, %1, 0f4B400000;" : "=f"(t) : "f"(x));        // floor(x) + 12582912
  const float xf = x - (t - 12582912.f);                              // in [0, 1)
  float p = fmaf(0.077119089663028717f, xf, 0.227564394474029541f);
  p = fmaf(p, xf, 0.695146143436431885f);
  p = fmaf(p, xf, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// kVarlen (sequence packing): several samples share one row of S frames; seg_start[b, i] is the
// first position of the sample that position i belongs to (padding: i itself), so query i sees keys
// seg_start[b, i] <= j <= i.  A query tile then starts at the key block that holds the first row's segment start
// instead of block 0 — packed short samples cost what they would cost alone — and the blocks that straddle a
// segment start are masked like the diagonal ones.
// kTmemO: the output tile accumulates in TMEM across the key blocks (PV_j issued with accumulate) instead of being
// folded into registers block by block.  The softmax threads then neither read PV_j back (64 tcgen05.ld columns + 64
// FMAs per row and block) nor hold 64 accumulator registers; the running maximum is LAZY: a row keeps exponentiating
// against its current reference until the block maximum exceeds it by more than 2^8, and only then is the TMEM
// accumulator rescaled in place (rare after the first blocks; P stays <= 256, exact in bf16's exponent range).
// kPoly: every kPoly-th score of a row (0: none) takes ex2_fma instead of the MUFU.
template <bool kVarlen, bool kTmemO, int kPoly = 0>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ o, float* __restrict__ lse, int S,
                   int H, int KV, int64_t ldo, float scale_log2, const int32_t* __restrict__ seg_start) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // FST
  uint64_t* kv_empty = bars + 5;      // FST
  uint64_t* s_full = bars + 9;        // 2
  uint64_t* p_full = bars + 11;       // 2 (one arrival per softmax warp)
  uint64_t* pv_full = bars + 13;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  static_assert(FST <= 4, "barrier slots");

  // programmatic dependent launch (common.cuh): no global memory is touched before pdl_wait().  Packed rows read
  // seg_start before the prologue, so they wait first; otherwise the prologue overlaps the previous kernel's tail.
  if (kVarlen) pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid in longest-processing-time-first order: all (head, batch) CTAs of the last query tile (most key blocks)
  // are dispatched before any of the next-to-last one, ... — the causal triangle then packs to ~95 % of the SM-time
  const int nq = (S + TQ - 1) / TQ;
  const int per_q = (int)gridDim.x / nq;       // H * B
  const int qb = nq - 1 - (int)blockIdx.x / per_q;
  const int hb = (int)blockIdx.x % per_q;
  const int h = hb % H, b = hb / H;
  const int kvh = h / (H / KV);
  const int q0 = qb * TQ;
  const int blk_end = min(2 * (qb + 1), (S + FK - 1) / FK);   // 64-key blocks up to the diagonal
  const int blk0 = kVarlen ? seg_start[(int64_t)b * S + q0] / FK : 0;   // first block any row of the tile can see
  const int nblk = blk_end - blk0;                            // blocks are walked as jj = 0..nblk-1, key block blk0 + jj

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < FST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_full[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_S = 0, COL_PV = 128;
  if (!kVarlen) pdl_wait();
  pdl_trigger();

  // warps 0 and 1 keep warp-uniform control flow and elect one lane per TMA / tcgen05 issue (see tc::elect_one)
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, TQ * THD * 2);
      tma_load_3d(smem + SM_Q, &tmQ, q_full, h * THD, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < nblk; ++j) {
      const int ks = j % FST;
      mbar_wait(&kv_empty[ks], ((j / FST) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[ks], 2 * FK * THD * 2);
        tma_load_3d(smem + SM_K + ks * (FK * THD * 2), &tmK, &kv_full[ks], kvh * THD, (blk0 + j) * FK, b);
        tma_load_3d(smem + SM_V + ks * (FK * THD * 2), &tmV, &kv_full[ks], kvh * THD, (blk0 + j) * FK, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(TQ, FK, 0, 0);    // S = Q K^T : both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(TQ, THD, 0, 1);  // PV = P V  : V is MN-major ([key][hd] rows)
      const uint32_t sq = smem_u32(smem + SM_Q);
      // one elected thread runs the whole issue loop: warp-wide waits + elect + __syncwarp around every MMA block cost
      // ~65 cycles each even when the barrier is open, and the tensor pipe's queue is too shallow to hide that behind
      // 51-cycle N=64 MMAs (tools/ubench/mma_rate.cu k4)
      if (elect_one()) {
        int ks_s = 0, ks_v = 0;                    // K/V ring stage of the next S block resp. the next PV block
        uint32_t ks_ph = 0;                        // parity of kv_full[ks_s]
        auto issue_s = [&](int j) {
          const int st = j & 1;
          const uint32_t sk = smem_u32(smem + SM_K + ks_s * (FK * THD * 2));
          const uint64_t ad = make_smem_desc(sq, 16, 1024), bd = make_smem_desc(sk, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < THD / 16; ++kk)
            umma_bf16(tmem_base + COL_S + st * FK, ad + (uint64_t)((kk * 32) >> 4), bd + (uint64_t)((kk * 32) >> 4),
                      idesc_s, kk ? 1u : 0u);
          umma_commit(&s_full[st]);
          if (++ks_s == FST) { ks_s = 0; ks_ph ^= 1; }
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_s(0);
        for (int j = 0; j < nblk; ++j) {
          const int st = j & 1;
          if (j + 1 < nblk) {
            mbar_wait(&kv_full[ks_s], ks_ph);
            tc_fence_after();
            issue_s(j + 1);  // S buffer st^1 was drained before p_full(j-1) completed (waited last iteration)
          }
          mbar_wait(&p_full[st], (j >> 1) & 1);
          tc_fence_after();
          const uint32_t sv = smem_u32(smem + SM_V + ks_v * (FK * THD * 2));
          // P: read straight from TMEM (bf16 pairs written back over the S columns, 8 cells per 16 keys) — no smem
          // round trip.  V: 16 key-rows x 128 B per K step.  The in-order tensor pipe runs PV(j) before S(j+2), which
          // overwrites the same columns.
          const uint64_t bd = make_smem_desc(sv, 64 * FK * 2, 1024);
#pragma unroll
          for (int kk = 0; kk < FK / 16; ++kk)
            umma_bf16_ts(tmem_base + COL_PV + (kTmemO ? 0 : st * THD), tmem_base + COL_S + st * FK + kk * 8,
                         bd + (uint64_t)((kk * 16 * 128) >> 4), idesc_pv, (kk || (kTmemO && j > 0)) ? 1u : 0u);
          umma_commit(&pv_full[st]);
          umma_commit(&kv_empty[ks_v]);
          if (++ks_v == FST) ks_v = 0;
        }
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;           // row inside the tile == TMEM lane
    const int qi = q0 + r;                    // query index inside the sequence
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // packing: this row's first visible key and the latest segment start inside the tile (blocks below it need masks)
    const int lo = kVarlen ? seg_start[(int64_t)b * S + min(qi, S - 1)] : 0;
    const int lo_max = kVarlen ? seg_start[(int64_t)b * S + min(q0 + TQ - 1, S - 1)] : 0;
    float m = -INFINITY, l = 0.f, corr_prev = 1.f;
    float oacc[THD];
#pragma unroll
    for (int i = 0; i < THD; ++i) oacc[i] = 0.f;

    auto fold = [&](int j, float corr) {      // oacc = oacc * corr + PV_j
      const int st = j & 1;
      mbar_wait(&pv_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[THD];
      __syncwarp();
#pragma unroll
      for (int c = 0; c < THD; c += 32) tmem_ld32(lane_addr + COL_PV + st * THD + c, v + c);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < THD; ++i) oacc[i] = fmaf(oacc[i], corr, __uint_as_float(v[i]));
    };

    // ---- kTmemO: lazy running maximum, accumulator rescaled in TMEM only when a row's reference moves
    float m_used = -INFINITY;                  // the reference the row's exponentials are taken against (log2 domain)
    auto block_tmem = [&](int j, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      const int st = j & 1;
      const int kbase = (blk0 + j) * FK;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + COL_S + st * FK;
      uint32_t v[FK];
      __syncwarp();
#pragma unroll
      for (int c = 0; c < FK; c += 32) tmem_ld32(s_addr + c, v + c);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < FK; ++i) {
        if (DIAG && ((kbase + i > qi) || (kVarlen && kbase + i < lo))) v[i] = 0xff800000u;   // -inf
        m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
      }
      const float mrs = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
      // move the reference when the block maximum is more than 2^8 above it (or when the row had seen no key yet)
      const bool move = mrs > m_used + 8.f;                       // (-inf + 8 = -inf: any finite maximum moves it)
      const float m_new = move ? mrs : m_used;
      const float corr = move ? ex2(m_used - m_new) : 1.f;        // exp2(-inf) = 0 on an accumulator that is still 0
      if (j > 0 && __any_sync(0xffffffffu, move)) {
        // PV(j-1) — the last MMA that wrote the accumulator — has completed; PV(j) cannot start before this warp
        // arrives on p_full(j) below
        mbar_wait(&pv_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t ov[32];
#pragma unroll
        for (int c = 0; c < THD; c += 32) {
          __syncwarp();
          tmem_ld32(lane_addr + COL_PV + c, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
          tmem_st16(lane_addr + COL_PV + c, ov);
          tmem_st16(lane_addr + COL_PV + c + 16, ov + 16);
        }
      }
      m_used = m_new;
      const float mref = (m_new == -INFINITY) ? 0.f : m_new;      // packing: a row that still sees no key
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < FK; c += 32) {
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = fmaf(__uint_as_float(v[c + i]), scale_log2, -mref);
          p[i] = (kPoly > 0 && (i % (kPoly > 0 ? kPoly : 1)) == (kPoly > 0 ? kPoly : 1) - 1) ? ex2_fma(x) : ex2(x);
          rs4[i & 3] += p[i];
        }
        uint32_t pw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pw[i] = pack_bf16(p[2 * i], p[2 * i + 1]);
        tmem_st16(s_addr + (c >> 1), pw);
      }
      l = l * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    };

    // one KV block of the online softmax; DIAG is the block on the causal diagonal (the only one that needs masks)
    auto block = [&](int j, auto diag_tag) {
      if (kTmemO) { block_tmem(j, diag_tag); return; }
      constexpr bool DIAG = decltype(diag_tag)::value;
      const int st = j & 1;
      const int kbase = (blk0 + j) * FK;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + COL_S + st * FK;
      // the whole 128-column S row in registers: 4 tcgen05.ld in flight, ONE wait (a single TMEM round trip)
      uint32_t v[FK];
      __syncwarp();
#pragma unroll
      for (int c = 0; c < FK; c += 32) tmem_ld32(s_addr + c, v + c);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < FK; ++i) {
        if (DIAG && ((kbase + i > qi) || (kVarlen && kbase + i < lo))) v[i] = 0xff800000u;   // -inf
        m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
      }
      const float mr = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));   // scale > 0: scaling commutes with max
      float mx = fmaxf(m, mr * scale_log2);
      // packing: a row whose segment starts later has seen no key yet (max still -inf): exponentials against 0 instead
      // (exp2(-inf - 0) = 0 for the masked scores and for the correction of the still-empty accumulator)
      const float mref = (kVarlen && DIAG && mx == -INFINITY) ? 0.f : mx;
      const float corr = ex2(m - mref);
      m = mx;
      mx = mref;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < FK; c += 32) {
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          p[i] = ex2(fmaf(__uint_as_float(v[c + i]), scale_log2, -mx));   // exp2(-inf) = 0 for masked keys
          rs4[i & 3] += p[i];
        }
        uint32_t pw[16];                      // bf16 pairs (key, key+1): the A operand of the PV MMA, kept in TMEM
#pragma unroll
        for (int i = 0; i < 16; ++i) pw[i] = pack_bf16(p[2 * i], p[2 * i + 1]);
        tmem_st16(s_addr + (c >> 1), pw);
      }
      l = l * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      if (j > 0) fold(j - 1, corr_prev);
      corr_prev = corr;
    };
    // blocks that reach past the first query row of the tile (kbase + 63 > q0) need the causal mask: the last two;
    // packing: so do the blocks that start below the latest segment start of the tile
    if (!kVarlen) {
      const int nfull = min(nblk, q0 / FK);
      for (int j = 0; j < nfull; ++j) block(j, std::false_type{});
      for (int j = nfull; j < nblk; ++j) block(j, std::true_type{});
    } else {
      for (int j = 0; j < nblk; ++j) {
        const int kb = (blk0 + j) * FK;
        if (kb < lo_max || kb + FK - 1 > q0) block(j, std::true_type{}); else block(j, std::false_type{});
      }
    }
    if (kTmemO) {
      mbar_wait(&pv_full[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);      // the last PV: the accumulator is complete
      tc_fence_after();
      uint32_t ov[THD];
      __syncwarp();
#pragma unroll
      for (int c = 0; c < THD; c += 32) tmem_ld32(lane_addr + COL_PV + c, ov + c);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < THD; ++i) oacc[i] = __uint_as_float(ov[i]);
      m = m_used;
    } else {
      fold(nblk - 1, corr_prev);
    }
    if (qi < S) {
      const float inv = 1.f / l;
      bf16* op = o + ((int64_t)b * S + qi) * ldo + (int64_t)h * THD;
#pragma unroll
      for (int c = 0; c < THD; c += 8) {
        uint4 w;
        w.x = pack_bf16(oacc[c + 0] * inv, oacc[c + 1] * inv); w.y = pack_bf16(oacc[c + 2] * inv, oacc[c + 3] * inv);
        w.z = pack_bf16(oacc[c + 4] * inv, oacc[c + 5] * inv); w.w = pack_bf16(oacc[c + 6] * inv, oacc[c + 7] * inv);
        *reinterpret_cast<uint4*>(op + c) = w;
      }
      lse[((int64_t)b * H + h) * S + qi] = m * kLn2 + logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// =====================================================================================================================
// Backward on tcgen05.  Two kernels that each own their output rows (no atomics, deterministic):
//   dQ kernel   : CTA = 128 queries of one head; per 64-key sub-block  S = Q K^T, dP = dO V^T  (TMEM, triple
//                 buffered)  ->  dS = P o (dP - delta) * scale, written back over the dP columns as bf16 pairs  ->
//                 dQ += dS K with dS read as the A operand from TMEM, accumulated in TMEM over the whole key range.
//   dKdV kernel : CTA = 128 keys of one kv head; per (q head of the group, 64-query sub-block)  S^T = K Q^T,
//                 dP^T = V dO^T  ->  P^T, dS^T back into the same TMEM columns  ->  dV += P^T dO,  dK += dS^T Q.
// P / dS never go through shared memory: the 128 x 64 MMAs with both operands in smem were operand-feed-bound
// (6 KB of smem per 32 math cycles) and the compute warps' tile stores competed for the same 128 B/clk
// (ncu: pipe_tc 56 % busy for 25 % of math) — with A in TMEM the backward went from 286 to 239 us at B=2, S=2048.
// The MMA warp runs NB = 3 sub-blocks ahead of the 8 compute warps: S/dP(u + 3) is issued right behind the MMAs that
// consume P / dS of sub-block u from the same TMEM buffer — tcgen05.mma instructions of one thread execute in issue order,
// so no "buffer free" barrier (and no pipe drain per sub-block) is needed.
// The Q / dO / K / V tiles are loaded once per use by TMA as [rows][64] 128B-swizzled tiles and serve BOTH as a
// K-major operand (rows = M or N, hd = K) and as an MN-major B operand (hd = N, rows = K): same bytes, two descriptors.
// Sixteen compute warps: the four warps of a TMEM lane quadrant take 16 columns of a sub-block each, and every warp
// issues the TMEM loads of sub-block u+1 before the math of sub-block u.  Measured (tools/ubench/tmem_rate.cu): one warp
// reads TMEM at 16-19 B/clk, 8 warps in lock step at ~110 B/clk/SM (the loads of a 64 KB sub-block then take as long as its
// 8192 exponentials on the MUFU, and the two phases did not overlap), 16 warps at 155-190 B/clk/SM.  No row reductions
// are needed (lse and delta come from the forward / the delta pre-pass).
constexpr int kBwdCW = 16;                      // compute warps: four per TMEM lane quadrant
constexpr int kBwdThreads = 64 + kBwdCW * 32;   // warp0 TMA, warp1 MMA, warps 2..17 compute
constexpr int SUB = 64;                         // columns (keys resp. queries) per pipelined sub-block
constexpr int PC = SUB / (kBwdCW / 4);          // columns of a sub-block per compute warp (16)
static_assert(PC == 16, "the compute warps use the x16 TMEM load / x8 store shapes");

// D[128 x 64] (+)= A[TMEM, 128 lanes x 64 k as bf16 pairs] * B, B = [64 rows x 64] smem tile used MN-major.
// Each of the four compute warps of a lane quadrant wrote its 16 k (8 cells) at the start of its own 16-column quarter:
// k-step kk lives at column 16 * kk.
__device__ __forceinline__ void issue_atmem_bmn(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem, bool accumulate_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, THD, 0, 1);
  const uint64_t bd = make_smem_desc(b_smem, 64 * 128 * 2, 1024);
#pragma unroll
  for (int kk = 0; kk < SUB / 16; ++kk)
    umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(kk * PC), bd + (uint64_t)((kk * 16 * 128) >> 4), idesc,
                 (accumulate_first || kk) ? 1u : 0u);
}

// D[128 x 64] = A[128 x 64] * B[64 x 64]^T, both K-major tiles (reduction over head_dim)
__device__ __forceinline__ void issue_nt_64(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem) {
  constexpr uint32_t idesc = make_idesc_bf16(128, SUB, 0, 0);
  const uint64_t ad = make_smem_desc(a_smem, 16, 1024), bd = make_smem_desc(b_smem, 16, 1024);
#pragma unroll
  for (int kk = 0; kk < THD / 16; ++kk)
    umma_bf16(d_tmem, ad + (uint64_t)((kk * 32) >> 4), bd + (uint64_t)((kk * 32) >> 4), idesc, kk ? 1u : 0u);
}

// D[128 x 64] = A[TMEM: 128 lanes x 64 k as bf16 pairs in 32 consecutive columns] * B[64 x 64]^T (K-major smem tile)
__device__ __forceinline__ void issue_ts_nt_64(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem) {
  constexpr uint32_t idesc = make_idesc_bf16(128, SUB, 0, 0);
  const uint64_t bd = make_smem_desc(b_smem, 16, 1024);
#pragma unroll
  for (int kk = 0; kk < THD / 16; ++kk)
    umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(kk * 8), bd + (uint64_t)((kk * 32) >> 4), idesc, kk ? 1u : 0u);
}

// Copies this thread's 16-element slice of row `r` of a [128][64] bf16 tile (128B-swizzled, as TMA wrote it) into 8 TMEM
// columns as (k, k+1) pairs: the A-operand layout of a K-major A in TMEM.  `part` selects elements 16*part .. 16*part+15.
__device__ __forceinline__ void smem_row_to_tmem_a(const uint8_t* tile, int r, int part, uint32_t taddr) {
  const uint8_t* row = tile + r * 128;
  const uint4 c0 = *reinterpret_cast<const uint4*>(row + (((2 * part) ^ (r & 7)) << 4));
  const uint4 c1 = *reinterpret_cast<const uint4*>(row + (((2 * part + 1) ^ (r & 7)) << 4));
  const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  tmem_st8(taddr, w);
}

constexpr int DQ_Q = 0;                              // 16 KB  Q tile
constexpr int DQ_DO = DQ_Q + TQ * THD * 2;           // 16 KB  dO tile
constexpr int KST = 3;                               // K/V (resp. Q/dO) TMA ring depth of the backward kernels
constexpr int NB = 3;                                // TMEM S/dP sub-block buffers: the MMA warp runs two ahead
constexpr int DQ_K = DQ_DO + TQ * THD * 2;           // KST x 16 KB (128-key tiles = two sub-blocks each)
constexpr int DQ_V = DQ_K + KST * TK * THD * 2;      // KST x 16 KB
constexpr int DQ_BAR = DQ_V + KST * TK * THD * 2;    // (dS goes back into TMEM, not through smem)
constexpr int kDqSmem = DQ_BAR + 256 + 1024;

template <bool kVarlen>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                      const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dq, int S,
                      int H, int KV, int64_t lddq, float scale, const float* __restrict__ rope_cache,
                      const int32_t* __restrict__ seg_start) {
  PROF_ONLY(const long long p_entry = clock64();)
  if (kVarlen) pdl_wait();            // (see attn_fwd_tc_kernel)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_BAR);
  uint64_t* q_full = bars;            // Q + dO landed
  uint64_t* kv_full = bars + 1;       // [KST]
  uint64_t* kv_empty = bars + 4;      // [KST]
  uint64_t* sdp_full = bars + 7;      // [NB] S and dP sub-block ready in TMEM
  uint64_t* ds_full = bars + 13;      // [NB] dS written back over the dP columns as bf16 pairs (8 warp arrivals)
  uint64_t* qa_full = bars + 16;      // Q and dO copied into TMEM as A operands (kBwdCW warp arrivals)
  uint64_t* acc_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  static_assert(KST == 3 && NB == 3, "barrier slots");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (S + TQ - 1) / TQ;            // 1-D grid, longest query tiles first (see the forward kernel)
  const int per_q = (int)gridDim.x / nq;
  const int qb = nq - 1 - (int)blockIdx.x / per_q;
  const int hb = (int)blockIdx.x % per_q;
  const int h = hb % H, b = hb / H;
  const int kvh = h / (H / KV);
  const int q0 = qb * TQ;
  // packing: the first 128-key tile any row of this query tile can see (see attn_fwd_tc_kernel)
  const int jt0 = kVarlen ? seg_start[(int64_t)b * S + q0] / TK : 0;
  const int nblk = qb + 1 - jt0;
  const int nsub = 2 * nblk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_full[i], kBwdCW);
    }
    mbar_init(qa_full, kBwdCW);
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (!kVarlen) pdl_wait();
  pdl_trigger();
  PROF_ONLY(if (blockIdx.x == 0 && threadIdx.x == 64) g_attn_prof[64] = clock64() - p_entry;)   // setup done
  // 192 (S) + 192 (dP) + 64 (dQ) + 32 (Q as A operand) + 32 (dO as A operand) = 512 columns
  constexpr uint32_t COL_S = 0, COL_DP = NB * SUB, COL_DQ = 2 * NB * SUB, COL_QA = COL_DQ + THD, COL_DOA = COL_QA + THD / 2;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * TQ * THD * 2);
      tma_load_3d(smem + DQ_Q, &tmQ, q_full, h * THD, q0, b);
      tma_load_3d(smem + DQ_DO, &tmDO, q_full, h * THD, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < nblk; ++j) {
      const int st = j % KST;
      mbar_wait(&kv_empty[st], ((j / KST) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[st], 2 * TK * THD * 2);
        tma_load_3d(smem + DQ_K + st * (TK * THD * 2), &tmK, &kv_full[st], kvh * THD, (jt0 + j) * TK, b);
        tma_load_3d(smem + DQ_V + st * (TK * THD * 2), &tmV, &kv_full[st], kvh * THD, (jt0 + j) * TK, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    {
      PROF_ONLY(long long pw0 = 0, pw1 = 0, pw2 = 0; const long long pt0 = clock64();)
      // S/dP(u + NB) overwrites the TMEM buffer of sub-block u.  No "buffer free" barrier is needed: it is issued after
      // dQ(u) — the last reader of that buffer — and tcgen05.mma instructions of one thread execute in issue order (the
      // forward kernel relies on the same property); the compute warps' loads of S/dP(u) completed before ds_full(u).
      // The whole loop runs in ONE elected thread: a warp-wide mbarrier wait + elect + __syncwarp per MMA block costs
      // ~65 cycles even when the barrier is already open (tools/ubench/mma_rate.cu k4: 668 vs 516 cycles per sub-block),
      // and the tensor pipe's queue is too shallow to ride that out with 51-cycle N=64 MMAs.
      if (elect_one()) {
        // ring positions are carried incrementally (no divisions by KST / NB on the issue path): the producer side `p`
        // walks S/dP(u + NB), the consumer side `c` walks dQ(u)
        int p_st = 0, p_bb = 0, p_hk = 0;
        uint32_t p_stph = 0;
        auto issue_sdp = [&]() {
          if (p_hk == 0) {
            PROF_WAIT(pw0, mbar_wait(&kv_full[p_st], p_stph))
            tc_fence_after();
          }
          const uint32_t sk = smem_u32(smem + DQ_K + p_st * (TK * THD * 2)) + p_hk * (SUB * 128);
          const uint32_t sv = smem_u32(smem + DQ_V + p_st * (TK * THD * 2)) + p_hk * (SUB * 128);
          // Q and dO are read as A operands out of TMEM (copied there once per CTA): an N=64 MMA with both operands in
          // smem is operand-fetch-bound (6 KB per 32 math cycles: 51 cycles measured, 41 with A in TMEM) and shares the
          // 128 B/clk with the TMA writes of the K/V ring
          issue_ts_nt_64(tmem_base + COL_S + p_bb * SUB, tmem_base + COL_QA, sk);     // S  = Q K_sub^T
          issue_ts_nt_64(tmem_base + COL_DP + p_bb * SUB, tmem_base + COL_DOA, sv);   // dP = dO V_sub^T
          umma_commit(&sdp_full[p_bb]);
          if (++p_bb == NB) p_bb = 0;
          if (p_hk) { if (++p_st == KST) { p_st = 0; p_stph ^= 1; } }
          p_hk ^= 1;
        };
        mbar_wait(qa_full, 0);
        tc_fence_after();
        PROF_ONLY(if (blockIdx.x == 0) g_attn_prof[65] = clock64() - p_entry;)   // Q/dO in TMEM
        for (int u = 0; u < NB && u < nsub; ++u) issue_sdp();
        PROF_ONLY(if (blockIdx.x == 0) g_attn_prof[66] = clock64() - p_entry;)   // first S/dP issued
        int c_st = 0, c_bb = 0, c_hk = 0;
        uint32_t c_bbph = 0;
        for (int u = 0; u < nsub; ++u) {
          PROF_WAIT(pw2, mbar_wait(&ds_full[c_bb], c_bbph))
          tc_fence_after();
          const uint32_t sk = smem_u32(smem + DQ_K + c_st * (TK * THD * 2)) + c_hk * (SUB * 128);
          issue_atmem_bmn(tmem_base + COL_DQ, tmem_base + COL_DP + c_bb * SUB, sk, u > 0);   // dQ += dS K_sub (dS in TMEM)
          if (c_hk) { umma_commit(&kv_empty[c_st]); if (++c_st == KST) c_st = 0; }
          if (u == nsub - 1) umma_commit(acc_full);
          if (u + NB < nsub) issue_sdp();
          if (++c_bb == NB) { c_bb = 0; c_bbph ^= 1; }
          c_hk ^= 1;
        }
      }
      __syncwarp();
      PROF_ONLY(if (blockIdx.x == 0 && lane == 0) {
        g_attn_prof[0] = clock64() - pt0; g_attn_prof[1] = pw0; g_attn_prof[2] = pw1; g_attn_prof[3] = pw2; g_attn_prof[4] = nsub;
      })
    }
  } else {
    PROF_ONLY(long long pw0 = 0, pw1 = 0, pw2 = 0; const long long pt0 = clock64();)
    const int quad = warp & 3, part = (warp - 2) >> 2;          // TMEM lane quadrant, 16-column quarter of a sub-block
    const int r = quad * 32 + lane, qi = q0 + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * PC);
    const float scale_log2 = scale * kLog2e;
    const int64_t li = ((int64_t)b * H + h) * S + qi;
    const float L2 = (qi < S) ? lse[li] * kLog2e : INFINITY;   // +inf => P = 0 for rows past the sequence end
    const float Dls = (qi < S) ? delta[li] * scale : 0.f;
    const int lo = kVarlen ? seg_start[(int64_t)b * S + min(qi, S - 1)] : 0;
    const int lo_max = kVarlen ? seg_start[(int64_t)b * S + min(q0 + TQ - 1, S - 1)] : 0;
    int fb = 0, bb = 0;                            // TMEM buffer of the sub-block being fetched / computed
    uint32_t fph = 0;
    int kbase = jt0 * TK + part * PC;              // this warp's first key of the sub-block being computed
    // S / dP of the next sub-block into registers (asynchronous: the caller waits with tmem_ld_wait())
    auto fetch = [&](uint32_t (&s_)[PC], uint32_t (&d_)[PC]) {
      PROF_WAIT(pw0, mbar_wait(&sdp_full[fb], fph))
      tc_fence_after();
      __syncwarp();
      tmem_ld16(lane_addr + COL_S + fb * SUB, s_);
      tmem_ld16(lane_addr + COL_DP + fb * SUB, d_);
      if (++fb == NB) { fb = 0; fph ^= 1; }
    };
    auto compute = [&](const uint32_t (&s_)[PC], const uint32_t (&d_)[PC], auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      PROF_ONLY(const long long pm0 = clock64();)
      float f[PC];
#pragma unroll
      for (int i = 0; i < PC; ++i) {
        float x = fmaf(__uint_as_float(s_[i]), scale_log2, -L2);
        if (DIAG && ((kbase + i > qi) || (kVarlen && kbase + i < lo))) x = -INFINITY;
        f[i] = ex2(x) * fmaf(__uint_as_float(d_[i]), scale, -Dls);     // P * (dP - delta) * scale
      }
      // dS as bf16 pairs (key, key+1) over the first 8 of this warp's own 16 dP columns: the A operand of dQ += dS K
      uint32_t dw[PC / 2];
#pragma unroll
      for (int j = 0; j < PC / 2; ++j) dw[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
      tmem_st8(lane_addr + COL_DP + bb * SUB, dw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[bb]);
      PROF_ONLY(pw2 += clock64() - pm0;)
      if (++bb == NB) bb = 0;
      kbase += SUB;
    };
    auto step = [&](const uint32_t (&s_)[PC], const uint32_t (&d_)[PC], bool diag) {
      if (diag) compute(s_, d_, std::true_type{});
      else compute(s_, d_, std::false_type{});
    };
    // Q and dO rows -> TMEM (A operands of S = Q K^T and dP = dO V^T), once per CTA
    mbar_wait(q_full, 0);
    smem_row_to_tmem_a(smem + DQ_Q, r, part, tmem_base + ((uint32_t)(quad * 32) << 16) + COL_QA + (uint32_t)(part * 8));
    smem_row_to_tmem_a(smem + DQ_DO, r, part, tmem_base + ((uint32_t)(quad * 32) << 16) + COL_DOA + (uint32_t)(part * 8));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(qa_full);
    uint32_t sa[PC], da[PC], sb[PC], db[PC];
    fetch(sa, da);
    tmem_ld_wait();
    PROF_ONLY(if (blockIdx.x == 0 && threadIdx.x == 64) g_attn_prof[67] = clock64() - p_entry;)   // first S/dP in registers
    for (int t = 0; t < nblk; ++t) {               // two sub-blocks per 128-key tile; the last tile is the diagonal one
      const bool last = t == nblk - 1;
      fetch(sb, db);
      step(sa, da, last || (kVarlen && kbase - part * PC < lo_max));
      PROF_WAIT(pw1, tmem_ld_wait())
      if (!last) fetch(sa, da);
      step(sb, db, last || (kVarlen && kbase - part * PC < lo_max));
      PROF_WAIT(pw1, tmem_ld_wait())
    }
    PROF_ONLY(if (blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 7)) {
      long long* o = g_attn_prof + (warp == 2 ? 8 : 16);
      o[0] = clock64() - pt0; o[1] = pw0; o[2] = pw1; o[3] = pw2;
    })
    PROF_ONLY(if (blockIdx.x == 0 && threadIdx.x == 64) g_attn_prof[68] = clock64() - p_entry;)   // last dS written
    mbar_wait(acc_full, 0);
    tc_fence_after();
    PROF_ONLY(if (blockIdx.x == 0 && threadIdx.x == 64) g_attn_prof[69] = clock64() - p_entry;)   // dQ complete
    {
      uint32_t v[PC];
      __syncwarp();
      tmem_ld16(lane_addr + COL_DQ, v);
      tmem_ld_wait();
      if (qi < S) {
        bf16* dp_ = dq + ((int64_t)b * S + qi) * lddq + (int64_t)h * THD + part * PC;
#pragma unroll
        for (int c = 0; c < PC; c += 8) {
          uint32_t w[4];
          w[0] = pack_bf16(__uint_as_float(v[c + 0]), __uint_as_float(v[c + 1]));
          w[1] = pack_bf16(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
          w[2] = pack_bf16(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5]));
          w[3] = pack_bf16(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7]));
          // gradient w.r.t. the un-rotated q: the inverse rotation of the (bf16) gradient, as csm_rope(inverse) would do
          // (packing: positions restart at every segment)
          if (rope_cache) rope_rotate8(w, rope_cache + (int64_t)(qi - lo) * THD, (part * PC + c) >> 1, -1.f);
          *reinterpret_cast<uint4*>(dp_ + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  PROF_ONLY(if (blockIdx.x == 0 && threadIdx.x == 64) g_attn_prof[70] = clock64() - p_entry;)   // exit
}

constexpr int DK_K = 0;                              // 16 KB  K tile
constexpr int DK_V = DK_K + TK * THD * 2;            // 16 KB  V tile
constexpr int DK_Q = DK_V + TK * THD * 2;            // KST x 16 KB Q tiles (128 queries = two sub-blocks each)
constexpr int DK_DO = DK_Q + KST * TQ * THD * 2;     // KST x 16 KB dO tiles
constexpr int DK_LD = DK_DO + KST * TQ * THD * 2;    // 2 x (128 lse*log2e + 128 delta*scale) floats
                                                     // (P^T / dS^T go back into TMEM, not through smem)
constexpr int DK_BAR = DK_LD + 2 * 256 * 4;
constexpr int kDkSmem = DK_BAR + 256 + 1024;

// kVarlen: key j is seen by queries j <= i < seg_end[b, j] (the end of j's sample), so a key tile walks the query tiles
// only up to the end of its last key's segment and masks the sub-blocks that reach past its first key's segment end.
template <bool kVarlen>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dk,
                        bf16* __restrict__ dv, int S, int H, int KV, int64_t lddk, int64_t lddv, float scale,
                        const float* __restrict__ rope_cache, const int32_t* __restrict__ seg_start,
                        const int32_t* __restrict__ seg_end) {
  if (kVarlen) pdl_wait();            // (see attn_fwd_tc_kernel)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DK_BAR);
  uint64_t* kv_full = bars;           // K + V landed
  uint64_t* qd_full = bars + 1;       // [KST] Q_i + dO_i tiles landed
  uint64_t* qd_empty = bars + 4;      // [KST]
  uint64_t* sdp_full = bars + 7;      // [NB] S^T and dP^T sub-block ready
  uint64_t* pt_full = bars + 13;      // [NB] P^T and dS^T written back into the buffer as bf16 (8 warp arrivals)
  uint64_t* acc_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (S + TK - 1) / TK;           // 1-D grid, longest first: key tile 0 meets every query tile
  const int per_k = (int)gridDim.x / nkb;      // KV * B
  const int kvb = (int)blockIdx.x / per_k;
  const int kvh = ((int)blockIdx.x % per_k) % KV, b = ((int)blockIdx.x % per_k) / KV;
  const int rep = H / KV;
  const int k0 = kvb * TK;
  // packing: query tiles end with the segment of the tile's last key
  const int nqb = kVarlen ? (seg_end[(int64_t)b * S + min(k0 + TK - 1, S - 1)] + TQ - 1) / TQ : (S + TQ - 1) / TQ;
  const int nq_iter = nqb - kvb;
  const int total = rep * nq_iter;      // 128-query tiles
  const int nsub = 2 * total;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1);
      mbar_init(&sdp_full[i], 1);
      mbar_init(&pt_full[i], kBwdCW);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (!kVarlen) pdl_wait();
  pdl_trigger();
  constexpr uint32_t COL_ST = 0, COL_DPT = NB * SUB, COL_DK = 2 * NB * SUB, COL_DV = 2 * NB * SUB + THD;   // 512 columns

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * TK * THD * 2);
      tma_load_3d(smem + DK_K, &tmK, kv_full, kvh * THD, k0, b);
      tma_load_3d(smem + DK_V, &tmV, kv_full, kvh * THD, k0, b);
    }
    __syncwarp();
    int hh = 0, qi_ = 0;                          // it = hh * nq_iter + qi_ without divisions
    for (int it = 0; it < total; ++it) {
      const int st = it % KST;
      const int h = kvh * rep + hh, qb = kvb + qi_;
      mbar_wait(&qd_empty[st], ((it / KST) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&qd_full[st], 2 * TQ * THD * 2);
        tma_load_3d(smem + DK_Q + st * (TQ * THD * 2), &tmQ, &qd_full[st], h * THD, qb * TQ, b);
        tma_load_3d(smem + DK_DO + st * (TQ * THD * 2), &tmDO, &qd_full[st], h * THD, qb * TQ, b);
      }
      __syncwarp();
      if (++qi_ == nq_iter) { qi_ = 0; ++hh; }
    }
  } else if (warp == 1) {
    {
      PROF_ONLY(long long pw0 = 0, pw1 = 0, pw2 = 0; const long long pt0 = clock64();)
      const uint32_t sk = smem_u32(smem + DK_K), sv = smem_u32(smem + DK_V);
      // S^T/dP^T(u + NB) is issued right after dV/dK(u), the last readers of the buffer it overwrites: in-order execution
      // of one thread's tcgen05.mma instructions replaces a "buffer free" barrier; one elected thread runs the whole
      // loop (see the dQ kernel)
      if (elect_one()) {
        int p_st = 0, p_bb = 0, p_hq = 0;          // ring positions carried incrementally (see the dQ kernel)
        uint32_t p_stph = 0;
        auto issue_sdp = [&]() {
          if (p_hq == 0) {
            PROF_WAIT(pw0, mbar_wait(&qd_full[p_st], p_stph))
            tc_fence_after();
          }
          const uint32_t sq = smem_u32(smem + DK_Q + p_st * (TQ * THD * 2)) + p_hq * (SUB * 128);
          const uint32_t sdo = smem_u32(smem + DK_DO + p_st * (TQ * THD * 2)) + p_hq * (SUB * 128);
          issue_nt_64(tmem_base + COL_ST + p_bb * SUB, sk, sq);      // S^T  = K Q_sub^T
          issue_nt_64(tmem_base + COL_DPT + p_bb * SUB, sv, sdo);    // dP^T = V dO_sub^T
          umma_commit(&sdp_full[p_bb]);
          if (++p_bb == NB) p_bb = 0;
          if (p_hq) { if (++p_st == KST) { p_st = 0; p_stph ^= 1; } }
          p_hq ^= 1;
        };
        mbar_wait(kv_full, 0);
        tc_fence_after();
        for (int u = 0; u < NB && u < nsub; ++u) issue_sdp();
        int c_st = 0, c_bb = 0, c_hq = 0;
        uint32_t c_bbph = 0;
        for (int u = 0; u < nsub; ++u) {
          PROF_WAIT(pw2, mbar_wait(&pt_full[c_bb], c_bbph))
          tc_fence_after();
          const uint32_t sq = smem_u32(smem + DK_Q + c_st * (TQ * THD * 2)) + c_hq * (SUB * 128);
          const uint32_t sdo = smem_u32(smem + DK_DO + c_st * (TQ * THD * 2)) + c_hq * (SUB * 128);
          // A operands straight from TMEM: P^T / dS^T were written back (bf16 pairs) over the S^T / dP^T columns they
          // came from — no shared-memory round trip for the 128 x 64 tiles, and the MMAs read only B from smem
          issue_atmem_bmn(tmem_base + COL_DV, tmem_base + COL_ST + c_bb * SUB, sdo, u > 0);    // dV += P^T dO
          issue_atmem_bmn(tmem_base + COL_DK, tmem_base + COL_DPT + c_bb * SUB, sq, u > 0);    // dK += dS^T Q
          if (c_hq) { umma_commit(&qd_empty[c_st]); if (++c_st == KST) c_st = 0; }
          if (u == nsub - 1) umma_commit(acc_full);
          if (u + NB < nsub) issue_sdp();
          if (++c_bb == NB) { c_bb = 0; c_bbph ^= 1; }
          c_hq ^= 1;
        }
      }
      __syncwarp();
      PROF_ONLY(if (blockIdx.x == 0 && lane == 0) {
        g_attn_prof[32] = clock64() - pt0; g_attn_prof[33] = pw0; g_attn_prof[34] = pw1; g_attn_prof[35] = pw2; g_attn_prof[36] = nsub;
      })
    }
  } else {
    PROF_ONLY(long long pw0 = 0, pw1 = 0, pw2 = 0, pw3 = 0; const long long pt0 = clock64();)
    const int quad = warp & 3, part = (warp - 2) >> 2;          // TMEM lane quadrant, 16-column quarter of a sub-block
    const int r = quad * 32 + lane, kj = k0 + r;               // this thread's key
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t lane_addr = lane_base + (uint32_t)(part * PC);
    const float scale_log2 = scale * kLog2e;
    const int ctid = threadIdx.x - 64;                         // 0..511 among the compute warps
    // per 128-query tile: lse*log2e (+inf past the sequence end => P = 0) and delta*scale, staged through smem once
    // per CTA by the first 256 compute threads; the global load for tile it+1 is issued one tile early so its latency
    // hides behind tile it's math (raw value only: scaling it here would make the thread wait for the load right away)
    const bool stager = ctid < 256;
    const float ld_mul = ctid < 128 ? kLog2e : scale;
    const int hi = kVarlen ? seg_end[(int64_t)b * S + min(kj, S - 1)] : S;              // this key's last query + 1
    const int hi_min = kVarlen ? seg_end[(int64_t)b * S + min(k0, S - 1)] : S;          // earliest segment end of the tile
    // (head, query tile) of the 128-query tile whose lse / delta this thread fetches next: carried incrementally — the
    // compute warps are a latency-bound serial chain, an integer division per sub-block shows up 1:1 in the kernel time
    auto load_ld = [&](int hh, int qt) -> float {
      const int q = (kvb + qt) * TQ + (ctid & 127);
      const int64_t li = ((int64_t)b * H + (kvh * rep + hh)) * S + q;
      if (q >= S) return ctid < 128 ? INFINITY : 0.f;
      return __ldg((ctid < 128 ? lse : delta) + li);
    };
    float ld_next = stager ? load_ld(0, 0) : 0.f;
    int fb = 0, bb = 0;                            // TMEM buffer of the sub-block being fetched / computed
    uint32_t fph = 0;
    // S^T / dP^T of the next sub-block into registers (asynchronous: the caller waits with tmem_ld_wait())
    auto fetch = [&](uint32_t (&s_)[PC], uint32_t (&d_)[PC]) {
      PROF_WAIT(pw0, mbar_wait(&sdp_full[fb], fph))
      tc_fence_after();
      __syncwarp();
      tmem_ld16(lane_addr + COL_ST + fb * SUB, s_);
      tmem_ld16(lane_addr + COL_DPT + fb * SUB, d_);
      if (++fb == NB) { fb = 0; fph ^= 1; }
    };
    auto compute = [&](const uint32_t (&s_)[PC], const uint32_t (&d_)[PC], int qb, int hq, const float* sLD, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      PROF_ONLY(const long long pm0 = clock64();)
      const int col0 = hq * SUB + part * PC;                   // first query column (inside the 128-query tile)
      const int qbase = qb * TQ + col0;
      const float4* L4 = reinterpret_cast<const float4*>(sLD + col0);
      const float4* D4 = reinterpret_cast<const float4*>(sLD + 128 + col0);
      // bf16 pairs (q, q+1) back into the first 8 of this warp's own 16 columns of each buffer
      uint32_t pw[PC / 2], dw[PC / 2];
#pragma unroll
      for (int i4 = 0; i4 < PC / 4; ++i4) {
        const float4 Lq = L4[i4], Dq = D4[i4];
        const float Ls[4] = {Lq.x, Lq.y, Lq.z, Lq.w}, Ds[4] = {Dq.x, Dq.y, Dq.z, Dq.w};
        float pf[4], df[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int i = i4 * 4 + t;
          float x = fmaf(__uint_as_float(s_[i]), scale_log2, -Ls[t]);
          if (DIAG && ((kj > qbase + i) || (kVarlen && qbase + i >= hi))) x = -INFINITY;
          pf[t] = ex2(x);
          df[t] = pf[t] * fmaf(__uint_as_float(d_[i]), scale, -Ds[t]);
        }
        pw[2 * i4] = pack_bf16(pf[0], pf[1]);
        pw[2 * i4 + 1] = pack_bf16(pf[2], pf[3]);
        dw[2 * i4] = pack_bf16(df[0], df[1]);
        dw[2 * i4 + 1] = pack_bf16(df[2], df[3]);
      }
      tmem_st8(lane_addr + COL_ST + bb * SUB, pw);
      tmem_st8(lane_addr + COL_DPT + bb * SUB, dw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pt_full[bb]);
      PROF_ONLY(pw2 += clock64() - pm0;)
      if (++bb == NB) bb = 0;
    };
    auto step = [&](const uint32_t (&s_)[PC], const uint32_t (&d_)[PC], int qb, int hq, const float* sLD, bool diag) {
      if (diag) compute(s_, d_, qb, hq, sLD, std::true_type{});
      else compute(s_, d_, qb, hq, sLD, std::false_type{});
    };
    uint32_t sa[PC], da[PC], sb[PC], db[PC];
    fetch(sa, da);
    tmem_ld_wait();
    int par = 0;                                  // which half of the lse/delta staging buffer this tile uses
    for (int hh = 0; hh < rep; ++hh) {
      for (int qt = 0; qt < nq_iter; ++qt) {
        const int qb = kvb + qt;
        const bool last = hh == rep - 1 && qt == nq_iter - 1;
        float* sLD = reinterpret_cast<float*>(smem + DK_LD) + par * 256;
        par ^= 1;
        if (stager) sLD[ctid] = ld_next * ld_mul;
        PROF_WAIT(pw3, named_bar_sync(1, kBwdCW * 32))
        if (stager) {
          const bool wrap = qt + 1 == nq_iter;
          if (!last) ld_next = load_ld(wrap ? hh + 1 : hh, wrap ? 0 : qt + 1);
        }
        fetch(sb, db);
        step(sa, da, qb, 0, sLD, qt == 0 || (kVarlen && qb * TQ + SUB > hi_min));
        PROF_WAIT(pw1, tmem_ld_wait())
        if (!last) fetch(sa, da);
        step(sb, db, qb, 1, sLD, qt == 0 || (kVarlen && qb * TQ + 2 * SUB > hi_min));
        PROF_WAIT(pw1, tmem_ld_wait())
      }
    }
    PROF_ONLY(if (blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 7)) {
      long long* o = g_attn_prof + (warp == 2 ? 40 : 48);
      o[0] = clock64() - pt0; o[1] = pw0; o[2] = pw1; o[3] = pw2; o[4] = pw3;
    })
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // quarters 0, 1 store the two 32-column halves of the dK rows, quarters 2, 3 those of the dV rows
    {
      const int c0 = (part & 1) * 32;
      bf16* outp = part < 2 ? dk + ((int64_t)b * S + kj) * lddk + (int64_t)kvh * THD + c0
                            : dv + ((int64_t)b * S + kj) * lddv + (int64_t)kvh * THD + c0;
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(lane_base + (part < 2 ? COL_DK : COL_DV) + c0, v);
      tmem_ld_wait();
      if (kj < S) {
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 8) {
          uint32_t w[4];
          w[0] = pack_bf16(__uint_as_float(v[c8 + 0]), __uint_as_float(v[c8 + 1]));
          w[1] = pack_bf16(__uint_as_float(v[c8 + 2]), __uint_as_float(v[c8 + 3]));
          w[2] = pack_bf16(__uint_as_float(v[c8 + 4]), __uint_as_float(v[c8 + 5]));
          w[3] = pack_bf16(__uint_as_float(v[c8 + 6]), __uint_as_float(v[c8 + 7]));
          if (rope_cache && part < 2)                                                                // dK only
            rope_rotate8(w, rope_cache + (int64_t)(kj - (kVarlen ? seg_start[(int64_t)b * S + kj] : 0)) * THD,
                         (c0 + c8) >> 1, -1.f);
          *reinterpret_cast<uint4*>(outp + c8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// forward variant: 0 = output folded into registers block by block, 1 = output accumulated in TMEM with a lazy maximum,
// 2 / 3 / 4 = variant 1 with every 4th / 3rd / 2nd exponential on the FMA pipe (ex2_fma)
static std::atomic<int> g_attn_fwd_variant{1};
void attn_tc_set_fwd_variant(int v) { g_attn_fwd_variant.store(v); }

bool attn_tc_supported(int S, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k,
                       const void* v, const void* o) {
  return hd == THD && S >= 128 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && aligned16(q) &&
         aligned16(k) && aligned16(v) && aligned16(o);
}

int attn_fwd_tc_launch(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int H, int KV,
                       int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, const int32_t* seg_start,
                       cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = encode_tmap_bf16(&tq, q, (uint64_t)H * THD, S, B, ldq, (uint64_t)S * ldq, TQ))) return rc;
  if ((rc = encode_tmap_bf16(&tk, k, (uint64_t)KV * THD, S, B, ldk, (uint64_t)S * ldk, FK))) return rc;
  if ((rc = encode_tmap_bf16(&tv, v, (uint64_t)KV * THD, S, B, ldv, (uint64_t)S * ldv, FK))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<true, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) { set_error("attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
    configured = true;
  }
  dim3 grid(((S + TQ - 1) / TQ) * H * B);
  const int variant = g_attn_fwd_variant.load();
  cudaError_t le = cudaSuccess;
#define FWD(V, T, P) le = launch_k(attn_fwd_tc_kernel<V, T, P>, grid, dim3(kAttnThreads), kAttnSmem, st, 1, tq, tk, tv, (bf16*)o, \
                              lse, S, H, KV, ldo, scale * kLog2e, seg_start)
  if (seg_start) { if (variant == 0) FWD(true, false, 0); else if (variant == 1) FWD(true, true, 0); else FWD(true, true, 4); }
  else if (variant == 0) FWD(false, false, 0);
  else if (variant == 1) FWD(false, true, 0);
  else if (variant == 2) FWD(false, true, 4);
  else if (variant == 3) FWD(false, true, 3);
  else FWD(false, true, 2);
#undef FWD
  if (le != cudaSuccess) { set_error("attn_fwd_tc: launch failed: %s", cudaGetErrorString(le)); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("attn_fwd_tc");
  return CSM_OK;
}

int attn_delta_launch(const void* o, const void* dout, float* delta, int B, int S, int H, int64_t ldo,
                      cudaStream_t st);

int attn_bwd_tc_launch(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                       void* dq, void* dk, void* dv, float* delta, int B, int S, int H, int KV, int64_t ldq, int64_t ldk,
                       int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                       const float* rope_cache, const int32_t* seg_start, const int32_t* seg_end, cudaStream_t st) {
  CSM_REQUIRE(aligned16(dout) && aligned16(dq) && aligned16(dk) && aligned16(dv) && lddq % 8 == 0 && lddk % 8 == 0 &&
                  lddv % 8 == 0,
              CSM_ERR_ALIGN, "attn_bwd_tc: misaligned gradient buffers");
  int rc = attn_delta_launch(o, dout, delta, B, S, H, ldo, st);
  if (rc) return rc;
  CUtensorMap tq, tk, tv, tdo;
  if ((rc = encode_tmap_bf16(&tq, q, (uint64_t)H * THD, S, B, ldq, (uint64_t)S * ldq, TQ))) return rc;
  if ((rc = encode_tmap_bf16(&tk, k, (uint64_t)KV * THD, S, B, ldk, (uint64_t)S * ldk, TK))) return rc;
  if ((rc = encode_tmap_bf16(&tv, v, (uint64_t)KV * THD, S, B, ldv, (uint64_t)S * ldv, TK))) return rc;
  if ((rc = encode_tmap_bf16(&tdo, dout, (uint64_t)H * THD, S, B, ldo, (uint64_t)S * ldo, TQ))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkSmem);
    if (e != cudaSuccess) { set_error("attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
    configured = true;
  }
  dim3 gq(((S + TQ - 1) / TQ) * H * B);
  const bool varlen = seg_start != nullptr && seg_end != nullptr;
  const float* lse_c = lse;
  const float* delta_c = delta;
  const int32_t* no_seg = nullptr;
  cudaError_t le;
  if (varlen)
    le = launch_k(attn_bwd_dq_tc_kernel<true>, gq, dim3(kBwdThreads), kDqSmem, st, 1, tq, tk, tv, tdo, lse_c, delta_c, (bf16*)dq,
             S, H, KV, lddq, scale, rope_cache, seg_start);
  else
    le = launch_k(attn_bwd_dq_tc_kernel<false>, gq, dim3(kBwdThreads), kDqSmem, st, 1, tq, tk, tv, tdo, lse_c, delta_c,
             (bf16*)dq, S, H, KV, lddq, scale, rope_cache, no_seg);
  if (le != cudaSuccess) { set_error("attn_bwd_dq_tc: launch failed: %s", cudaGetErrorString(le)); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("attn_bwd_dq_tc");
  dim3 gk(((S + TK - 1) / TK) * KV * B);
  if (varlen)
    le = launch_k(attn_bwd_dkdv_tc_kernel<true>, gk, dim3(kBwdThreads), kDkSmem, st, 1, tq, tk, tv, tdo, lse_c, delta_c,
             (bf16*)dk, (bf16*)dv, S, H, KV, lddk, lddv, scale, rope_cache, seg_start, seg_end);
  else
    le = launch_k(attn_bwd_dkdv_tc_kernel<false>, gk, dim3(kBwdThreads), kDkSmem, st, 1, tq, tk, tv, tdo, lse_c, delta_c,
             (bf16*)dk, (bf16*)dv, S, H, KV, lddk, lddv, scale, rope_cache, no_seg, no_seg);
  if (le != cudaSuccess) { set_error("attn_bwd_dkdv_tc: launch failed: %s", cudaGetErrorString(le)); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("attn_bwd_dkdv_tc");
  return CSM_OK;
}

}  // namespace csm

#ifdef CSM_ATTN_PROF
extern "C" int csm_debug_attn_prof(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, csm::g_attn_prof, sizeof(long long) * 128);
}
#endif
