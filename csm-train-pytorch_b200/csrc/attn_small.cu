// Short-sequence causal GQA attention (seq <= 32, head_dim 64 / 128): the depth decoder's shape — 32 codebook
// positions per selected frame, 8 query heads over 2 KV heads of 128 (reference model.py:28-42, 184).
//
// One CTA per (sequence, kv head), one warp per query head of the group.  Everything is a 32 x 32 x HD problem, far too
// small for tcgen05 (128-row MMAs): the warps use mma.sync.m16n8k16 (bf16 in, fp32 accumulate) on operands staged in
// XOR-swizzled shared memory and read with ldmatrix:
//   forward : S = Q K^T (scores of a row stay in the quad's registers: no online softmax), P = softmax, O = P V
//   backward: phase 1 (warp = query head)   S, P, dP = dO V^T, delta = rowsum(P o dP), dS = P o (dP - delta) * scale,
//                                            dQ = dS K;  P and dS parked in smem as bf16
//             phase 2 (warp = 32 head-dim columns) dV = sum_h P_h^T dO_h, dK = sum_h dS_h^T Q_h — the sum over the
//                                            group's query heads happens inside the CTA: deterministic, no atomics.
// Tiles above the causal diagonal (rows 0-15 x keys 16-31) are skipped in every product.
// Semantics as attn_simt.cu: F.scaled_dot_product_attention(is_causal=True) + torchtune's GQA expansion.
#include "common.cuh"

namespace csm {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxS = 32;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D (16x8 fp32) += A (16x16 bf16, row) * B (16x8 bf16, col)
__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// [32][HD] bf16 tile, 16-byte chunks XOR-swizzled with the row (ldmatrix reads 8 rows x 16 B: conflict-free)
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int col) {   // col multiple of 8
  constexpr int CH = HD / 8;
  return (uint32_t)(row * (HD * 2) + ((((col >> 3) ^ (row & 7)) & (CH - 1)) << 4));
}
// [32][32] bf16 tile (P, dS): 64-byte rows, chunk ^ ((row >> 1) & 3)
__device__ __forceinline__ uint32_t ptile_off(int row, int col) {  // col multiple of 8
  return (uint32_t)(row * 64 + ((((col >> 3) ^ (row >> 1)) & 3) << 4));
}

template <int HD>
__device__ __forceinline__ void stage_tile(uint8_t* dst, const bf16* src, int64_t ld, int rows, int tid, int nthr) {
  constexpr int CH = HD / 8;
  for (int idx = tid; idx < kMaxS * CH; idx += nthr) {
    const int r = idx / CH, c = idx % CH;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < rows) val = ld_nc16(src + (int64_t)r * ld + c * 8);
    *reinterpret_cast<uint4*>(dst + tile_off<HD>(r, c * 8)) = val;
  }
}

// A fragments of a row-major [32][HD] tile: rows m0..m0+15, k columns k0..k0+15
template <int HD>
__device__ __forceinline__ void load_a(uint32_t tile, int m0, int k0, int lane, uint32_t* a) {
  ldsm_x4(tile + tile_off<HD>(m0 + (lane & 15), k0 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
}
// B fragments for TWO n-tiles (n0..n0+15) of B[k][n] = T[n][k]  (T row-major [n][HD], e.g. K in Q K^T): k0..k0+15
template <int HD>
__device__ __forceinline__ void load_b_nk(uint32_t tile, int n0, int k0, int lane, uint32_t* b) {
  ldsm_x4(tile + tile_off<HD>(n0 + (lane >> 4) * 8 + (lane & 7), k0 + ((lane >> 3) & 1) * 8), b[0], b[1], b[2], b[3]);
}
// B fragments for TWO n-tiles (n0..n0+15) of B[k][n] = T[k][n]  (T row-major [k][HD], e.g. V in P V): k0..k0+15
template <int HD>
__device__ __forceinline__ void load_b_kn(uint32_t tile, int k0, int n0, int lane, uint32_t* b) {
  ldsm_x4_t(tile + tile_off<HD>(k0 + ((lane >> 3) & 1) * 8 + (lane & 7), n0 + (lane >> 4) * 8), b[0], b[1], b[2], b[3]);
}
// A fragments of the TRANSPOSE of a [32 q][32 key] P tile: A[m = key][k = q], keys m0..m0+15, queries k0..k0+15
__device__ __forceinline__ void load_a_pt(uint32_t ptile, int m0, int k0, int lane, uint32_t* a) {
  const int mi = lane >> 3;
  ldsm_x4_t(ptile + ptile_off(k0 + (mi >> 1) * 8 + (lane & 7), m0 + (mi & 1) * 8), a[0], a[1], a[2], a[3]);
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// S = Q K^T for one head: s[mt][nt][4], m-tile 0 x key tiles 2,3 (above the diagonal) are left at zero and never used
template <int HD>
__device__ __forceinline__ void qk_scores(uint32_t sq, uint32_t sk, int lane, float (&s)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < HD; k0 += 16) {
    uint32_t a0[4], a1[4], b01[4], b23[4];
    load_a<HD>(sq, 0, k0, lane, a0);
    load_a<HD>(sq, 16, k0, lane, a1);
    load_b_nk<HD>(sk, 0, k0, lane, b01);
    load_b_nk<HD>(sk, 16, k0, lane, b23);
    mma16816(s[0][0], a0, b01[0], b01[1]);
    mma16816(s[0][1], a0, b01[2], b01[3]);
    mma16816(s[1][0], a1, b01[0], b01[1]);
    mma16816(s[1][1], a1, b01[2], b01[3]);
    mma16816(s[1][2], a1, b23[0], b23[1]);
    mma16816(s[1][3], a1, b23[2], b23[3]);
  }
}

// C fragments of a 16 x 16 block (two adjacent n-tiles) -> A fragment (bf16) of the same block
__device__ __forceinline__ void c_to_a(const float* c_lo, const float* c_hi, uint32_t* a) {
  a[0] = pack_bf16(c_lo[0], c_lo[1]);
  a[1] = pack_bf16(c_lo[2], c_lo[3]);
  a[2] = pack_bf16(c_hi[0], c_hi[1]);
  a[3] = pack_bf16(c_hi[2], c_hi[3]);
}

// out[32][HD] = X[32 x 32 as A fragments xa[mt][kt]] * T  (T row-major [32 keys][HD]); written as bf16 with `mul`
template <int HD>
__device__ __forceinline__ void x_times_tile(const uint32_t (&xa)[2][2][4], uint32_t st, int lane, bf16* out, int64_t ld,
                                             int rows, const float* mul_lo, const float* mul_hi) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int n0 = 0; n0 < HD; n0 += 32) {
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      uint32_t b01[4], b23[4];
      load_b_kn<HD>(st, kt * 16, n0, lane, b01);
      load_b_kn<HD>(st, kt * 16, n0 + 16, lane, b23);
      if (kt == 0) {                                   // rows 0-15 only see keys 0-15
        mma16816(acc[0][0], xa[0][0], b01[0], b01[1]);
        mma16816(acc[0][1], xa[0][0], b01[2], b01[3]);
        mma16816(acc[0][2], xa[0][0], b23[0], b23[1]);
        mma16816(acc[0][3], xa[0][0], b23[2], b23[3]);
      }
      mma16816(acc[1][0], xa[1][kt], b01[0], b01[1]);
      mma16816(acc[1][1], xa[1][kt], b01[2], b01[3]);
      mma16816(acc[1][2], xa[1][kt], b23[0], b23[1]);
      mma16816(acc[1][3], xa[1][kt], b23[2], b23[3]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = n0 + nt * 8 + 2 * t;
        if (r_lo < rows)
          *reinterpret_cast<uint32_t*>(out + (int64_t)r_lo * ld + col) =
              pack_bf16(acc[mt][nt][0] * mul_lo[mt], acc[mt][nt][1] * mul_lo[mt]);
        if (r_hi < rows)
          *reinterpret_cast<uint32_t*>(out + (int64_t)r_hi * ld + col) =
              pack_bf16(acc[mt][nt][2] * mul_hi[mt], acc[mt][nt][3] * mul_hi[mt]);
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(128)
attn_small_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      bf16* __restrict__ o, float* __restrict__ lse, int S, int H, int KV, int64_t ldq, int64_t ldk,
                      int64_t ldv, int64_t ldo, float scale) {
  constexpr int TB = kMaxS * HD * 2;               // bytes of one [32][HD] tile
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x / KV, kvh = blockIdx.x % KV, rep = H / KV;
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* Ks = smem;
  uint8_t* Vs = smem + TB;
  uint8_t* Qs = smem + 2 * TB;                     // [rep] tiles
  stage_tile<HD>(Ks, k + (int64_t)b * S * ldk + (int64_t)kvh * HD, ldk, S, tid, nthr);
  stage_tile<HD>(Vs, v + (int64_t)b * S * ldv + (int64_t)kvh * HD, ldv, S, tid, nthr);
  for (int r = 0; r < rep; ++r)
    stage_tile<HD>(Qs + r * TB, q + (int64_t)b * S * ldq + (int64_t)(kvh * rep + r) * HD, ldq, S, tid, nthr);
  __syncthreads();
  if (warp >= rep) return;
  const int h = kvh * rep + warp;
  const int g = lane >> 2, t = lane & 3;
  float s[2][4][4];
  qk_scores<HD>(smem_addr(Qs + warp * TB), smem_addr(Ks), lane, s);
  const float sc = scale * kLog2e;
  uint32_t pa[2][2][4];
  float inv_lo[2], inv_hi[2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (mt == 0 && nt >= 2) continue;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int key = nt * 8 + 2 * t + i;
        s[mt][nt][i] = key <= r_lo ? s[mt][nt][i] * sc : -INFINITY;
        s[mt][nt][2 + i] = key <= r_hi ? s[mt][nt][2 + i] * sc : -INFINITY;
        m_lo = fmaxf(m_lo, s[mt][nt][i]);
        m_hi = fmaxf(m_hi, s[mt][nt][2 + i]);
      }
    }
    m_lo = quad_max(m_lo);
    m_hi = quad_max(m_hi);
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool live = !(mt == 0 && nt >= 2);
        s[mt][nt][i] = live ? exp2f(s[mt][nt][i] - m_lo) : 0.f;
        s[mt][nt][2 + i] = live ? exp2f(s[mt][nt][2 + i] - m_hi) : 0.f;
        l_lo += s[mt][nt][i];
        l_hi += s[mt][nt][2 + i];
      }
    }
    l_lo = quad_sum(l_lo);
    l_hi = quad_sum(l_hi);
    inv_lo[mt] = 1.f / l_lo;
    inv_hi[mt] = 1.f / l_hi;
    if (t == 0) {
      if (r_lo < S) lse[((int64_t)b * H + h) * S + r_lo] = m_lo * kLn2 + logf(l_lo);
      if (r_hi < S) lse[((int64_t)b * H + h) * S + r_hi] = m_hi * kLn2 + logf(l_hi);
    }
    c_to_a(s[mt][0], s[mt][1], pa[mt][0]);
    c_to_a(s[mt][2], s[mt][3], pa[mt][1]);
  }
  x_times_tile<HD>(pa, smem_addr(Vs), lane, o + (int64_t)b * S * ldo + (int64_t)h * HD, ldo, S, inv_lo, inv_hi);
}

template <int HD>
__global__ void __launch_bounds__(128)
attn_small_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      const float* __restrict__ lse, const bf16* __restrict__ dout, bf16* __restrict__ dq,
                      bf16* __restrict__ dk, bf16* __restrict__ dv, int S, int H, int KV, int64_t ldq, int64_t ldk,
                      int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale) {
  constexpr int TB = kMaxS * HD * 2;
  constexpr int PB = kMaxS * kMaxS * 2;            // bytes of one [32][32] bf16 tile
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x / KV, kvh = blockIdx.x % KV, rep = H / KV;
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  uint8_t* Ks = smem;
  uint8_t* Vs = smem + TB;
  uint8_t* Qs = smem + 2 * TB;                     // [rep]
  uint8_t* Ds = Qs + rep * TB;                     // [rep]  dO
  uint8_t* Ps = Ds + rep * TB;                     // [rep]  P  (bf16 [q][key])
  uint8_t* Gs = Ps + rep * PB;                     // [rep]  dS
  stage_tile<HD>(Ks, k + (int64_t)b * S * ldk + (int64_t)kvh * HD, ldk, S, tid, nthr);
  stage_tile<HD>(Vs, v + (int64_t)b * S * ldv + (int64_t)kvh * HD, ldv, S, tid, nthr);
  for (int r = 0; r < rep; ++r) {
    const int h = kvh * rep + r;
    stage_tile<HD>(Qs + r * TB, q + (int64_t)b * S * ldq + (int64_t)h * HD, ldq, S, tid, nthr);
    stage_tile<HD>(Ds + r * TB, dout + (int64_t)b * S * ldo + (int64_t)h * HD, ldo, S, tid, nthr);
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;

  // ---- phase 1: warp = query head
  if (warp < rep) {
    const int h = kvh * rep + warp;
    float s[2][4][4], dp[2][4][4];
    qk_scores<HD>(smem_addr(Qs + warp * TB), smem_addr(Ks), lane, s);
    qk_scores<HD>(smem_addr(Ds + warp * TB), smem_addr(Vs), lane, dp);      // dP = dO V^T: same operand shapes
    const float sc = scale * kLog2e;
    uint32_t dsa[2][2][4];
    const float one[2] = {1.f, 1.f};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
      const float L_lo = r_lo < S ? lse[((int64_t)b * H + h) * S + r_lo] * kLog2e : INFINITY;   // +inf => P = 0
      const float L_hi = r_hi < S ? lse[((int64_t)b * H + h) * S + r_hi] * kLog2e : INFINITY;
      float d_lo = 0.f, d_hi = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int key = nt * 8 + 2 * t + i;
          const bool live = !(mt == 0 && nt >= 2);
          s[mt][nt][i] = (live && key <= r_lo) ? exp2f(fmaf(s[mt][nt][i], sc, -L_lo)) : 0.f;
          s[mt][nt][2 + i] = (live && key <= r_hi) ? exp2f(fmaf(s[mt][nt][2 + i], sc, -L_hi)) : 0.f;
          d_lo += s[mt][nt][i] * dp[mt][nt][i];
          d_hi += s[mt][nt][2 + i] * dp[mt][nt][2 + i];
        }
      }
      d_lo = quad_sum(d_lo);                       // delta = rowsum(P o dP) == rowsum(dO o O)
      d_hi = quad_sum(d_hi);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        // park P and dS (bf16) for phase 2; dS also stays in registers as the A operand of dQ = dS K
        const uint32_t p_lo = pack_bf16(s[mt][nt][0], s[mt][nt][1]), p_hi = pack_bf16(s[mt][nt][2], s[mt][nt][3]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          dp[mt][nt][i] = s[mt][nt][i] * (dp[mt][nt][i] - d_lo) * scale;
          dp[mt][nt][2 + i] = s[mt][nt][2 + i] * (dp[mt][nt][2 + i] - d_hi) * scale;
        }
        const uint32_t g_lo = pack_bf16(dp[mt][nt][0], dp[mt][nt][1]), g_hi = pack_bf16(dp[mt][nt][2], dp[mt][nt][3]);
        const int col = nt * 8 + 2 * t;
        uint8_t* pt = Ps + warp * PB;
        uint8_t* gt = Gs + warp * PB;
        *reinterpret_cast<uint32_t*>(pt + ptile_off(r_lo, col & ~7) + (col & 7) * 2) = p_lo;
        *reinterpret_cast<uint32_t*>(pt + ptile_off(r_hi, col & ~7) + (col & 7) * 2) = p_hi;
        *reinterpret_cast<uint32_t*>(gt + ptile_off(r_lo, col & ~7) + (col & 7) * 2) = g_lo;
        *reinterpret_cast<uint32_t*>(gt + ptile_off(r_hi, col & ~7) + (col & 7) * 2) = g_hi;
      }
      c_to_a(dp[mt][0], dp[mt][1], dsa[mt][0]);
      c_to_a(dp[mt][2], dp[mt][3], dsa[mt][1]);
    }
    x_times_tile<HD>(dsa, smem_addr(Ks), lane, dq + (int64_t)b * S * lddq + (int64_t)h * HD, lddq, S, one, one);
  }
  __syncthreads();

  // ---- phase 2: warp = 32 head-dim columns; dV = sum_h P_h^T dO_h, dK = sum_h dS_h^T Q_h
  for (int n0 = warp * 32; n0 < HD; n0 += nwarp * 32) {
    float ak[2][4][4], av[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) { ak[mt][nt][i] = 0.f; av[mt][nt][i] = 0.f; }
    for (int r = 0; r < rep; ++r) {
      const uint32_t pt = smem_addr(Ps + r * PB), gt = smem_addr(Gs + r * PB);
      const uint32_t qt = smem_addr(Qs + r * TB), dt = smem_addr(Ds + r * TB);
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {             // 16 queries per step
        uint32_t bq01[4], bq23[4], bd01[4], bd23[4];
        load_b_kn<HD>(qt, kt * 16, n0, lane, bq01);
        load_b_kn<HD>(qt, kt * 16, n0 + 16, lane, bq23);
        load_b_kn<HD>(dt, kt * 16, n0, lane, bd01);
        load_b_kn<HD>(dt, kt * 16, n0 + 16, lane, bd23);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {           // 16 keys per tile
          if (mt == 1 && kt == 0) continue;        // keys 16-31 never see queries 0-15
          uint32_t ap[4], ag[4];
          load_a_pt(pt, mt * 16, kt * 16, lane, ap);
          load_a_pt(gt, mt * 16, kt * 16, lane, ag);
          mma16816(av[mt][0], ap, bd01[0], bd01[1]);
          mma16816(av[mt][1], ap, bd01[2], bd01[3]);
          mma16816(av[mt][2], ap, bd23[0], bd23[1]);
          mma16816(av[mt][3], ap, bd23[2], bd23[3]);
          mma16816(ak[mt][0], ag, bq01[0], bq01[1]);
          mma16816(ak[mt][1], ag, bq01[2], bq01[3]);
          mma16816(ak[mt][2], ag, bq23[0], bq23[1]);
          mma16816(ak[mt][3], ag, bq23[2], bq23[3]);
        }
      }
    }
    bf16* dk0 = dk + (int64_t)b * S * lddk + (int64_t)kvh * HD;
    bf16* dv0 = dv + (int64_t)b * S * lddv + (int64_t)kvh * HD;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = n0 + nt * 8 + 2 * t;
        if (r_lo < S) {
          *reinterpret_cast<uint32_t*>(dk0 + (int64_t)r_lo * lddk + col) = pack_bf16(ak[mt][nt][0], ak[mt][nt][1]);
          *reinterpret_cast<uint32_t*>(dv0 + (int64_t)r_lo * lddv + col) = pack_bf16(av[mt][nt][0], av[mt][nt][1]);
        }
        if (r_hi < S) {
          *reinterpret_cast<uint32_t*>(dk0 + (int64_t)r_hi * lddk + col) = pack_bf16(ak[mt][nt][2], ak[mt][nt][3]);
          *reinterpret_cast<uint32_t*>(dv0 + (int64_t)r_hi * lddv + col) = pack_bf16(av[mt][nt][2], av[mt][nt][3]);
        }
      }
    }
  }
}

size_t fwd_smem_bytes(int hd, int rep) { return (size_t)(2 + rep) * kMaxS * hd * 2; }
size_t bwd_smem_bytes(int hd, int rep) {
  return (size_t)(2 + 2 * rep) * kMaxS * hd * 2 + (size_t)2 * rep * kMaxS * kMaxS * 2;
}

}  // namespace

bool attn_small_supported(int S, int H, int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                          const void* q, const void* k, const void* v, const void* o) {
  if (S > kMaxS || (hd != 64 && hd != 128) || KV <= 0 || H % KV != 0) return false;
  const int rep = H / KV;
  if (rep > 4) return false;             // one warp per query head of the group, 4 warps per CTA
  if ((ldq | ldk | ldv | ldo) & 7) return false;
  return aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o);
}

int attn_fwd_small_launch(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int H,
                          int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale,
                          cudaStream_t st) {
  const int rep = H / KV;
  const unsigned grid = (unsigned)(B * KV);
  const size_t smem = fwd_smem_bytes(hd, rep);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_small_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)fwd_smem_bytes(64, 4));
    cudaFuncSetAttribute(attn_small_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)fwd_smem_bytes(128, 4));
    attr_set = true;
  }
#define LAUNCH(HD) launch_k(attn_small_fwd_kernel<HD>, dim3(grid), dim3(128), smem, st, 1, (const bf16*)q, (const bf16*)k, \
      (const bf16*)v, (bf16*)o, lse, S, H, KV, ldq, ldk, ldv, ldo, scale)
  if (hd == 64) LAUNCH(64); else LAUNCH(128);
#undef LAUNCH
  CSM_CHECK_LAUNCH("attn_small_fwd");
  return CSM_OK;
}

int attn_bwd_small_launch(const void* q, const void* k, const void* v, const void* o, const float* lse,
                          const void* dout, void* dq, void* dk, void* dv, int B, int S, int H, int KV, int hd,
                          int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk,
                          int64_t lddv, float scale, cudaStream_t st) {
  (void)o;   // delta is recomputed as rowsum(P o dP)
  CSM_REQUIRE(((lddq | lddk | lddv) & 7) == 0 && aligned16(dq) && aligned16(dk) && aligned16(dv) && aligned16(dout),
              CSM_ERR_ALIGN, "attn_bwd (short-sequence kernel): gradients must be 16-byte aligned");
  const int rep = H / KV;
  const unsigned grid = (unsigned)(B * KV);
  const size_t smem = bwd_smem_bytes(hd, rep);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_small_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bwd_smem_bytes(64, 4));
    cudaFuncSetAttribute(attn_small_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bwd_smem_bytes(128, 4));
    attr_set = true;
  }
#define LAUNCH(HD) launch_k(attn_small_bwd_kernel<HD>, dim3(grid), dim3(128), smem, st, 1, (const bf16*)q, (const bf16*)k, \
      (const bf16*)v, (const float*)lse, (const bf16*)dout, (bf16*)dq, (bf16*)dk, (bf16*)dv, S, H, KV, ldq, ldk, ldv, ldo, \
      lddq, lddk, lddv, scale)
  if (hd == 64) LAUNCH(64); else LAUNCH(128);
#undef LAUNCH
  CSM_CHECK_LAUNCH("attn_small_bwd");
  return CSM_OK;
}

}  // namespace csm
