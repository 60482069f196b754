// K5: causal GQA flash attention for head_dim 64 (the CSM-1B backbone: 32 q heads / 8 kv heads).
// Forward, backward-dQ and backward-dK/dV kernels; 64x64 tiles, 4 warps, bf16 mma.sync.m16n8k16 with fp32
// accumulation, online softmax in the exp2 domain, cp.async double-buffered K/V (resp. Q/dO) tiles in
// XOR-swizzled shared memory, ldmatrix operand fetch.  No atomics: dQ and dK/dV are produced by two kernels that
// each own their output rows, so results are deterministic.
// (Decoder attention — head_dim 128 over 32 positions — and the tiny test model run attn_simt.cu.)
#include "common.cuh"

namespace csm {

namespace {

constexpr int BR = 64, BC = 64, HD = 64, NTHR = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A 64 x 64 bf16 tile in shared memory: row r is 128 bytes, its 16-byte chunk c lives at chunk (c ^ (r & 7)).
struct Tile {
  bf16* base;
  __device__ __forceinline__ uint32_t chunk_addr(int r, int c) const {
    return smem_addr(base) + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
  }
  __device__ __forceinline__ bf16* chunk_ptr(int r, int c) const { return base + r * 64 + ((c ^ (r & 7)) << 3); }
};

// global [rows, ld] (starting at row0, column col0) -> tile; rows >= nrows are zero-filled
__device__ __forceinline__ void load_tile(const Tile& t, const bf16* g, int64_t ld, int row0, int nrows, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * NTHR;  // 512 chunks
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < nrows;
    const bf16* src = g + (int64_t)(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(t.chunk_ptr(r, c), src, ok);
  }
}

// A fragments (16 rows x 64 cols = 4 k-steps) of the warp's rows [r0, r0+16)
__device__ __forceinline__ void load_a_frags(const Tile& t, int r0, int lane, uint32_t (*a)[4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(t.chunk_addr(r0 + (lane & 15), ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// C[16 x 64] += A(16 x 64, register frags over the tile's column dim) * T^T where T is a [64 n][64 k] tile
// (i.e. B[k][n] = T[n][k]: "NT" product, used for Q K^T, dO V^T, K Q^T, V dO^T)
__device__ __forceinline__ void mma_nt(float (*c)[4], const uint32_t (*a)[4], const Tile& t, int lane) {
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // pairs of 8-wide n tiles
      uint32_t b0, b1, b2, b3;
      ldsm_x4(t.chunk_addr(np * 16 + (mi >> 1) * 8 + rr, ks * 2 + (mi & 1)), b0, b1, b2, b3);
      mma16816(c[np * 2], a[ks], b0, b1);
      mma16816(c[np * 2 + 1], a[ks], b2, b3);
    }
  }
}

// C[16 x 64] += P(16 x 64 as A frags over the tile's ROW dim) * T where T is a [64 k][64 n] tile ("NN" product,
// used for P V, dS K, P^T dO, dS^T Q)
__device__ __forceinline__ void mma_nn(float (*c)[4], const uint32_t (*p)[4], const Tile& t, int lane) {
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {     // 16 tile rows per step
#pragma unroll
    for (int np = 0; np < 4; ++np) {   // pairs of 8-wide n tiles (columns)
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(t.chunk_addr(ks * 16 + (mi & 1) * 8 + rr, np * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma16816(c[np * 2], p[ks], b0, b1);
      mma16816(c[np * 2 + 1], p[ks], b2, b3);
    }
  }
}

// fp32 C fragments of a 16 x 64 tile -> bf16 A fragments (k dim = the 64 columns)
__device__ __forceinline__ void c_to_a(const float (*c)[4], uint32_t (*a)[4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

__device__ __forceinline__ void zero_acc(float (*c)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
}

// write the warp's 16 x 64 fp32 accumulators (scaled) as bf16 into a tile, then the CTA stores it coalesced
__device__ __forceinline__ void acc_to_tile(const Tile& t, int r0, int lane, const float (*c)[4], float s0, float s1) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(t.chunk_ptr(r0 + g, nt) + tq * 2) = pack_bf16(c[nt][0] * s0, c[nt][1] * s0);
    *reinterpret_cast<uint32_t*>(t.chunk_ptr(r0 + g + 8, nt) + tq * 2) = pack_bf16(c[nt][2] * s1, c[nt][3] * s1);
  }
}
__device__ __forceinline__ void store_tile(const Tile& t, bf16* g, int64_t ld, int row0, int nrows, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * NTHR;
    const int r = idx >> 3, c = idx & 7;
    if (row0 + r < nrows)
      *reinterpret_cast<uint4*>(g + (int64_t)(row0 + r) * ld + c * 8) = *reinterpret_cast<const uint4*>(t.chunk_ptr(r, c));
  }
}

// ------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(NTHR)
attn_fwd_mma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                    bf16* __restrict__ o, float* __restrict__ lse, int S, int H, int KV, int64_t ldq, int64_t ldk,
                    int64_t ldv, int64_t ldo, float scale_log2) {
  __shared__ __align__(128) bf16 sQ[BR * HD];
  __shared__ __align__(128) bf16 sK[2][BC * HD];
  __shared__ __align__(128) bf16 sV[2][BC * HD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = gridDim.x - 1 - blockIdx.x;  // longest rows first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (H / KV);
  const int q0 = qb * BR;
  const bf16* qg = q + (int64_t)b * S * ldq + (int64_t)h * HD;
  const bf16* kg = k + (int64_t)b * S * ldk + (int64_t)kvh * HD;
  const bf16* vg = v + (int64_t)b * S * ldv + (int64_t)kvh * HD;
  const Tile tQ{sQ};
  const int nkb = qb + 1;

  load_tile(tQ, qg, ldq, q0, S, tid);
  load_tile(Tile{sK[0]}, kg, ldk, 0, S, tid);
  load_tile(Tile{sV[0]}, vg, ldv, 0, S, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qa[4][4];
  load_a_frags(tQ, warp * 16, lane, qa);

  float oacc[8][4];
  zero_acc(oacc);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int g = lane >> 2, tq = lane & 3;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;

  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) {
      load_tile(Tile{sK[buf ^ 1]}, kg, ldk, (kb + 1) * BC, S, tid);
      load_tile(Tile{sV[buf ^ 1]}, vg, ldv, (kb + 1) * BC, S, tid);
      cp_async_commit();
    }
    float s[8][4];
    zero_acc(s);
    mma_nt(s, qa, Tile{sK[buf]}, lane);
    if (kb == nkb - 1) {  // diagonal block: causal mask (also hides keys beyond the sequence end)
      const int kbase = kb * BC;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c0 = kbase + nt * 8 + tq * 2;
        if (c0 > row0) s[nt][0] = -INFINITY;
        if (c0 + 1 > row0) s[nt][1] = -INFINITY;
        if (c0 > row1) s[nt][2] = -INFINITY;
        if (c0 + 1 > row1) s[nt][3] = -INFINITY;
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]) * scale_log2);
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]) * scale_log2);
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
    m0 = mx0; m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * scale_log2 - m0);
      s[nt][1] = exp2f(s[nt][1] * scale_log2 - m0);
      s[nt][2] = exp2f(s[nt][2] * scale_log2 - m1);
      s[nt][3] = exp2f(s[nt][3] * scale_log2 - m1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
      oacc[nt][0] *= c0; oacc[nt][1] *= c0; oacc[nt][2] *= c1; oacc[nt][3] *= c1;
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
    uint32_t pa[4][4];
    c_to_a(s, pa);
    mma_nn(oacc, pa, Tile{sV[buf]}, lane);
    if (kb + 1 < nkb) cp_async_wait<0>();
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // all warps are past their last read of sQ (qa lives in registers): reuse it to stage the output tile
  acc_to_tile(tQ, warp * 16, lane, oacc, 1.f / l0, 1.f / l1);
  if (tq == 0) {
    float* lp = lse + ((int64_t)b * H + h) * S;
    if (row0 < S) lp[row0] = m0 * kLn2 + logf(l0);
    if (row1 < S) lp[row1] = m1 * kLn2 + logf(l1);
  }
  __syncthreads();
  store_tile(tQ, o + (int64_t)b * S * ldo + (int64_t)h * HD, ldo, q0, S, tid);
}

// ------------------------------------------------------------------------------------------- delta = rowsum(dO * O)
// 8 lanes per (row, head): one 16-byte load of O and dO each per lane, 3 shuffles, consecutive lanes walk consecutive
// heads of a row so a warp reads 512 contiguous bytes per tensor; 4 independent pairs of loads in flight per lane.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta, int B, int S,
                  int H, int64_t ldo, int64_t lddo) {
  static_assert(HD == 64, "8 lanes x 8 elements per head");
  pdl_wait();
  pdl_trigger();
  const int sub = threadIdx.x & 7;
  const int64_t total = (int64_t)B * S * H;                       // (row, head) pairs, head fastest
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 3);
  int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  constexpr int U = 4;
  for (; w < total; w += U * stride) {
    uint4 a[U], d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t wu = w + u * stride;
      a[u] = d[u] = make_uint4(0, 0, 0, 0);
      if (wu < total) {
        const int64_t row = wu / H;
        const int h = (int)(wu % H);
        a[u] = ld_nc16(o + row * ldo + h * HD + sub * 8);
        d[u] = ld_nc16(dout + row * lddo + h * HD + sub * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = bf16_lo(a[u].x) * bf16_lo(d[u].x) + bf16_hi(a[u].x) * bf16_hi(d[u].x) +
                bf16_lo(a[u].y) * bf16_lo(d[u].y) + bf16_hi(a[u].y) * bf16_hi(d[u].y) +
                bf16_lo(a[u].z) * bf16_lo(d[u].z) + bf16_hi(a[u].z) * bf16_hi(d[u].z) +
                bf16_lo(a[u].w) * bf16_lo(d[u].w) + bf16_hi(a[u].w) * bf16_hi(d[u].w);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      const int64_t wu = w + u * stride;
      if (sub == 0 && wu < total) {
        const int64_t row = wu / H;
        const int h = (int)(wu % H);
        const int64_t b = row / S, i = row % S;
        delta[(b * H + h) * S + i] = s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- backward: dQ
__global__ void __launch_bounds__(NTHR)
attn_bwd_dq_mma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                       const bf16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                       bf16* __restrict__ dq, int S, int H, int KV, int64_t ldq, int64_t ldk, int64_t ldv,
                       int64_t lddo, int64_t lddq, float scale) {
  __shared__ __align__(128) bf16 sQ[BR * HD];
  __shared__ __align__(128) bf16 sK[2][BC * HD];
  __shared__ __align__(128) bf16 sV[2][BC * HD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = gridDim.x - 1 - blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (H / KV);
  const int q0 = qb * BR;
  const bf16* kg = k + (int64_t)b * S * ldk + (int64_t)kvh * HD;
  const bf16* vg = v + (int64_t)b * S * ldv + (int64_t)kvh * HD;
  const Tile tQ{sQ};
  const int nkb = qb + 1;
  const float scale_log2 = scale * kLog2e;

  // Q fragments, then dO fragments, through the same staging tile
  load_tile(tQ, q + (int64_t)b * S * ldq + (int64_t)h * HD, ldq, q0, S, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qa[4][4], da[4][4];
  load_a_frags(tQ, warp * 16, lane, qa);
  __syncthreads();
  load_tile(tQ, dout + (int64_t)b * S * lddo + (int64_t)h * HD, lddo, q0, S, tid);
  load_tile(Tile{sK[0]}, kg, ldk, 0, S, tid);
  load_tile(Tile{sV[0]}, vg, ldv, 0, S, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  load_a_frags(tQ, warp * 16, lane, da);

  const int g = lane >> 2, tq = lane & 3;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  const float* Lp = lse + ((int64_t)b * H + h) * S;
  const float* Dp = delta + ((int64_t)b * H + h) * S;
  const float L0 = (row0 < S) ? Lp[row0] * kLog2e : 0.f, L1 = (row1 < S) ? Lp[row1] * kLog2e : 0.f;
  const float D0 = (row0 < S) ? Dp[row0] : 0.f, D1 = (row1 < S) ? Dp[row1] : 0.f;

  float acc[8][4];
  zero_acc(acc);
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) {
      load_tile(Tile{sK[buf ^ 1]}, kg, ldk, (kb + 1) * BC, S, tid);
      load_tile(Tile{sV[buf ^ 1]}, vg, ldv, (kb + 1) * BC, S, tid);
      cp_async_commit();
    }
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_nt(s, qa, Tile{sK[buf]}, lane);
    mma_nt(dp, da, Tile{sV[buf]}, lane);
    const int kbase = kb * BC;
    const bool diag = (kb == nkb - 1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c0 = kbase + nt * 8 + tq * 2;
      float p0 = exp2f(s[nt][0] * scale_log2 - L0), p1 = exp2f(s[nt][1] * scale_log2 - L0);
      float p2 = exp2f(s[nt][2] * scale_log2 - L1), p3 = exp2f(s[nt][3] * scale_log2 - L1);
      if (diag) {
        if (c0 > row0) p0 = 0.f;
        if (c0 + 1 > row0) p1 = 0.f;
        if (c0 > row1) p2 = 0.f;
        if (c0 + 1 > row1) p3 = 0.f;
      }
      s[nt][0] = p0 * (dp[nt][0] - D0) * scale;
      s[nt][1] = p1 * (dp[nt][1] - D0) * scale;
      s[nt][2] = p2 * (dp[nt][2] - D1) * scale;
      s[nt][3] = p3 * (dp[nt][3] - D1) * scale;
    }
    uint32_t dsa[4][4];
    c_to_a(s, dsa);
    mma_nn(acc, dsa, Tile{sK[buf]}, lane);
    if (kb + 1 < nkb) cp_async_wait<0>();
    __syncthreads();
  }
  acc_to_tile(tQ, warp * 16, lane, acc, 1.f, 1.f);
  __syncthreads();
  store_tile(tQ, dq + (int64_t)b * S * lddq + (int64_t)h * HD, lddq, q0, S, tid);
}

// ------------------------------------------------------------------------------------------- backward: dK, dV
__global__ void __launch_bounds__(NTHR)
attn_bwd_dkdv_mma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                         const bf16* __restrict__ dout, const float* __restrict__ lse,
                         const float* __restrict__ delta, bf16* __restrict__ dk, bf16* __restrict__ dv, int S, int H,
                         int KV, int64_t ldq, int64_t ldk, int64_t ldv, int64_t lddo, int64_t lddk, int64_t lddv,
                         float scale) {
  __shared__ __align__(128) bf16 sQ[2][BR * HD];
  __shared__ __align__(128) bf16 sD[2][BR * HD];
  __shared__ float sL[2][BR], sDel[2][BR];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kvb = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int rep = H / KV;
  const int k0 = kvb * BC;
  const int nqb = (S + BR - 1) / BR;
  const float scale_log2 = scale * kLog2e;

  // this CTA's K and V tiles -> per-warp A fragments (16 keys x 64 dims), staged through sQ[0] / sD[0]
  load_tile(Tile{sQ[0]}, k + (int64_t)b * S * ldk + (int64_t)kvh * HD, ldk, k0, S, tid);
  load_tile(Tile{sD[0]}, v + (int64_t)b * S * ldv + (int64_t)kvh * HD, ldv, k0, S, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t ka[4][4], va[4][4];
  load_a_frags(Tile{sQ[0]}, warp * 16, lane, ka);
  load_a_frags(Tile{sD[0]}, warp * 16, lane, va);
  __syncthreads();

  const int g = lane >> 2, tq = lane & 3;
  const int key0 = k0 + warp * 16 + g, key1 = key0 + 8;
  float dkacc[8][4], dvacc[8][4];
  zero_acc(dkacc);
  zero_acc(dvacc);

  const int nq_iter = nqb - kvb;          // q blocks kvb .. nqb-1
  const int total = rep * nq_iter;
  auto issue = [&](int it, int buf) {
    const int hh = it / nq_iter, qb = kvb + it % nq_iter;
    const int h = kvh * rep + hh;
    load_tile(Tile{sQ[buf]}, q + (int64_t)b * S * ldq + (int64_t)h * HD, ldq, qb * BR, S, tid);
    load_tile(Tile{sD[buf]}, dout + (int64_t)b * S * lddo + (int64_t)h * HD, lddo, qb * BR, S, tid);
    if (tid < BR) {
      const int r = qb * BR + tid;
      const int64_t li = ((int64_t)b * H + h) * S + r;
      sL[buf][tid] = (r < S) ? lse[li] * kLog2e : INFINITY;   // +inf => P = 0 for rows beyond the sequence
      sDel[buf][tid] = (r < S) ? delta[li] : 0.f;
    }
    cp_async_commit();
  };
  issue(0, 0);
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    cp_async_wait<0>();
    __syncthreads();                       // tile `buf` landed; everyone is done with tile buf^1
    if (it + 1 < total) issue(it + 1, buf ^ 1);
    const int qb = kvb + it % nq_iter;
    const Tile tQ{sQ[buf]}, tD{sD[buf]};
    float st[8][4], dpt[8][4];             // S^T and dP^T: rows = this warp's keys, cols = the 64 queries
    zero_acc(st);
    zero_acc(dpt);
    mma_nt(st, ka, tQ, lane);
    mma_nt(dpt, va, tD, lane);
    const int qbase = qb * BR;
    const bool diag = (qb == kvb);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + tq * 2;       // query column within the tile
      const float La = sL[buf][c], Lb = sL[buf][c + 1];
      const float Da = sDel[buf][c], Db = sDel[buf][c + 1];
      float p0 = exp2f(st[nt][0] * scale_log2 - La), p1 = exp2f(st[nt][1] * scale_log2 - Lb);
      float p2 = exp2f(st[nt][2] * scale_log2 - La), p3 = exp2f(st[nt][3] * scale_log2 - Lb);
      if (diag) {                          // key > query => masked
        const int qa_ = qbase + c;
        if (key0 > qa_) p0 = 0.f;
        if (key0 > qa_ + 1) p1 = 0.f;
        if (key1 > qa_) p2 = 0.f;
        if (key1 > qa_ + 1) p3 = 0.f;
      }
      st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
      dpt[nt][0] = p0 * (dpt[nt][0] - Da) * scale;
      dpt[nt][1] = p1 * (dpt[nt][1] - Db) * scale;
      dpt[nt][2] = p2 * (dpt[nt][2] - Da) * scale;
      dpt[nt][3] = p3 * (dpt[nt][3] - Db) * scale;
    }
    uint32_t pa[4][4];
    c_to_a(st, pa);
    mma_nn(dvacc, pa, tD, lane);           // dV += P^T dO
    c_to_a(dpt, pa);
    mma_nn(dkacc, pa, tQ, lane);           // dK += dS^T Q
  }
  __syncthreads();
  acc_to_tile(Tile{sQ[0]}, warp * 16, lane, dkacc, 1.f, 1.f);
  acc_to_tile(Tile{sD[0]}, warp * 16, lane, dvacc, 1.f, 1.f);
  __syncthreads();
  store_tile(Tile{sQ[0]}, dk + (int64_t)b * S * lddk + (int64_t)kvh * HD, lddk, k0, S, tid);
  store_tile(Tile{sD[0]}, dv + (int64_t)b * S * lddv + (int64_t)kvh * HD, lddv, k0, S, tid);
}

}  // namespace

static unsigned delta_grid(int64_t pairs) {
  const int64_t want = (pairs + 32 * 4 - 1) / (32 * 4);           // 32 pairs per CTA pass, 4 passes per thread
  const int64_t cap = (int64_t)num_sms() * 8;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

int attn_delta_launch(const void* o, const void* dout, float* delta, int B, int S, int H, int64_t ldo,
                      cudaStream_t st) {
  const int64_t rows = (int64_t)B * H * S;
  if (launch_k(attn_delta_kernel, dim3(delta_grid(rows)), dim3(256), 0, st, 1, (const bf16*)o, (const bf16*)dout, delta, B,
               S, H, ldo, ldo) != cudaSuccess) { set_error("attn_delta: launch failed"); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("attn_delta");
  return CSM_OK;
}

bool attn_mma_supported(int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo) {
  return hd == HD && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0;
}

int attn_fwd_mma_launch(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int H, int KV,
                        int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, cudaStream_t st) {
  (void)hd;
  CSM_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), CSM_ERR_ALIGN, "attn_fwd: misaligned");
  dim3 grid((S + BR - 1) / BR, H, B);
  attn_fwd_mma_kernel<<<grid, NTHR, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, lse, S, H, KV,
                                             ldq, ldk, ldv, ldo, scale * kLog2e);
  CSM_CHECK_LAUNCH("attn_fwd_mma");
  return CSM_OK;
}

int attn_bwd_mma_launch(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                        void* dq, void* dk, void* dv, float* delta, int B, int S, int H, int KV, int hd, int64_t ldq,
                        int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                        cudaStream_t st) {
  (void)hd;
  CSM_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(dout) && aligned16(dq) &&
                  aligned16(dk) && aligned16(dv) && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
              CSM_ERR_ALIGN, "attn_bwd: misaligned");
  const int64_t rows = (int64_t)B * H * S;
  attn_delta_kernel<<<delta_grid(rows), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, delta, B, S, H, ldo, ldo);
  CSM_CHECK_LAUNCH("attn_delta");
  dim3 gq((S + BR - 1) / BR, H, B);
  attn_bwd_dq_mma_kernel<<<gq, NTHR, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, lse,
                                              delta, (bf16*)dq, S, H, KV, ldq, ldk, ldv, ldo, lddq, scale);
  CSM_CHECK_LAUNCH("attn_bwd_dq_mma");
  dim3 gk((S + BC - 1) / BC, KV, B);
  attn_bwd_dkdv_mma_kernel<<<gk, NTHR, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, lse,
                                                delta, (bf16*)dk, (bf16*)dv, S, H, KV, ldq, ldk, ldv, ldo, lddk, lddv,
                                                scale);
  CSM_CHECK_LAUNCH("attn_bwd_dkdv_mma");
  return CSM_OK;
}

}  // namespace csm
