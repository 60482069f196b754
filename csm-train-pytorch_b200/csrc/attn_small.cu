// Short-sequence causal GQA attention (seq <= 32, head_dim 64 / 128): the depth decoder's shape — 32 codebook
// positions per selected frame, 8 query heads over 2 KV heads of 128 (reference model.py:28-42, 184).
//
// One CTA per (sequence, kv head).  K and V of that head (<= 32 x 128 bf16 each) live in shared memory; every query
// row of the `rep` query heads sharing the kv head is owned by a PAIR of lanes (one half of the head dim each), so a
// score costs one shuffle instead of a 5-step warp reduction, all 32 scores of a row stay in registers (no online
// softmax needed) and K/V rows are read as 16-byte shared-memory broadcasts.  The backward does dQ row-parallel in
// the same mapping, parks P and dS (fp32) in shared memory, then does dK/dV column-parallel with the sum over the
// `rep` query heads inside the CTA — deterministic, no atomics, one launch.
// Semantics as attn_simt.cu: F.scaled_dot_product_attention(is_causal=True) + torchtune's GQA expansion.
#include "common.cuh"

namespace csm {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxS = 32;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// copies `rows` rows of HD bf16 (row stride ld elements) into dense smem [kMaxS][HD]; rows >= `rows` are zeroed
template <int HD>
__device__ __forceinline__ void stage_rows(bf16* dst, const bf16* src, int64_t ld, int rows, int tid, int nthr) {
  constexpr int CH = HD / 8;
  for (int idx = tid; idx < kMaxS * CH; idx += nthr) {
    const int r = idx / CH, c = idx % CH;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < rows) val = *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + c * 8);
    *reinterpret_cast<uint4*>(dst + r * HD + c * 8) = val;
  }
}

// dot of a register half-row with a shared-memory half-row (HH elements)
template <int HH>
__device__ __forceinline__ float dot_half(const float* a, const bf16* b) {
  float p0 = 0.f, p1 = 0.f;
#pragma unroll
  for (int c = 0; c < HH / 8; ++c) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(b + c * 8), f);
    p0 += a[c * 8 + 0] * f[0]; p1 += a[c * 8 + 1] * f[1];
    p0 += a[c * 8 + 2] * f[2]; p1 += a[c * 8 + 3] * f[3];
    p0 += a[c * 8 + 4] * f[4]; p1 += a[c * 8 + 5] * f[5];
    p0 += a[c * 8 + 6] * f[6]; p1 += a[c * 8 + 7] * f[7];
  }
  return p0 + p1;
}

template <int HH>
__device__ __forceinline__ void axpy_half(float* acc, float w, const bf16* b) {
#pragma unroll
  for (int c = 0; c < HH / 8; ++c) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(b + c * 8), f);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[c * 8 + t] += w * f[t];
  }
}

template <int HH>
__device__ __forceinline__ void load_half_global(float* dst, const bf16* src, bool valid) {
#pragma unroll
  for (int c = 0; c < HH / 8; ++c) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (valid) u = *reinterpret_cast<const uint4*>(src + c * 8);
    unpack8(u, dst + c * 8);
  }
}

template <int HH>
__device__ __forceinline__ void store_half_global(bf16* dst, const float* src, float mul) {
#pragma unroll
  for (int c = 0; c < HH / 8; ++c) {
    uint4 u;
    u.x = pack_bf16(src[c * 8 + 0] * mul, src[c * 8 + 1] * mul);
    u.y = pack_bf16(src[c * 8 + 2] * mul, src[c * 8 + 3] * mul);
    u.z = pack_bf16(src[c * 8 + 4] * mul, src[c * 8 + 5] * mul);
    u.w = pack_bf16(src[c * 8 + 6] * mul, src[c * 8 + 7] * mul);
    *reinterpret_cast<uint4*>(dst + c * 8) = u;
  }
}

// thread -> (local query head, row, half of the head dim); a warp holds 16 rows x 2 halves of one head
struct RowMap {
  int hl, row, half, jmax;
  __device__ RowMap(int tid, int S) {
    const int w = tid >> 5, lane = tid & 31;
    hl = w >> 1;
    row = (w & 1) * 16 + (lane & 15);
    half = lane >> 4;
    jmax = min(S - 1, (w & 1) * 16 + 15);   // warp-uniform loop bound
  }
};

template <int HD>
__global__ void __launch_bounds__(256)
attn_small_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      bf16* __restrict__ o, float* __restrict__ lse, int S, int H, int KV, int64_t ldq, int64_t ldk,
                      int64_t ldv, int64_t ldo, float scale) {
  constexpr int HH = HD / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
  bf16* Vs = Ks + kMaxS * HD;
  const int b = blockIdx.x / KV, kvh = blockIdx.x % KV, rep = H / KV;
  const int tid = threadIdx.x;
  stage_rows<HD>(Ks, k + (int64_t)b * S * ldk + (int64_t)kvh * HD, ldk, S, tid, blockDim.x);
  stage_rows<HD>(Vs, v + (int64_t)b * S * ldv + (int64_t)kvh * HD, ldv, S, tid, blockDim.x);
  const RowMap m(tid, S);
  const int h = kvh * rep + m.hl;
  const bool valid = m.row < S;
  float qf[HH];
  load_half_global<HH>(qf, q + ((int64_t)b * S + m.row) * ldq + (int64_t)h * HD + m.half * HH, valid);
  const float sc = scale * kLog2e;
#pragma unroll
  for (int t = 0; t < HH; ++t) qf[t] *= sc;
  __syncthreads();
  float s[kMaxS];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxS; ++j) {
    s[j] = -INFINITY;
    if (j <= m.jmax) {
      float part = dot_half<HH>(qf, Ks + j * HD + m.half * HH);
      part += __shfl_xor_sync(0xffffffffu, part, 16);
      if (j <= m.row) s[j] = part;
    }
    mx = fmaxf(mx, s[j]);
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxS; ++j) {
    s[j] = exp2f(s[j] - mx);     // exp2(-inf) = 0 for masked entries
    l += s[j];
  }
  float acc[HH];
#pragma unroll
  for (int t = 0; t < HH; ++t) acc[t] = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxS; ++j)
    if (j <= m.jmax) axpy_half<HH>(acc, s[j], Vs + j * HD + m.half * HH);
  if (valid) {
    store_half_global<HH>(o + ((int64_t)b * S + m.row) * ldo + (int64_t)h * HD + m.half * HH, acc, 1.f / l);
    if (m.half == 0) lse[((int64_t)b * H + h) * S + m.row] = mx * kLn2 + logf(l);
  }
}

template <int HD>
__global__ void __launch_bounds__(256, 1)
attn_small_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      const bf16* __restrict__ o, const float* __restrict__ lse, const bf16* __restrict__ dout,
                      bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv, int S, int H, int KV,
                      int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv,
                      float scale) {
  constexpr int HH = HD / 2;
  constexpr int PS = kMaxS + 1;     // padded row of the P / dS tiles
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / KV, kvh = blockIdx.x % KV, rep = H / KV;
  const int tid = threadIdx.x, nthr = blockDim.x;
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
  bf16* Vs = Ks + kMaxS * HD;
  bf16* Qs = Vs + kMaxS * HD;                       // [rep][32][HD]
  bf16* Ds = Qs + rep * kMaxS * HD;                 // [rep][32][HD]  (dO)
  float* Ps = reinterpret_cast<float*>(Ds + rep * kMaxS * HD);   // [rep][32][33]
  float* Gs = Ps + rep * kMaxS * PS;                // dS
  stage_rows<HD>(Ks, k + (int64_t)b * S * ldk + (int64_t)kvh * HD, ldk, S, tid, nthr);
  stage_rows<HD>(Vs, v + (int64_t)b * S * ldv + (int64_t)kvh * HD, ldv, S, tid, nthr);
  for (int r = 0; r < rep; ++r) {
    const int h = kvh * rep + r;
    stage_rows<HD>(Qs + r * kMaxS * HD, q + (int64_t)b * S * ldq + (int64_t)h * HD, ldq, S, tid, nthr);
    stage_rows<HD>(Ds + r * kMaxS * HD, dout + (int64_t)b * S * ldo + (int64_t)h * HD, ldo, S, tid, nthr);
  }
  __syncthreads();

  // ---- phase 1: row-parallel — P, dS (to smem) and dQ
  {
    const RowMap m(tid, S);
    const int h = kvh * rep + m.hl;
    const bool valid = m.row < S;
    float vec[HH];
    const bf16* qrow = Qs + (m.hl * kMaxS + m.row) * HD + m.half * HH;
    const bf16* drow = Ds + (m.hl * kMaxS + m.row) * HD + m.half * HH;
#pragma unroll
    for (int c = 0; c < HH / 8; ++c) unpack8(*reinterpret_cast<const uint4*>(qrow + c * 8), vec + c * 8);
    const float L = valid ? lse[((int64_t)b * H + h) * S + m.row] * kLog2e : 0.f;
    const float sc = scale * kLog2e;
    float p[kMaxS];
#pragma unroll
    for (int j = 0; j < kMaxS; ++j) {
      p[j] = 0.f;
      if (j <= m.jmax) {
        float part = dot_half<HH>(vec, Ks + j * HD + m.half * HH);
        part += __shfl_xor_sync(0xffffffffu, part, 16);
        if (j <= m.row && valid) p[j] = exp2f(part * sc - L);
      }
    }
    float* prow = Ps + (m.hl * kMaxS + m.row) * PS;
    float* grow = Gs + (m.hl * kMaxS + m.row) * PS;
    if (m.half == 0) {
#pragma unroll
      for (int j = 0; j < kMaxS; ++j) prow[j] = p[j];
    }
    // delta = rowsum(dO * O)
#pragma unroll
    for (int c = 0; c < HH / 8; ++c) unpack8(*reinterpret_cast<const uint4*>(drow + c * 8), vec + c * 8);
    float delta = 0.f;
    {
      const bf16* orow = o + ((int64_t)b * S + m.row) * ldo + (int64_t)h * HD + m.half * HH;
#pragma unroll
      for (int c = 0; c < HH / 8; ++c) {
        float f[8];
        uint4 u = make_uint4(0, 0, 0, 0);
        if (valid) u = *reinterpret_cast<const uint4*>(orow + c * 8);
        unpack8(u, f);
#pragma unroll
        for (int t = 0; t < 8; ++t) delta += vec[c * 8 + t] * f[t];
      }
      delta += __shfl_xor_sync(0xffffffffu, delta, 16);
    }
#pragma unroll
    for (int j = 0; j < kMaxS; ++j) {
      if (j <= m.jmax) {
        float part = dot_half<HH>(vec, Vs + j * HD + m.half * HH);
        part += __shfl_xor_sync(0xffffffffu, part, 16);
        p[j] = p[j] * (part - delta) * scale;      // dS
      }
    }
    if (m.half == 1) {
#pragma unroll
      for (int j = 0; j < kMaxS; ++j) grow[j] = p[j];
    }
    float acc[HH];
#pragma unroll
    for (int t = 0; t < HH; ++t) acc[t] = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxS; ++j)
      if (j <= m.jmax) axpy_half<HH>(acc, p[j], Ks + j * HD + m.half * HH);
    if (valid)
      store_half_global<HH>(dq + ((int64_t)b * S + m.row) * lddq + (int64_t)h * HD + m.half * HH, acc, 1.f);
  }
  __syncthreads();

  // ---- phase 2: column-parallel — dK[j] = sum_h sum_{i>=j} dS[i][j] Q[i], dV[j] = sum_h sum_{i>=j} P[i][j] dO[i]
  // a work item is a pair of keys (j, 31 - j) x 8 head-dim elements: every item walks the same 33 query rows per
  // head, so the causal triangle is balanced across the threads
  constexpr int CW = 8;
  constexpr int NC = HD / CW;
  for (int item = tid; item < (kMaxS / 2) * NC; item += nthr) {
    const int jlo = item / NC, c = item % NC;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const int j = side ? kMaxS - 1 - jlo : jlo;
      if (j >= S) continue;
      float ak[CW], av[CW];
#pragma unroll
      for (int t = 0; t < CW; ++t) { ak[t] = 0.f; av[t] = 0.f; }
      for (int r = 0; r < rep; ++r) {
        const float* Pr = Ps + r * kMaxS * PS + j;
        const float* Gr = Gs + r * kMaxS * PS + j;
        const bf16* Qr = Qs + r * kMaxS * HD + c * CW;
        const bf16* Dr = Ds + r * kMaxS * HD + c * CW;
#pragma unroll 4
        for (int i = j; i < S; ++i) {
          const float ds = Gr[i * PS], pp = Pr[i * PS];
          float f[8];
          unpack8(*reinterpret_cast<const uint4*>(Qr + i * HD), f);
#pragma unroll
          for (int t = 0; t < 8; ++t) ak[t] += ds * f[t];
          unpack8(*reinterpret_cast<const uint4*>(Dr + i * HD), f);
#pragma unroll
          for (int t = 0; t < 8; ++t) av[t] += pp * f[t];
        }
      }
      store_half_global<CW>(dk + ((int64_t)b * S + j) * lddk + (int64_t)kvh * HD + c * CW, ak, 1.f);
      store_half_global<CW>(dv + ((int64_t)b * S + j) * lddv + (int64_t)kvh * HD + c * CW, av, 1.f);
    }
  }
}

size_t bwd_smem_bytes(int hd, int rep) {
  return (size_t)(2 + 2 * rep) * kMaxS * hd * sizeof(bf16) + (size_t)2 * rep * kMaxS * (kMaxS + 1) * sizeof(float);
}

}  // namespace

bool attn_small_supported(int S, int H, int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                          const void* q, const void* k, const void* v, const void* o) {
  if (S > kMaxS || (hd != 64 && hd != 128) || KV <= 0 || H % KV != 0) return false;
  const int rep = H / KV;
  if (rep > 4) return false;             // 64 threads per query head, 256 per CTA
  if ((ldq | ldk | ldv | ldo) & 7) return false;
  return aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o);
}

int attn_fwd_small_launch(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int H,
                          int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale,
                          cudaStream_t st) {
  const int rep = H / KV;
  const unsigned grid = (unsigned)(B * KV), block = (unsigned)(rep * 64);
  const size_t smem = (size_t)2 * kMaxS * hd * sizeof(bf16);
#define LAUNCH(HD) attn_small_fwd_kernel<HD><<<grid, block, smem, st>>>((const bf16*)q, (const bf16*)k, \
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
  CSM_REQUIRE(((lddq | lddk | lddv) & 7) == 0 && aligned16(dq) && aligned16(dk) && aligned16(dv) && aligned16(dout),
              CSM_ERR_ALIGN, "attn_bwd (short-sequence kernel): gradients must be 16-byte aligned");
  const int rep = H / KV;
  const unsigned grid = (unsigned)(B * KV), block = (unsigned)(rep * 64);
  const size_t smem = bwd_smem_bytes(hd, rep);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_small_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bwd_smem_bytes(64, 4));
    cudaFuncSetAttribute(attn_small_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bwd_smem_bytes(128, 4));
    attr_set = true;
  }
#define LAUNCH(HD) attn_small_bwd_kernel<HD><<<grid, block, smem, st>>>((const bf16*)q, (const bf16*)k, \
      (const bf16*)v, (const bf16*)o, lse, (const bf16*)dout, (bf16*)dq, (bf16*)dk, (bf16*)dv, S, H, KV, ldq, ldk, \
      ldv, ldo, lddq, lddk, lddv, scale)
  if (hd == 64) LAUNCH(64); else LAUNCH(128);
#undef LAUNCH
  CSM_CHECK_LAUNCH("attn_small_bwd");
  return CSM_OK;
}

}  // namespace csm
