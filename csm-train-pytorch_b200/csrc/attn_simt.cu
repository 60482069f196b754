// Small-shape causal GQA attention (fwd, bwd-dQ, bwd-dKdV): one warp per query (or key) row, the head
// dimension spread over the lanes, online softmax in fp32.  Serves head_dim < 64 (tiny test model) and is
// the deterministic, atomic-free reference implementation inside the library; head_dim 64/128 run the
// tensor-core kernels in attn_mma.cu.
// Semantics: F.scaled_dot_product_attention(is_causal=True) with torchtune's GQA head expansion
// (oracle/torchtune_shim.py MultiHeadAttention).
#include "common.cuh"

namespace csm {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <int DPL>
__device__ __forceinline__ void load_row(const bf16* p, int lane, int hd, float* out) {
#pragma unroll
  for (int t = 0; t < DPL; ++t) {
    const int d = lane * DPL + t;
    out[t] = (d < hd) ? __bfloat162float(p[d]) : 0.f;
  }
}

template <int DPL>
__global__ void __launch_bounds__(128)
attn_fwd_simt_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                     bf16* __restrict__ o, float* __restrict__ lse, int B, int S, int H, int KV, int hd,
                     int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * H * S) return;
  const int i = (int)(w % S);
  const int h = (int)((w / S) % H);
  const int b = (int)(w / ((int64_t)S * H));
  const int kvh = h / (H / KV);
  float qv[DPL], acc[DPL];
  load_row<DPL>(q + ((int64_t)b * S + i) * ldq + (int64_t)h * hd, lane, hd, qv);
#pragma unroll
  for (int t = 0; t < DPL; ++t) { qv[t] *= scale * kLog2e; acc[t] = 0.f; }
  float m = -INFINITY, l = 0.f;
  const bf16* kb = k + (int64_t)b * S * ldk + (int64_t)kvh * hd;
  const bf16* vb = v + (int64_t)b * S * ldv + (int64_t)kvh * hd;
  for (int j0 = 0; j0 <= i; j0 += 4) {
    float s[4], vv[4][DPL];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      float kk[DPL];
      float part = 0.f;
      if (j <= i) {
        load_row<DPL>(kb + (int64_t)j * ldk, lane, hd, kk);
        load_row<DPL>(vb + (int64_t)j * ldv, lane, hd, vv[u]);
#pragma unroll
        for (int t = 0; t < DPL; ++t) part += qv[t] * kk[t];
      } else {
#pragma unroll
        for (int t = 0; t < DPL; ++t) vv[u][t] = 0.f;
      }
      s[u] = part;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = warp_sum(s[u]);
      if (j0 + u > i) s[u] = -INFINITY;
    }
    const float mn = fmaxf(fmaxf(fmaxf(m, s[0]), fmaxf(s[1], s[2])), s[3]);
    const float corr = exp2f(m - mn);
    l *= corr;
#pragma unroll
    for (int t = 0; t < DPL; ++t) acc[t] *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float p = exp2f(s[u] - mn);
      l += p;
#pragma unroll
      for (int t = 0; t < DPL; ++t) acc[t] += p * vv[u][t];
    }
    m = mn;
  }
  const float inv = 1.f / l;
  bf16* op = o + ((int64_t)b * S + i) * ldo + (int64_t)h * hd;
#pragma unroll
  for (int t = 0; t < DPL; ++t) {
    const int d = lane * DPL + t;
    if (d < hd) op[d] = __float2bfloat16_rn(acc[t] * inv);
  }
  if (lane == 0) lse[((int64_t)b * H + h) * S + i] = m * kLn2 + logf(l);
}

template <int DPL>
__global__ void __launch_bounds__(128)
attn_bwd_dq_simt_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                        const bf16* __restrict__ o, const float* __restrict__ lse, const bf16* __restrict__ dout,
                        bf16* __restrict__ dq, float* __restrict__ delta, int B, int S, int H, int KV, int hd,
                        int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * H * S) return;
  const int i = (int)(w % S);
  const int h = (int)((w / S) % H);
  const int b = (int)(w / ((int64_t)S * H));
  const int kvh = h / (H / KV);
  float qv[DPL], dov[DPL], ov[DPL], acc[DPL];
  load_row<DPL>(q + ((int64_t)b * S + i) * ldq + (int64_t)h * hd, lane, hd, qv);
  load_row<DPL>(dout + ((int64_t)b * S + i) * ldo + (int64_t)h * hd, lane, hd, dov);
  load_row<DPL>(o + ((int64_t)b * S + i) * ldo + (int64_t)h * hd, lane, hd, ov);
  float dl = 0.f;
#pragma unroll
  for (int t = 0; t < DPL; ++t) { dl += dov[t] * ov[t]; acc[t] = 0.f; }
  dl = warp_sum(dl);
  const int64_t li = ((int64_t)b * H + h) * S + i;
  if (lane == 0) delta[li] = dl;
  const float L = lse[li];
  const bf16* kb = k + (int64_t)b * S * ldk + (int64_t)kvh * hd;
  const bf16* vb = v + (int64_t)b * S * ldv + (int64_t)kvh * hd;
  for (int j0 = 0; j0 <= i; j0 += 2) {
    float s[2], dp[2], kk[2][DPL];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + u;
      s[u] = 0.f; dp[u] = 0.f;
      if (j <= i) {
        float vv[DPL];
        load_row<DPL>(kb + (int64_t)j * ldk, lane, hd, kk[u]);
        load_row<DPL>(vb + (int64_t)j * ldv, lane, hd, vv);
#pragma unroll
        for (int t = 0; t < DPL; ++t) { s[u] += qv[t] * kk[u][t]; dp[u] += dov[t] * vv[t]; }
      } else {
#pragma unroll
        for (int t = 0; t < DPL; ++t) kk[u][t] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      s[u] = warp_sum(s[u]);
      dp[u] = warp_sum(dp[u]);
      const float p = (j0 + u <= i) ? __expf(s[u] * scale - L) : 0.f;
      const float ds = p * (dp[u] - dl) * scale;
#pragma unroll
      for (int t = 0; t < DPL; ++t) acc[t] += ds * kk[u][t];
    }
  }
  bf16* dp_ = dq + ((int64_t)b * S + i) * lddq + (int64_t)h * hd;
#pragma unroll
  for (int t = 0; t < DPL; ++t) {
    const int d = lane * DPL + t;
    if (d < hd) dp_[d] = __float2bfloat16_rn(acc[t]);
  }
}

template <int DPL>
__global__ void __launch_bounds__(128)
attn_bwd_dkdv_simt_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                          const float* __restrict__ lse, const bf16* __restrict__ dout,
                          const float* __restrict__ delta, bf16* __restrict__ dk, bf16* __restrict__ dv, int B,
                          int S, int H, int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                          int64_t lddk, int64_t lddv, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * KV * S) return;
  const int j = (int)(w % S);
  const int kvh = (int)((w / S) % KV);
  const int b = (int)(w / ((int64_t)S * KV));
  const int rep = H / KV;
  float kk[DPL], vv[DPL], ak[DPL], av[DPL];
  load_row<DPL>(k + ((int64_t)b * S + j) * ldk + (int64_t)kvh * hd, lane, hd, kk);
  load_row<DPL>(v + ((int64_t)b * S + j) * ldv + (int64_t)kvh * hd, lane, hd, vv);
#pragma unroll
  for (int t = 0; t < DPL; ++t) { ak[t] = 0.f; av[t] = 0.f; }
  for (int hh = 0; hh < rep; ++hh) {
    const int h = kvh * rep + hh;
    const bf16* qb = q + (int64_t)b * S * ldq + (int64_t)h * hd;
    const bf16* db = dout + (int64_t)b * S * ldo + (int64_t)h * hd;
    const float* Lb = lse + ((int64_t)b * H + h) * S;
    const float* Db = delta + ((int64_t)b * H + h) * S;
    for (int i = j; i < S; ++i) {
      float qv[DPL], dov[DPL];
      load_row<DPL>(qb + (int64_t)i * ldq, lane, hd, qv);
      load_row<DPL>(db + (int64_t)i * ldo, lane, hd, dov);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int t = 0; t < DPL; ++t) { s += qv[t] * kk[t]; dp += dov[t] * vv[t]; }
      s = warp_sum(s);
      dp = warp_sum(dp);
      const float p = __expf(s * scale - Lb[i]);
      const float ds = p * (dp - Db[i]) * scale;
#pragma unroll
      for (int t = 0; t < DPL; ++t) { ak[t] += ds * qv[t]; av[t] += p * dov[t]; }
    }
  }
  bf16* dkp = dk + ((int64_t)b * S + j) * lddk + (int64_t)kvh * hd;
  bf16* dvp = dv + ((int64_t)b * S + j) * lddv + (int64_t)kvh * hd;
#pragma unroll
  for (int t = 0; t < DPL; ++t) {
    const int d = lane * DPL + t;
    if (d < hd) { dkp[d] = __float2bfloat16_rn(ak[t]); dvp[d] = __float2bfloat16_rn(av[t]); }
  }
}

int attn_fwd_simt_launch(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int H,
                         int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale,
                         cudaStream_t st) {
  const int64_t warps = (int64_t)B * H * S;
  const unsigned grid = (unsigned)((warps + 3) / 4);
#define LAUNCH(DPL) attn_fwd_simt_kernel<DPL><<<grid, 128, 0, st>>>((const bf16*)q, (const bf16*)k, \
      (const bf16*)v, (bf16*)o, lse, B, S, H, KV, hd, ldq, ldk, ldv, ldo, scale)
  if (hd <= 32) LAUNCH(1); else if (hd <= 64) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
  CSM_CHECK_LAUNCH("attn_fwd_simt");
  return CSM_OK;
}

int attn_bwd_simt_launch(const void* q, const void* k, const void* v, const void* o, const float* lse,
                         const void* dout, void* dq, void* dk, void* dv, float* delta, int B, int S, int H, int KV,
                         int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk,
                         int64_t lddv, float scale, cudaStream_t st) {
  {
    const int64_t warps = (int64_t)B * H * S;
    const unsigned grid = (unsigned)((warps + 3) / 4);
#define LAUNCH(DPL) attn_bwd_dq_simt_kernel<DPL><<<grid, 128, 0, st>>>((const bf16*)q, (const bf16*)k, \
      (const bf16*)v, (const bf16*)o, lse, (const bf16*)dout, (bf16*)dq, delta, B, S, H, KV, hd, ldq, ldk, ldv, \
      ldo, lddq, scale)
    if (hd <= 32) LAUNCH(1); else if (hd <= 64) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
    CSM_CHECK_LAUNCH("attn_bwd_dq_simt");
  }
  {
    const int64_t warps = (int64_t)B * KV * S;
    const unsigned grid = (unsigned)((warps + 3) / 4);
#define LAUNCH(DPL) attn_bwd_dkdv_simt_kernel<DPL><<<grid, 128, 0, st>>>((const bf16*)q, (const bf16*)k, \
      (const bf16*)v, lse, (const bf16*)dout, delta, (bf16*)dk, (bf16*)dv, B, S, H, KV, hd, ldq, ldk, ldv, ldo, \
      lddk, lddv, scale)
    if (hd <= 32) LAUNCH(1); else if (hd <= 64) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
    CSM_CHECK_LAUNCH("attn_bwd_dkdv_simt");
  }
  return CSM_OK;
}

}  // namespace csm
