// KV-cache decode attention for Model.generate_frame (reference model.py:140-195; torchtune MultiHeadAttention with
// kv_cache, restated in oracle/torchtune_shim.py): ONE new query position per sample attends to the kv_len cached
// positions (itself included) of its sample.  GQA: the H / KV query heads that share a KV head are processed by the same
// CTA, so every cached K / V row is read once.  HBM / latency-bound streaming kernel: 8 warps split the keys, each keeps
// an online softmax per query head, partials are merged through shared memory.  fp32 math, bf16 in / out.
#include "common.cuh"

namespace csm {

namespace {

constexpr int kDecThreads = 256;
constexpr int kMaxGroup = 8;      // query heads per KV head (CSM-1B: 4)
constexpr int kMaxVpl = 4;        // head_dim <= 128: up to 4 elements per lane

template <int VPL>
__global__ void __launch_bounds__(kDecThreads)
attn_decode_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
                   bf16* __restrict__ o, int H, int KV, int hd, int kv_len, int64_t ldq, int64_t ldo,
                   int64_t cache_batch_stride, int64_t cache_row_stride, float scale) {
  __shared__ float s_m[8][kMaxGroup], s_l[8][kMaxGroup];
  __shared__ float s_acc[8][kMaxGroup][32 * kMaxVpl];
  const int b = blockIdx.x / KV, kvh = blockIdx.x % KV;
  const int G = H / KV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = lane * VPL;                       // this lane's slice of the head dimension
  const bool live = d0 < hd;
  float qv[kMaxGroup][VPL];
#pragma unroll
  for (int g = 0; g < kMaxGroup; ++g)
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      qv[g][i] = (g < G && live && d0 + i < hd)
                     ? __bfloat162float(q[(int64_t)b * ldq + (int64_t)(kvh * G + g) * hd + d0 + i]) * scale : 0.f;
  float m[kMaxGroup], l[kMaxGroup], acc[kMaxGroup][VPL];
#pragma unroll
  for (int g = 0; g < kMaxGroup; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[g][i] = 0.f;
  }
  const bf16* kb = kc + (int64_t)b * cache_batch_stride + (int64_t)kvh * hd;
  const bf16* vb = vc + (int64_t)b * cache_batch_stride + (int64_t)kvh * hd;
  for (int j = warp; j < kv_len; j += 8) {
    float kx[VPL], vx[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const bool ok = live && d0 + i < hd;
      kx[i] = ok ? __bfloat162float(kb[(int64_t)j * cache_row_stride + d0 + i]) : 0.f;
      vx[i] = ok ? __bfloat162float(vb[(int64_t)j * cache_row_stride + d0 + i]) : 0.f;
    }
#pragma unroll
    for (int g = 0; g < kMaxGroup; ++g) {
      if (g >= G) break;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) s += qv[g][i] * kx[i];
      s = warp_sum(s);
      const float mn = fmaxf(m[g], s);
      const float corr = __expf(m[g] - mn), p = __expf(s - mn);
      l[g] = l[g] * corr + p;
#pragma unroll
      for (int i = 0; i < VPL; ++i) acc[g][i] = acc[g][i] * corr + p * vx[i];
      m[g] = mn;
    }
  }
  for (int g = 0; g < G; ++g) {
    if (lane == 0) { s_m[warp][g] = m[g]; s_l[warp][g] = l[g]; }
#pragma unroll
    for (int i = 0; i < VPL; ++i) s_acc[warp][g][d0 + i] = acc[g][i];
  }
  __syncthreads();
  // merge the 8 partial softmaxes: thread t handles (g, d) pairs
  for (int idx = threadIdx.x; idx < G * hd; idx += kDecThreads) {
    const int g = idx / hd, d = idx % hd;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) mm = fmaxf(mm, s_m[w][g]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float c = (s_m[w][g] == -INFINITY) ? 0.f : __expf(s_m[w][g] - mm);
      num += c * s_acc[w][g][d];
      den += c * s_l[w][g];
    }
    o[(int64_t)b * ldo + (int64_t)(kvh * G + g) * hd + d] = __float2bfloat16_rn(num / den);
  }
}

}  // namespace

}  // namespace csm

using namespace csm;

extern "C" int csm_attn_decode(const void* q, const void* k_cache, const void* v_cache, void* o, int32_t batch,
                               int32_t heads, int32_t kv_heads, int32_t head_dim, int32_t kv_len, int64_t ldq,
                               int64_t ldo, int64_t cache_batch_stride, int64_t cache_row_stride, float scale,
                               csm_stream_t stream) {
  CSM_REQUIRE(batch >= 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && heads / kv_heads <= kMaxGroup &&
                  head_dim > 0 && head_dim <= 32 * kMaxVpl && kv_len > 0,
              CSM_ERR_SHAPE, "attn_decode: bad shape B=%d H=%d KV=%d hd=%d kv_len=%d", batch, heads, kv_heads, head_dim,
              kv_len);
  CSM_REQUIRE(q && k_cache && v_cache && o, CSM_ERR_SHAPE, "attn_decode: null pointer");
  if (batch == 0) return CSM_OK;
  const int vpl = (head_dim + 31) / 32;
  const unsigned grid = (unsigned)(batch * kv_heads);
  cudaStream_t st = as_stream(stream);
#define LAUNCH(V)                                                                                                    \
  attn_decode_kernel<V><<<grid, kDecThreads, 0, st>>>((const bf16*)q, (const bf16*)k_cache, (const bf16*)v_cache,     \
                                                      (bf16*)o, heads, kv_heads, head_dim, kv_len, ldq, ldo,          \
                                                      cache_batch_stride, cache_row_stride, scale)
  if (vpl <= 1) LAUNCH(1); else if (vpl == 2) LAUNCH(2); else if (vpl == 3) LAUNCH(3); else LAUNCH(4);
#undef LAUNCH
  CSM_CHECK_LAUNCH("attn_decode");
  return CSM_OK;
}
