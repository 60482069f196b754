// RMSNorm fwd/bwd, interleaved-pair RoPE, SwiGLU fwd/bwd and small conversion helpers.
// All HBM-bound: 16-byte vector accesses, fp32 math, one rounding per bf16 the reference rounds
// (oracle/torchtune_shim.py restates torchtune 0.4.0's RMSNorm / Llama3ScaledRoPE / FeedForward).
#include "common.cuh"

namespace csm {

constexpr int kNormThreads = 256;
constexpr int kNormMaxChunks = 4;  // dim <= 256*8*4 = 8192

__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  float t = (lane < nw) ? smem[lane] : 0.f;
  return warp_sum(t);
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// 8 consecutive elements of a row that is stored in bf16 or (the fp32 residual stream) in fp32
template <bool XF32>
__device__ __forceinline__ void load8x(const void* base, int64_t elem, float* f) {
  if (XF32) {
    const float* p = reinterpret_cast<const float*>(base) + elem;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack8(ld_nc16(reinterpret_cast<const bf16*>(base) + elem), f);
  }
}

// XF32: x is the fp32 residual stream; the normalised value is then rounded ONCE, after the scale (torchtune's
// `.type_as(x)` is a no-op for an fp32 x); for a bf16 x the reference's two roundings are kept.
template <bool XF32>
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_fwd_kernel(const void* __restrict__ x, const bf16* __restrict__ scale, bf16* __restrict__ y,
                   float* __restrict__ rstd, int64_t rows, int D, float eps) {
  __shared__ float red[32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    float v[kNormMaxChunks][8];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
      const int d0 = (c * kNormThreads + threadIdx.x) * 8;
      if (d0 < D) {
        load8x<XF32>(x, r * D + d0, v[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += v[c][i] * v[c][i];
      }
    }
    ss = block_sum(ss, red);
    const float rs = rsqrtf(ss / (float)D + eps);
    if (threadIdx.x == 0 && rstd) rstd[r] = rs;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
      const int d0 = (c * kNormThreads + threadIdx.x) * 8;
      if (d0 < D) {
        float f[8], s[8];
        unpack8(*reinterpret_cast<const uint4*>(scale + d0), s);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = XF32 ? v[c][i] * rs * s[i] : round_bf16(v[c][i] * rs) * s[i];
        *reinterpret_cast<uint4*>(y + r * D + d0) = pack8(f);
      }
    }
  }
}

template <bool XF32>
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const void* __restrict__ x, const bf16* __restrict__ scale,
                   const float* __restrict__ rstd, const bf16* __restrict__ dres, bf16* __restrict__ dx,
                   float* __restrict__ dscale, int64_t rows, int D) {
  __shared__ float red[32];
  float ds[kNormMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kNormMaxChunks; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) ds[c][i] = 0.f;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float rs = rstd[r];
    float xh[kNormMaxChunks][8], gs[kNormMaxChunks][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
      const int d0 = (c * kNormThreads + threadIdx.x) * 8;
      if (d0 < D) {
        float g[8], s[8];
        load8x<XF32>(x, r * D + d0, xh[c]);
        unpack8(*reinterpret_cast<const uint4*>(dy + r * D + d0), g);
        unpack8(*reinterpret_cast<const uint4*>(scale + d0), s);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[c][i] *= rs;
          ds[c][i] += g[i] * xh[c][i];
          gs[c][i] = g[i] * s[i];
          dot += gs[c][i] * xh[c][i];
        }
      }
    }
    dot = block_sum(dot, red) / (float)D;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
      const int d0 = (c * kNormThreads + threadIdx.x) * 8;
      if (d0 < D) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (gs[c][i] - xh[c][i] * dot);
        if (dres) {
          float e[8];
          unpack8(*reinterpret_cast<const uint4*>(dres + r * D + d0), e);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += e[i];
        }
        *reinterpret_cast<uint4*>(dx + r * D + d0) = pack8(o);
      }
    }
  }
  if (dscale) {
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
      const int d0 = (c * kNormThreads + threadIdx.x) * 8;
      if (d0 < D) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dscale + d0 + i, ds[c][i]);
      }
    }
  }
}

// ---- warp-per-row RMSNorm for D = 256 * VPL (CSM-1B: 2048 and 1024).  No block barriers, VPL independent 16-byte
// loads per tensor in flight per lane (the CTA-per-row kernels above keep one and are latency-bound at ~2.5 TB/s).
template <int VPL, bool XF32>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_warp_kernel(const void* __restrict__ x, const bf16* __restrict__ scale, bf16* __restrict__ y,
                        float* __restrict__ rstd, int64_t rows, float eps) {
  constexpr int D = 256 * VPL;
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nw) {
    float v[VPL][8];
#pragma unroll
    for (int c = 0; c < VPL; ++c) load8x<XF32>(x, r * D + (c * 32 + lane) * 8, v[c]);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < VPL; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += v[c][i] * v[c][i];
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / (float)D + eps);
    if (lane == 0 && rstd) rstd[r] = rs;
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      float f[8], s[8];
      unpack8(*reinterpret_cast<const uint4*>(scale + (c * 32 + lane) * 8), s);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = XF32 ? v[c][i] * rs * s[i] : round_bf16(v[c][i] * rs) * s[i];
      *reinterpret_cast<uint4*>(y + r * D + (c * 32 + lane) * 8) = pack8(f);
    }
  }
}

// DS: the scale gradient is wanted (full fine-tune).  Its partial sums (8 * VPL per lane) live in WARP-PRIVATE shared
// memory (conflict-free float4 planes), not in registers: with them in registers the kernel sat at 255 registers — one
// 8-warp CTA per SM, 24-29 us per 4096 x 2048 call, half the HBM rate; now both variants run 4-warp CTAs, three per SM
// (15 us without DS).  The CTA folds its four copies at the end: D global fp32 atomics per CTA.
template <int VPL, bool XF32, bool DS>
__global__ void __launch_bounds__(128, 3)
rmsnorm_bwd_warp_kernel(const bf16* __restrict__ dy, const void* __restrict__ x, const bf16* __restrict__ scale,
                        const float* __restrict__ rstd, const bf16* __restrict__ dres, bf16* __restrict__ dx,
                        float* __restrict__ dscale, int64_t rows) {
  constexpr int D = 256 * VPL;
  // element (c, lane, i) of a warp's copy sits in float4 slot (2c + i/4) * 32 + lane: a quarter-warp's float4 accesses
  // cover 128 contiguous bytes
  __shared__ float4 sds[DS ? 4 * (D / 4) : 1];
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4* const my = sds + (DS ? warp * (D / 4) : 0);
  if (DS) {
#pragma unroll
    for (int q = 0; q < 2 * VPL; ++q) my[q * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; r < rows; r += nw) {
    float vx[VPL][8];
    uint4 vg[VPL], ve[VPL];
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      load8x<XF32>(x, r * D + (c * 32 + lane) * 8, vx[c]);
      vg[c] = ld_nc16(dy + r * D + (c * 32 + lane) * 8);
      if (dres) ve[c] = ld_nc16(dres + r * D + (c * 32 + lane) * 8);
    }
    const float rs = rstd[r];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      float xh[8], g[8], s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xh[i] = vx[c][i];
      unpack8(vg[c], g);
      unpack8(*reinterpret_cast<const uint4*>(scale + (c * 32 + lane) * 8), s);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[i] *= rs;
        dot += g[i] * s[i] * xh[i];
      }
      if (DS) {
        float4 a = my[(2 * c) * 32 + lane], b = my[(2 * c + 1) * 32 + lane];
        a.x += g[0] * xh[0]; a.y += g[1] * xh[1]; a.z += g[2] * xh[2]; a.w += g[3] * xh[3];
        b.x += g[4] * xh[4]; b.y += g[5] * xh[5]; b.z += g[6] * xh[6]; b.w += g[7] * xh[7];
        my[(2 * c) * 32 + lane] = a;
        my[(2 * c + 1) * 32 + lane] = b;
      }
    }
    dot = warp_sum(dot) / (float)D;
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      float xh[8], g[8], s[8], o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xh[i] = vx[c][i];
      unpack8(vg[c], g);
      unpack8(*reinterpret_cast<const uint4*>(scale + (c * 32 + lane) * 8), s);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs * (g[i] * s[i] - xh[i] * rs * dot);
      if (dres) {
        float e[8];
        unpack8(ve[c], e);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += e[i];
      }
      *reinterpret_cast<uint4*>(dx + r * D + (c * 32 + lane) * 8) = pack8(o);
    }
  }
  if (DS) {
    __syncthreads();
    const float* flat = reinterpret_cast<const float*>(sds);
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      const int q = j >> 2, e = j & 3, ln = q & 31, ch = q >> 5;            // float4 slot -> (c, lane, i)
      const int col = ((ch >> 1) * 32 + ln) * 8 + (ch & 1) * 4 + e;
      atomicAdd(dscale + col, flat[j] + flat[D + j] + flat[2 * D + j] + flat[3 * D + j]);
    }
  }
}

// one thread per 8 bf16 (= 4 rotation pairs)
__global__ void __launch_bounds__(256)
rope_kernel(bf16* __restrict__ x, const float* __restrict__ cache, int64_t rows, int seq_len, int heads,
            int hd, int64_t ldx, float sgn, const int32_t* __restrict__ pos_of_row) {
  const int vec_per_head = hd / 8;
  const int64_t total = rows * heads * vec_per_head;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int vv = (int)(i % vec_per_head);
    const int64_t t = i / vec_per_head;
    const int hh = (int)(t % heads);
    const int64_t r = t / heads;
    const int pos = pos_of_row ? pos_of_row[r] : (int)(r % seq_len);
    bf16* p = x + r * ldx + (int64_t)hh * hd + vv * 8;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(p), f);
    const float4* cs = reinterpret_cast<const float4*>(cache + ((int64_t)pos * (hd / 2) + vv * 4) * 2);
    const float4 c01 = cs[0], c23 = cs[1];
    const float co[4] = {c01.x, c01.z, c23.x, c23.z};
    const float si[4] = {c01.y * sgn, c01.w * sgn, c23.y * sgn, c23.w * sgn};
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // separate roundings (no FMA contraction): bit-identical to the fp32 torch formula of Llama3ScaledRoPE
      o[2 * j] = __fsub_rn(__fmul_rn(f[2 * j], co[j]), __fmul_rn(f[2 * j + 1], si[j]));
      o[2 * j + 1] = __fadd_rn(__fmul_rn(f[2 * j + 1], co[j]), __fmul_rn(f[2 * j], si[j]));
    }
    *reinterpret_cast<uint4*>(p) = pack8(o);
  }
}

__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const bf16* __restrict__ gate, const bf16* __restrict__ up, bf16* __restrict__ out,
                  int64_t rows, int64_t cols, int64_t ldg, int64_t ldu, int64_t ldo) {
  const int64_t vc = cols / 8, total = rows * vc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vc, c = (i % vc) * 8;
    float g[8], u[8];
    unpack8(*reinterpret_cast<const uint4*>(gate + r * ldg + c), g);
    unpack8(*reinterpret_cast<const uint4*>(up + r * ldu + c), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = round_bf16(g[j] / (1.f + __expf(-g[j]))) * u[j];
    *reinterpret_cast<uint4*>(out + r * ldo + c) = pack8(g);
  }
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ gate, const bf16* __restrict__ up,
                  bf16* __restrict__ dgate, bf16* __restrict__ dup, int64_t rows, int64_t cols, int64_t ldo,
                  int64_t ldg, int64_t ldu, int64_t lddg, int64_t lddu) {
  const int64_t vc = cols / 8, total = rows * vc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vc, c = (i % vc) * 8;
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(*reinterpret_cast<const uint4*>(gate + r * ldg + c), g);
    unpack8(*reinterpret_cast<const uint4*>(up + r * ldu + c), u);
    unpack8(*reinterpret_cast<const uint4*>(dout + r * ldo + c), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = 1.f / (1.f + __expf(-g[j]));
      du[j] = d[j] * g[j] * s;
      dg[j] = d[j] * u[j] * s * (1.f + g[j] * (1.f - s));
    }
    *reinterpret_cast<uint4*>(dgate + r * lddg + c) = pack8(dg);
    *reinterpret_cast<uint4*>(dup + r * lddu + c) = pack8(du);
  }
}

__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n, float scale, int acc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = src[i] * scale;
    if (acc) v += __bfloat162float(dst[i]);
    dst[i] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(256)
add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, int64_t n) {
  const int64_t nv = n / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(*reinterpret_cast<const uint4*>(a + i * 8), x);
    unpack8(*reinterpret_cast<const uint4*>(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    *reinterpret_cast<uint4*>(o + i * 8) = pack8(x);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = nv * 8; i < n; ++i) o[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

static inline unsigned grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t g = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace csm

using namespace csm;

// LoRA input dropout (reference lora.py:87-90: x_for_lora = dropout(x, p)): out = (accumulate ? out : 0) + x o keep / (1 - p).
// The keep mask is a counter-based hash of (seed, salt, element index) — nothing is stored: the backward calls the same
// kernel with the same seed to re-apply the mask (to x for dA, to dts A for dx).  The seed lives in DEVICE memory (the
// trainer bumps it once per step), so a replayed CUDA graph draws a fresh mask every step.
__device__ __forceinline__ uint32_t lora_hash(uint64_t key, uint64_t idx) {
  uint64_t z = key + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

__global__ void __launch_bounds__(256)
lora_dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int64_t rows, int64_t cols, int64_t ldx,
                    int64_t ldo, float p, const int64_t* __restrict__ seed, int64_t salt, int accumulate) {
  const uint64_t key = ((uint64_t)(*seed) * 0xD1342543DE82EF95ull) ^ ((uint64_t)salt * 0xA0761D6478BD642Full);
  const uint32_t thresh = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
  const float inv = 1.f / (1.f - p);
  const int64_t vc = cols / 8, total = rows * vc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vc, c = (i % vc) * 8;
    float f[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(x + r * ldx + c), f);
    if (accumulate) unpack8(*reinterpret_cast<const uint4*>(out + r * ldo + c), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = lora_hash(key, (uint64_t)(r * cols + c + j)) >= thresh ? f[j] * inv : 0.f;
      o[j] = accumulate ? o[j] + v : v;
    }
    *reinterpret_cast<uint4*>(out + r * ldo + c) = pack8(o);
  }
}

// Multi-adapter LoRA (several speakers' adapters side by side in one low-rank tail): t [rows, cols]
// holds, per adapted projection, `adapters` blocks of `rank` columns; a row keeps only the block of ITS adapter
// (ids[row]; a negative id keeps nothing: base model only).  Applied to t = s x A^T after the skinny GEMM and to
// dts = s dy B before the dA / dx GEMMs, so one base GEMM serves every speaker in the batch.
__global__ void __launch_bounds__(256)
lora_mask_rows_kernel(bf16* __restrict__ t, int64_t ldt, int64_t rows, int cols, const int32_t* __restrict__ ids,
                      int rank, int adapters) {
  const int64_t total = rows * cols;
  const int span = rank * adapters;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    if ((c % span) / rank != ids[r]) t[r * ldt + c] = __float2bfloat16_rn(0.f);
  }
}

extern "C" int csm_rmsnorm_fwd(const void* x, const void* scale, void* y, float* rstd, int64_t rows,
                               int32_t dim, float eps, int32_t x_dtype, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && dim > 0 && dim % 8 == 0 && dim <= kNormThreads * 8 * kNormMaxChunks, CSM_ERR_SHAPE,
              "rmsnorm_fwd: dim=%d must be a multiple of 8 and <= %d", dim, kNormThreads * 8 * kNormMaxChunks);
  CSM_REQUIRE(x_dtype == CSM_DT_BF16 || x_dtype == CSM_DT_F32, CSM_ERR_SHAPE, "rmsnorm_fwd: bad x_dtype %d", x_dtype);
  CSM_REQUIRE(aligned16(x) && aligned16(y) && aligned16(scale), CSM_ERR_ALIGN, "rmsnorm_fwd: misaligned pointer");
  if (rows == 0) return CSM_OK;
  const bool f32 = x_dtype == CSM_DT_F32;
  cudaStream_t st = as_stream(stream);
  if ((dim == 2048 || dim == 1024) && rows >= 64) {
    const int64_t ctas = (rows + 7) / 8, cap = (int64_t)num_sms() * 8;
    const unsigned g = (unsigned)(ctas < cap ? ctas : cap);
    const bf16* sc = (const bf16*)scale;
    cudaError_t e;
#define NF(V, F) e = launch_k(rmsnorm_fwd_warp_kernel<V, F>, dim3(g), dim3(256), 0, st, 1, x, sc, (bf16*)y, rstd, rows, eps)
    if (dim == 2048 && f32) NF(8, true);
    else if (dim == 2048) NF(8, false);
    else if (f32) NF(4, true);
    else NF(4, false);
#undef NF
    if (e != cudaSuccess) { set_error("rmsnorm_fwd: launch failed: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
    CSM_CHECK_LAUNCH("rmsnorm_fwd");
    return CSM_OK;
  }
  unsigned grid = (unsigned)(rows < (int64_t)num_sms() * 8 ? rows : (int64_t)num_sms() * 8);
  if (f32) rmsnorm_fwd_kernel<true><<<grid, kNormThreads, 0, st>>>(x, (const bf16*)scale, (bf16*)y, rstd, rows, dim, eps);
  else rmsnorm_fwd_kernel<false><<<grid, kNormThreads, 0, st>>>(x, (const bf16*)scale, (bf16*)y, rstd, rows, dim, eps);
  CSM_CHECK_LAUNCH("rmsnorm_fwd");
  return CSM_OK;
}

extern "C" int csm_rmsnorm_bwd(const void* dy, const void* x, const void* scale, const float* rstd,
                               const void* dres, void* dx, float* dscale_f32, int64_t rows, int32_t dim,
                               int32_t x_dtype, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && dim > 0 && dim % 8 == 0 && dim <= kNormThreads * 8 * kNormMaxChunks, CSM_ERR_SHAPE,
              "rmsnorm_bwd: bad dim=%d", dim);
  CSM_REQUIRE(x_dtype == CSM_DT_BF16 || x_dtype == CSM_DT_F32, CSM_ERR_SHAPE, "rmsnorm_bwd: bad x_dtype %d", x_dtype);
  CSM_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(scale) && aligned16(dres),
              CSM_ERR_ALIGN, "rmsnorm_bwd: misaligned pointer");
  if (rows == 0) return CSM_OK;
  const bool f32 = x_dtype == CSM_DT_F32;
  cudaStream_t st = as_stream(stream);
  const bf16 *gy = (const bf16*)dy, *sc = (const bf16*)scale, *dr = (const bf16*)dres;
  if ((dim == 2048 || dim == 1024) && rows >= 64) {
    // one row per warp, 4-warp CTAs; with dscale the grid is capped at three CTAs per SM (D global atomics per CTA)
    // (whole rows per warp, the same number for every warp: e.g. 4096 rows -> 342 CTAs x 4 warps x 3 rows)
    const unsigned g4 = (unsigned)((rows + 3) / 4);
    const int64_t cap_warps = 4ll * 3 * num_sms();
    const int64_t rows_per_warp = (rows + cap_warps - 1) / cap_warps;
    const unsigned gds = (unsigned)((rows + 4 * rows_per_warp - 1) / (4 * rows_per_warp));
    cudaError_t e;
#define NB(V, F, D_, G) e = launch_k(rmsnorm_bwd_warp_kernel<V, F, D_>, dim3(G), dim3(128), 0, st, 1, gy, x, sc, rstd, dr, \
                                  (bf16*)dx, dscale_f32, rows)
    if (dscale_f32) {
      if (dim == 2048 && f32) NB(8, true, true, gds);
      else if (dim == 2048) NB(8, false, true, gds);
      else if (f32) NB(4, true, true, gds);
      else NB(4, false, true, gds);
    } else {
      if (dim == 2048 && f32) NB(8, true, false, g4);
      else if (dim == 2048) NB(8, false, false, g4);
      else if (f32) NB(4, true, false, g4);
      else NB(4, false, false, g4);
    }
#undef NB
    if (e != cudaSuccess) { set_error("rmsnorm_bwd: launch failed: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
    CSM_CHECK_LAUNCH("rmsnorm_bwd");
    return CSM_OK;
  }
  unsigned grid = (unsigned)(rows < (int64_t)num_sms() * 2 ? rows : (int64_t)num_sms() * 2);
  if (f32) rmsnorm_bwd_kernel<true><<<grid, kNormThreads, 0, st>>>(gy, x, sc, rstd, dr, (bf16*)dx, dscale_f32, rows, dim);
  else rmsnorm_bwd_kernel<false><<<grid, kNormThreads, 0, st>>>(gy, x, sc, rstd, dr, (bf16*)dx, dscale_f32, rows, dim);
  CSM_CHECK_LAUNCH("rmsnorm_bwd");
  return CSM_OK;
}

extern "C" int csm_rope(void* x, const float* cache, int64_t rows, int32_t seq_len, int32_t heads,
                        int32_t head_dim, int64_t ldx, int32_t inverse, const int32_t* positions, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && seq_len > 0 && heads > 0 && head_dim > 0 && head_dim % 8 == 0 && ldx % 8 == 0,
              CSM_ERR_SHAPE, "rope: head_dim=%d and ldx=%lld must be multiples of 8", head_dim, (long long)ldx);
  CSM_REQUIRE(aligned16(x) && aligned16(cache), CSM_ERR_ALIGN, "rope: misaligned pointer");
  if (rows == 0) return CSM_OK;
  const int64_t total = rows * heads * (head_dim / 8);
  rope_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((bf16*)x, cache, rows, seq_len, heads,
                                                                   head_dim, ldx, inverse ? -1.f : 1.f, positions);
  CSM_CHECK_LAUNCH("rope");
  return CSM_OK;
}

extern "C" int csm_swiglu_fwd(const void* gate, const void* up, void* out, int64_t rows, int64_t cols,
                              int64_t ldg, int64_t ldu, int64_t ldo, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && ldg % 8 == 0 && ldu % 8 == 0 && ldo % 8 == 0,
              CSM_ERR_SHAPE, "swiglu_fwd: cols and strides must be multiples of 8");
  CSM_REQUIRE(aligned16(gate) && aligned16(up) && aligned16(out), CSM_ERR_ALIGN, "swiglu_fwd: misaligned");
  if (rows == 0) return CSM_OK;
  swiglu_fwd_kernel<<<grid_for(rows * (cols / 8), 256), 256, 0, as_stream(stream)>>>(
      (const bf16*)gate, (const bf16*)up, (bf16*)out, rows, cols, ldg, ldu, ldo);
  CSM_CHECK_LAUNCH("swiglu_fwd");
  return CSM_OK;
}

extern "C" int csm_swiglu_bwd(const void* dout, const void* gate, const void* up, void* dgate, void* dup,
                              int64_t rows, int64_t cols, int64_t ldo, int64_t ldg, int64_t ldu, int64_t lddg,
                              int64_t lddu, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && ldg % 8 == 0 && ldu % 8 == 0 && ldo % 8 == 0 &&
                  lddg % 8 == 0 && lddu % 8 == 0,
              CSM_ERR_SHAPE, "swiglu_bwd: cols and strides must be multiples of 8");
  CSM_REQUIRE(aligned16(gate) && aligned16(up) && aligned16(dout) && aligned16(dgate) && aligned16(dup),
              CSM_ERR_ALIGN, "swiglu_bwd: misaligned");
  if (rows == 0) return CSM_OK;
  swiglu_bwd_kernel<<<grid_for(rows * (cols / 8), 256), 256, 0, as_stream(stream)>>>(
      (const bf16*)dout, (const bf16*)gate, (const bf16*)up, (bf16*)dgate, (bf16*)dup, rows, cols, ldo, ldg,
      ldu, lddg, lddu);
  CSM_CHECK_LAUNCH("swiglu_bwd");
  return CSM_OK;
}

extern "C" int csm_f32_to_bf16(const float* src, void* dst, int64_t n, float scale, int32_t accumulate,
                               csm_stream_t stream) {
  if (n <= 0) return CSM_OK;
  f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(src, (bf16*)dst, n, scale, accumulate);
  CSM_CHECK_LAUNCH("f32_to_bf16");
  return CSM_OK;
}

extern "C" int csm_add_bf16(const void* a, const void* b, void* out, int64_t n, csm_stream_t stream) {
  if (n <= 0) return CSM_OK;
  CSM_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), CSM_ERR_ALIGN, "add_bf16: misaligned");
  add_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, as_stream(stream)>>>((const bf16*)a, (const bf16*)b,
                                                                           (bf16*)out, n);
  CSM_CHECK_LAUNCH("add_bf16");
  return CSM_OK;
}


extern "C" int csm_lora_mask_rows(void* t, int64_t ldt, int64_t rows, int32_t cols, const int32_t* adapter_ids,
                                  int32_t rank, int32_t adapters, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && cols > 0 && rank > 0 && adapters > 0 && cols % (rank * adapters) == 0 && ldt >= cols,
              CSM_ERR_SHAPE, "lora_mask_rows: cols=%d must be a whole number of %d x %d adapter blocks", cols, adapters,
              rank);
  CSM_REQUIRE(t && adapter_ids, CSM_ERR_SHAPE, "lora_mask_rows: null pointer");
  if (rows == 0) return CSM_OK;
  const int64_t total = rows * cols;
  const unsigned grid = (unsigned)((total + 255) / 256 < (int64_t)num_sms() * 8 ? (total + 255) / 256 : (int64_t)num_sms() * 8);
  lora_mask_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>((bf16*)t, ldt, rows, cols, adapter_ids, rank, adapters);
  CSM_CHECK_LAUNCH("lora_mask_rows");
  return CSM_OK;
}


extern "C" int csm_lora_dropout(const void* x, void* out, int64_t rows, int64_t cols, int64_t ldx, int64_t ldo, float p,
                                const int64_t* seed_dev, int64_t salt, int32_t accumulate, csm_stream_t stream) {
  CSM_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && p >= 0.f && p < 1.f && seed_dev,
              CSM_ERR_SHAPE, "lora_dropout: cols / strides must be multiples of 8 and 0 <= p < 1");
  CSM_REQUIRE(aligned16(x) && aligned16(out), CSM_ERR_ALIGN, "lora_dropout: misaligned pointer");
  if (rows == 0) return CSM_OK;
  const int64_t total = rows * (cols / 8);
  lora_dropout_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const bf16*)x, (bf16*)out, rows, cols, ldx,
                                                                          ldo, p, seed_dev, salt, accumulate);
  CSM_CHECK_LAUNCH("lora_dropout");
  return CSM_OK;
}
