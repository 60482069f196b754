// global-norm gradient clipping + AdamW in two multi-tensor passes (reference trainer.py:269-278:
// clip_grad_norm_(max_norm) then torch.optim.AdamW.step with per-group learning rates).
//   pass 1  sqnorm_multi_kernel : sum of squares of every gradient -> one device float (fp32 atomics per block)
//   pass 2  adamw_multi_kernel  : reads that float, derives the clip coefficient min(1, max_norm / (norm + 1e-6)) itself
//                                 and applies AdamW with the clipped gradient — the gradients are never rewritten
//                                 (torch: _foreach_norm + _foreach_mul_ over all gradients + the AdamW pass).
// HBM-bound: 2 B/param read in pass 1; pass 2 moves 14 B/param with bf16 state (p, m, v read+write, g read) and
// 28 B/param in the DEFAULT mode: fp32 master weights + fp32 moments (master, m, v fp32 read+write, g read, bf16 p
// written).  The reference trains fp32 parameters with fp32 AdamW (trainer.py:107,166-173) at lr 1e-5 x 0.1: an update
// of ~1e-6 on |w| ~ 0.02 is 1/60 of a bf16 half-ulp, so without the master copy the backbone would never move, and a
// bf16 exp_avg_sq stalls at beta2 = 0.999 (0.1 % per step is below its half-ulp).
// Parameters / gradients may be STRIDED 2-D views (rows of `inner` contiguous elements, `stride` apart): the LoRA B
// factors live as blocks of one block-diagonal tail operand (csm/autograd.py).  State tensors are always dense.
// Tensor lists travel BY VALUE in the kernel parameters (chunks of kMaxTensors), so a captured CUDA graph keeps them.
// The step counter and the squared norm live in device memory: nothing here depends on a host-side value that changes
// between graph replays.
#include "common.cuh"

namespace csm {

namespace {

constexpr int kMaxTensors = 24;
constexpr int kChunk = 16384;          // elements per block

struct TensorList {
  void* p[kMaxTensors];
  const void* g[kMaxTensors];
  void* m[kMaxTensors];
  void* v[kMaxTensors];
  float* master[kMaxTensors];          // fp32 master copy of p (master mode), dense
  int64_t numel[kMaxTensors];
  int64_t inner[kMaxTensors];          // contiguous run of p / g (== numel when dense)
  int64_t p_stride[kMaxTensors];       // distance between runs, in elements
  int64_t g_stride[kMaxTensors];
  int first_block[kMaxTensors + 1];    // blocks [first_block[t], first_block[t+1]) belong to tensor t
  float lr[kMaxTensors];
  float wd[kMaxTensors];
  int n;
};

__device__ __forceinline__ int find_tensor(const TensorList& tl, int block) {
  int t = 0;
  while (t + 1 < tl.n && block >= tl.first_block[t + 1]) ++t;
  return t;
}

__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

__global__ void __launch_bounds__(256)
sqnorm_multi_kernel(const __grid_constant__ TensorList tl, float* __restrict__ out_sq) {
  __shared__ float red[8];
  const int t = find_tensor(tl, blockIdx.x);
  const int64_t n = tl.numel[t];
  const int64_t start = (int64_t)(blockIdx.x - tl.first_block[t]) * kChunk;
  const int64_t end = start + kChunk < n ? start + kChunk : n;
  const bf16* g = reinterpret_cast<const bf16*>(tl.g[t]);
  float acc = 0.f;
  const int64_t inner = tl.inner[t], gs = tl.g_stride[t];
  if (inner != n) {                    // strided rows (small LoRA factors): scalar walk
    for (int64_t j = start + threadIdx.x; j < end; j += blockDim.x) {
      const float x = __bfloat162float(g[(j / inner) * gs + (j % inner)]);
      acc += x * x;
    }
  } else if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    constexpr int U = 4;
    int64_t i = start + (int64_t)threadIdx.x * 8;
    for (; i + (U - 1) * 2048 + 8 <= end; i += U * 2048) {
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k) u[k] = ld_nc16(g + i + k * 2048);
#pragma unroll
      for (int k = 0; k < U; ++k) {
        float f[8];
        unpack8f(u[k], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += f[e] * f[e];
      }
    }
    for (; i + 8 <= end; i += 2048) {
      float f[8];
      unpack8f(ld_nc16(g + i), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += f[e] * f[e];
    }
    // ragged tail of the chunk (numel not a multiple of 8)
    const int64_t tail = start + ((end - start) / 8) * 8;
    for (int64_t j = tail + threadIdx.x; j < end; j += blockDim.x) {
      const float x = __bfloat162float(g[j]);
      acc += x * x;
    }
  } else {
    for (int64_t j = start + threadIdx.x; j < end; j += blockDim.x) {
      const float x = __bfloat162float(g[j]);
      acc += x * x;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = red[threadIdx.x];
    s += __shfl_xor_sync(0xffu, s, 4);
    s += __shfl_xor_sync(0xffu, s, 2);
    s += __shfl_xor_sync(0xffu, s, 1);
    if (threadIdx.x == 0) atomicAdd(out_sq, s);
  }
}

struct AdamScalars {
  float beta1, beta2, eps, max_norm;   // max_norm <= 0: no clipping
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float lr, float wd, float b1, float b2,
                                          float eps, float inv_bc1, float inv_sqrt_bc2) {
  p *= 1.f - lr * wd;                                  // decoupled weight decay (torch.optim.AdamW)
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p -= lr * inv_bc1 * (m / denom);
}

__device__ __forceinline__ void ld8_f32(const float* src, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8_f32(float* dst, const float* f) {
  *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// kMaster: p (bf16) is the rounded image of an fp32 master copy that carries the update.  kF32: fp32 moments.
template <bool kMaster, bool kF32>
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const __grid_constant__ TensorList tl, const float* __restrict__ sq_norm,
                   const float* __restrict__ step_ptr, const AdamScalars sc) {
  const int t = find_tensor(tl, blockIdx.x);
  const int64_t n = tl.numel[t];
  const int64_t start = (int64_t)(blockIdx.x - tl.first_block[t]) * kChunk;
  const int64_t end = start + kChunk < n ? start + kChunk : n;
  bf16* p = reinterpret_cast<bf16*>(tl.p[t]);
  const bf16* g = reinterpret_cast<const bf16*>(tl.g[t]);
  float* w32 = tl.master[t];
  bf16* m16 = reinterpret_cast<bf16*>(tl.m[t]);
  bf16* v16 = reinterpret_cast<bf16*>(tl.v[t]);
  float* m32 = reinterpret_cast<float*>(tl.m[t]);
  float* v32 = reinterpret_cast<float*>(tl.v[t]);
  const float lr = tl.lr[t], wd = tl.wd[t];
  const int64_t inner = tl.inner[t], ps = tl.p_stride[t], gs = tl.g_stride[t];
  const bool dense = inner == n;
  float clip = 1.f;
  if (sc.max_norm > 0.f) {
    const float norm = sqrtf(*sq_norm);
    clip = fminf(1.f, sc.max_norm / (norm + 1e-6f));   // torch.nn.utils.clip_grad_norm_
  }
  const float step = *step_ptr;
  const float inv_bc1 = 1.f / (1.f - powf(sc.beta1, step));
  const float inv_sqrt_bc2 = rsqrtf(1.f - powf(sc.beta2, step));
  // 8-element vectors never straddle a run when inner and the strides are multiples of 8 (chunk starts are too)
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(tl.m[t]) |
                     reinterpret_cast<uintptr_t>(tl.v[t]) | reinterpret_cast<uintptr_t>(w32)) & 15u) == 0 &&
                   (dense || (((inner | ps | gs) & 7) == 0));
  if (vec) {
    for (int64_t i = start + (int64_t)threadIdx.x * 8; i + 8 <= end; i += 2048) {
      int64_t po = i, go = i;
      if (!dense) {
        const int64_t r = i / inner, c = i - r * inner;
        po = r * ps + c;
        go = r * gs + c;
      }
      float pf[8], gf[8], mf[8], vf[8];
      if (kMaster) ld8_f32(w32 + i, pf); else unpack8f(*reinterpret_cast<const uint4*>(p + po), pf);
      unpack8f(ld_nc16(g + go), gf);
      if (kF32) { ld8_f32(m32 + i, mf); ld8_f32(v32 + i, vf); }
      else { unpack8f(*reinterpret_cast<const uint4*>(m16 + i), mf); unpack8f(*reinterpret_cast<const uint4*>(v16 + i), vf); }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        adam_elem(pf[e], gf[e] * clip, mf[e], vf[e], lr, wd, sc.beta1, sc.beta2, sc.eps, inv_bc1, inv_sqrt_bc2);
      if (kMaster) st8_f32(w32 + i, pf);
      *reinterpret_cast<uint4*>(p + po) = pack8f(pf);
      if (kF32) { st8_f32(m32 + i, mf); st8_f32(v32 + i, vf); }
      else { *reinterpret_cast<uint4*>(m16 + i) = pack8f(mf); *reinterpret_cast<uint4*>(v16 + i) = pack8f(vf); }
    }
  }
  const int64_t scalar_from = vec ? start + ((end - start) / 8) * 8 : start;
  for (int64_t j = scalar_from + threadIdx.x; j < end; j += blockDim.x) {
    const int64_t po = dense ? j : (j / inner) * ps + (j % inner);
    const int64_t go = dense ? j : (j / inner) * gs + (j % inner);
    float pf = kMaster ? w32[j] : __bfloat162float(p[po]);
    float mf = kF32 ? m32[j] : __bfloat162float(m16[j]);
    float vf = kF32 ? v32[j] : __bfloat162float(v16[j]);
    adam_elem(pf, __bfloat162float(g[go]) * clip, mf, vf, lr, wd, sc.beta1, sc.beta2, sc.eps, inv_bc1, inv_sqrt_bc2);
    if (kMaster) w32[j] = pf;
    p[po] = __float2bfloat16_rn(pf);
    if (kF32) { m32[j] = mf; v32[j] = vf; }
    else { m16[j] = __float2bfloat16_rn(mf); v16[j] = __float2bfloat16_rn(vf); }
  }
}

__global__ void optim_prepare_kernel(float* sq_norm, float* step) {
  *sq_norm = 0.f;
  *step += 1.f;
}

// walks the host-side tensor table in chunks of kMaxTensors (and <= 65535 * 16 blocks) and calls launch(chunk)
template <typename F>
int for_each_chunk(int n, const int64_t* numel, F&& fill_and_launch) {
  int i = 0;
  while (i < n) {
    TensorList tl;
    tl.n = 0;
    int blocks = 0;
    while (i < n && tl.n < kMaxTensors) {
      const int64_t nb = (numel[i] + kChunk - 1) / kChunk;
      if (numel[i] <= 0) { ++i; continue; }
      if (tl.n > 0 && blocks + nb > (1 << 20)) break;
      tl.first_block[tl.n] = blocks;
      tl.numel[tl.n] = numel[i];
      int rc = fill_and_launch(tl, tl.n, i, /*launch=*/false, 0);
      if (rc) return rc;
      blocks += (int)nb;
      ++tl.n;
      ++i;
    }
    if (tl.n == 0) continue;
    tl.first_block[tl.n] = blocks;
    int rc = fill_and_launch(tl, 0, 0, /*launch=*/true, blocks);
    if (rc) return rc;
  }
  return CSM_OK;
}

}  // namespace

}  // namespace csm

using namespace csm;

extern "C" int csm_adamw_clip_step_v2(void* const* params, const void* const* grads, void* const* exp_avg,
                                      void* const* exp_avg_sq, void* const* master, const int64_t* numel,
                                      const int64_t* inner, const int64_t* p_stride, const int64_t* g_stride,
                                      const float* lr, const float* weight_decay, int32_t n_tensors, float beta1,
                                      float beta2, float eps, float max_norm, int32_t state_fp32, float* step_dev,
                                      float* sq_norm_dev, csm_stream_t stream) {
  CSM_REQUIRE(n_tensors >= 0 && step_dev && sq_norm_dev, CSM_ERR_SHAPE, "adamw_clip_step: bad arguments");
  CSM_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps > 0.f, CSM_ERR_SHAPE,
              "adamw_clip_step: betas must be in [0, 1) and eps > 0");
  for (int i = 0; i < n_tensors; ++i) {
    CSM_REQUIRE(!inner || (inner[i] >= 1 && numel[i] % inner[i] == 0), CSM_ERR_SHAPE,
                "adamw_clip_step: tensor %d: numel %lld is not a whole number of runs of %lld", i, (long long)numel[i],
                (long long)(inner ? inner[i] : 0));
    CSM_REQUIRE(!master || master[i], CSM_ERR_SHAPE, "adamw_clip_step: tensor %d has no master copy", i);
  }
  cudaStream_t st = as_stream(stream);
  optim_prepare_kernel<<<1, 1, 0, st>>>(sq_norm_dev, step_dev);     // zero the norm accumulator, step += 1
  CSM_CHECK_LAUNCH("optim_prepare");
  if (n_tensors == 0) return CSM_OK;
  auto fill_layout = [&](TensorList& tl, int slot, int i) {
    const bool strided = inner && inner[i] != numel[i];
    tl.inner[slot] = strided ? inner[i] : numel[i];
    tl.p_stride[slot] = strided ? p_stride[i] : numel[i];
    tl.g_stride[slot] = strided ? g_stride[i] : numel[i];
  };
  int rc = CSM_OK;
  if (max_norm > 0.f) {
    rc = for_each_chunk(n_tensors, numel, [&](TensorList& tl, int slot, int i, bool launch, int blocks) -> int {
      if (!launch) { tl.g[slot] = grads[i]; fill_layout(tl, slot, i); return CSM_OK; }
      sqnorm_multi_kernel<<<blocks, 256, 0, st>>>(tl, sq_norm_dev);
      CSM_CHECK_LAUNCH("sqnorm_multi");
      return CSM_OK;
    });
    if (rc) return rc;
  }
  const AdamScalars sc{beta1, beta2, eps, max_norm};
  return for_each_chunk(n_tensors, numel, [&](TensorList& tl, int slot, int i, bool launch, int blocks) -> int {
    if (!launch) {
      tl.p[slot] = params[i]; tl.g[slot] = grads[i]; tl.m[slot] = exp_avg[i]; tl.v[slot] = exp_avg_sq[i];
      tl.master[slot] = master ? reinterpret_cast<float*>(master[i]) : nullptr;
      tl.lr[slot] = lr[i]; tl.wd[slot] = weight_decay[i];
      fill_layout(tl, slot, i);
      return CSM_OK;
    }
    if (master && state_fp32) adamw_multi_kernel<true, true><<<blocks, 256, 0, st>>>(tl, sq_norm_dev, step_dev, sc);
    else if (master) adamw_multi_kernel<true, false><<<blocks, 256, 0, st>>>(tl, sq_norm_dev, step_dev, sc);
    else if (state_fp32) adamw_multi_kernel<false, true><<<blocks, 256, 0, st>>>(tl, sq_norm_dev, step_dev, sc);
    else adamw_multi_kernel<false, false><<<blocks, 256, 0, st>>>(tl, sq_norm_dev, step_dev, sc);
    CSM_CHECK_LAUNCH("adamw_multi");
    return CSM_OK;
  });
}

extern "C" int csm_adamw_clip_step(void* const* params, const void* const* grads, void* const* exp_avg,
                                   void* const* exp_avg_sq, const int64_t* numel, const float* lr,
                                   const float* weight_decay, int32_t n_tensors, float beta1, float beta2, float eps,
                                   float max_norm, float* step_dev, float* sq_norm_dev, csm_stream_t stream) {
  return csm_adamw_clip_step_v2(params, grads, exp_avg, exp_avg_sq, nullptr, numel, nullptr, nullptr, nullptr, lr,
                                weight_decay, n_tensors, beta1, beta2, eps, max_norm, 0, step_dev, sq_norm_dev, stream);
}
