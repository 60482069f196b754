// K1: 33-way embedding gather-sum with padding mask (+ scatter-add backward) and the decoder-input
// gather.  HBM-bound byte work: one CTA per frame, 16-byte coalesced row reads, up to 8 independent
// loads in flight per thread, fp32 accumulation in codebook order, one bf16 rounding.
// Reference semantics: src/csm/models/model.py:206-217 + src/csm/training/utils.py:85-87.
#include "common.cuh"

namespace csm {

constexpr int kEmbThreads = 256;

// kPacked: the compact device format of the data pipeline (csm/data/frames.py::pack_tokens) —
// `tokens` is int32 [n, C+1] of PRE-OFFSET table rows (audio column c: id + c * V, text column: id) and `mask` one
// uint64 per frame whose bit c is the mask of column c: 140 instead of 297 bytes per frame over PCIe and from HBM.
template <bool kPacked>
__global__ void __launch_bounds__(kEmbThreads)
embed_gather_sum_fwd_kernel(const void* __restrict__ tokens_v, const void* __restrict__ mask_v,
                            const bf16* __restrict__ audio_emb, const bf16* __restrict__ text_emb,
                            bf16* __restrict__ h, int64_t* __restrict__ idx_out,
                            uint8_t* __restrict__ mask_out, int32_t* __restrict__ status, int C, int64_t V,
                            int64_t Vt, int D) {
  extern __shared__ int64_t s_idx[];  // C+1 table row indices, -1 = contributes nothing
  const int64_t n = blockIdx.x;
  const int W = C + 1;
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    int64_t idx;
    bool m, ok;
    if (kPacked) {
      idx = reinterpret_cast<const int32_t*>(tokens_v)[n * W + c];
      m = (reinterpret_cast<const uint64_t*>(mask_v)[n] >> c) & 1ull;
      const int64_t t = (c < C) ? idx - (int64_t)c * V : idx;
      ok = (t >= 0) && (t < ((c < C) ? V : Vt));
    } else {
      const int64_t t = reinterpret_cast<const int64_t*>(tokens_v)[n * W + c];
      m = reinterpret_cast<const uint8_t*>(mask_v)[n * W + c] != 0;
      idx = (c < C) ? t + (int64_t)c * V : t;  // integer half of the op: bit-exact
      ok = (t >= 0) && (t < ((c < C) ? V : Vt));
    }
    if (idx_out) idx_out[n * W + c] = idx;
    if (mask_out) mask_out[n * W + c] = m ? 1 : 0;
    if (!ok && status) *status = 1;
    s_idx[c] = (m && ok) ? idx : -1;
  }
  __syncthreads();
  for (int d0 = threadIdx.x * 8; d0 < D; d0 += blockDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int c0 = 0; c0 < C; c0 += 8) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const int64_t idx = (c < C) ? s_idx[c] : -1;
        v[j] = (idx >= 0) ? ld_nc16(audio_emb + idx * D + d0) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0] += bf16_lo(v[j].x); acc[1] += bf16_hi(v[j].x);
        acc[2] += bf16_lo(v[j].y); acc[3] += bf16_hi(v[j].y);
        acc[4] += bf16_lo(v[j].z); acc[5] += bf16_hi(v[j].z);
        acc[6] += bf16_lo(v[j].w); acc[7] += bf16_hi(v[j].w);
      }
    }
    const int64_t ti = s_idx[C];
    if (ti >= 0) {
      const uint4 v = ld_nc16(text_emb + ti * D + d0);
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x);
      acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z);
      acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(h + n * D + d0) = o;
  }
}

__device__ __forceinline__ void red_add_bf16x8(bf16* dst, uint4 v) {
  bf162* p = reinterpret_cast<bf162*>(dst);
  atomicAdd(p + 0, *reinterpret_cast<bf162*>(&v.x));
  atomicAdd(p + 1, *reinterpret_cast<bf162*>(&v.y));
  atomicAdd(p + 2, *reinterpret_cast<bf162*>(&v.z));
  atomicAdd(p + 3, *reinterpret_cast<bf162*>(&v.w));
}

template <bool kPacked>
__global__ void __launch_bounds__(kEmbThreads)
embed_gather_sum_bwd_kernel(const void* __restrict__ tokens_v, const void* __restrict__ mask_v,
                            const bf16* __restrict__ dh, bf16* __restrict__ d_audio, bf16* __restrict__ d_text,
                            int C, int64_t V, int64_t Vt, int D) {
  extern __shared__ int64_t s_idx[];
  const int64_t n = blockIdx.x;
  const int W = C + 1;
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    if (kPacked) {
      const int64_t idx = reinterpret_cast<const int32_t*>(tokens_v)[n * W + c];
      const int64_t t = (c < C) ? idx - (int64_t)c * V : idx;
      const bool ok = ((reinterpret_cast<const uint64_t*>(mask_v)[n] >> c) & 1ull) && (t >= 0) && (t < ((c < C) ? V : Vt));
      s_idx[c] = ok ? idx : -1;
    } else {
      const int64_t t = reinterpret_cast<const int64_t*>(tokens_v)[n * W + c];
      const bool ok = reinterpret_cast<const uint8_t*>(mask_v)[n * W + c] && (t >= 0) && (t < ((c < C) ? V : Vt));
      s_idx[c] = ok ? ((c < C) ? t + (int64_t)c * V : t) : -1;
    }
  }
  __syncthreads();
  for (int d0 = threadIdx.x * 8; d0 < D; d0 += blockDim.x * 8) {
    const uint4 g = *reinterpret_cast<const uint4*>(dh + n * D + d0);
    if (d_audio) {
      for (int c = 0; c < C; ++c) {
        const int64_t idx = s_idx[c];
        if (idx >= 0) red_add_bf16x8(d_audio + idx * D + d0, g);
      }
    }
    if (d_text && s_idx[C] >= 0) red_add_bf16x8(d_text + s_idx[C] * D + d0, g);
  }
}

// one CTA per (frame f, slot j): slot 0 copies the backbone state, slot 1+i copies emb(i, c_i)
__global__ void __launch_bounds__(128)
decoder_input_fwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ audio_emb,
                         const int64_t* __restrict__ targets, const int64_t* __restrict__ frame_idx,
                         bf16* __restrict__ x, int64_t seq, int64_t tgt_len, int C, int64_t V, int D) {
  const int64_t f = blockIdx.x;
  const int j = blockIdx.y;
  const int64_t b = frame_idx[2 * f], p = frame_idx[2 * f + 1];
  const bf16* src;
  if (j == 0) {
    src = h + (b * seq + p) * D;
  } else {
    int64_t t = targets[(b * tgt_len + p) * C + (j - 1)];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    src = audio_emb + (t + (int64_t)(j - 1) * V) * D;
  }
  bf16* dst = x + (f * C + j) * D;
  for (int d0 = threadIdx.x * 8; d0 < D; d0 += blockDim.x * 8)
    *reinterpret_cast<uint4*>(dst + d0) = ld_nc16(src + d0);
}

__global__ void __launch_bounds__(128)
decoder_input_bwd_kernel(const bf16* __restrict__ dx, const int64_t* __restrict__ targets,
                         const int64_t* __restrict__ frame_idx, bf16* __restrict__ dh,
                         bf16* __restrict__ d_audio, int64_t seq, int64_t tgt_len, int C, int64_t V, int D) {
  const int64_t f = blockIdx.x;
  const int j = blockIdx.y;
  const int64_t b = frame_idx[2 * f], p = frame_idx[2 * f + 1];
  bf16* dst;
  if (j == 0) {
    dst = dh + (b * seq + p) * D;
  } else {
    if (!d_audio) return;
    int64_t t = targets[(b * tgt_len + p) * C + (j - 1)];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    dst = d_audio + (t + (int64_t)(j - 1) * V) * D;
  }
  const bf16* src = dx + (f * C + j) * D;
  for (int d0 = threadIdx.x * 8; d0 < D; d0 += blockDim.x * 8)
    red_add_bf16x8(dst + d0, *reinterpret_cast<const uint4*>(src + d0));
}

}  // namespace csm

using namespace csm;

extern "C" int csm_embed_gather_sum_fwd(const int64_t* tokens, const uint8_t* mask, const void* audio_emb,
                                        const void* text_emb, void* h, int64_t* idx_out, uint8_t* mask_out,
                                        int32_t* status, int64_t n_frames, int32_t codebooks,
                                        int64_t audio_vocab, int64_t text_vocab, int32_t dim,
                                        csm_stream_t stream) {
  CSM_REQUIRE(n_frames >= 0 && codebooks > 0 && dim > 0 && dim % 8 == 0, CSM_ERR_SHAPE,
              "embed_gather_sum_fwd: bad shape n=%lld C=%d D=%d (D must be a multiple of 8)",
              (long long)n_frames, codebooks, dim);
  CSM_REQUIRE(aligned16(audio_emb) && aligned16(text_emb) && aligned16(h), CSM_ERR_ALIGN,
              "embed_gather_sum_fwd: tables and output must be 16-byte aligned");
  if (n_frames == 0) return CSM_OK;
  embed_gather_sum_fwd_kernel<false><<<(unsigned)n_frames, kEmbThreads, (codebooks + 1) * sizeof(int64_t),
                                       as_stream(stream)>>>(tokens, mask, (const bf16*)audio_emb, (const bf16*)text_emb,
                                                            (bf16*)h, idx_out, mask_out, status, codebooks,
                                                            audio_vocab, text_vocab, dim);
  CSM_CHECK_LAUNCH("embed_gather_sum_fwd");
  return CSM_OK;
}

extern "C" int csm_embed_gather_sum_packed_fwd(const int32_t* rows, const uint64_t* mask_bits, const void* audio_emb,
                                               const void* text_emb, void* h, int32_t* status, int64_t n_frames,
                                               int32_t codebooks, int64_t audio_vocab, int64_t text_vocab, int32_t dim,
                                               csm_stream_t stream) {
  CSM_REQUIRE(n_frames >= 0 && codebooks > 0 && codebooks < 64 && dim > 0 && dim % 8 == 0, CSM_ERR_SHAPE,
              "embed_gather_sum_packed_fwd: bad shape n=%lld C=%d D=%d (C + 1 mask bits must fit 64, D % 8 == 0)",
              (long long)n_frames, codebooks, dim);
  CSM_REQUIRE((int64_t)codebooks * audio_vocab < (1ll << 31) && text_vocab < (1ll << 31), CSM_ERR_SHAPE,
              "embed_gather_sum_packed_fwd: table rows do not fit int32");
  CSM_REQUIRE(aligned16(audio_emb) && aligned16(text_emb) && aligned16(h), CSM_ERR_ALIGN,
              "embed_gather_sum_packed_fwd: tables and output must be 16-byte aligned");
  if (n_frames == 0) return CSM_OK;
  embed_gather_sum_fwd_kernel<true><<<(unsigned)n_frames, kEmbThreads, (codebooks + 1) * sizeof(int64_t),
                                      as_stream(stream)>>>(rows, mask_bits, (const bf16*)audio_emb, (const bf16*)text_emb,
                                                           (bf16*)h, nullptr, nullptr, status, codebooks, audio_vocab,
                                                           text_vocab, dim);
  CSM_CHECK_LAUNCH("embed_gather_sum_packed_fwd");
  return CSM_OK;
}

extern "C" int csm_embed_gather_sum_bwd(const int64_t* tokens, const uint8_t* mask, const void* dh,
                                        void* d_audio_emb, void* d_text_emb, int64_t n_frames,
                                        int32_t codebooks, int64_t audio_vocab, int64_t text_vocab,
                                        int32_t dim, csm_stream_t stream) {
  CSM_REQUIRE(n_frames >= 0 && codebooks > 0 && dim > 0 && dim % 8 == 0, CSM_ERR_SHAPE,
              "embed_gather_sum_bwd: bad shape");
  CSM_REQUIRE(aligned16(dh), CSM_ERR_ALIGN, "embed_gather_sum_bwd: dh must be 16-byte aligned");
  if (n_frames == 0 || (!d_audio_emb && !d_text_emb)) return CSM_OK;
  embed_gather_sum_bwd_kernel<false><<<(unsigned)n_frames, kEmbThreads, (codebooks + 1) * sizeof(int64_t),
                                       as_stream(stream)>>>(tokens, mask, (const bf16*)dh, (bf16*)d_audio_emb,
                                                            (bf16*)d_text_emb, codebooks, audio_vocab, text_vocab, dim);
  CSM_CHECK_LAUNCH("embed_gather_sum_bwd");
  return CSM_OK;
}

extern "C" int csm_embed_gather_sum_packed_bwd(const int32_t* rows, const uint64_t* mask_bits, const void* dh,
                                               void* d_audio_emb, void* d_text_emb, int64_t n_frames, int32_t codebooks,
                                               int64_t audio_vocab, int64_t text_vocab, int32_t dim,
                                               csm_stream_t stream) {
  CSM_REQUIRE(n_frames >= 0 && codebooks > 0 && codebooks < 64 && dim > 0 && dim % 8 == 0, CSM_ERR_SHAPE,
              "embed_gather_sum_packed_bwd: bad shape");
  CSM_REQUIRE(aligned16(dh), CSM_ERR_ALIGN, "embed_gather_sum_packed_bwd: dh must be 16-byte aligned");
  if (n_frames == 0 || (!d_audio_emb && !d_text_emb)) return CSM_OK;
  embed_gather_sum_bwd_kernel<true><<<(unsigned)n_frames, kEmbThreads, (codebooks + 1) * sizeof(int64_t),
                                      as_stream(stream)>>>(rows, mask_bits, (const bf16*)dh, (bf16*)d_audio_emb,
                                                           (bf16*)d_text_emb, codebooks, audio_vocab, text_vocab, dim);
  CSM_CHECK_LAUNCH("embed_gather_sum_packed_bwd");
  return CSM_OK;
}

extern "C" int csm_decoder_input_fwd(const void* h, const void* audio_emb, const int64_t* targets,
                                     const int64_t* frame_idx, void* x, int64_t n_sel, int64_t seq,
                                     int64_t tgt_len, int32_t codebooks, int64_t audio_vocab, int32_t dim,
                                     csm_stream_t stream) {
  CSM_REQUIRE(n_sel >= 0 && codebooks > 0 && dim % 8 == 0, CSM_ERR_SHAPE, "decoder_input_fwd: bad shape");
  if (n_sel == 0) return CSM_OK;
  decoder_input_fwd_kernel<<<dim3((unsigned)n_sel, codebooks), 128, 0, as_stream(stream)>>>(
      (const bf16*)h, (const bf16*)audio_emb, targets, frame_idx, (bf16*)x, seq, tgt_len, codebooks,
      audio_vocab, dim);
  CSM_CHECK_LAUNCH("decoder_input_fwd");
  return CSM_OK;
}

extern "C" int csm_decoder_input_bwd(const void* dx, const int64_t* targets, const int64_t* frame_idx, void* dh,
                                     void* d_audio_emb, int64_t n_sel, int64_t seq, int64_t tgt_len,
                                     int32_t codebooks, int64_t audio_vocab, int32_t dim, csm_stream_t stream) {
  CSM_REQUIRE(n_sel >= 0 && codebooks > 0 && dim % 8 == 0, CSM_ERR_SHAPE, "decoder_input_bwd: bad shape");
  if (n_sel == 0) return CSM_OK;
  decoder_input_bwd_kernel<<<dim3((unsigned)n_sel, codebooks), 128, 0, as_stream(stream)>>>(
      (const bf16*)dx, targets, frame_idx, (bf16*)dh, (bf16*)d_audio_emb, seq, tgt_len, codebooks,
      audio_vocab, dim);
  CSM_CHECK_LAUNCH("decoder_input_bwd");
  return CSM_OK;
}
