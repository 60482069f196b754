/* csm_b200.h — C ABI of libcsm_b200.so: the sm_100a kernels behind the CSM training step.
 *
 * Drop-in boundary.  The reference (imaginateit/csm-train-pytorch) is pure Python
 * and reaches its arithmetic through torch / torchtune calls; each entry point below replaces one of
 * those call sites (cited per function, paths relative to /root/reference).  A maintainer binds them
 * with ctypes (INTEGRATION.md shows the stub); the product binding is csm-train-pytorch_b200/csm/_lib.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current device unless marked host; the caller owns all
 *     buffers (inputs, outputs, workspace); nothing here allocates, frees or retains device memory.
 *   - bf16 storage, fp32 accumulation. "ld*" are row strides in ELEMENTS. Rows must be 16-byte aligned
 *     for the tensor-core paths; other shapes run the scalar-FMA small-shape kernels (tiny test model).
 *   - all calls are asynchronous on `stream` (a cudaStream_t) and never synchronise the host.
 *   - return 0 on success; <0 on error: -1 bad shape/argument, -2 misaligned pointer or stride,
 *     -3 unsupported device (needs sm_100), -4 CUDA launch/runtime error. csm_last_error() returns a
 *     thread-local message.  There is NO CPU fallback.
 */
#ifndef CSM_B200_H
#define CSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* csm_stream_t; /* cudaStream_t */

#define CSM_ABI_VERSION 2

#define CSM_OK 0
#define CSM_ERR_SHAPE (-1)
#define CSM_ERR_ALIGN (-2)
#define CSM_ERR_ARCH (-3)
#define CSM_ERR_CUDA (-4)

#define CSM_DT_BF16 0
#define CSM_DT_F32 1
/* OR'ed into csm_gemm_bf16's c_dtype next to CSM_DT_F32: the residual R is fp32 as well (the fp32 residual stream of
 * the transformer stacks: h = x + attn(...), out = h + mlp(...) never round to bf16 between layers) */
#define CSM_DT_RES_F32 2

/* GEMM back-end selector (csm_gemm_bf16 `backend`): AUTO picks tcgen05 when the shape is tileable. */
#define CSM_GEMM_AUTO 0
#define CSM_GEMM_SIMT 1
#define CSM_GEMM_TCGEN05 2

int csm_abi_version(void);
const char* csm_last_error(void);
/* 1 if the current device is compute capability 10.x, else 0 (no CUDA call fails on a CPU-only host: returns 0). */
int csm_device_supported(void);
/* number of kernel launches issued through this library by the calling process so far */
int64_t csm_launch_count(void);

/* ---- A2: Model._embed_tokens + mask-mul + sum (src/csm/models/model.py:206-217, src/csm/training/utils.py:85-87)
 * h[n,:] = sum_{c<=C} mask[n,c] * table_c[idx[n,c],:], summed in fp32 in order c=0..C, rounded once to bf16.
 * idx[n,c] = tokens[n,c] + c*audio_vocab (c<C, into audio_emb); idx[n,C] = tokens[n,C] (into text_emb).
 * idx_out / mask_out (nullable) export the integer index and effective mask the kernel used (bit-exact parity hook).
 * status (nullable int32[1]) is set to 1 if any index was out of range (the row is then treated as masked). */
int csm_embed_gather_sum_fwd(const int64_t* tokens, const uint8_t* mask, const void* audio_emb,
                             const void* text_emb, void* h, int64_t* idx_out, uint8_t* mask_out,
                             int32_t* status, int64_t n_frames, int32_t codebooks, int64_t audio_vocab,
                             int64_t text_vocab, int32_t dim, csm_stream_t stream);
/* backward of the above: d_table_c[idx[n,c],:] += mask[n,c]*dh[n,:] (bf16 atomics; either table grad may be NULL). */
int csm_embed_gather_sum_bwd(const int64_t* tokens, const uint8_t* mask, const void* dh, void* d_audio_emb,
                             void* d_text_emb, int64_t n_frames, int32_t codebooks, int64_t audio_vocab,
                             int64_t text_vocab, int32_t dim, csm_stream_t stream);

/* The same two ops on the COMPACT device format of the data pipeline (the reference's collate,
 * training_data.py:379-408, feeds int64 tokens and a bool mask: 297 bytes per frame): rows int32 [n_frames, C+1] holds
 * PRE-OFFSET table rows (audio column c: id + c*audio_vocab, text column: id), mask_bits one uint64 per frame whose bit
 * c is the mask of column c — 140 bytes per frame over PCIe and out of HBM.  Same summation order: h is bit-identical
 * to csm_embed_gather_sum_fwd on the unpacked batch.  Needs C < 64 and table rows that fit int32. */
int csm_embed_gather_sum_packed_fwd(const int32_t* rows, const uint64_t* mask_bits, const void* audio_emb,
                                    const void* text_emb, void* h, int32_t* status, int64_t n_frames, int32_t codebooks,
                                    int64_t audio_vocab, int64_t text_vocab, int32_t dim, csm_stream_t stream);
int csm_embed_gather_sum_packed_bwd(const int32_t* rows, const uint64_t* mask_bits, const void* dh, void* d_audio_emb,
                                    void* d_text_emb, int64_t n_frames, int32_t codebooks, int64_t audio_vocab,
                                    int64_t text_vocab, int32_t dim, csm_stream_t stream);

/* ---- A7 decoder input assembly (model.py:176,189-191 teacher-forced; _embed_audio model.py:202-204)
 * x[f,0,:] = h[b_f*seq + p_f,:]; x[f,1+i,:] = audio_emb[targets[b_f,p_f,i] + i*audio_vocab,:], i < C-1.
 * frame_idx int64 [n_sel,2] = (b,p); targets int64 [batch, tgt_len, C]. */
int csm_decoder_input_fwd(const void* h, const void* audio_emb, const int64_t* targets, const int64_t* frame_idx,
                          void* x, int64_t n_sel, int64_t seq, int64_t tgt_len, int32_t codebooks,
                          int64_t audio_vocab, int32_t dim, csm_stream_t stream);
/* dh[b*seq+p,:] += dx[f,0,:]; d_audio_emb[row,:] += dx[f,1+i,:] (d_audio_emb nullable). */
int csm_decoder_input_bwd(const void* dx, const int64_t* targets, const int64_t* frame_idx, void* dh,
                          void* d_audio_emb, int64_t n_sel, int64_t seq, int64_t tgt_len, int32_t codebooks,
                          int64_t audio_vocab, int32_t dim, csm_stream_t stream);

/* ---- torchtune RMSNorm (call sites model.py:13-25 norm_eps; restated in oracle/torchtune_shim.py)
 * y = bf16(x * rsqrt(mean(x^2)+eps)) * scale ; rstd[rows] fp32 is saved for backward.
 * x_dtype: CSM_DT_BF16, or CSM_DT_F32 when x is the fp32 residual stream (then y = bf16(x * rstd * scale): torchtune's
 * `.type_as(x)` does not round an fp32 x).  y, dy, dres and dx are bf16 in both cases. */
int csm_rmsnorm_fwd(const void* x, const void* scale, void* y, float* rstd, int64_t rows, int32_t dim,
                    float eps, int32_t x_dtype, csm_stream_t stream);
/* dx = dres + d(rmsnorm)/dx (dres nullable); dscale_f32[dim] += sum_rows dy*xhat (nullable; fp32 atomics). */
int csm_rmsnorm_bwd(const void* dy, const void* x, const void* scale, const float* rstd, const void* dres,
                    void* dx, float* dscale_f32, int64_t rows, int32_t dim, int32_t x_dtype, csm_stream_t stream);

/* ---- torchtune Llama3ScaledRoPE, interleaved pairs, fp32 math (restated in oracle/torchtune_shim.py)
 * in place on x[rows, heads, head_dim] with row stride ldx; position = positions[row] when `positions` (int32 [rows],
 * nullable; sequence packing: positions restart with every packed sample) is given, else row % seq_len;
 * cache fp32 [max_seq, head_dim/2, 2] = (cos, sin). inverse != 0 applies the transpose rotation (backward). */
int csm_rope(void* x, const float* cache, int64_t rows, int32_t seq_len, int32_t heads, int32_t head_dim,
             int64_t ldx, int32_t inverse, const int32_t* positions, csm_stream_t stream);

/* ---- F.linear / torch.mm / their autograd (model.py:124-126,187; every torchtune projection)
 * C[M,N] (=|+=) alpha * op(A)[M,K] * op(B)[K,N] (+ A2[M,K2] * B2[N,K2]^T) (+ R[M,N])
 *   transA == 0: A stored [M,K] row-major (lda);  transA != 0: stored [K,M] row-major.
 *   transB == 0: B stored [N,K] row-major (nn.Linear weight layout); transB != 0: stored [K,N] row-major.
 *   A2/B2 (nullable): LoRA low-rank tail (K2 <= 256: up to four extra K blocks), stored with the SAME orientation as A/B (A2 is [M,K2] or,
 *   when transA, [K2,M]; B2 is [N,K2] or, when transB, [K2,N]); fused into the main loop as one more K block
 *   (lora.py:82-105: y = x W0^T + (alpha/r)(x A^T) B^T; several adapters side by side: csm_lora_mask_rows).
 *   R (nullable, bf16, ldr): residual added in the epilogue.  c_dtype: CSM_DT_BF16 | CSM_DT_F32.
 *   accumulate != 0: C += result (read-modify-write in c_dtype). */
int csm_gemm_bf16(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int32_t transA, int32_t transB,
                  int32_t c_dtype, int32_t accumulate, float alpha, const void* A2, const void* B2,
                  int64_t K2, int64_t lda2, int64_t ldb2, int32_t backend, csm_stream_t stream);

/* ---- optimiser step of the reference trainers (trainer.py:269-278): torch.nn.utils.clip_grad_norm_(max_norm) followed
 * by torch.optim.AdamW.step (decoupled weight decay, bias correction), as one squared-norm pass and one update pass over
 * n_tensors bf16 tensors.  All seven tables are HOST arrays of n_tensors entries (device pointers / sizes / per-tensor
 * learning rate and weight decay).  step_dev: device float, incremented by this call; sq_norm_dev: device float, holds
 * the squared global gradient norm afterwards.  max_norm <= 0 disables clipping. */
int csm_adamw_clip_step(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                        const int64_t* numel, const float* lr, const float* weight_decay, int32_t n_tensors, float beta1,
                        float beta2, float eps, float max_norm, float* step_dev, float* sq_norm_dev, csm_stream_t stream);

/* Same step with the reference's precision (fp32 parameters + fp32 AdamW, trainer.py:107,166-173) and strided operands:
 *   master      (nullable table) per-tensor fp32 master copy of the parameter: the update is applied to it and the bf16
 *               parameter is re-derived by one rounding;  state_fp32 != 0: exp_avg / exp_avg_sq are fp32 (else bf16).
 *   inner / p_stride / g_stride (nullable tables): tensor i is numel[i] / inner[i] runs of inner[i] contiguous elements,
 *               p_stride[i] (parameter) resp. g_stride[i] (gradient) elements apart; inner[i] == numel[i] means dense.
 *               master and the moments are always dense.  28 B/param in master + fp32-state mode. */
int csm_adamw_clip_step_v2(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                           void* const* master, const int64_t* numel, const int64_t* inner, const int64_t* p_stride,
                           const int64_t* g_stride, const float* lr, const float* weight_decay, int32_t n_tensors,
                           float beta1, float beta2, float eps, float max_norm, int32_t state_fp32, float* step_dev,
                           float* sq_norm_dev, csm_stream_t stream);

/* ---- fused q|k|v projection + RoPE: C[M,N] = A[M,K] B[N,K]^T (+ A2 B2^T), then columns [0, rope_cols) — heads of
 * head_dim — are rotated by position (row % seq_len) exactly as csm_rope would (bf16 rounding of the projection first).
 * N must be a multiple of 32; returns CSM_ERR_SHAPE for shapes the tcgen05 GEMM does not take (then: csm_gemm_bf16 +
 * csm_rope). */
int csm_gemm_bf16_rope(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                       int64_t ldc, const void* A2, const void* B2, int64_t K2, int64_t lda2, int64_t ldb2,
                       const float* rope_cache, int32_t seq_len, int32_t rope_cols, int32_t head_dim,
                       const int32_t* positions /* nullable int32 [M]: see csm_rope */, csm_stream_t stream);

/* ---- skinny GEMM with a long reduction (LoRA t = x A^T, dA = dt^T x, dB = dy^T t): same operand conventions as
 * csm_gemm_bf16 (no residual / tail / accumulate), C bf16 = alpha * op(A) op(B).  The reduction is split into `splits`
 * groups computed by separate CTAs (fp32 partials in `workspace`, csm_gemm_splitk_workspace_bytes()) and summed by a
 * second kernel.  K must be a multiple of splits * 64. */
size_t csm_gemm_splitk_workspace_bytes(int64_t M, int64_t N, int32_t splits);
int csm_gemm_bf16_splitk(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                         int64_t ldc, int32_t transA, int32_t transB, float alpha, int32_t splits, void* workspace,
                         size_t workspace_bytes, csm_stream_t stream);

/* ---- torchtune FeedForward activation: out = silu(gate) * up (elementwise, fp32 math) */
int csm_swiglu_fwd(const void* gate, const void* up, void* out, int64_t rows, int64_t cols, int64_t ldg,
                   int64_t ldu, int64_t ldo, csm_stream_t stream);
int csm_swiglu_bwd(const void* dout, const void* gate, const void* up, void* dgate, void* dup, int64_t rows,
                   int64_t cols, int64_t ldo, int64_t ldg, int64_t ldu, int64_t lddg, int64_t lddu,
                   csm_stream_t stream);

/* ---- torchtune FeedForward w2(silu(w1 x) * w3 x) (llama3_2 builder, reference model.py:11-42) with the SwiGLU
 * fused into the GEMM epilogues.  w13 = [w1; w3] packed [2*inter, K] row-major (gate rows first).
 *   fwd: gate_up[M, 2*inter] = x w13^T (+ A2 B2^T LoRA tail), act[M, inter] = bf16(silu(gate)) * up — one launch.
 *   bwd: dgate_up[M, 2*inter] from dact = dy w2 (+ A2 B2 tail; w2 is [K, inter] row-major), dact never materialised.
 * csm_gemm_swiglu_supported() == 0 (small shapes): call csm_gemm_bf16 + csm_swiglu_{fwd,bwd} instead. */
int csm_gemm_swiglu_supported(int64_t M, int64_t inter, int64_t K);
int csm_gemm_swiglu_fwd(const void* x, const void* w13, void* gate_up, void* act, int64_t M, int64_t inter, int64_t K,
                        int64_t ldx, int64_t ldw, int64_t ldgu, int64_t ldact, const void* A2, const void* B2,
                        int64_t K2, int64_t lda2, int64_t ldb2, csm_stream_t stream);
int csm_gemm_swiglu_bwd(const void* dy, const void* w2, const void* gate_up, void* dgate_up, int64_t M, int64_t inter,
                        int64_t K, int64_t lddy, int64_t ldw, int64_t ldgu, int64_t lddgu, const void* A2,
                        const void* B2, int64_t K2, int64_t lda2, int64_t ldb2, csm_stream_t stream);

/* ---- torchtune MultiHeadAttention core = F.scaled_dot_product_attention(is_causal=True) with GQA
 * q [batch*seq, heads*hd] (ldq), k/v [batch*seq, kv_heads*hd], o like q; kv head j serves q heads
 * j*(heads/kv_heads) .. (j+1)*(heads/kv_heads)-1.  lse fp32 [batch, heads, seq] saved for backward. */
int csm_attn_causal_gqa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t batch,
                            int32_t seq, int32_t heads, int32_t kv_heads, int32_t head_dim, int64_t ldq,
                            int64_t ldk, int64_t ldv, int64_t ldo, float scale, csm_stream_t stream);
/* test hook: force the attention back-end. 0 = automatic (tcgen05 for head_dim 64 and seq >= 128, the
 * short-sequence kernel for seq <= 32 and head_dim 64/128 (depth decoder), mma.sync for other head_dim 64 shapes,
 * scalar otherwise), 1 = scalar, 2 = mma.sync, 3 = tcgen05, 4 = short-sequence (3/4: error if unsupported). */
void csm_set_attn_backend(int32_t backend);
/* A/B hook for the tcgen05 attention forward: 1 (default) the output tile accumulates in TMEM across the key blocks
 * with a lazily updated row maximum, 0 = folded into registers block by block (the round-1 kernel). */
void csm_set_attn_fwd_variant(int32_t variant);
/* test hook: CTA-pair (tcgen05 cta_group::2, 256-row tiles) mode of the tensor-core GEMM. -1 = automatic (large
 * plain GEMMs only), 0 = never, 1 = whenever the shape allows it. */
void csm_set_gemm_cta_pair_mode(int32_t mode);
/* 1: the persistent GEMM draws every tile after a CTA's first from a global counter — robust when another kernel
 * (NCCL's all-reduce overlapping the backward under data parallelism) holds SMs: a CTA that starts late finds less
 * work; the data-parallel full-fine-tune trainer turns it on.  0 (default) = static round-robin assignment, ~1 us per
 * launch cheaper when the GEMM has the GPU to itself.  Needs the scratch registered with csm_gemm_set_streamk_workspace (the counters live in its flag area);
 * without it the static assignment is used. */
void csm_set_gemm_dynamic_tiles(int32_t mode);
/* Stream-K scratch of the CTA-pair GEMM (fp32 partial tiles + self-resetting flags).  The caller allocates
 * csm_gemm_streamk_workspace_bytes() bytes of ZEROED device memory once per device and registers it; with no workspace
 * registered the GEMM never cuts tiles.  One buffer per device: GEMMs that may use it must be issued on one stream.
 * csm_set_gemm_streamk_mode: 0 never cut tiles (default: no gain measured on a power-capped B200), 1 cut tiles when
 * the last wave of whole tiles would be badly filled (only in a -DCSM_GEMM_EXPERIMENTS build; a no-op otherwise).  The
 * default build uses the scratch for the dynamic tile scheduler's counters only. */
/* 1 when the library was built with -DCSM_GEMM_EXPERIMENTS: stream-K and the narrow-tail MMAs are compiled into the GEMM
 * kernels and the two mode setters below act; 0 (the default build): both are compiled out and the setters are no-ops. */
int csm_gemm_experiments_compiled(void);
size_t csm_gemm_streamk_workspace_bytes(void);
void csm_gemm_set_streamk_workspace(void* workspace, size_t bytes);
void csm_set_gemm_streamk_mode(int32_t mode);
/* Experimental (-DCSM_GEMM_EXPERIMENTS builds only; a no-op otherwise), default 0 (measured: bit-identical, no gain —
 * profiles/r1_ctest_narrow_tail.txt): 1 issues the MMAs of a ragged last column tile (K-major B
 * operand; e.g. the 3 leftover columns of the 2051-wide audio heads) with N rounded up to 16 instead of the full tile
 * width.  Results are identical; only tensor-pipe time changes. */
void csm_set_gemm_narrow_tail_mode(int32_t mode);
/* Persistent kernels (GEMM, fused CE) size their grids for (SM count - n): leaves n SMs to a collective that runs
 * concurrently on another stream (data-parallel gradient all-reduce overlapped with backward).  Default 0. */
void csm_set_reserved_sms(int32_t n);
/* workspace: csm_attn_bwd_workspace_bytes() bytes (delta[batch,heads,seq] fp32 + fp32 dk/dv staging). */
size_t csm_attn_bwd_workspace_bytes(int32_t batch, int32_t seq, int32_t heads, int32_t kv_heads,
                                    int32_t head_dim);
int csm_attn_causal_gqa_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse,
                            const void* dout, void* dq, void* dk, void* dv, int32_t batch, int32_t seq,
                            int32_t heads, int32_t kv_heads, int32_t head_dim, int64_t ldq, int64_t ldk,
                            int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                            void* workspace, size_t workspace_bytes, csm_stream_t stream);
/* Same, returning the gradients w.r.t. the UN-rotated q and k: the inverse RoPE (rope_cache: fp32
 * [seq][head_dim/2][cos, sin], position = row within the sequence) is applied to the bf16 dq / dk in the kernels'
 * store epilogues, exactly as csm_rope(inverse=1) would afterwards.  tcgen05 kernels only (head_dim 64, seq >= 128);
 * CSM_ERR_SHAPE otherwise (then: csm_attn_causal_gqa_bwd + csm_rope). */
int csm_attn_causal_gqa_bwd_rope(const void* q, const void* k, const void* v, const void* o, const float* lse,
                                 const void* dout, void* dq, void* dk, void* dv, int32_t batch, int32_t seq,
                                 int32_t heads, int32_t kv_heads, int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv,
                                 int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                                 const float* rope_cache, void* workspace, size_t workspace_bytes, csm_stream_t stream);

/* ---- codebook0_head + F.cross_entropy (utils.py:98-107) and audio_head[i-1] + CE (model.py:187, A7)
 * `groups` independent heads g: logits_g = H_g[M,K] * W_g; loss_rows[g*M+m] = lse - logit[target];
 * H_g = H + g*h_group_stride (row stride ldh), W_g = W + g*w_group_stride, stored [V,K] (transW==0,
 * nn.Linear layout, ldw) or [K,V] (transW!=0, audio_head layout); targets[m*tgt_row_stride + g*tgt_group_stride].
 * Full logits are never written on the tcgen05 path (per-tile online-softmax partials only).
 * lse fp32 [groups*M] is saved for backward. */
size_t csm_linear_ce_workspace_bytes(int64_t M, int64_t V, int64_t K, int32_t groups);
int csm_linear_ce_fwd(const void* H, const void* W, const int64_t* targets, float* loss_rows, float* lse,
                      int64_t M, int64_t V, int64_t K, int32_t groups, int64_t ldh, int64_t h_group_stride,
                      int64_t ldw, int64_t w_group_stride, int32_t transW, int64_t tgt_row_stride,
                      int64_t tgt_group_stride, void* workspace, size_t workspace_bytes, int32_t backend,
                      csm_stream_t stream);
/* dH_g (bf16, lddh / dh_group_stride) = s * (softmax - onehot) * W_g^T ; dW_g (nullable, bf16, same layout as W,
 * accumulated when dw_accumulate) = s * H_g^T (softmax - onehot), with s = grad_scale * (grad_scale_dev ? *grad_scale_dev : 1)
 * (the device scalar is the upstream autograd gradient, read on the device: no host sync).
 * A negative target marks an ignored row (loss 0, zero gradient) like F.cross_entropy's ignore_index. */
int csm_linear_ce_bwd(const void* H, const void* W, const int64_t* targets, const float* lse, float grad_scale,
                      const float* grad_scale_dev, void* dH, void* dW, int32_t dw_accumulate, int64_t M, int64_t V, int64_t K,
                      int32_t groups, int64_t ldh, int64_t h_group_stride, int64_t ldw, int64_t w_group_stride,
                      int32_t transW, int64_t tgt_row_stride, int64_t tgt_group_stride, int64_t lddh,
                      int64_t dh_group_stride, void* workspace, size_t workspace_bytes, int32_t backend,
                      csm_stream_t stream);

/* ---- sequence packing (the reference pads every sample to the batch maximum,
 * training_data.py:379-408): several samples share one row of `seq` frames and attention is block-diagonal causal.
 * seg_start / seg_end: int32 [batch, seq], first position and last position + 1 of the sample that position i of row b
 * belongs to (a padding frame: i and i + 1): query i sees keys seg_start[b,i] <= j <= i.  Same operand layout as
 * csm_attn_causal_gqa_fwd / _bwd_rope; tcgen05 kernels only (head_dim 64, seq >= 128) — CSM_ERR_SHAPE otherwise.  A query
 * tile starts at the key block of its first row's segment start, so packed short samples cost what they cost alone.
 * With rope_cache (nullable) dq / dk come back un-rotated, positions counted from each sample's start. */
int csm_attn_varlen_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t batch, int32_t seq,
                        int32_t heads, int32_t kv_heads, int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv,
                        int64_t ldo, float scale, const int32_t* seg_start, csm_stream_t stream);
int csm_attn_varlen_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                        void* dq, void* dk, void* dv, int32_t batch, int32_t seq, int32_t heads, int32_t kv_heads,
                        int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk,
                        int64_t lddv, float scale, const float* rope_cache, const int32_t* seg_start,
                        const int32_t* seg_end, void* workspace, size_t workspace_bytes, csm_stream_t stream);

/* ---- KV-cache decode attention of Model.generate_frame (model.py:140-195; torchtune attention with kv_cache): one new
 * query position per sample, q [batch, heads*head_dim] (row stride ldq), against the first kv_len cached positions of
 * that sample: caches [batch, max_seq, kv_heads*head_dim] bf16 (cache_row_stride between positions, cache_batch_stride
 * between samples; keys already rotated).  o [batch, heads*head_dim] bf16.  head_dim <= 128, heads/kv_heads <= 8. */
int csm_attn_decode(const void* q, const void* k_cache, const void* v_cache, void* o, int32_t batch, int32_t heads,
                    int32_t kv_heads, int32_t head_dim, int32_t kv_len, int64_t ldq, int64_t ldo,
                    int64_t cache_batch_stride, int64_t cache_row_stride, float scale, csm_stream_t stream);

/* ---- multi-adapter LoRA batching (GPU-native form of MultiSpeakerLoRATrainer, multi_speaker_lora.py:276-300,378-438):
 * the adapters of `adapters` speakers sit side by side in one low-rank tail (per adapted projection: `adapters` blocks of
 * `rank` columns).  Zeroes, in place, every column block of row i of t[rows, cols] (bf16, row stride ldt) that does not
 * belong to adapter_ids[i] (int32; negative: no adapter).  Applied to t = s x A^T and to dts = s dy B. */
int csm_lora_mask_rows(void* t, int64_t ldt, int64_t rows, int32_t cols, const int32_t* adapter_ids, int32_t rank,
                       int32_t adapters, csm_stream_t stream);

/* ---- LoRA input dropout (lora.py:87-90): out[rows, cols] (=|+=) x o keep / (1 - p), keep = hash(*seed_dev, salt, element
 * index) >= p.  Stateless: the backward re-applies the same mask by calling again with the same seed and salt (to x for
 * dA, to dts A — accumulated into dx — for the input gradient).  seed_dev: int64 in device memory (bumped per step). */
int csm_lora_dropout(const void* x, void* out, int64_t rows, int64_t cols, int64_t ldx, int64_t ldo, float p,
                     const int64_t* seed_dev, int64_t salt, int32_t accumulate, csm_stream_t stream);

/* ---- tall-skinny products of the LoRA path (lora.py:87-105 and its autograd; trainer call sites
 * lora_trainer.py:150-178): one dimension R <= 64 (even), the other operand a long stream read once.
 *   rowdot: T[M, R] (bf16, ldt) = alpha * X[M, K] . W, W stored [R, K] (w_is_kr == 0: t = s x A^T) or [K, R]
 *           (w_is_kr != 0: dts = s dy B).  K a multiple of 32, X rows 16-byte aligned.
 *   coldot: G = alpha * X[N, C]^T Tm[N, R], stored [C, R] (out_is_rk == 0: dB = dy^T t) or [R, C] (out_is_rk != 0:
 *           dA = dts^T x).  C a multiple of 8, X rows 16-byte aligned.
 * fp32 accumulation, fixed summation order (bit-reproducible), no workspace.  csm_skinny_supported(kind 0 = rowdot /
 * 1 = coldot, same operands) == 0: use csm_gemm_bf16 / csm_gemm_bf16_splitk.  csm_set_skinny_mode(0) disables (A/B). */
int csm_skinny_supported(int32_t kind, const void* X, const void* W, const void* out, int64_t rows, int64_t cols,
                         int64_t R, int64_t ldx, int64_t ldw, int64_t ldo, int32_t layout);
int csm_skinny_rowdot(const void* X, const void* W, void* T, int64_t M, int64_t K, int64_t R, int64_t ldx, int64_t ldw,
                      int64_t ldt, int32_t w_is_kr, float alpha, csm_stream_t stream);
int csm_skinny_coldot(const void* X, const void* Tm, void* G, int64_t N, int64_t C, int64_t R, int64_t ldx, int64_t ldt,
                      int64_t ldg, int32_t out_is_rk, float alpha, csm_stream_t stream);
void csm_set_skinny_mode(int32_t mode);
/* Programmatic dependent launch between consecutive kernels of the step (default 0; 1, or CSM_PDL=1 in the environment,
 * turns it on): each kernel's launch latency and prologue overlap the tail of its predecessor; every kernel waits
 * (griddepcontrol.wait) before its first global-memory access, so results are unchanged.  Measured neutral to slightly
 * negative on the CUDA-graph-replayed step of a power-capped B200, hence off. */
void csm_set_pdl(int32_t on);

/* ---- small helpers used by the training step */
/* dst_bf16[i] (=|+=) src_f32[i] * scale */
int csm_f32_to_bf16(const float* src, void* dst, int64_t n, float scale, int32_t accumulate, csm_stream_t stream);
/* out_bf16 = a_bf16 + b_bf16 (elementwise, fp32 add) */
int csm_add_bf16(const void* a, const void* b, void* out, int64_t n, csm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CSM_B200_H */
