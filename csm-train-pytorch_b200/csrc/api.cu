// extern "C" surface that is not tied to one kernel file: error state, device probe, GEMM/attention dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace csm {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
// Programmatic dependent launch between consecutive kernels of the step (common.cuh); CSM_PDL=1 turns it on.
// OFF by default — measured on B200 (round 2, profiles/r2b_summary.md): the CSM-1B LoRA step replayed as a CUDA graph
// takes 24.2 ms with it (dependents triggered right after each kernel's wait), 23.8-24.0 ms with the implicit trigger
// at kernel exit, 23.7-24.1 ms without: inside a graph the launch gaps it hides are already small, and the GPU is
// power-capped, so idle gaps are not lost throughput.
static int pdl_env_default() {
  const char* e = getenv("CSM_PDL");
  return (e && e[0] == '1') ? 1 : 0;
}
std::atomic<int> g_pdl{pdl_env_default()};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// SMs the persistent kernels size their grids for.  csm_set_reserved_sms(n) keeps n SMs free for a concurrently running
// collective (data-parallel full fine-tune: NCCL's all-reduce CTAs overlap the backward; a persistent 148-CTA GEMM whose
// last CTAs must wait for an SM held by NCCL would take twice as long).
static std::atomic<int> g_reserved_sms{0};
int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      sms = 148;
  }
  const int r = g_reserved_sms.load(std::memory_order_relaxed);
  return sms - r > 8 ? sms - r : 8;
}

int gemm_simt_launch(const void*, const void*, void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                     int64_t, int64_t, int, int, int, int, float, const void*, const void*, int64_t, int64_t,
                     int64_t, cudaStream_t);
bool gemm_tc_supported(const void* A, const void* B, const void* C, const void* R, int64_t M, int64_t N, int64_t K,
                       int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                       const void* A2, const void* B2, int64_t K2, int64_t lda2, int64_t ldb2);
int gemm_tc_launch(const void*, const void*, void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                   int64_t, int64_t, int, int, int, int, float, const void*, const void*, int64_t, int64_t, int64_t,
                   cudaStream_t);
int attn_fwd_simt_launch(const void*, const void*, const void*, void*, float*, int, int, int, int, int, int64_t,
                         int64_t, int64_t, int64_t, float, cudaStream_t);
int attn_bwd_simt_launch(const void*, const void*, const void*, const void*, const float*, const void*, void*,
                         void*, void*, float*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t,
                         int64_t, int64_t, float, cudaStream_t);
bool attn_mma_supported(int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo);
int attn_fwd_mma_launch(const void*, const void*, const void*, void*, float*, int, int, int, int, int, int64_t,
                        int64_t, int64_t, int64_t, float, cudaStream_t);
int attn_bwd_mma_launch(const void*, const void*, const void*, const void*, const float*, const void*, void*,
                        void*, void*, float*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t,
                        int64_t, int64_t, float, cudaStream_t);

bool attn_small_supported(int S, int H, int KV, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                          const void* q, const void* k, const void* v, const void* o);
int attn_fwd_small_launch(const void*, const void*, const void*, void*, float*, int, int, int, int, int, int64_t,
                          int64_t, int64_t, int64_t, float, cudaStream_t);
int attn_bwd_small_launch(const void*, const void*, const void*, const void*, const float*, const void*, void*,
                          void*, void*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                          int64_t, float, cudaStream_t);
bool attn_tc_supported(int S, int hd, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k,
                       const void* v, const void* o);
int attn_fwd_tc_launch(const void*, const void*, const void*, void*, float*, int, int, int, int, int64_t, int64_t,
                       int64_t, int64_t, float, const int32_t*, cudaStream_t);
int attn_bwd_tc_launch(const void*, const void*, const void*, const void*, const float*, const void*, void*, void*,
                       void*, float*, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                       float, const float*, const int32_t*, const int32_t*, cudaStream_t);
static std::atomic<int> g_attn_backend{0};  // 0 auto, 1 scalar, 2 mma.sync, 3 tcgen05, 4 short-sequence
void attn_tc_set_fwd_variant(int v);

int gemm_dispatch(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                  int accumulate, float alpha, const void* A2, const void* B2, int64_t K2, int64_t lda2,
                  int64_t ldb2, int backend, cudaStream_t stream) {
  CSM_REQUIRE(M >= 0 && N >= 0 && K >= 0, CSM_ERR_SHAPE, "gemm: negative dimension");
  CSM_REQUIRE(c_dtype >= 0 && c_dtype <= (CSM_DT_F32 | CSM_DT_RES_F32), CSM_ERR_SHAPE, "gemm: bad c_dtype %d", c_dtype);
  CSM_REQUIRE(!(c_dtype & CSM_DT_RES_F32) || (R && (c_dtype & CSM_DT_F32)), CSM_ERR_SHAPE,
              "gemm: an fp32 residual needs R and an fp32 output");
  CSM_REQUIRE((A2 == nullptr) == (B2 == nullptr), CSM_ERR_SHAPE, "gemm: A2 and B2 must be given together");
  if (M == 0 || N == 0) return CSM_OK;
  const bool tc_ok = gemm_tc_supported(A, B, C, R, M, N, K, lda, ldb, ldc, ldr, transA, transB, c_dtype, A2, B2,
                                       A2 ? K2 : 0, lda2, ldb2);
  if (backend == CSM_GEMM_TCGEN05) {
    CSM_REQUIRE(tc_ok, CSM_ERR_ALIGN,
                "gemm: tcgen05 path needs 16-byte aligned pointers, leading dimensions that are multiples of 8 "
                "and K >= 16 (got M=%lld N=%lld K=%lld lda=%lld ldb=%lld ldc=%lld)",
                (long long)M, (long long)N, (long long)K, (long long)lda, (long long)ldb, (long long)ldc);
  }
  if (tc_ok && backend != CSM_GEMM_SIMT)
    return gemm_tc_launch(A, B, C, R, M, N, K, lda, ldb, ldc, ldr, transA, transB, c_dtype, accumulate, alpha, A2,
                          B2, A2 ? K2 : 0, lda2, ldb2, stream);
  if (backend != CSM_GEMM_SIMT && (double)M * (double)N * (double)K > 1e9) {
    // a large GEMM on the scalar kernel is a performance bug, never silent: say why (first few times only)
    static std::atomic<int> warned{0};
    if (warned.fetch_add(1) < 8)
      fprintf(stderr,
              "[csm_b200] warning: GEMM M=%lld N=%lld K=%lld (transA=%d transB=%d) runs on the scalar kernel: "
              "A=%p lda=%lld B=%p ldb=%lld A2=%p lda2=%lld B2=%p ldb2=%lld K2=%lld\n",
              (long long)M, (long long)N, (long long)K, transA, transB, A, (long long)lda, B, (long long)ldb, A2,
              (long long)lda2, B2, (long long)ldb2, (long long)K2);
  }
  return gemm_simt_launch(A, B, C, R, M, N, K, lda, ldb, ldc, ldr, transA, transB, c_dtype, accumulate, alpha, A2,
                          B2, K2, lda2, ldb2, stream);
}

}  // namespace csm

using namespace csm;

extern "C" int csm_abi_version(void) { return CSM_ABI_VERSION; }
extern "C" void csm_set_attn_backend(int32_t backend) { g_attn_backend.store(backend); }
extern "C" void csm_set_attn_fwd_variant(int32_t v) { csm::attn_tc_set_fwd_variant(v); }
extern "C" void csm_set_reserved_sms(int32_t n) { csm::g_reserved_sms.store(n < 0 ? 0 : n); }
namespace csm {
void gemm_tc_set_cta_pair_mode(int m);
size_t gemm_tc_streamk_workspace_bytes();
void gemm_tc_set_streamk_workspace(void* ptr, size_t bytes);
void gemm_tc_set_streamk_mode(int m);
void gemm_tc_set_narrow_tail_mode(int m);
void gemm_tc_set_dynamic_tiles(int m);
int gemm_tc_experiments_compiled();
}  // namespace csm
extern "C" int csm_gemm_experiments_compiled(void) { return csm::gemm_tc_experiments_compiled(); }
extern "C" size_t csm_gemm_streamk_workspace_bytes(void) { return csm::gemm_tc_streamk_workspace_bytes(); }
extern "C" void csm_gemm_set_streamk_workspace(void* workspace, size_t bytes) {
  csm::gemm_tc_set_streamk_workspace(workspace, bytes);
}
extern "C" void csm_set_gemm_streamk_mode(int32_t mode) { csm::gemm_tc_set_streamk_mode(mode); }
extern "C" void csm_set_gemm_narrow_tail_mode(int32_t mode) { csm::gemm_tc_set_narrow_tail_mode(mode); }
extern "C" void csm_set_gemm_cta_pair_mode(int32_t mode) { csm::gemm_tc_set_cta_pair_mode(mode); }
extern "C" void csm_set_gemm_dynamic_tiles(int32_t mode) { csm::gemm_tc_set_dynamic_tiles(mode); }
extern "C" void csm_set_pdl(int32_t on) { csm::g_pdl.store(on ? 1 : 0); }
extern "C" const char* csm_last_error(void) { return g_err; }
extern "C" int64_t csm_launch_count(void) { return g_launches.load(); }

extern "C" int csm_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

extern "C" int csm_gemm_bf16(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                             int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int32_t transA, int32_t transB,
                             int32_t c_dtype, int32_t accumulate, float alpha, const void* A2, const void* B2,
                             int64_t K2, int64_t lda2, int64_t ldb2, int32_t backend, csm_stream_t stream) {
  return gemm_dispatch(A, B, C, R, M, N, K, lda, ldb, ldc, ldr, transA, transB, c_dtype, accumulate, alpha, A2, B2,
                       K2, lda2, ldb2, backend, as_stream(stream));
}

namespace csm {
bool gemm_tc_swiglu_supported(int64_t M, int64_t inter, int64_t K);
int gemm_tc_swiglu_fwd(const void*, const void*, void*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                       int64_t, const void*, const void*, int64_t, int64_t, int64_t, cudaStream_t);
int gemm_tc_swiglu_bwd(const void*, const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                       int64_t, int64_t, const void*, const void*, int64_t, int64_t, int64_t, cudaStream_t);
}  // namespace csm

namespace csm {
bool gemm_tc_splitk_supported(int64_t M, int64_t N, int64_t K, int splits);
int gemm_tc_splitk(const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, float,
                   int, float*, cudaStream_t);
}  // namespace csm

namespace csm {
int gemm_tc_launch_rope(const void*, const void*, void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                        int64_t, int64_t, int, int, int, int, float, const void*, const void*, int64_t, int64_t, int64_t,
                        const float*, int, int, int, const int32_t*, cudaStream_t);
}  // namespace csm

extern "C" int csm_gemm_bf16_rope(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                                  int64_t ldb, int64_t ldc, const void* A2, const void* B2, int64_t K2, int64_t lda2,
                                  int64_t ldb2, const float* rope_cache, int32_t seq_len, int32_t rope_cols,
                                  int32_t head_dim, const int32_t* rope_pos, csm_stream_t stream) {
  CSM_REQUIRE(csm_device_supported() == 1, CSM_ERR_ARCH, "gemm_rope: needs an sm_100 device");
  CSM_REQUIRE(rope_cache && aligned16(rope_cache) && seq_len > 0 && head_dim >= 8 && head_dim % 8 == 0 &&
                  rope_cols >= 0 && rope_cols <= N && rope_cols % head_dim == 0,
              CSM_ERR_SHAPE, "gemm_rope: bad RoPE geometry (cols=%d head_dim=%d)", rope_cols, head_dim);
  // the rotation lives in the vectorised store path: full 32-column chunks of a bf16 output with an aligned stride
  CSM_REQUIRE(N % 32 == 0 && ldc % 8 == 0 && aligned16(C) &&
                  gemm_tc_supported(A, B, C, nullptr, M, N, K, lda, ldb, ldc, 0, 0, 0, CSM_DT_BF16, A2, B2, A2 ? K2 : 0,
                                    lda2, ldb2),
              CSM_ERR_SHAPE, "gemm_rope: shape not supported by the tcgen05 GEMM (use csm_gemm_bf16 + csm_rope)");
  return gemm_tc_launch_rope(A, B, C, nullptr, M, N, K, lda, ldb, ldc, 0, 0, 0, CSM_DT_BF16, 0, 1.f, A2, A2 ? B2 : nullptr,
                             A2 ? K2 : 0, lda2, ldb2, rope_cache, seq_len, rope_cols, head_dim, rope_pos,
                             as_stream(stream));
}

extern "C" size_t csm_gemm_splitk_workspace_bytes(int64_t M, int64_t N, int32_t splits) {
  return (size_t)splits * (size_t)M * (size_t)N * sizeof(float) + 256;
}

extern "C" int csm_gemm_bf16_splitk(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                                    int64_t ldb, int64_t ldc, int32_t transA, int32_t transB, float alpha,
                                    int32_t splits, void* workspace, size_t workspace_bytes, csm_stream_t stream) {
  CSM_REQUIRE(csm_device_supported() == 1, CSM_ERR_ARCH, "gemm_splitk: needs an sm_100 device");
  CSM_REQUIRE(gemm_tc_splitk_supported(M, N, K, splits), CSM_ERR_SHAPE,
              "gemm_splitk: K=%lld must split into %d groups of a multiple of 64", (long long)K, splits);
  CSM_REQUIRE(gemm_tc_supported(A, B, C, nullptr, M, N, K / splits, lda, ldb, ldc, 0, transA, transB, CSM_DT_BF16,
                                nullptr, nullptr, 0, 0, 0) ||
                  (aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0),
              CSM_ERR_ALIGN, "gemm_splitk: operands must be 16-byte aligned with strides that are multiples of 8");
  CSM_REQUIRE(workspace && aligned16(workspace) && workspace_bytes >= csm_gemm_splitk_workspace_bytes(M, N, splits),
              CSM_ERR_SHAPE, "gemm_splitk: workspace too small");
  return gemm_tc_splitk(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, alpha, splits,
                        reinterpret_cast<float*>(workspace), as_stream(stream));
}

static bool swiglu_args_ok(const void* a, const void* b, const void* c, const void* d, int64_t l0, int64_t l1,
                           int64_t l2, int64_t l3, const void* a2, const void* b2, int64_t K2, int64_t lda2,
                           int64_t ldb2) {
  if (!aligned16(a) || !aligned16(b) || !aligned16(c) || !aligned16(d)) return false;
  if ((l0 | l1 | l2 | l3) & 7) return false;
  if (a2 && (!b2 || !aligned16(a2) || !aligned16(b2) || K2 < 1 || K2 > 256 || (lda2 & 7) || (ldb2 & 7))) return false;
  return true;
}

extern "C" int csm_gemm_swiglu_supported(int64_t M, int64_t inter, int64_t K) {
  return (csm_device_supported() == 1 && gemm_tc_swiglu_supported(M, inter, K)) ? 1 : 0;
}

extern "C" int csm_gemm_swiglu_fwd(const void* x, const void* w13, void* gate_up, void* act, int64_t M, int64_t inter,
                                   int64_t K, int64_t ldx, int64_t ldw, int64_t ldgu, int64_t ldact, const void* A2,
                                   const void* B2, int64_t K2, int64_t lda2, int64_t ldb2, csm_stream_t stream) {
  CSM_REQUIRE(gemm_tc_swiglu_supported(M, inter, K), CSM_ERR_SHAPE,
              "gemm_swiglu_fwd: shape M=%lld I=%lld K=%lld is below the CTA-pair tile grid; use gemm + swiglu",
              (long long)M, (long long)inter, (long long)K);
  CSM_REQUIRE(swiglu_args_ok(x, w13, gate_up, act, ldx, ldw, ldgu, ldact, A2, B2, K2, lda2, ldb2), CSM_ERR_ALIGN,
              "gemm_swiglu_fwd: operands must be 16-byte aligned with strides that are multiples of 8");
  return gemm_tc_swiglu_fwd(x, w13, gate_up, act, M, inter, K, ldx, ldw, ldgu, ldact, A2, A2 ? B2 : nullptr, A2 ? K2 : 0,
                            lda2, ldb2, as_stream(stream));
}

extern "C" int csm_gemm_swiglu_bwd(const void* dy, const void* w2, const void* gate_up, void* dgate_up, int64_t M,
                                   int64_t inter, int64_t K, int64_t lddy, int64_t ldw, int64_t ldgu, int64_t lddgu,
                                   const void* A2, const void* B2, int64_t K2, int64_t lda2, int64_t ldb2,
                                   csm_stream_t stream) {
  CSM_REQUIRE(gemm_tc_swiglu_supported(M, inter, K), CSM_ERR_SHAPE,
              "gemm_swiglu_bwd: shape M=%lld I=%lld K=%lld is below the CTA-pair tile grid; use gemm + swiglu_bwd",
              (long long)M, (long long)inter, (long long)K);
  CSM_REQUIRE(swiglu_args_ok(dy, w2, gate_up, dgate_up, lddy, ldw, ldgu, lddgu, A2, B2, K2, lda2, ldb2), CSM_ERR_ALIGN,
              "gemm_swiglu_bwd: operands must be 16-byte aligned with strides that are multiples of 8");
  return gemm_tc_swiglu_bwd(dy, w2, gate_up, dgate_up, M, inter, K, lddy, ldw, ldgu, lddgu, A2, A2 ? B2 : nullptr,
                            A2 ? K2 : 0, lda2, ldb2, as_stream(stream));
}

extern "C" int csm_attn_causal_gqa_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                                       int32_t batch, int32_t seq, int32_t heads, int32_t kv_heads,
                                       int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                       float scale, csm_stream_t stream) {
  CSM_REQUIRE(batch >= 0 && seq > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && head_dim > 0 &&
                  head_dim <= 128,
              CSM_ERR_SHAPE, "attn_fwd: bad shape B=%d S=%d H=%d KV=%d hd=%d", batch, seq, heads, kv_heads, head_dim);
  if (batch == 0) return CSM_OK;
  const int be = g_attn_backend.load();
  if ((be == 0 || be == 3) && attn_tc_supported(seq, head_dim, ldq, ldk, ldv, ldo, q, k, v, o))
    return attn_fwd_tc_launch(q, k, v, o, lse, batch, seq, heads, kv_heads, ldq, ldk, ldv, ldo, scale, nullptr,
                              as_stream(stream));
  CSM_REQUIRE(be != 3, CSM_ERR_SHAPE, "attn_fwd: shape not supported by the tcgen05 kernel (hd=64, seq>=128)");
  if ((be == 0 || be == 4) && attn_small_supported(seq, heads, kv_heads, head_dim, ldq, ldk, ldv, ldo, q, k, v, o))
    return attn_fwd_small_launch(q, k, v, o, lse, batch, seq, heads, kv_heads, head_dim, ldq, ldk, ldv, ldo, scale,
                                 as_stream(stream));
  CSM_REQUIRE(be != 4, CSM_ERR_SHAPE, "attn_fwd: shape not supported by the short-sequence kernel (seq<=32, hd 64/128)");
  if (be != 1 && attn_mma_supported(head_dim, ldq, ldk, ldv, ldo))
    return attn_fwd_mma_launch(q, k, v, o, lse, batch, seq, heads, kv_heads, head_dim, ldq, ldk, ldv, ldo, scale,
                               as_stream(stream));
  return attn_fwd_simt_launch(q, k, v, o, lse, batch, seq, heads, kv_heads, head_dim, ldq, ldk, ldv, ldo, scale,
                              as_stream(stream));
}

extern "C" int csm_attn_causal_gqa_bwd_rope(const void* q, const void* k, const void* v, const void* o,
                                            const float* lse, const void* dout, void* dq, void* dk, void* dv,
                                            int32_t batch, int32_t seq, int32_t heads, int32_t kv_heads,
                                            int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                            int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                                            const float* rope_cache, void* workspace, size_t workspace_bytes,
                                            csm_stream_t stream) {
  CSM_REQUIRE(batch > 0 && seq > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && rope_cache &&
                  aligned16(rope_cache),
              CSM_ERR_SHAPE, "attn_bwd_rope: bad arguments");
  CSM_REQUIRE(g_attn_backend.load() == 0 || g_attn_backend.load() == 3, CSM_ERR_SHAPE,
              "attn_bwd_rope: only the tcgen05 kernels fuse the inverse RoPE");
  CSM_REQUIRE(attn_tc_supported(seq, head_dim, ldq, ldk, ldv, ldo, q, k, v, o), CSM_ERR_SHAPE,
              "attn_bwd_rope: shape not supported by the tcgen05 kernels (use csm_attn_causal_gqa_bwd + csm_rope)");
  CSM_REQUIRE(workspace && workspace_bytes >= (size_t)batch * heads * seq * sizeof(float), CSM_ERR_SHAPE,
              "attn_bwd_rope: workspace too small");
  return attn_bwd_tc_launch(q, k, v, o, lse, dout, dq, dk, dv, reinterpret_cast<float*>(workspace), batch, seq, heads,
                            kv_heads, ldq, ldk, ldv, ldo, lddq, lddk, lddv, scale, rope_cache, nullptr, nullptr,
                            as_stream(stream));
}

// ---- sequence packing: several samples per row, block-diagonal causal attention
extern "C" int csm_attn_varlen_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t batch,
                                   int32_t seq, int32_t heads, int32_t kv_heads, int32_t head_dim, int64_t ldq,
                                   int64_t ldk, int64_t ldv, int64_t ldo, float scale, const int32_t* seg_start,
                                   csm_stream_t stream) {
  CSM_REQUIRE(batch > 0 && seq > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && seg_start, CSM_ERR_SHAPE,
              "attn_varlen_fwd: bad arguments");
  CSM_REQUIRE(attn_tc_supported(seq, head_dim, ldq, ldk, ldv, ldo, q, k, v, o), CSM_ERR_SHAPE,
              "attn_varlen_fwd: packed attention runs on the tcgen05 kernels only (head_dim 64, rows of >= 128 frames, "
              "16-byte aligned operands)");
  return attn_fwd_tc_launch(q, k, v, o, lse, batch, seq, heads, kv_heads, ldq, ldk, ldv, ldo, scale, seg_start,
                            as_stream(stream));
}

extern "C" int csm_attn_varlen_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse,
                                   const void* dout, void* dq, void* dk, void* dv, int32_t batch, int32_t seq,
                                   int32_t heads, int32_t kv_heads, int32_t head_dim, int64_t ldq, int64_t ldk,
                                   int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                                   const float* rope_cache, const int32_t* seg_start, const int32_t* seg_end,
                                   void* workspace, size_t workspace_bytes, csm_stream_t stream) {
  CSM_REQUIRE(batch > 0 && seq > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && seg_start && seg_end,
              CSM_ERR_SHAPE, "attn_varlen_bwd: bad arguments");
  CSM_REQUIRE(!rope_cache || aligned16(rope_cache), CSM_ERR_ALIGN, "attn_varlen_bwd: misaligned RoPE table");
  CSM_REQUIRE(attn_tc_supported(seq, head_dim, ldq, ldk, ldv, ldo, q, k, v, o), CSM_ERR_SHAPE,
              "attn_varlen_bwd: packed attention runs on the tcgen05 kernels only (head_dim 64, rows of >= 128 frames)");
  CSM_REQUIRE(workspace && workspace_bytes >= (size_t)batch * heads * seq * sizeof(float), CSM_ERR_SHAPE,
              "attn_varlen_bwd: workspace too small");
  return attn_bwd_tc_launch(q, k, v, o, lse, dout, dq, dk, dv, reinterpret_cast<float*>(workspace), batch, seq, heads,
                            kv_heads, ldq, ldk, ldv, ldo, lddq, lddk, lddv, scale, rope_cache, seg_start, seg_end,
                            as_stream(stream));
}

extern "C" size_t csm_attn_bwd_workspace_bytes(int32_t batch, int32_t seq, int32_t heads, int32_t kv_heads,
                                               int32_t head_dim) {
  (void)kv_heads; (void)head_dim;
  return (size_t)batch * heads * seq * sizeof(float) + 256;
}

extern "C" int csm_attn_causal_gqa_bwd(const void* q, const void* k, const void* v, const void* o,
                                       const float* lse, const void* dout, void* dq, void* dk, void* dv,
                                       int32_t batch, int32_t seq, int32_t heads, int32_t kv_heads,
                                       int32_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                       int64_t lddq, int64_t lddk, int64_t lddv, float scale, void* workspace,
                                       size_t workspace_bytes, csm_stream_t stream) {
  CSM_REQUIRE(batch >= 0 && seq > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && head_dim > 0 &&
                  head_dim <= 128,
              CSM_ERR_SHAPE, "attn_bwd: bad shape");
  if (batch == 0) return CSM_OK;
  CSM_REQUIRE(workspace && workspace_bytes >= csm_attn_bwd_workspace_bytes(batch, seq, heads, kv_heads, head_dim),
              CSM_ERR_SHAPE, "attn_bwd: workspace too small");
  float* delta = reinterpret_cast<float*>(workspace);
  const int be = g_attn_backend.load();
  if ((be == 0 || be == 3) && attn_tc_supported(seq, head_dim, ldq, ldk, ldv, ldo, q, k, v, o))
    return attn_bwd_tc_launch(q, k, v, o, lse, dout, dq, dk, dv, delta, batch, seq, heads, kv_heads, ldq, ldk, ldv, ldo,
                              lddq, lddk, lddv, scale, nullptr, nullptr, nullptr, as_stream(stream));
  if ((be == 0 || be == 4) && attn_small_supported(seq, heads, kv_heads, head_dim, ldq, ldk, ldv, ldo, q, k, v, o))
    return attn_bwd_small_launch(q, k, v, o, lse, dout, dq, dk, dv, batch, seq, heads, kv_heads, head_dim, ldq, ldk,
                                 ldv, ldo, lddq, lddk, lddv, scale, as_stream(stream));
  CSM_REQUIRE(be != 4, CSM_ERR_SHAPE, "attn_bwd: shape not supported by the short-sequence kernel");
  if (be != 1 && attn_mma_supported(head_dim, ldq, ldk, ldv, ldo))
    return attn_bwd_mma_launch(q, k, v, o, lse, dout, dq, dk, dv, delta, batch, seq, heads, kv_heads, head_dim, ldq,
                               ldk, ldv, ldo, lddq, lddk, lddv, scale, as_stream(stream));
  return attn_bwd_simt_launch(q, k, v, o, lse, dout, dq, dk, dv, delta, batch, seq, heads, kv_heads, head_dim, ldq,
                              ldk, ldv, ldo, lddq, lddk, lddv, scale, as_stream(stream));
}
