// Cross-entropy row kernels + the linear+CE drivers.
//   * small-shape path: scalar GEMM into an fp32 logits workspace, then row-wise CE (tiny test model only)
//   * tcgen05 path (ce_tc.cu): logits tiles live in TMEM; only per-tile (max, sumexp, target-logit)
//     partials are written, combined here by ce_combine_kernel.
// Reference: F.cross_entropy at src/csm/training/utils.py:102-105; per-codebook heads model.py:187.
#include "common.cuh"

namespace csm {

int gemm_dispatch(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                  int accumulate, float alpha, const void* A2, const void* B2, int64_t K2, int64_t lda2,
                  int64_t ldb2, int backend, cudaStream_t stream);
int linear_ce_tc_fwd(const void* H, const void* W, const int64_t* targets, float* loss_rows, float* lse, int64_t M,
                     int64_t V, int64_t K, int groups, int64_t ldh, int64_t hgs, int64_t ldw, int64_t wgs,
                     int transW, int64_t trs, int64_t tgs, void* ws, size_t ws_bytes, cudaStream_t st);
int linear_ce_tc_bwd_dlogits(const void* H, const void* W, const int64_t* targets, const float* lse,
                             float grad_scale, const float* grad_scale_dev, void* dlogits, int64_t ldd, int64_t M,
                             int64_t V, int64_t K, int groups, int64_t ldh, int64_t hgs, int64_t ldw, int64_t wgs,
                             int64_t trs, int64_t tgs, cudaStream_t st);
int gemm_tc_grouped(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int groups, int64_t lda,
                    int64_t ags, int64_t ldb, int64_t bgs, int64_t ldc, int64_t cgs, int transA, int transB,
                    int accumulate, cudaStream_t st);
bool linear_ce_tc_supported(int64_t M, int64_t V, int64_t K, int64_t ldh, int64_t ldw, int transW,
                            const void* H, const void* W);
size_t linear_ce_tc_workspace(int64_t M, int64_t V, int groups);

__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float t = (lane < nw) ? sm[lane] : (is_max ? -INFINITY : 0.f);
  return is_max ? warp_max(t) : warp_sum(t);
}

__global__ void __launch_bounds__(128)
ce_rows_fwd_kernel(const float* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ targets,
                   int64_t tgt_stride, float* __restrict__ loss, float* __restrict__ lse, int64_t V) {
  __shared__ float sm[32];
  const int64_t r = blockIdx.x;
  const float* row = logits + r * ldl;
  float mx = -INFINITY;
  for (int64_t c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = block_reduce(mx, sm, true);
  float s = 0.f;
  for (int64_t c = threadIdx.x; c < V; c += blockDim.x) s += __expf(row[c] - mx);
  s = block_reduce(s, sm, false);
  if (threadIdx.x == 0) {
    const float L = mx + logf(s);
    const int64_t t = targets[r * tgt_stride];
    lse[r] = L;
    loss[r] = (t < 0 || t >= V) ? 0.f : L - row[t];  // negative target == ignored row
  }
}

__global__ void __launch_bounds__(128)
ce_rows_bwd_kernel(const float* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ targets,
                   int64_t tgt_stride, const float* __restrict__ lse, float gscale,
                   const float* __restrict__ gscale_dev, bf16* __restrict__ dlogits, int64_t ldd, int64_t V) {
  const int64_t r = blockIdx.x;
  const float* row = logits + r * ldl;
  const float L = lse[r];
  const int64_t t = targets[r * tgt_stride];
  if (gscale_dev) gscale *= *gscale_dev;
  if (t < 0 || t >= V) gscale = 0.f;
  for (int64_t c = threadIdx.x; c < ldd; c += blockDim.x) {
    float g = 0.f;
    if (c < V) g = gscale * (__expf(row[c] - L) - (c == t ? 1.f : 0.f));
    dlogits[r * ldd + c] = __float2bfloat16_rn(g);
  }
}

// combine per-N-tile online-softmax partials: part[(g*M+m)*nt + t] = {max, sumexp, target_logit or -inf}
__global__ void __launch_bounds__(256)
ce_combine_kernel(const float4* __restrict__ part, int nt, float* __restrict__ loss, float* __restrict__ lse,
                  int64_t rows) {
  pdl_wait();
  pdl_trigger();
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float mx = -INFINITY;
  for (int t = 0; t < nt; ++t) mx = fmaxf(mx, part[r * nt + t].x);
  float s = 0.f, tl = -INFINITY;
  for (int t = 0; t < nt; ++t) {
    const float4 p = part[r * nt + t];
    s += p.y * __expf(p.x - mx);
    tl = fmaxf(tl, p.z);
  }
  const float L = mx + logf(s);
  lse[r] = L;
  loss[r] = (tl == -INFINITY) ? 0.f : L - tl;  // no tile saw the target: ignored row (target < 0)
}

int ce_combine_launch(const void* part, int nt, float* loss, float* lse, int64_t rows, cudaStream_t st) {
  // always with the PDL attribute: the combine is scheduled while the CE GEMM drains and starts the moment it completes
  if (launch_k(ce_combine_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), 0, st, -1, (const float4*)part, nt, loss,
               lse, rows) != cudaSuccess) { set_error("ce_combine: launch failed"); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("ce_combine");
  return CSM_OK;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace csm

using namespace csm;

extern "C" size_t csm_linear_ce_workspace_bytes(int64_t M, int64_t V, int64_t K, int32_t groups) {
  (void)K;
  const int64_t v8 = round_up(V, 8);
  size_t generic = (size_t)M * V * sizeof(float) + (size_t)M * v8 * sizeof(bf16) + 256;
  size_t tc = linear_ce_tc_workspace(M, V, groups) + (size_t)groups * M * v8 * sizeof(bf16) + 256;
  return generic > tc ? generic : tc;
}

extern "C" int csm_linear_ce_fwd(const void* H, const void* W, const int64_t* targets, float* loss_rows,
                                 float* lse, int64_t M, int64_t V, int64_t K, int32_t groups, int64_t ldh,
                                 int64_t h_group_stride, int64_t ldw, int64_t w_group_stride, int32_t transW,
                                 int64_t tgt_row_stride, int64_t tgt_group_stride, void* workspace,
                                 size_t workspace_bytes, int32_t backend, csm_stream_t stream) {
  CSM_REQUIRE(M >= 0 && V > 0 && K > 0 && groups > 0, CSM_ERR_SHAPE, "linear_ce_fwd: bad shape");
  if (M == 0) return CSM_OK;
  CSM_REQUIRE(workspace && workspace_bytes >= csm_linear_ce_workspace_bytes(M, V, K, groups), CSM_ERR_SHAPE,
              "linear_ce_fwd: workspace too small (%zu < %zu)", workspace_bytes,
              csm_linear_ce_workspace_bytes(M, V, K, groups));
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = linear_ce_tc_supported(M, V, K, ldh, ldw, transW, H, W);
  CSM_REQUIRE(backend != CSM_GEMM_TCGEN05 || tc_ok, CSM_ERR_SHAPE,
              "linear_ce_fwd: shape/alignment not supported by the tcgen05 path");
  if (tc_ok && backend != CSM_GEMM_SIMT)
    return linear_ce_tc_fwd(H, W, targets, loss_rows, lse, M, V, K, groups, ldh, h_group_stride, ldw,
                            w_group_stride, transW, tgt_row_stride, tgt_group_stride, workspace, workspace_bytes, st);
  float* logits = reinterpret_cast<float*>(workspace);
  for (int g = 0; g < groups; ++g) {
    const bf16* Hg = (const bf16*)H + g * h_group_stride;
    const bf16* Wg = (const bf16*)W + g * w_group_stride;
    int rc = gemm_dispatch(Hg, Wg, logits, nullptr, M, V, K, ldh, ldw, V, 0, 0, transW, CSM_DT_F32, 0, 1.f,
                           nullptr, nullptr, 0, 0, 0, CSM_GEMM_SIMT, st);
    if (rc) return rc;
    ce_rows_fwd_kernel<<<(unsigned)M, 128, 0, st>>>(logits, V, targets + g * tgt_group_stride, tgt_row_stride,
                                                    loss_rows + (int64_t)g * M, lse + (int64_t)g * M, V);
    CSM_CHECK_LAUNCH("ce_rows_fwd");
  }
  return CSM_OK;
}

extern "C" int csm_linear_ce_bwd(const void* H, const void* W, const int64_t* targets, const float* lse,
                                 float grad_scale, const float* grad_scale_dev, void* dH, void* dW, int32_t dw_accumulate, int64_t M,
                                 int64_t V, int64_t K, int32_t groups, int64_t ldh, int64_t h_group_stride,
                                 int64_t ldw, int64_t w_group_stride, int32_t transW, int64_t tgt_row_stride,
                                 int64_t tgt_group_stride, int64_t lddh, int64_t dh_group_stride,
                                 void* workspace, size_t workspace_bytes, int32_t backend, csm_stream_t stream) {
  CSM_REQUIRE(M >= 0 && V > 0 && K > 0 && groups > 0, CSM_ERR_SHAPE, "linear_ce_bwd: bad shape");
  if (M == 0) return CSM_OK;
  CSM_REQUIRE(workspace && workspace_bytes >= csm_linear_ce_workspace_bytes(M, V, K, groups), CSM_ERR_SHAPE,
              "linear_ce_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int64_t v8 = round_up(V, 8);
  const bool tc_ok = linear_ce_tc_supported(M, V, K, ldh, ldw, transW, H, W);
  CSM_REQUIRE(backend != CSM_GEMM_TCGEN05 || tc_ok, CSM_ERR_SHAPE,
              "linear_ce_bwd: shape/alignment not supported by the tcgen05 path");
  const bool use_tc = tc_ok && backend != CSM_GEMM_SIMT;
  // workspace: [fp32 logits (generic only)] [bf16 dlogits M x v8]
  char* wsp = reinterpret_cast<char*>(workspace);
  float* logits = reinterpret_cast<float*>(wsp);
  size_t off = use_tc ? 0 : ((size_t)M * V * sizeof(float) + 255) / 256 * 256;
  bf16* dlog = reinterpret_cast<bf16*>(wsp + off);
  const int gb = use_tc ? CSM_GEMM_AUTO : CSM_GEMM_SIMT;
  if (use_tc) {
    // tcgen05 path: every head in one launch per stage — dlogits [groups, M, v8], then grouped dH / dW GEMMs
    int rc = linear_ce_tc_bwd_dlogits(H, W, targets, lse, grad_scale, grad_scale_dev, dlog, v8, M, V, K, groups, ldh,
                                      h_group_stride, ldw, w_group_stride, tgt_row_stride, tgt_group_stride, st);
    if (rc) return rc;
    if (dH) {  // dH_g[M,K] = dlogits_g[M,V] * W_g   (W_g stored [V,K]: the [K_red, N_out] layout => transB)
      rc = gemm_tc_grouped(dlog, W, dH, M, K, V, groups, v8, M * v8, ldw, w_group_stride, lddh, dh_group_stride, 0, 1,
                           0, st);
      if (rc) return rc;
    }
    if (dW) {  // dW_g[V,K] = dlogits_g^T[V,M] * H_g[M,K]
      rc = gemm_tc_grouped(dlog, H, dW, V, K, M, groups, v8, M * v8, ldh, h_group_stride, ldw, w_group_stride, 1, 1,
                           dw_accumulate, st);
      if (rc) return rc;
    }
    return CSM_OK;
  }
  for (int g = 0; g < groups; ++g) {
    const bf16* Hg = (const bf16*)H + g * h_group_stride;
    const bf16* Wg = (const bf16*)W + g * w_group_stride;
    const int64_t* tg = targets + g * tgt_group_stride;
    const float* lg = lse + (int64_t)g * M;
    int rc = gemm_dispatch(Hg, Wg, logits, nullptr, M, V, K, ldh, ldw, V, 0, 0, transW, CSM_DT_F32, 0, 1.f, nullptr,
                           nullptr, 0, 0, 0, CSM_GEMM_SIMT, st);
    if (rc) return rc;
    ce_rows_bwd_kernel<<<(unsigned)M, 128, 0, st>>>(logits, V, tg, tgt_row_stride, lg, grad_scale, grad_scale_dev,
                                                    dlog, v8, V);
    CSM_CHECK_LAUNCH("ce_rows_bwd");
    if (dH) {
      // dH[M,K] = dlogits[M,V] * W ; W stored [V,K] (transW==0) is the [K_red=V, N_out=K] row-major layout
      rc = gemm_dispatch(dlog, Wg, (bf16*)dH + g * dh_group_stride, nullptr, M, K, V, v8, ldw, lddh, 0, 0,
                         transW ? 0 : 1, CSM_DT_BF16, 0, 1.f, nullptr, nullptr, 0, 0, 0, gb, st);
      if (rc) return rc;
    }
    if (dW) {
      bf16* dWg = (bf16*)dW + g * w_group_stride;
      if (!transW)  // dW[V,K] = dlogits^T[V,M] * H[M,K]
        rc = gemm_dispatch(dlog, Hg, dWg, nullptr, V, K, M, v8, ldh, ldw, 0, 1, 1, CSM_DT_BF16, dw_accumulate,
                           1.f, nullptr, nullptr, 0, 0, 0, gb, st);
      else  // dW[K,V] = H^T[K,M] * dlogits[M,V]
        rc = gemm_dispatch(Hg, dlog, dWg, nullptr, K, V, M, ldh, v8, ldw, 0, 1, 1, CSM_DT_BF16, dw_accumulate,
                           1.f, nullptr, nullptr, 0, 0, 0, gb, st);
      if (rc) return rc;
    }
  }
  return CSM_OK;
}
