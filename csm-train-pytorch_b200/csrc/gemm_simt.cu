// Scalar-FMA small-shape GEMM: any M/N/K, any operand major-ness, fp32 accumulate.
// This is the path for shapes below the tcgen05 tile minima (the tiny test model: D=32/16, hd=8) and
// for unaligned leading dimensions; the CSM-1B shapes run gemm_tc.cu.  Same semantics as csm_gemm_bf16.
#include "common.cuh"

namespace csm {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

struct SimtGemmArgs {
  const bf16* A; const bf16* B; void* C; const bf16* R;
  int64_t M, N, K, sAm, sAk, sBn, sBk, ldc, ldr;
  const bf16* A2; const bf16* B2; int64_t K2, lda2, ldb2;
  int c_f32, r_f32, accumulate, transA, transB;
  float alpha;
};

__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(SimtGemmArgs a) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t Ktot = a.K + a.K2;
  for (int64_t k0 = 0; k0 < Ktot; k0 += SG_BK) {
    // ---- stage A tile [BK][BM]
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int mm, kk;
      if (!a.transA) { kk = tid % 16; mm = tid / 16 + 16 * p; } else { mm = tid % 64; kk = tid / 64 + 4 * p; }
      const int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < a.M && k < Ktot) {
        v = (k < a.K) ? __bfloat162float(a.A[m * a.sAm + k * a.sAk])
                      : __bfloat162float(a.transA ? a.A2[(k - a.K) * a.lda2 + m] : a.A2[m * a.lda2 + (k - a.K)]);
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int nn, kk;
      if (!a.transB) { kk = tid % 16; nn = tid / 16 + 16 * p; } else { nn = tid % 64; kk = tid / 64 + 4 * p; }
      const int64_t n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < a.N && k < Ktot) {
        v = (k < a.K) ? __bfloat162float(a.B[n * a.sBn + k * a.sBk])
                      : __bfloat162float(a.transB ? a.B2[(k - a.K) * a.ldb2 + n] : a.B2[n * a.ldb2 + (k - a.K)]);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = a.alpha * acc[i][j];
      if (a.R) v += a.r_f32 ? reinterpret_cast<const float*>(a.R)[m * a.ldr + n] : __bfloat162float(a.R[m * a.ldr + n]);
      if (a.c_f32) {
        float* c = reinterpret_cast<float*>(a.C) + m * a.ldc + n;
        *c = a.accumulate ? (*c + v) : v;
      } else {
        bf16* c = reinterpret_cast<bf16*>(a.C) + m * a.ldc + n;
        if (a.accumulate) v += __bfloat162float(*c);
        *c = __float2bfloat16_rn(v);
      }
    }
  }
}

int gemm_simt_launch(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                     int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                     int accumulate, float alpha, const void* A2, const void* B2, int64_t K2, int64_t lda2,
                     int64_t ldb2, cudaStream_t stream) {
  SimtGemmArgs a;
  a.A = (const bf16*)A; a.B = (const bf16*)B; a.C = C; a.R = (const bf16*)R;
  a.M = M; a.N = N; a.K = K;
  a.sAm = transA ? 1 : lda; a.sAk = transA ? lda : 1;
  a.sBn = transB ? 1 : ldb; a.sBk = transB ? ldb : 1;
  a.ldc = ldc; a.ldr = ldr;
  a.A2 = (const bf16*)A2; a.B2 = (const bf16*)B2; a.K2 = (A2 && B2) ? K2 : 0; a.lda2 = lda2; a.ldb2 = ldb2;
  a.c_f32 = (c_dtype & CSM_DT_F32) != 0; a.r_f32 = (c_dtype & CSM_DT_RES_F32) != 0; a.accumulate = accumulate; a.transA = transA; a.transB = transB;
  a.alpha = alpha;
  dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM));
  gemm_simt_kernel<<<grid, SG_THREADS, 0, stream>>>(a);
  CSM_CHECK_LAUNCH("gemm_simt");
  return CSM_OK;
}

}  // namespace csm
