// Standalone C-ABI check (no Python, no torch): links libcsm_b200.so and
//   1. checks the fused audio-head CE (31 heads, 232 rows, V = 2051, K = 1024) against a host computation of a few rows,
//   2. checks that csm_set_gemm_narrow_tail_mode(1) leaves the fused-CE forward, its backward (dH) and a plain
//      ragged-N GEMM bit-identical, in the single-CTA and the CTA-pair kernels,
//   3. times the CE forward in both modes (CUDA events, 256 MB L2 flush between iterations, median of 15).
// Build + run: tools/ctest/run.sh   (needs one B200; a few seconds)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/csm_b200.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)
#define CSM(x)                                                                             \
  do {                                                                                     \
    int r_ = (x);                                                                          \
    if (r_ != 0) {                                                                         \
      printf("csm error %d (%s) at %s:%d\n", r_, csm_last_error(), __FILE__, __LINE__);    \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    p[i] = __float2bfloat16(((h & 0xffff) / 32768.0f - 1.0f) * scale);
  }
}
__global__ void fill_targets(int64_t* t, size_t n, int V) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    t[i] = (int64_t)(((uint32_t)i * 2246822519u >> 8) % (uint32_t)V);
}

static float bf(const __nv_bfloat16& x) { return __bfloat162float(x); }

int main() {
  if (csm_device_supported() != 1) { printf("no sm_100 device\n"); return 1; }
  const int G = 31, M = 232, MAXM = 512, V = 2051, K = 1024, LDC = 2056;
  __nv_bfloat16 *H, *W, *dH[2], *C[2];
  int64_t* T;
  float *loss[2], *lse[2];
  void* ws;
  uint8_t* flush;
  const size_t nH = (size_t)G * MAXM * K, nW = (size_t)G * V * K;
  CK(cudaMalloc(&H, nH * 2)); CK(cudaMalloc(&W, nW * 2)); CK(cudaMalloc(&T, (size_t)G * MAXM * 8));
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&dH[i], nH * 2)); CK(cudaMalloc(&C[i], (size_t)M * LDC * 2));
    CK(cudaMalloc(&loss[i], (size_t)G * MAXM * 4)); CK(cudaMalloc(&lse[i], (size_t)G * MAXM * 4));
  }
  const size_t wsb = csm_linear_ce_workspace_bytes(MAXM, V, K, G);
  CK(cudaMalloc(&ws, wsb));
  const size_t fl = 256u << 20;
  CK(cudaMalloc(&flush, fl));
  fill_bf16<<<1024, 256>>>(H, nH, 17u, 1.0f);
  fill_bf16<<<4096, 256>>>(W, nW, 99u, 0.0625f);
  fill_targets<<<64, 256>>>(T, (size_t)G * MAXM, V);
  CK(cudaDeviceSynchronize());

  auto ce_fwd = [&](int i) {
    return csm_linear_ce_fwd(H, W, T, loss[i], lse[i], M, V, K, G, K, (int64_t)M * K, K, (int64_t)V * K, 0, 1, M, ws, wsb,
                             2, nullptr);
  };
  auto ce_bwd = [&](int i) {
    return csm_linear_ce_bwd(H, W, T, lse[0], 1.0f / (G * M), nullptr, dH[i], nullptr, 0, M, V, K, G, K, (int64_t)M * K, K,
                             (int64_t)V * K, 0, 1, M, K, (int64_t)M * K, ws, wsb, 2, nullptr);
  };
  auto gemm = [&](int i) {
    CK(cudaMemset(C[i], 0, (size_t)M * LDC * 2));
    return csm_gemm_bf16(H, W, C[i], nullptr, M, V, K, K, K, LDC, 0, 0, 0, CSM_DT_BF16, 0, 1.0f, nullptr, nullptr, 0, 0, 0,
                         2, nullptr);
  };

  int bad = 0;
  // ---- 1. default mode against the host, rows 0 and 231 of heads 0 and 30
  csm_set_gemm_narrow_tail_mode(0);
  CSM(ce_fwd(0));
  CK(cudaDeviceSynchronize());
  {
    std::vector<float> hl((size_t)G * M);
    CK(cudaMemcpy(hl.data(), loss[0], hl.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<__nv_bfloat16> hrow(K), wmat((size_t)V * K);
    for (int g : {0, 30}) {
      CK(cudaMemcpy(wmat.data(), W + (size_t)g * V * K, wmat.size() * 2, cudaMemcpyDeviceToHost));
      for (int m : {0, 231}) {
        CK(cudaMemcpy(hrow.data(), H + ((size_t)g * M + m) * K, K * 2, cudaMemcpyDeviceToHost));
        int64_t t;
        CK(cudaMemcpy(&t, T + (size_t)g * M + m, 8, cudaMemcpyDeviceToHost));
        double mx = -1e30, tl = 0;
        std::vector<double> lg(V);
        for (int v = 0; v < V; ++v) {
          double s = 0;
          for (int k = 0; k < K; ++k) s += (double)bf(hrow[k]) * (double)bf(wmat[(size_t)v * K + k]);
          lg[v] = s; mx = std::max(mx, s);
          if (v == t) tl = s;
        }
        double se = 0;
        for (int v = 0; v < V; ++v) se += std::exp(lg[v] - mx);
        const double ref = mx + std::log(se) - tl, got = hl[(size_t)g * M + m];
        const bool ok = std::fabs(ref - got) <= 2e-3 * std::max(1.0, std::fabs(ref));
        printf("host check head %d row %d: ref %.6f got %.6f %s\n", g, m, ref, got, ok ? "ok" : "MISMATCH");
        bad += !ok;
      }
    }
  }
  // ---- 2. narrow-tail mode bit-identical (pair mode auto / never / forced)
  for (int pair : {-1, 0, 1}) {
    csm_set_gemm_cta_pair_mode(pair);
    for (int mode = 0; mode < 2; ++mode) {
      csm_set_gemm_narrow_tail_mode(mode);
      CK(cudaMemset(loss[mode], 0xff, (size_t)G * M * 4)); CK(cudaMemset(lse[mode], 0xff, (size_t)G * M * 4));
      CK(cudaMemset(dH[mode], 0xff, nH * 2));
      CSM(ce_fwd(mode));
      CSM(ce_bwd(mode));
      CSM(gemm(mode));
    }
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> a, b;
    auto same = [&](const void* p0, const void* p1, size_t n) {
      a.resize(n); b.resize(n);
      cudaMemcpy(a.data(), p0, n, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), p1, n, cudaMemcpyDeviceToHost);
      return std::memcmp(a.data(), b.data(), n) == 0;
    };
    const bool s1 = same(loss[0], loss[1], (size_t)G * M * 4), s2 = same(lse[0], lse[1], (size_t)G * M * 4);
    const bool s3 = same(dH[0], dH[1], nH * 2), s4 = same(C[0], C[1], (size_t)M * LDC * 2);
    printf("pair mode %2d: narrow tail vs full tile  loss %s  lse %s  dH %s  gemm %s\n", pair, s1 ? "same" : "DIFF",
           s2 ? "same" : "DIFF", s3 ? "same" : "DIFF", s4 ? "same" : "DIFF");
    bad += !(s1 && s2 && s3 && s4);
  }
  csm_set_gemm_cta_pair_mode(-1);
  // ---- 3. timing of the CE forward and the dlogits + dH backward, both modes
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int which = 0; which < 2; ++which) {
    for (int mode = 0; mode < 2; ++mode) {
      csm_set_gemm_narrow_tail_mode(mode);
      std::vector<float> ts;
      for (int it = 0; it < 18; ++it) {
        CK(cudaMemsetAsync(flush, it, fl));
        CK(cudaEventRecord(e0));
        if (which == 0) CSM(ce_fwd(mode)); else CSM(ce_bwd(mode));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 3) ts.push_back(ms);
      }
      std::sort(ts.begin(), ts.end());
      printf("%s narrow_tail=%d: median %.1f us, min %.1f us (15 timed, L2 flushed)\n",
             which == 0 ? "ce_fwd (partials + combine)" : "ce_bwd (dlogits + dH, no dW)", mode, ts[ts.size() / 2] * 1e3,
             ts[0] * 1e3);
    }
  }
  csm_set_gemm_narrow_tail_mode(0);
  // ---- 4. where the ridge is: CE forward time against the number of rows per head (weights streamed: 130.2 MB)
  for (int pair : {-1, 0}) {
    csm_set_gemm_cta_pair_mode(pair);
    for (int rows : {32, 64, 128, 192, 232, 256, 384, 512}) {
      std::vector<float> ts;
      for (int it = 0; it < 13; ++it) {
        CK(cudaMemsetAsync(flush, it, fl));
        CK(cudaEventRecord(e0));
        CSM(csm_linear_ce_fwd(H, W, T, loss[1], lse[1], rows, V, K, G, K, (int64_t)rows * K, K, (int64_t)V * K, 0, 1, rows,
                              ws, wsb, 2, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 3) ts.push_back(ms);
      }
      std::sort(ts.begin(), ts.end());
      const double us = ts[ts.size() / 2] * 1e3;
      printf("ce_fwd rows/head %3d pair mode %2d: %.1f us  (%.0f GB/s of weights, %.0f TFLOP/s useful)\n", rows, pair, us,
             130.2e6 / (us * 1e-6) / 1e9, 2.0 * G * rows * (double)V * K / (us * 1e-6) / 1e12);
    }
  }
  csm_set_gemm_cta_pair_mode(-1);
  printf(bad ? "RESULT: FAIL (%d)\n" : "RESULT: PASS\n", bad);
  return bad ? 4 : 0;
}
