// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16) for the operand shapes the attention kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --cudart=shared -o mma_rate mma_rate.cu && ./mma_rate
// One CTA per SM; one thread issues `iters` groups of `per` MMAs followed by one tcgen05.commit and (optionally) waits
// for that commit before the next group.  Reports cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../csm-train-pytorch_b200/csrc/tc_common.cuh"
using namespace csm::tc;

// mode: 0 = both K-major (A 128x16, B Nx16), 1 = B MN-major
__global__ void __launch_bounds__(128, 1) k(int N, int mode, int per, int iters, int wait_each, int ncommit, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[3];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, mode);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      for (int m = 0; m < per; ++m) {
        const int kk = m & 3;
        uint64_t ad = make_smem_desc(sa + kk * 32, 16, 1024);
        uint64_t bd = mode == 0 ? make_smem_desc(sb + kk * 32, 16, 1024) : make_smem_desc(sb + kk * 16 * 128, 64 * 128 * 2, 1024);
        umma_bf16(tm + ((m >> 2) & 1) * 256, ad, bd, idesc, m & 3 ? 1u : 0u);
      }
      for (int c = 0; c < ncommit; ++c) umma_commit(&bar[c ? 1 : 0]);
      if (wait_each) { mbar_wait(&bar[0], ph); ph ^= 1; }
    }
    umma_commit(&bar[2]);
    mbar_wait(&bar[2], 0);        // drain: everything issued above has completed
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}


// variant: the whole warp runs the loop (uniform control flow), only the tcgen05 instructions are elect-guarded
template <int N, int MODE, int PER>
__global__ void __launch_bounds__(128, 1) k2(int iters, int ncommit, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[3];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, MODE);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    const uint64_t ad0 = make_smem_desc(sa, 16, 1024);
    const uint64_t bd0 = MODE == 0 ? make_smem_desc(sb, 16, 1024) : make_smem_desc(sb, 64 * 128 * 2, 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < PER; ++m) {
          const int kk = m & 3;
          const uint64_t ad = ad0 + (uint64_t)((kk * 32) >> 4);
          const uint64_t bd = bd0 + (uint64_t)((MODE == 0 ? kk * 32 : kk * 16 * 128) >> 4);
          umma_bf16(tm + ((m >> 2) & 1) * 256, ad, bd, idesc, m & 3 ? 1u : 0u);
        }
        for (int c = 0; c < ncommit; ++c) umma_commit(&bar[c ? 1 : 0]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar[2]);
    __syncwarp();
    mbar_wait(&bar[2], 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 32) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int MODE, int PER>
void run2(const char* what, int nc, long long* d, int iters = 64) {
  cudaFuncSetAttribute(k2<N, MODE, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k2<N, MODE, PER><<<148, 128, 100 * 1024>>>(iters, nc, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  printf("[elect] %-52s %8.1f cycles/MMA  (%lld cycles, %d MMAs)\n", what, (double)cyc / (iters * PER), cyc, iters * PER);
}


// ---- k3: one iteration = the MMAs of one 64-column sub-block of the attention backward, issued back to back (no waits),
// next to 8 other warps that (cmode) 0 sleep at a barrier, 1 poll an mbarrier that never completes, 2 run the compute
// warps' TMEM chain (2 x tcgen05.ld.x32 -> wait -> 32 ex2 -> tcgen05.st.x16 -> wait), 3 stream 16 KB of st.shared per
// ~600 cycles (what the TMA producer writes per sub-block).
//   PAT 0: 8 SS N=64            (S = Q K^T, dP = dO V^T)
//   PAT 1: 8 SS N=64 + 4 TS     (+ dQ += dS K: A from TMEM, B MN-major)          = dq kernel
//   PAT 2: 8 SS N=64 + 8 TS     (+ dV += P^T dO, dK += dS^T Q)                   = dkdv kernel
//   PAT 3: 4 SS N=128           (S, dP for 128 columns at once: HALF an iteration's worth per 64 columns)
//   PAT 4: 8 TS N=64 (B K-major) + 4 TS (B MN-major): Q / dO resident in TMEM as the A operands
//   PAT 5: 4 TS (B MN-major) only
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int PAT>
__global__ void __launch_bounds__(320, 1) k3(int iters, int cmode, long long* out, float* sink, int wdepth = 0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[3];
  __shared__ uint64_t gbar[8];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); for (int i = 0; i < 8; ++i) mbar_init(&gbar[i], 1); mbar_fence_init(); done = 0; }
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const int wd = wdepth < 0 ? -wdepth : wdepth;
    constexpr uint32_t id_ss64 = make_idesc_bf16(128, 64, 0, 0), id_ss128 = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id_ts_mn = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384), sc = smem_u32(smem + 32768);
    const uint64_t ad0 = make_smem_desc(sa, 16, 1024), bd0 = make_smem_desc(sb, 16, 1024);
    const uint64_t cd0 = make_smem_desc(sc, 64 * 128 * 2, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t buf = (uint32_t)(it % 3) * 64;      // S at [0,192), dP at [192,384): three sub-block buffers each
      if (elect_one()) {
        if constexpr (PAT <= 2) {
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tm + p * 192 + buf, ad0 + (uint64_t)((kk * 32) >> 4), bd0 + (uint64_t)((kk * 32) >> 4), id_ss64, kk ? 1u : 0u);
        }
        if constexpr (PAT == 3) {
          if (it & 1) {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tm + p * 192, ad0 + (uint64_t)((kk * 32) >> 4), bd0 + (uint64_t)((kk * 32) >> 4), id_ss128, kk ? 1u : 0u);
          }
        }
        if constexpr (PAT == 4) {
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(tm + p * 192 + buf, tm + 448 + p * 32 + kk * 8, bd0 + (uint64_t)((kk * 32) >> 4), id_ss64, kk ? 1u : 0u);
        }
        if constexpr (PAT == 1 || PAT == 2 || PAT == 4 || PAT == 5) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tm + 384, tm + 192 + buf + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), cd0 + (uint64_t)((kk * 16 * 128) >> 4),
                         id_ts_mn, 1u);
        }
        if constexpr (PAT == 2) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tm + 448, tm + buf + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), cd0 + (uint64_t)((kk * 16 * 128) >> 4),
                         id_ts_mn, 1u);
        }
        umma_commit(wd ? &gbar[it & 7] : &bar[0]);
      }
      __syncwarp();
      if (wd && it >= wd) {          // wait for the group issued wd iterations ago (barrier (it - wd) & 7, its ((it - wd) >> 3)-th use)
        const int g = it - wd;
        if (wdepth > 0) mbar_wait(&gbar[g & 7], (g >> 3) & 1);
        else {
          uint32_t ok = 0;
          while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(&gbar[g & 7])), "r"((uint32_t)((g >> 3) & 1)) : "memory");
        }
      }
    }
    if (elect_one()) umma_commit(&bar[2]);
    __syncwarp();
    mbar_wait(&bar[2], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    done = 1;
  } else if (warp >= 2) {
    if (cmode == 1) {
      while (!done) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(&bar[1])), "r"(0u) : "memory");
      }
    } else if (cmode == 2) {
      const uint32_t la = tm + ((uint32_t)((warp & 3) * 32) << 16);
      const uint32_t c0 = (uint32_t)((warp - 2) >> 2) * 32;
      long long n = 0;
      while (!done) {
        uint32_t s_[32], d_[32];
        tmem_ld32(la + c0, s_);
        tmem_ld32(la + 192 + c0, d_);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = ex2f(fmaf(__uint_as_float(s_[2 * j]), 0.18f, -3.f)) * fmaf(__uint_as_float(d_[2 * j]), 0.125f, -1.f);
          const float b = ex2f(fmaf(__uint_as_float(s_[2 * j + 1]), 0.18f, -3.f)) * fmaf(__uint_as_float(d_[2 * j + 1]), 0.125f, -1.f);
          w[j] = packbf(a, b);
        }
        tmem_st16(la + 64 + c0, w);      // into a column range no MMA of this benchmark reads as A
        tmem_st_wait();
        ++n;
      }
      if (blockIdx.x == 0 && lane == 0 && warp == 2) out[1] = n;
    } else if (cmode == 3) {
      uint4* dst = reinterpret_cast<uint4*>(smem + 49152);        // 16 KB scratch behind the operand tiles
      while (!done) {
        const long long t = clock64();
        for (int i = threadIdx.x - 64; i < 1024; i += 256) dst[i] = make_uint4(i, i, i, i);
        while (clock64() - t < 600 && !done) {}
      }
      if (dst[lane].x == 0xdeadbeefu) sink[0] = 1.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int PAT>
void run3(const char* what, long long* d, float* sink) {
  const int iters = 2048;
  cudaFuncSetAttribute(k3<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* cm[] = {"other warps asleep", "8 warps polling an mbarrier", "8 warps on the TMEM ld/ex2/st chain", "8 warps streaming 16 KB st.shared / 600 cyc"};
  for (int cmode = 0; cmode < 4; ++cmode) {
    for (int rep = 0; rep < 2; ++rep) {
      k3<PAT><<<148, 320, 100 * 1024>>>(iters, cmode, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
    printf("[k3] %-46s | %-44s %7.1f cycles per sub-block", what, cm[cmode], (double)r[0] / iters * (PAT == 3 ? 1.0 : 1.0));
    if (cmode == 2) printf("   (chain: %.1f cycles per iteration)", (double)r[0] / (double)(r[1] ? r[1] : 1));
    printf("\n");
  }
}

// ---- k4: what each piece of the real MMA-warp loop costs.  One iteration = 8 SS N=64 + 4 TS (the dQ kernel's sub-block);
// flags: 1 = an (immediately successful) mbarrier wait before each of the two MMA blocks, 2 = tcgen05.fence::after_thread_sync
// after each wait, 4 = the two blocks under separate elect_one() + __syncwarp(), 8 = a second commit per iteration,
// 16 = the whole loop runs in ONE elected thread (no per-iteration elect / syncwarp)
template <int FLAGS>
__global__ void __launch_bounds__(320, 1) k4(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t id_ss64 = make_idesc_bf16(128, 64, 0, 0), id_ts_mn = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384), sc = smem_u32(smem + 32768);
    const uint64_t ad0 = make_smem_desc(sa, 16, 1024), bd0 = make_smem_desc(sb, 16, 1024);
    const uint64_t cd0 = make_smem_desc(sc, 64 * 128 * 2, 1024);
    auto ss = [&](uint32_t buf) {
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tm + p * 192 + buf, ad0 + (uint64_t)((kk * 32) >> 4), bd0 + (uint64_t)((kk * 32) >> 4), id_ss64, kk ? 1u : 0u);
    };
    auto ts = [&](uint32_t buf) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16_ts(tm + 384, tm + 192 + buf + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), cd0 + (uint64_t)((kk * 16 * 128) >> 4), id_ts_mn, 1u);
    };
    const long long t0 = clock64();
    if constexpr (FLAGS & 16) {
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
          const uint32_t buf = (uint32_t)(it % 3) * 64;
          if constexpr (FLAGS & 1) mbar_wait(&bar[1], 1);
          if constexpr (FLAGS & 2) tc_fence_after();
          ss(buf);
          umma_commit(&bar[0]);
          if constexpr (FLAGS & 1) mbar_wait(&bar[2], 1);
          if constexpr (FLAGS & 2) tc_fence_after();
          ts(buf);
          if constexpr (FLAGS & 8) umma_commit(&bar[3]);
        }
      }
      __syncwarp();
    } else {
      for (int it = 0; it < iters; ++it) {
        const uint32_t buf = (uint32_t)(it % 3) * 64;
        if constexpr (FLAGS & 1) mbar_wait(&bar[1], 1);
        if constexpr (FLAGS & 2) tc_fence_after();
        if constexpr (FLAGS & 4) {
          if (elect_one()) { ss(buf); umma_commit(&bar[0]); }
          __syncwarp();
          if constexpr (FLAGS & 1) mbar_wait(&bar[2], 1);
          if constexpr (FLAGS & 2) tc_fence_after();
          if (elect_one()) { ts(buf); if constexpr (FLAGS & 8) umma_commit(&bar[3]); }
          __syncwarp();
        } else {
          if (elect_one()) { ss(buf); umma_commit(&bar[0]); ts(buf); if constexpr (FLAGS & 8) umma_commit(&bar[3]); }
          __syncwarp();
        }
      }
    }
    if (elect_one()) umma_commit(&bar[2]);
    __syncwarp();
    if constexpr (FLAGS & 1) { /* bar[2] is used as an always-open gate above: drain through a spin on the clock instead */
      const long long t = clock64(); while (clock64() - t < 4000) {}
    } else {
      mbar_wait(&bar[2], 0);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0 - ((FLAGS & 1) ? 4000 : 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int FLAGS>
void run4(const char* what, long long* d) {
  const int iters = 2048;
  cudaFuncSetAttribute(k4<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k4<FLAGS><<<148, 320, 100 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long r; cudaMemcpy(&r, d, 8, cudaMemcpyDeviceToHost);
  printf("[k4] %-86s %7.1f cycles per sub-block\n", what, (double)r / iters);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  float* sink; cudaMalloc(&sink, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct { int N, mode, per, wait, nc; const char* what; } cfg[] = {
    {64, 0, 4, 0, 0, "N=64, 4 MMA/iter, NO commit in loop"},
    {64, 0, 4, 0, 1, "N=64, 4/commit"},
    {64, 0, 16, 0, 1, "N=64, 16/commit"},
    {64, 0, 32, 0, 1, "N=64, 32/commit"},
    {256, 0, 4, 0, 0, "N=256, 4 MMA/iter, NO commit in loop"},
    {256, 0, 4, 0, 1, "N=256, 4/commit"},
    {256, 0, 8, 0, 1, "N=256, 8/commit"},
    {256, 0, 16, 0, 1, "N=256, 16/commit"},
    {128, 0, 4, 0, 0, "N=128, 4 MMA/iter, NO commit in loop"},
  };
  for (auto& c : cfg) {
    const int iters = 64;   // keep <= a few parities: no-wait mode re-waits every phase in order
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 128, 100 * 1024>>>(c.N, c.mode, c.per, iters, c.wait, c.nc, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %8.1f cycles/MMA  (%lld cycles, %d MMAs)\n", c.what, (double)cyc / (iters * c.per), cyc, iters * c.per);
  }
  run2<64, 0, 4>("N=64 K-major 4/commit", 1, d);
  run2<64, 1, 4>("N=64 MN-major 4/commit", 1, d);
  run2<64, 0, 8>("N=64 K-major 8/commit", 1, d);
  run2<64, 0, 4>("N=64 K-major 4 + 3 commits", 3, d);
  run2<128, 0, 4>("N=128 4/commit", 1, d);
  run2<256, 0, 4>("N=256 4/commit", 1, d);
  run2<256, 0, 4>("N=256 4 + 2 commits", 2, d);
  // sustained: the same loops for 8x and 64x as many MMAs
  for (int it : {512, 4096}) {
    printf("-- %d iterations\n", it);
    run2<64, 0, 8>("N=64 K-major 8/commit", 1, d, it);
    run2<64, 1, 8>("N=64 MN-major 8/commit", 1, d, it);
    run2<128, 0, 8>("N=128 8/commit", 1, d, it);
    run2<256, 0, 8>("N=256 8/commit", 1, d, it);
  }
  {
    cudaFuncSetAttribute(k3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int wdpt : {0, 1, 2, 3, 4, -1, -2, -3}) {
      for (int rep = 0; rep < 2; ++rep) { k3<1><<<148, 320, 100 * 1024>>>(2048, 0, d, sink, wdpt); cudaDeviceSynchronize(); }
      long long r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
      printf("[k3w] 8 SS + 4 TS per group, MMA warp waits for the commit of the group issued %d groups earlier (%s): %7.1f cycles per group\n",
             wdpt < 0 ? -wdpt : wdpt, wdpt == 0 ? "no wait" : wdpt > 0 ? "try_wait" : "test_wait spin", (double)r[0] / 2048);
    }
  }
  run4<0>("8 SS + 4 TS, one elect block, one commit", d);
  run4<8>("+ second commit", d);
  run4<4>("two elect blocks", d);
  run4<4 | 1>("two elect blocks, open mbarrier wait before each", d);
  run4<4 | 1 | 2>("two elect blocks, wait + tcgen05.fence::after before each", d);
  run4<4 | 1 | 2 | 8>("two elect blocks, wait + fence before each, second commit (= the kernel's loop)", d);
  run4<16>("whole loop in ONE elected thread", d);
  run4<16 | 1>("whole loop in one elected thread, open wait before each block", d);
  run4<16 | 1 | 2 | 8>("whole loop in one elected thread, wait + fence before each block, second commit", d);
  run3<0>("8 SS N=64", d, sink);
  run3<1>("8 SS N=64 + 4 TS (dq kernel)", d, sink);
  run3<2>("8 SS N=64 + 8 TS (dkdv kernel)", d, sink);
  run3<3>("4 SS N=128 per 64 columns (8 per 2 sub-blocks)", d, sink);
  run3<4>("8 TS K-major B + 4 TS (Q/dO resident in TMEM)", d, sink);
  run3<5>("4 TS only", d, sink);
  return 0;
}
