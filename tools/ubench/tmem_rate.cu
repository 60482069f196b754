// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing them, and the
// cost of the softmax-style chain  ld -> wait -> (ex2 work) -> st -> wait  the attention kernels' compute warps run.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --cudart=shared -o tmem_rate tmem_rate.cu && ./tmem_rate
// One CTA per SM (grid = SM count), `warps` warps; warp w owns TMEM lanes (w % 4) * 32 .. +31 and columns
// ((w / 4) * 32 * k) % 512.  Reports bytes per cycle and SM (max over the CTA's warps of the clock64 span).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../csm-train-pytorch_b200/csrc/tc_common.cuh"
using namespace csm::tc;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// mode 0: x32 loads, wait after every `burst` loads; mode 1: x16 loads; mode 2: x16 stores;
// mode 3: chain  2 x ld32 -> wait -> 32 ex2 + fma -> st16 -> wait  (one dq-kernel sub-block chunk per iteration)
// mode 4: same chain with 16-column chunks (2 x ld16, 16 ex2, 8-cell store emulated by st16 of half the data)
// mma_duty > 0: one extra warp (the last) issues, for as long as the other warps run, groups of tcgen05 MMAs shaped like one
// sub-block of the attention backward (2 x [128x64x64, both operands in smem] + mma_ts x [128x64x64, A from TMEM]) into
// columns the chain does not touch, waits for the group's commit and then idles so that the tensor pipe is busy roughly
// mma_duty percent of the time.
__global__ void __launch_bounds__(1024, 1) k(int mode, int iters, int burst, int mma_duty, int mma_ts, long long* out, float* sink) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar[2];
  __shared__ volatile int done;
  __shared__ uint32_t slot;
  __shared__ long long span[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = (blockDim.x >> 5) - (mma_duty > 0 ? 1 : 0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  if (mma_duty > 0) {
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); done = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t col0 = (uint32_t)((warp >> 2) * 64) % 448;
  uint32_t acc = 0;
  float facc = 0.f;
  __syncthreads();
  if (warp == nw) {   // the MMA warp (only exists when mma_duty > 0).  Warp-uniform control flow + elect_one: tcgen05
                      // instructions under `if (lane == 0)` get a waterfall loop each (profiles/r1_summary.md 2a)
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384), sc = smem_u32(smem + 32768);
    constexpr uint32_t idesc_nt = make_idesc_bf16(128, 64, 0, 0), idesc_ts = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t ad0 = make_smem_desc(sa, 16, 1024), bd0 = make_smem_desc(sb, 16, 1024);
    const uint64_t cd0 = make_smem_desc(sc, 64 * 128 * 2, 1024);
    long long groups = 0;
    const uint32_t tm0 = slot;   // keep the TMEM base in a register: the asm memory clobbers would reload it per MMA
    const long long g00 = clock64();
    uint32_t ph0 = 0, ph1 = 0;
    while (!done) {
      const long long g0 = clock64();
      const int bsel = (int)(groups & 1);
      if (elect_one()) {
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tm0 + 256 + (uint32_t)(p * 64), ad0 + (uint64_t)((kk * 32) >> 4), bd0 + (uint64_t)((kk * 32) >> 4), idesc_nt,
                      kk ? 1u : 0u);
        if (mma_ts >= 1) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tm0 + 384, tm0 + 256 + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), cd0 + (uint64_t)((kk * 16 * 128) >> 4),
                         idesc_ts, 1u);
        }
        if (mma_ts >= 2) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tm0 + 448, tm0 + 320 + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), cd0 + (uint64_t)((kk * 16 * 128) >> 4),
                         idesc_ts, 1u);
        }
        umma_commit(&bar[bsel]);
      }
      __syncwarp();
      ++groups;
      if (mma_duty == 100) {      // pipelined: wait for group g-1 after issuing group g (throughput)
        if (groups >= 2) {
          if (bsel) { mbar_wait(&bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&bar[1], ph1); ph1 ^= 1; }
        }
      } else {                    // wait for every group's own commit (latency), then idle to the requested duty cycle
        if (bsel) { mbar_wait(&bar[1], ph1); ph1 ^= 1; } else { mbar_wait(&bar[0], ph0); ph0 ^= 1; }
        const long long g1 = clock64();
        const long long idle = (g1 - g0) * (100 - mma_duty) / mma_duty;
        while (clock64() - g1 < idle && !done) {}
      }
    }
    const long long busy = clock64() - g00;
    if (blockIdx.x == 0 && lane == 0) { out[1] = busy; out[2] = groups; }
    __syncwarp();
  }
  const long long t0 = clock64();
  if (warp == nw) {
  } else if (mode == 7) {
    while (clock64() - t0 < 2000000) __nanosleep(2000);
  } else if (mode == 0) {
    uint32_t r[32];
    for (int it = 0; it < iters; it += burst) {
      for (int b = 0; b < burst; ++b) {
        tmem_ld32(tm + col0 + (uint32_t)((b & 1) * 32), r);
        if (b == burst - 1) tmem_ld_wait();
        acc ^= r[b & 31];
      }
    }
  } else if (mode == 1) {
    uint32_t r[16];
    for (int it = 0; it < iters; it += burst) {
      for (int b = 0; b < burst; ++b) {
        ld16(tm + col0 + (uint32_t)((b & 3) * 16), r);
        if (b == burst - 1) tmem_ld_wait();
        acc ^= r[b & 15];
      }
    }
  } else if (mode == 2) {
    uint32_t r[16];
    for (int i = 0; i < 16; ++i) r[i] = lane + i;
    for (int it = 0; it < iters; it += burst) {
      for (int b = 0; b < burst; ++b) {
        tmem_st16(tm + col0 + (uint32_t)((b & 3) * 16), r);
        if (b == burst - 1) tmem_st_wait();
      }
    }
  } else if (mode == 3) {
    for (int it = 0; it < iters; ++it) {
      uint32_t s[32], d[32];
      tmem_ld32(tm + col0, s);
      tmem_ld32(tm + col0 + 32, d);
      tmem_ld_wait();
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = ex2(fmaf(__uint_as_float(s[2 * j]), 0.18f, -3.f)) * fmaf(__uint_as_float(d[2 * j]), 0.125f, -1.f);
        const float b = ex2(fmaf(__uint_as_float(s[2 * j + 1]), 0.18f, -3.f)) * fmaf(__uint_as_float(d[2 * j + 1]), 0.125f, -1.f);
        w[j] = pack_bf16(a, b);
      }
      tmem_st16(tm + col0 + 32, w);
      tmem_st_wait();
    }
  } else {
    for (int it = 0; it < iters; ++it) {
      uint32_t s[16], d[16];
      ld16(tm + col0, s);
      ld16(tm + col0 + 16, d);
      tmem_ld_wait();
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = ex2(fmaf(__uint_as_float(s[2 * j]), 0.18f, -3.f)) * fmaf(__uint_as_float(d[2 * j]), 0.125f, -1.f);
        const float b = ex2(fmaf(__uint_as_float(s[2 * j + 1]), 0.18f, -3.f)) * fmaf(__uint_as_float(d[2 * j + 1]), 0.125f, -1.f);
        w[j] = pack_bf16(a, b);
        w[j + 8] = w[j];
      }
      tmem_st16(tm + col0 + 16, w);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp < nw) span[warp] = t1 - t0;
  if (acc == 0x12345u || facc == 1.f) sink[threadIdx.x] = (float)acc;
  if (mma_duty > 0 && warp < nw) {
    named_bar_sync(1, nw * 32);
    if (threadIdx.x == 0) done = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long m = 0;
    for (int i = 0; i < nw; ++i) m = span[i] > m ? span[i] : m;
    out[0] = m;
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, 4096);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096;
  printf("# SMs %d, %d iterations per warp; B/clk/SM of TMEM traffic (mode 3/4: cycles per iteration and elements/clk/SM)\n", sms, iters);
  const char* names[] = {"ld 32x32b.x32", "ld 32x32b.x16", "st 32x32b.x16", "chain 32 cols", "chain 16 cols"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int burst : {1, 2, 4}) {
      if (mode >= 3 && burst != 1) continue;
      for (int warps : {1, 4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) k<<<sms, warps * 32, 0>>>(mode, iters, burst, 0, 0, out, sink);
        long long cyc = 0;
        cudaError_t e = cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        const double bytes_per_it = mode == 0 ? 4096 : (mode == 1 || mode == 2) ? 2048 : mode == 3 ? 8192 + 2048 : 4096 + 2048;
        if (mode < 3)
          printf("%-14s burst %d warps %2d: %9lld cycles  %7.1f B/clk/SM\n", names[mode], burst, warps, cyc,
                 bytes_per_it * iters * warps / (double)cyc);
        else
          printf("%-14s warps %2d: %7.1f cycles/iteration  %6.2f elements/clk/SM  (%7.1f B/clk/SM of TMEM traffic)\n",
                 names[mode], warps, (double)cyc / iters, (mode == 3 ? 1024.0 : 512.0) * iters * warps / (double)cyc,
                 bytes_per_it * iters * warps / (double)cyc);
      }
    }
  }
  // the same chains next to a tensor pipe that is busy mma_duty % of the time
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  const char* n2[] = {"", "", "", "chain 32 cols", "chain 16 cols", "", "", "idle warps"};
  for (int mode : {7, 3, 4})
    for (int ts : {0, 1, 2})
      for (int duty : {50, 100})
        for (int warps : {8, 16}) {
          if (mode == 7 && warps == 16) continue;
          for (int rep = 0; rep < 2; ++rep) k<<<sms, (warps + 1) * 32, 50 * 1024>>>(mode, iters, 1, duty, ts, out, sink);
          long long r[3] = {0, 0, 0};
          cudaError_t e = cudaMemcpy(r, out, 24, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
          printf("%-14s warps %2d + MMA warp, group = 2 SS + %d TS products of 128x64x64, %s: %7.1f cycles/iteration  %6.2f elements/clk/SM;"
                 "  %7.1f cycles per MMA group (%lld groups)\n",
                 n2[mode], warps, ts, duty == 100 ? "pipelined (throughput)" : "wait each, 50% duty (latency x2)",
                 (double)r[0] / iters, (mode == 3 ? 1024.0 : 512.0) * iters * warps / (double)r[0],
                 r[2] ? (double)r[1] / r[2] : 0.0, r[2]);
        }
  return 0;
}
