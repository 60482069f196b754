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


__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
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
void run2(const char* what, int nc, long long* d) {
  const int iters = 64;
  cudaFuncSetAttribute(k2<N, MODE, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k2<N, MODE, PER><<<148, 128, 100 * 1024>>>(iters, nc, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  printf("[elect] %-52s %8.1f cycles/MMA  (%lld cycles, %d MMAs)\n", what, (double)cyc / (iters * PER), cyc, iters * PER);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
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
  return 0;
}
