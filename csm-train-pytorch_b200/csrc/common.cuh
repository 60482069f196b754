// Shared helpers for libcsm_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <utility>

#include "../../include/csm_b200.h"

namespace csm {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
int num_sms();

inline cudaStream_t as_stream(csm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define CSM_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ::csm::set_error(__VA_ARGS__);        \
      return (code);                        \
    }                                       \
  } while (0)

// call after every kernel launch
#define CSM_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    ::csm::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ::csm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return CSM_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

// Programmatic dependent launch (csm_set_pdl / CSM_PDL, off by default — see api.cu for the measurement): a kernel launched through launch_k() may be scheduled while the
// previous kernel of the stream is still draining; every such kernel calls pdl_wait() before its first global-memory
// access (reads AND writes: the caching allocator reuses buffers in stream order) and pdl_trigger() right after, so at
// most one successor is parked behind a running kernel.  Its launch latency and prologue (barrier init, TMEM
// allocation, tensor-map prefetch) then overlap the predecessor's tail instead of following its completion.
extern std::atomic<int> g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// <<<grid, block, smem, st>>> with optional cluster width and the PDL attribute.  Only for kernels that call pdl_wait().
// cluster_x < 0: |cluster_x| is the cluster width and the PDL attribute is set regardless of the global switch (a small
// kernel that directly follows its producer, e.g. the CE combine behind the CE GEMM: its launch overlaps the GEMM's tail).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                            Args&&... args) {
  const bool force_pdl = cluster_x < 0;
  if (force_pdl) cluster_x = -cluster_x;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (force_pdl || g_pdl.load(std::memory_order_relaxed)) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Llama3ScaledRoPE on 8 consecutive bf16 of one head (4 interleaved pairs (x[2j], x[2j+1]), first pair index jg):
// `cache_pos` points at the [head_dim/2][cos, sin] fp32 row of this position.  Separate roundings (no FMA
// contraction): bit-identical to the fp32 torch formula and to rope_kernel; sgn = -1 gives the inverse rotation.
__device__ __forceinline__ void rope_rotate8(uint32_t* w, const float* cache_pos, int jg, float sgn) {
  const float4 c01 = *reinterpret_cast<const float4*>(cache_pos + jg * 2);
  const float4 c23 = *reinterpret_cast<const float4*>(cache_pos + jg * 2 + 4);
  const float co[4] = {c01.x, c01.z, c23.x, c23.z};
  const float si[4] = {c01.y * sgn, c01.w * sgn, c23.y * sgn, c23.w * sgn};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x0 = bf16_lo(w[j]), x1 = bf16_hi(w[j]);
    w[j] = pack_bf16(__fsub_rn(__fmul_rn(x0, co[j]), __fmul_rn(x1, si[j])),
                     __fadd_rn(__fmul_rn(x1, co[j]), __fmul_rn(x0, si[j])));
  }
}

// 16-byte streaming load that does not pollute L1 (read-once data)
__device__ __forceinline__ uint4 ld_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace csm
