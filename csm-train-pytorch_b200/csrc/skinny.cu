// Tall-skinny bf16 products of the LoRA path (reference: csm/mlx/components/lora.py:87-105 forward, autograd of it):
//   rowdot:  T[M, R]  = alpha * X[M, K] . W      W given as [R, K] (t = x A^T) or as [K, R] (dts = dy B)
//   coldot:  G        = alpha * X[N, C]^T T[N, R]   written as [C, R] (dB = dy^T t) or as [R, C] (dA = dts^T x)
// with R <= 64.  On the 128-wide tcgen05 tiles these are 1/8-filled MMAs behind a full TMA / TMEM / split-reduction
// pipeline (10-13 us + a 3.4 us reduction launch each, ~80 launches per CSM-1B LoRA step); they are pure streaming
// problems: one pass over X (16-25 MB, usually L2-resident: the previous kernel wrote it).  Here X goes from global
// memory straight into mma.sync fragments with 16-byte loads and no shared-memory staging:
//   * the reduction index of an MMA may be permuted freely as long as both operands use the same permutation, so a lane
//     takes 8 CONSECUTIVE elements of the reduction dimension (one 16-byte load) and feeds them to two k16 MMAs;
//   * where the reduction runs over rows (coldot) the two halves of a fragment register come from two different rows:
//     two 16-byte row loads + byte permutes give the eight column fragments; the output column permutation this implies
//     is undone when the (tiny) result is written.
// fp32 accumulation in registers, warps of a CTA split the reduction and are summed in a fixed order through shared
// memory; coldot splits the rows over a thread-block CLUSTER and sums the partial tiles through distributed shared
// memory in rank order — no atomics, no workspace, bit-reproducible.
#include "common.cuh"

namespace csm {

namespace {

__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lo_pair(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); }  // (a.lo, b.lo)
__device__ __forceinline__ uint32_t hi_pair(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }  // (a.hi, b.hi)
__device__ __forceinline__ uint32_t word_of(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
// Loads and MMAs are all `asm volatile`: volatile asms keep their program order, so the loads of a whole round (U chunks)
// are issued back to back before the first MMA that consumes them — one memory round trip per round.  (Left to the
// scheduler, ptxas sinks every load next to its MMA to save registers: one round trip per chunk, 2-3x slower.)
__device__ __forceinline__ uint32_t ld_u32(const bf16* p) {
  uint32_t r;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld16(const bf16* p) {          // small re-used operand: may stay in L1
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
  return v;
}

constexpr int kSkWarps = 8;
constexpr int kSkThreads = kSkWarps * 32;

// ------------------------------------------------------------------------------------------------ rowdot
// One CTA = 16 rows of X; its 8 warps take the 32-element chunks of K round-robin (so the CTA walks 512 contiguous
// bytes of every row per round).  NT = ceil(R / 8) column tiles.  WKR: W is [K, R] (R contiguous).
template <int NT, bool WKR>
__global__ void __launch_bounds__(kSkThreads, 2)
skinny_rowdot_kernel(const bf16* __restrict__ X, const bf16* __restrict__ W, bf16* __restrict__ T, int64_t M, int K,
                     int R, int64_t ldx, int64_t ldw, int64_t ldt, float alpha) {
  __shared__ float red[kSkWarps][16][NT * 8 + 1];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * 16;
  const int64_t ra = min(row0 + g, M - 1), rb = min(row0 + g + 8, M - 1);     // clamped: rows >= M are never stored
  const bf16* xa_p = X + ra * ldx + 8 * t;
  const bf16* xb_p = X + rb * ldx + 8 * t;
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const int chunks = K >> 5;
  // U chunks per round: all their loads are issued before the first MMA (one memory round trip per U chunks)
  constexpr int U = NT <= 2 ? 4 : (NT <= 4 ? 2 : 1);
  constexpr int NQ = (NT + 1) / 2;
  for (int c0 = warp; c0 < chunks; c0 += kSkWarps * U) {
    uint4 xa[U], xb[U];
    uint4 wv[WKR ? 1 : U][WKR ? 1 : NT];          // [R, K] operand: 8 consecutive k of row 8j + g
    uint32_t wk[WKR ? U : 1][WKR ? NQ : 1][8];    // [K, R] operand: rows k0 + 8t + i at the column pair (16q + 2g, +1)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u * kSkWarps;
      const bool ok = c < chunks;
      const int k0 = (ok ? c : warp) << 5;        // (clamped: a chunk past the end contributes zeros)
      xa[u] = ld_nc16(xa_p + k0);
      xb[u] = ld_nc16(xb_p + k0);
      if (!ok) xa[u] = xb[u] = make_uint4(0, 0, 0, 0);
      if (!WKR) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int n = 8 * j + g;
          wv[WKR ? 0 : u][WKR ? 0 : j] = n < R ? ld16(W + (int64_t)n * ldw + k0 + 8 * t) : make_uint4(0, 0, 0, 0);
        }
      } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int col = 16 * q + 2 * g;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            wk[WKR ? u : 0][WKR ? q : 0][i] = col < R ? ld_u32(W + (int64_t)(k0 + 8 * t + i) * ldw + col) : 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!WKR) {
        // column tile j = rows 8j .. 8j+7 of W; lane (g, t) holds W[8j + g][k0 + 8t .. + 7]
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint4 w = wv[WKR ? 0 : u][WKR ? 0 : j];
          mma16816(acc[j], xa[u].x, xb[u].x, xa[u].y, xb[u].y, w.x, w.y);
          mma16816(acc[j], xa[u].z, xb[u].z, xa[u].w, xb[u].w, w.z, w.w);
        }
      } else {
        // W rows are reduction indices: the column pair (16q + 2g, +1) gives the fragments of the two column tiles 2q
        // (even physical columns) and 2q + 1 (odd ones)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const uint32_t* w = wk[WKR ? u : 0][WKR ? q : 0];
          mma16816(acc[2 * q], xa[u].x, xb[u].x, xa[u].y, xb[u].y, lo_pair(w[0], w[1]), lo_pair(w[2], w[3]));
          mma16816(acc[2 * q], xa[u].z, xb[u].z, xa[u].w, xb[u].w, lo_pair(w[4], w[5]), lo_pair(w[6], w[7]));
          if (2 * q + 1 < NT) {
            mma16816(acc[2 * q + 1], xa[u].x, xb[u].x, xa[u].y, xb[u].y, hi_pair(w[0], w[1]), hi_pair(w[2], w[3]));
            mma16816(acc[2 * q + 1], xa[u].z, xb[u].z, xa[u].w, xb[u].w, hi_pair(w[4], w[5]), hi_pair(w[6], w[7]));
          }
        }
      }
    }
  }
  // accumulator (row g | g + 8, logical column 2t | 2t + 1 of tile j) -> physical column
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int l = 2 * t + e;
      const int col = WKR ? 16 * (j >> 1) + 2 * l + (j & 1) : 8 * j + l;
      red[warp][g][col] = acc[j][e];
      red[warp][g + 8][col] = acc[j][2 + e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * NT * 8; i += kSkThreads) {
    const int r = i / (NT * 8), col = i - r * (NT * 8);
    if (row0 + r >= M || col >= R) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kSkWarps; ++w) s += red[w][r][col];
    T[(row0 + r) * ldt + col] = __float2bfloat16_rn(s * alpha);
  }
}

// ------------------------------------------------------------------------------------------------ coldot
// One CTA = 64 columns of X x all R, over the rows of its cluster rank; warps take 16-row steps round-robin.
// MT = ceil(R / 16) row tiles of the output (the MMA's M dimension is R, its N dimension the columns of X).
// ORK: output stored [R, C] (else [C, R]).
template <int MT, bool ORK>
__global__ void __launch_bounds__(kSkThreads, MT == 1 ? 2 : 1)      // R <= 16: two CTAs per SM, the whole grid in one wave
skinny_coldot_kernel(const bf16* __restrict__ X, const bf16* __restrict__ Tm, bf16* __restrict__ G, int64_t N, int C,
                     int R, int64_t ldx, int64_t ldt, int64_t ldg, int splits, float alpha) {
  __shared__ float part[64][MT * 16 + 1];
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint32_t rank = splits > 1 ? cluster_rank() : 0u;
  const int col0 = ((int)blockIdx.x / splits) * 64;
  const int64_t rows_per = (((N + splits - 1) / splits) + 15) / 16 * 16;
  const int64_t n_begin = (int64_t)rank * rows_per, n_end = min(N, n_begin + rows_per);
  const int xcol = col0 + 8 * g;
  const bool col_ok = xcol < C;                      // C % 8 == 0: a lane's 8 columns are all in or all out
  float acc[MT][8][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;

  const int64_t steps = (n_end > n_begin) ? (n_end - n_begin + 15) / 16 : 0;
  // U 16-row steps per round: all their loads are issued before the first MMA
  constexpr int U = 2;
  for (int64_t s0 = warp; s0 < steps; s0 += kSkWarps * U) {
    uint4 xr[U][4];
    uint32_t tr[U][MT][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // this lane's rows: n0, n0 + 1, n0 + 8, n0 + 9 (a step past the end has n0 >= n_end: all zeros)
      const int64_t n0 = n_begin + (s0 + (int64_t)u * kSkWarps) * 16 + 2 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t n = n0 + (i & 1) + (i >> 1) * 8;
        xr[u][i] = (col_ok && n < n_end) ? ld_nc16(X + n * ldx + xcol) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const int r = 16 * m + 2 * g;              // physical rows r (logical row g) and r + 1 (logical row g + 8)
          tr[u][m][i] = (r < R && n < n_end) ? ld_u32(Tm + n * ldt + r) : 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        a[m][0] = lo_pair(tr[u][m][0], tr[u][m][1]); a[m][1] = hi_pair(tr[u][m][0], tr[u][m][1]);
        a[m][2] = lo_pair(tr[u][m][2], tr[u][m][3]); a[m][3] = hi_pair(tr[u][m][2], tr[u][m][3]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {                  // column tile j = physical column 8g' + j of lane group g'
        const uint32_t w0 = word_of(xr[u][0], j >> 1), w1 = word_of(xr[u][1], j >> 1);
        const uint32_t w8 = word_of(xr[u][2], j >> 1), w9 = word_of(xr[u][3], j >> 1);
        const uint32_t b0 = (j & 1) ? hi_pair(w0, w1) : lo_pair(w0, w1);
        const uint32_t b1 = (j & 1) ? hi_pair(w8, w9) : lo_pair(w8, w9);
#pragma unroll
        for (int m = 0; m < MT; ++m) mma16816(acc[m][j], a[m][0], a[m][1], a[m][2], a[m][3], b0, b1);
      }
    }
  }
  // CTA partial: the warps add their accumulators in warp order (fixed summation order)
  for (int i = threadIdx.x; i < 64 * (MT * 16 + 1); i += kSkThreads) (&part[0][0])[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < kSkWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 8 * (2 * t + e) + j;     // logical column 2t + e of tile j
            part[col][16 * m + 2 * g] += acc[m][j][e];           // logical row g
            part[col][16 * m + 2 * g + 1] += acc[m][j][2 + e];   // logical row g + 8
          }
    }
    __syncthreads();
  }
  if (splits > 1) cluster_barrier();
  // every rank finishes a slice of the 64 columns: partials summed in rank order
  const int cols_per = 64 / splits;
  for (int i = threadIdx.x; i < cols_per * R; i += kSkThreads) {
    int col, r;
    if (ORK) { r = i / cols_per; col = (int)rank * cols_per + (i - r * cols_per); }
    else { col = (int)rank * cols_per + i / R; r = i - (i / R) * R; }
    if (col0 + col >= C) continue;
    float s = 0.f;
    if (splits > 1) {
      for (int q = 0; q < splits; ++q) s += ld_dsmem_f32(&part[col][r], (uint32_t)q);
    } else {
      s = part[col][r];
    }
    const bf16 o = __float2bfloat16_rn(s * alpha);
    if (ORK) G[(int64_t)r * ldg + col0 + col] = o;
    else G[(int64_t)(col0 + col) * ldg + r] = o;
  }
  if (splits > 1) cluster_barrier();                 // nobody leaves while a peer may still read its partial
}

}  // namespace

static std::atomic<int> g_skinny_mode{1};
void skinny_set_mode(int m) { g_skinny_mode.store(m); }

bool skinny_rowdot_supported(const void* X, const void* W, const void* T, int64_t M, int64_t K, int64_t R, int64_t ldx,
                             int64_t ldw, int64_t ldt, int w_kr) {
  if (g_skinny_mode.load() == 0) return false;
  if (M < 1 || R < 2 || R > 64 || (R & 1) || K < 64 || (K & 31) || K >= (1ll << 31)) return false;
  if (!aligned16(X) || (ldx & 7)) return false;
  if (w_kr) { if ((reinterpret_cast<uintptr_t>(W) & 3) || (ldw & 1)) return false; }
  else if (!aligned16(W) || (ldw & 7)) return false;
  (void)T; (void)ldt;
  return true;
}

int skinny_rowdot_launch(const void* X, const void* W, void* T, int64_t M, int64_t K, int64_t R, int64_t ldx,
                         int64_t ldw, int64_t ldt, int w_kr, float alpha, cudaStream_t st) {
  // [K, R] operand: column tiles come in (even, odd physical column) pairs, so their count is rounded up to even
  const int nt = w_kr ? 2 * (int)((R + 15) / 16) : (int)((R + 7) / 8);
  const dim3 grid((unsigned)((M + 15) / 16)), block(kSkThreads);
  cudaError_t e = cudaSuccess;
#define RD(NT_)                                                                                                        \
  e = w_kr ? launch_k(skinny_rowdot_kernel<NT_, true>, grid, block, 0, st, 1, (const bf16*)X, (const bf16*)W, (bf16*)T, \
                      M, (int)K, (int)R, ldx, ldw, ldt, alpha)                                                         \
           : launch_k(skinny_rowdot_kernel<NT_, false>, grid, block, 0, st, 1, (const bf16*)X, (const bf16*)W,          \
                      (bf16*)T, M, (int)K, (int)R, ldx, ldw, ldt, alpha)
  switch (nt) {
    case 1: RD(1); break;
    case 2: RD(2); break;
    case 3: RD(3); break;
    case 4: RD(4); break;
    case 5: case 6: RD(6); break;
    default: RD(8); break;
  }
#undef RD
  if (e != cudaSuccess) { set_error("skinny_rowdot: launch failed: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("skinny_rowdot");
  return CSM_OK;
}

bool skinny_coldot_supported(const void* X, const void* Tm, const void* G, int64_t N, int64_t C, int64_t R, int64_t ldx,
                             int64_t ldt, int64_t ldg, int out_rk) {
  if (g_skinny_mode.load() == 0) return false;
  if (N < 16 || R < 2 || R > 64 || (R & 1) || C < 8 || (C & 7) || C >= (1ll << 31)) return false;
  if (!aligned16(X) || (ldx & 7)) return false;
  if ((reinterpret_cast<uintptr_t>(Tm) & 3) || (ldt & 1)) return false;
  (void)G; (void)ldg; (void)out_rk;
  return true;
}

int skinny_coldot_launch(const void* X, const void* Tm, void* G, int64_t N, int64_t C, int64_t R, int64_t ldx,
                         int64_t ldt, int64_t ldg, int out_rk, float alpha, cudaStream_t st) {
  const int mt = (int)((R + 15) / 16);
  const int colblocks = (int)((C + 63) / 64);
  // row splits (= cluster width): enough CTAs for the machine, at least 16 rows per warp and split
  int splits = colblocks >= 96 ? 2 : colblocks >= 40 ? 4 : 8;
  while (splits > 1 && N / splits < 16 * kSkWarps) splits >>= 1;
  const dim3 grid((unsigned)(colblocks * splits)), block(kSkThreads);
  cudaError_t e = cudaSuccess;
#define CD(MT_)                                                                                                        \
  e = out_rk ? launch_k(skinny_coldot_kernel<MT_, true>, grid, block, 0, st, splits, (const bf16*)X, (const bf16*)Tm,  \
                        (bf16*)G, N, (int)C, (int)R, ldx, ldt, ldg, splits, alpha)                                     \
             : launch_k(skinny_coldot_kernel<MT_, false>, grid, block, 0, st, splits, (const bf16*)X, (const bf16*)Tm, \
                        (bf16*)G, N, (int)C, (int)R, ldx, ldt, ldg, splits, alpha)
  switch (mt) {
    case 1: CD(1); break;
    case 2: CD(2); break;
    case 3: CD(3); break;
    default: CD(4); break;
  }
#undef CD
  if (e != cudaSuccess) { set_error("skinny_coldot: launch failed: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("skinny_coldot");
  return CSM_OK;
}

}  // namespace csm

using namespace csm;

extern "C" void csm_set_skinny_mode(int32_t mode) { csm::skinny_set_mode(mode); }

extern "C" int csm_skinny_supported(int32_t kind, const void* X, const void* W, const void* out, int64_t rows,
                                    int64_t cols, int64_t R, int64_t ldx, int64_t ldw, int64_t ldo, int32_t layout) {
  if (csm_device_supported() != 1) return 0;
  if (kind == 0) return skinny_rowdot_supported(X, W, out, rows, cols, R, ldx, ldw, ldo, layout) ? 1 : 0;
  return skinny_coldot_supported(X, W, out, rows, cols, R, ldx, ldw, ldo, layout) ? 1 : 0;
}

extern "C" int csm_skinny_rowdot(const void* X, const void* W, void* T, int64_t M, int64_t K, int64_t R, int64_t ldx,
                                 int64_t ldw, int64_t ldt, int32_t w_is_kr, float alpha, csm_stream_t stream) {
  CSM_REQUIRE(csm_device_supported() == 1, CSM_ERR_ARCH, "skinny_rowdot: needs an sm_100 device");
  CSM_REQUIRE(skinny_rowdot_supported(X, W, T, M, K, R, ldx, ldw, ldt, w_is_kr), CSM_ERR_SHAPE,
              "skinny_rowdot: needs 2 <= R <= 64 (even), K a multiple of 32, 16-byte aligned X rows (M=%lld K=%lld R=%lld)",
              (long long)M, (long long)K, (long long)R);
  return skinny_rowdot_launch(X, W, T, M, K, R, ldx, ldw, ldt, w_is_kr, alpha, as_stream(stream));
}

extern "C" int csm_skinny_coldot(const void* X, const void* Tm, void* G, int64_t N, int64_t C, int64_t R, int64_t ldx,
                                 int64_t ldt, int64_t ldg, int32_t out_is_rk, float alpha, csm_stream_t stream) {
  CSM_REQUIRE(csm_device_supported() == 1, CSM_ERR_ARCH, "skinny_coldot: needs an sm_100 device");
  CSM_REQUIRE(skinny_coldot_supported(X, Tm, G, N, C, R, ldx, ldt, ldg, out_is_rk), CSM_ERR_SHAPE,
              "skinny_coldot: needs 2 <= R <= 64 (even), C a multiple of 8, 16-byte aligned X rows (N=%lld C=%lld R=%lld)",
              (long long)N, (long long)C, (long long)R);
  return skinny_coldot_launch(X, Tm, G, N, C, R, ldx, ldt, ldg, out_is_rk, alpha, as_stream(stream));
}
