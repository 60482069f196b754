// K4: persistent warp-specialised bf16 GEMM on the 5th-gen tensor cores.
//   warp 0      : TMA producer — 128B-swizzled A/B tiles into a kStages-deep smem ring (mbarrier full/empty)
//   warp 1      : TMEM allocator + tcgen05.mma issuer; accumulators live in TMEM, double-buffered
//   warps 2..9  : epilogue — tcgen05.ld the 128 x BN fp32 tile, apply {alpha, residual, accumulate, RoPE | SwiGLU fwd /
//                 bwd | CE partial | CE dlogits}, store through a swizzled smem tile (whole 64-byte row segments).
//                 Overlaps the next tile's main loop through the second TMEM buffer.
// Producer and MMA warps keep warp-uniform control flow and elect one lane per TMA / tcgen05 instruction
// (tc::elect_one): issued under `if (lane == 0)` every UTCHMMA costs an R2UR/ELECT/BRA.U.ANY waterfall loop.
// Two tilings: single CTA (128 x BN, BN 128/256) and CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256:
// each CTA loads its own 128 rows of A and half of the B tile, the leader CTA issues the MMAs for both, loads of both
// CTAs complete on the leader's barrier, commits are multicast to both).  The pair moves a third less operand data
// per flop and is what the big shapes use (93.6 % tensor-pipe activity vs 71.6 % for the single-CTA tiling).
// Operands may be K-major or MN-major (transposed storage) — the backward GEMMs (dgrad / wgrad) read the same
// tensors the forward wrote, no transposes are materialised.  An optional extra K-block (A2/B2) carries the
// LoRA low-rank term inside the main loop.  `groups` batches independent problems (31 audio heads, or the K-slices of
// a split reduction) in one launch.  A stream-K schedule exists for badly filled last waves (off: no gain measured).
#include "tc_common.cuh"

namespace csm {

using namespace tc;

constexpr int BM = 128, BK = 64;
// Two measured-and-rejected experiments (stream-K tile cutting, narrow MMAs on ragged last column tiles: both correct,
// neither faster on a power-capped B200 — profiles/r1_summary.md, r2_summary.md) are compiled OUT of the kernels unless
// the library is built with -DCSM_GEMM_EXPERIMENTS (CSM_EXTRA_NVCC_FLAGS of build.sh); csm_gemm_experiments_compiled()
// reports it, and their mode setters are then no-ops.
#ifdef CSM_GEMM_EXPERIMENTS
constexpr bool kExperiments = true;
#else
constexpr bool kExperiments = false;
#endif
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
// Two epilogue warps per TMEM lane quadrant split the tile's columns (the fused SwiGLU epilogues must finish inside
// one main loop, ~9 us at K = 2048); the CE epilogues keep whole rows thread-local, so only warps 2..5 work there.
__host__ __device__ constexpr int epi_warps(int epi) { return (epi == 1 || epi == 2) ? 4 : 8; }

// EPI_SWIGLU_FWD (CTA-pair mode only): the B tile of a pair is [128 gate rows | 128 up rows] of the packed w1|w3
//   weight (CTA 0 loads the gate rows, CTA 1 the up rows), so one accumulator row holds gate and up of the same 128
//   columns: the epilogue writes gate, up (saved for backward) and act = silu(gate) * up — the swiglu kernel and its
//   re-read of the [M, 2I] buffer disappear.
// EPI_SWIGLU_BWD: the w2 dgrad GEMM's epilogue reads gate/up and writes dgate/dup directly (dact never exists).
enum { EPI_STORE = 0, EPI_CE_PARTIAL = 1, EPI_CE_DLOGITS = 2, EPI_SWIGLU_FWD = 3, EPI_SWIGLU_BWD = 4 };

struct GemmTcParams {
  int64_t M, N, K;          // per-group problem
  int groups;
  int num_m, num_n;         // tile grid per group
  int k_blocks;             // ceil(K / 64) main blocks
  int has_tail;             // extra k-blocks from (A2, B2): ceil(K2 / 64), K2 <= 256 (several LoRA adapters side by side)
  // EPI_STORE
  void* C; const bf16* R;
  int64_t ldc, ldr, c_group_stride, r_group_stride;
  int c_f32, r_f32, accumulate;
  float alpha;
  // SwiGLU epilogues: C = gate|up buffer [M, 2*inter] (fwd: written, bwd: R = gate|up read, C = dgate|dup written),
  // C2 = act [M, inter] (fwd)
  void* C2; int64_t ldc2; int64_t inter;
  // stream-K (CTA-pair EPI_STORE only): the tiles' k-blocks are dealt out evenly to the pairs; a tile cut between two
  // pairs is finished by the pair that owns its last k-block, which adds the other pair's fp32 partial (sk_ws) first
  // RoPE in the store epilogue (fused q|k|v projection): columns [0, rope_cols) are heads of rope_hd that get rotated
  // by the position row % rope_seq; the linear output is rounded to bf16 first, as the unfused rope kernel would read it
  const float* rope_cache; int rope_seq, rope_cols, rope_hd;
  const int32_t* rope_pos;  // optional [M]: position of every row (sequence packing: positions restart per sample)
  int streamk;
  // off by default (no gain measured): issue the MMAs of a ragged last column tile with N rounded up to 16 instead of BN
  int narrow_tail;
  // dynamic tile scheduler (whole-tile schedules): tile_ctr[0] = next tile index (atomicAdd), tile_ctr[1] = CTAs that
  // have finished; the last CTA re-zeroes both.  nullptr: static round-robin assignment (tile = unit + i * units).
  // A persistent grid with a static assignment assumes every CTA is co-resident from the start; when another kernel
  // holds SMs (NCCL's all-reduce CTAs overlapping the backward under data parallelism) the CTAs that start late still
  // own a full share of tiles and the GEMM takes up to twice as long.  With the counter a late CTA simply finds less
  // (or no) work left.
  int* tile_ctr;
  float* sk_ws;             // [pairs][2 CTAs][BN cols][128 rows] fp32
  int* sk_flags;            // [pairs][2 CTAs][2]: partial-ready count, readers-done count (self-resetting)
  // CE epilogues
  const int64_t* targets; int64_t tgt_row_stride, tgt_group_stride;
  float4* ce_part;          // [groups*M, num_n]
  const float* lse; float gscale; const float* gscale_dev;
};

// kCta2: CTA-pair mode (cluster of 2, tcgen05 cta_group::2).  The pair computes a 256 x BN tile: each CTA keeps its
// own 128 rows of A and HALF of the B tile in smem (32 KB per stage instead of 48 KB -> 6 stages, a third less
// L2->SM traffic per flop), the leader CTA issues M=256 MMAs, each CTA's 128 x BN accumulator lives in its own TMEM.
template <int BN, bool kCta2> struct GemmCfg {
  static constexpr int kBRows = kCta2 ? BN / 2 : BN;            // B-tile rows held by this CTA
  static constexpr int kStages = (BN == 256 && !kCta2) ? 4 : 6;
  static constexpr int kStageBytes = BM * BK * 2 + kBRows * BK * 2;
  static constexpr int kStoreStageBytes = 8 * 2048;           // one 32 x 64 B staging tile per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kStoreStageBytes;
  static constexpr int kTmemCols = 2 * BN;
};

// sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5: one MUFU instead of exp + full-precision divide (abs error ~1e-6 x 2^-11,
// far below the bf16 rounding applied to every value it feeds)
__device__ __forceinline__ float fast_sigmoid(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(t, 0.5f, 0.5f);
}

// Writes a 32-row x 32-column bf16 tile that the warp holds one ROW PER LANE (the TMEM layout) to global memory.
// Stored straight from registers every lane would write 16 B into a different row: 32 half-filled 32-byte sectors per
// instruction (ncu: l1tex->xbar write bytes = 2x the tile).  Staged through a 2 KB per-warp smem tile (XOR-swizzled
// 16-byte chunks, conflict-free both ways) each instruction writes eight complete 64-byte row segments instead.
__device__ __forceinline__ void store_tile_32x32(uint32_t stage_saddr, bf16* tile_origin, int64_t ld, int lane,
                                                 const uint32_t* packed /*16 words = this lane's row*/, int rows_valid) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t a = stage_saddr + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(packed[q * 4 + 0]), "r"(packed[q * 4 + 1]),
                 "r"(packed[q * 4 + 2]), "r"(packed[q * 4 + 3])
                 : "memory");
  }
  __syncwarp();
  const int chunk = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2);
    const uint32_t a = stage_saddr + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    if (row < rows_valid) *reinterpret_cast<uint4*>(tile_origin + (int64_t)row * ld + chunk * 8) = v;
  }
  __syncwarp();
}

// The mirror image for loads: `co[i]` holds the 16-byte piece (row i*8 + lane/4, chunk lane%4) of a 32 x 32 bf16 tile
// (fetched with whole-64-byte-row-segment loads); returns this lane's own row (16 words) through the staging tile.
__device__ __forceinline__ void rows_from_coalesced(uint32_t stage_saddr, int lane, const uint4* co, uint32_t* row_words) {
  const int chunk = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2);
    const uint32_t a = stage_saddr + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(co[i].x), "r"(co[i].y), "r"(co[i].z), "r"(co[i].w)
                 : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t a = stage_saddr + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(row_words[q * 4 + 0]), "=r"(row_words[q * 4 + 1]), "=r"(row_words[q * 4 + 2]), "=r"(row_words[q * 4 + 3])
                 : "r"(a)
                 : "memory");
  }
  __syncwarp();
}

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& g, int& mb, int& nb) {
  const int per_group = num_m * num_n;
  g = t / per_group;
  int r = t - g * per_group;
  // grouped rasterisation: 16 m-blocks share each sweep over n so A and B tiles stay L2-resident
  constexpr int GM = 16;
  const int band = r / (GM * num_n);
  const int first_m = band * GM;
  const int band_m = min(GM, num_m - first_m);
  const int in_band = r - band * GM * num_n;
  mb = first_m + in_band % band_m;
  nb = in_band / band_m;
}

// One unit of a CTA's (pair's) persistent loop: the k-blocks [kb0, kb1) of output tile `tile`.
//   kind 0: the whole tile.   kind 1 (stream-K): the HEAD part of a tile that the next pair finishes — its fp32
//   accumulator goes to the workspace.   kind 2 (stream-K): the TAIL part — adds the previous pair's partial, then the
//   normal epilogue.  Every pair does its kind-1 item first and its kind-2 item second.
struct WorkItem { int tile, kb0, kb1, kind; };

__device__ __forceinline__ bool next_item(int streamk, int unit, int n_units, int num_tiles, int kb_total, int idx,
                                          WorkItem& w) {
  if (!streamk) {
    w.tile = unit + idx * n_units; w.kb0 = 0; w.kb1 = kb_total; w.kind = 0;
    return w.tile < num_tiles;
  }
  const long long U = (long long)num_tiles * kb_total;
  const long long u0 = U * unit / n_units, u1 = U * (unit + 1) / n_units;
  const int tA = (int)(u0 / kb_total), ka = (int)(u0 % kb_total);
  const int tC = (int)(u1 / kb_total), kc = (int)(u1 % kb_total);
  const int head = kc > 0 ? 1 : 0;                         // this pair starts a tile it does not finish
  const int f0 = ka == 0 ? tA : tA + 1;                    // first whole tile
  const int nfull = tC - f0;
  // order: the head part first (its partial is what the NEXT pair waits for), then the tail part this pair finishes
  // (the previous pair's partial is on its way; the fix-up epilogue then hides behind the whole tiles), then whole tiles
  if (idx < head) { w.tile = tC; w.kb0 = 0; w.kb1 = kc; w.kind = 1; return true; }
  idx -= head;
  const int tail = ka > 0 ? 1 : 0;
  if (idx < tail) { w.tile = tA; w.kb0 = ka; w.kb1 = kb_total; w.kind = 2; return true; }
  idx -= tail;
  if (idx < nfull) { w.tile = f0 + idx; w.kb0 = 0; w.kb1 = kb_total; w.kind = 0; return true; }
  return false;
}

template <bool kTransA, bool kTransB, int BN, int kEpi, bool kCta2>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const GemmTcParams p) {
  using Cfg = GemmCfg<BN, kCta2>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* tq_full = bars + 2 * kStages + 5;       // [4] tile-queue slot published (leader's producer -> everyone)
  uint64_t* tq_empty = tq_full + 4;                 // [4] ... read by every consumer warp (of both CTAs of a pair)
  volatile uint32_t* tq = reinterpret_cast<volatile uint32_t*>(tq_empty + 4);   // [4] (tag << 20) | (tile + 1)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA-pair mode: the scheduling unit is the pair; it walks (m-block pair, n-block) super-tiles in lockstep
  const uint32_t rank = kCta2 ? cluster_ctarank() : 0u;
  const int tile_m = kCta2 ? (p.num_m + 1) / 2 : p.num_m;
  const int num_tiles = p.groups * tile_m * p.num_n;
  const int first_tile = kCta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kCta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kb_total = p.k_blocks + p.has_tail;
  const int sk = (kExperiments && kCta2 && kEpi == EPI_STORE) ? p.streamk : 0;
  const bool dyn = p.tile_ctr != nullptr && !sk && num_tiles < (1 << 20) - 1;
  // consumers of a tile-queue slot: producer warp + epilogue warps of every CTA, MMA warp of the leader
  constexpr uint32_t kTqConsumers = kCta2 ? 3 + 2 * epi_warps(kEpi) : 2 + epi_warps(kEpi);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.has_tail) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    // pair mode: the leader's MMA warp waits for the epilogue warps of BOTH CTAs
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], epi_warps(kEpi) * (kCta2 ? 2 : 1));
    }
    for (int q = 0; q < 4; ++q) { mbar_init(&tq_full[q], 1); mbar_init(&tq_empty[q], kTqConsumers); tq[q] = 0; }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (kCta2) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols); else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (kCta2) cluster_sync();      // the peer's barriers are initialised before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) touched no global memory and may have run
  // while the previous kernel of the stream was still draining (programmatic dependent launch, common.cuh)
  pdl_wait();
  pdl_trigger();

  // Dynamic schedule.  Item 0 of every CTA (pair) is its static tile (no start-up latency); item it >= 1 is the tile
  // index the LEADER's producer warp drew from the global counter and wrote, tagged with `it`, into queue slot
  // (it - 1) % 4 of its shared memory.  The word is self-validating ((tag << 20) | (tile + 1), tile + 1 == 0: no work
  // left), so the peer CTA of a pair needs no release/acquire hand-off: its producer warp polls the leader's word through
  // DSMEM and republishes it locally.  Consumers arrive on the LEADER's tq_empty (the peer's remotely, relaxed).
  auto tq_tag = [](int it) -> uint32_t { return (uint32_t)((it & 0x7ff) + 1); };
  auto take_tile = [&](int it, WorkItem& w) -> bool {
    if (!dyn) return next_item(sk, first_tile, tile_step, num_tiles, kb_total, it, w);
    w.kb0 = 0; w.kb1 = kb_total; w.kind = 0;
    if (it == 0) { w.tile = first_tile; return first_tile < num_tiles; }
    const int slot = (it - 1) & 3;
    const uint32_t ph = (uint32_t)((it - 1) >> 2) & 1u;
    mbar_wait(&tq_full[slot], ph);
    const int t = (int)(tq[slot] & 0xfffffu) - 1;
    __syncwarp();
    if (lane == 0) {
      if (kCta2 && rank != 0) mbar_arrive_cluster(mapa(smem_u32(&tq_empty[slot]), 0)); else mbar_arrive(&tq_empty[slot]);
    }
    w.tile = t;
    return t >= 0;
  };
  // the same for a caller that is ONE thread (the MMA issuer): no warp-level synchronisation around the arrive
  auto take_tile_1t = [&](int it, WorkItem& w) -> bool {
    if (!dyn) return next_item(sk, first_tile, tile_step, num_tiles, kb_total, it, w);
    w.kb0 = 0; w.kb1 = kb_total; w.kind = 0;
    if (it == 0) { w.tile = first_tile; return first_tile < num_tiles; }
    const int slot = (it - 1) & 3;
    mbar_wait(&tq_full[slot], (uint32_t)((it - 1) >> 2) & 1u);
    const int t = (int)(tq[slot] & 0xfffffu) - 1;
    mbar_arrive(&tq_empty[slot]);                 // (only the leader CTA's MMA warp runs)
    w.tile = t;
    return t >= 0;
  };

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    WorkItem w;
    // publishes item `it` (>= 1) into this CTA's queue; returns through tq_open whether more work may follow.
    // Leader: draws the tile from the counter.  Peer: relays the leader's word.
    bool tq_open = dyn;
    auto publish = [&](int it) {
      const int slot = (it - 1) & 3;
      const uint32_t ph = (uint32_t)((it - 1) >> 2) & 1u;
      uint32_t word = 0;
      if (!kCta2 || rank == 0) {
        mbar_wait(&tq_empty[slot], ph ^ 1);        // every consumer (of both CTAs) has read what the slot held before
        if (elect_one()) {
          const int t = tile_step + atomicAdd(p.tile_ctr, 1);
          word = (tq_tag(it) << 20) | (uint32_t)(t < num_tiles ? t + 1 : 0);
          tq[slot] = word;
          mbar_arrive(&tq_full[slot]);
        }
      } else {
        if (elect_one()) {
          const uint32_t src = mapa(smem_u32(const_cast<uint32_t*>(&tq[slot])), 0);
          do {
            asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(word) : "r"(src) : "memory");
          } while ((word >> 20) != tq_tag(it));
          tq[slot] = word;
          mbar_arrive(&tq_full[slot]);
        }
      }
      __syncwarp();
      if ((tq[slot] & 0xfffffu) == 0) tq_open = false;
    };
    for (int it = 0;; ++it) {
      if (!take_tile(it, w)) break;
      int g, mb, nb;
      tile_coords(w.tile, tile_m, p.num_n, g, mb, nb);
      if (kCta2) mb = 2 * mb + (int)rank;
      const int m0 = mb * BM;
      // pair mode: this CTA's half of the B tile (SwiGLU forward: rank 0 = gate rows, rank 1 = up rows)
      const int n0 = kEpi == EPI_SWIGLU_FWD ? (int)rank * (int)p.inter + nb * Cfg::kBRows
                                            : nb * BN + (kCta2 ? (int)rank * Cfg::kBRows : 0);
      for (int kb = w.kb0; kb < w.kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          const bool tail = kb >= p.k_blocks;
          const CUtensorMap* ma = tail ? &tmA2 : &tmA;
          const CUtensorMap* mbp = tail ? &tmB2 : &tmB;
          const int k0 = tail ? (kb - p.k_blocks) * BK : kb * BK;
          if (!kCta2) {
            mbar_expect_tx(&full[stage], Cfg::kStageBytes);
            if (!kTransA) {
              tma_load_3d(sa, ma, &full[stage], k0, m0, g);                     // box {64 k, 128 m}
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)                                 // box {64 m, 64 k} x2
                tma_load_3d(sa + c * (64 * BK * 2), ma, &full[stage], m0 + c * 64, k0, g);
            }
            if (!kTransB) {
              tma_load_3d(sb, mbp, &full[stage], k0, n0, g);                    // box {64 k, BN n}
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)                                 // box {64 n, 64 k} x BN/64
                tma_load_3d(sb + c * (64 * BK * 2), mbp, &full[stage], n0 + c * 64, k0, g);
            }
          } else {
            // both CTAs' bytes are counted by the LEADER's full barrier (its MMA warp is the only consumer)
            const uint32_t fbar = mapa(smem_u32(&full[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
            if (!kTransA) {
              tma_load_3d_2sm(sa, ma, fbar, k0, m0, g);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_3d_2sm(sa + c * (64 * BK * 2), ma, fbar, m0 + c * 64, k0, g);
            }
            if (!kTransB) {
              tma_load_3d_2sm(sb, mbp, fbar, k0, n0, g);                        // box {64 k, BN/2 n}
            } else {
#pragma unroll
              for (int c = 0; c < Cfg::kBRows / 64; ++c)
                tma_load_3d_2sm(sb + c * (64 * BK * 2), mbp, fbar, n0 + c * 64, k0, g);
            }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        // next item: drawn once this tile's loads fill the ring (from then on the producer would only wait for free
        // slots, so the ~1 us of the atomic costs nothing; drawn right after the first k-block it starved the MMAs at
        // the start of every tile: +4 % on the K = 8192 GEMMs)
        if (tq_open && kb == min(w.kb0 + kStages - 1, w.kb1 - 1)) publish(it + 1);
      }
    }
    // This unit has made its last draw (the one that found no tile).  Once every drawing unit has, nobody touches the
    // counter again in this launch: the last one re-arms it for the next launch on this slot — here, under the last
    // tile's MMAs and epilogue, not at kernel exit (that cost ~2 us per launch).
    if (dyn && rank == 0 && elect_one()) {
      if (atomicAdd(p.tile_ctr + 1, 1) == tile_step - 1) {
        p.tile_ctr[0] = 0;
        p.tile_ctr[1] = 0;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected thread runs the whole loop =====================
    // (a warp-wide mbarrier wait + elect + __syncwarp around every k-block costs ~65 cycles even when the barrier is
    // already open — tools/ubench/mma_rate.cu k4 — and delays the hand-over of finished accumulators and smem stages)
    constexpr uint32_t idesc = make_idesc_bf16(kCta2 ? 2 * BM : BM, BN, kTransA ? 1 : 0, kTransB ? 1 : 0);
    if (rank == 0 && elect_one()) {                  // (pair mode: the leader issues for both CTAs)
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      WorkItem w;
      for (int it = 0; take_tile_1t(it, w); ++it) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t tile_idesc = idesc;
        if (kExperiments && !kTransB && (kEpi == EPI_STORE || kEpi == EPI_CE_PARTIAL || kEpi == EPI_CE_DLOGITS) &&
            p.narrow_tail) {
          // K-major B: smem rows are output columns, rows past N are TMA zero fill.  A pair takes N/2 rows from each
          // CTA (columns [0, N/2) from the leader's half), so the narrow form needs every valid column in the leader.
          int g, mb, nb;
          tile_coords(w.tile, tile_m, p.num_n, g, mb, nb);
          const int valid = (int)min((int64_t)BN, p.N - (int64_t)nb * BN);
          const int n_eff = kCta2 ? (valid > BN / 2 ? BN : ((2 * valid + 15) & ~15)) : ((valid + 15) & ~15);
          tile_idesc = make_idesc_bf16(kCta2 ? 2 * BM : BM, n_eff, kTransA ? 1 : 0, 0);
        }
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + BM * BK * 2;
          // K-major: 8-row groups 1024 B apart, K advance = 32 B inside the swizzle atom.
          // MN-major: 64-wide MN chunks (64 x BK x 2 = 8192 B apart), K advance = 16 rows x 128 B.
          const uint64_t adesc = kTransA ? make_smem_desc(sa, 64 * BK * 2, 1024) : make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = kTransB ? make_smem_desc(sb, 64 * BK * 2, 1024) : make_smem_desc(sb, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ad = adesc + (uint64_t)((kTransA ? kk * 16 * 128 : kk * 32) >> 4);
            const uint64_t bd = bdesc + (uint64_t)((kTransB ? kk * 16 * 128 : kk * 32) >> 4);
            const uint32_t accum = ((kb - w.kb0) | kk) ? 1u : 0u;      // the first MMA of a work item overwrites
            if (kCta2) umma_bf16_2sm(d_tmem, ad, bd, tile_idesc, accum);
            else umma_bf16(d_tmem, ad, bd, tile_idesc, accum);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (kCta2) umma_commit_2sm(&empty[stage], 3); else umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (kCta2) umma_commit_2sm(&tfull[acc], 3); else umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps, TMEM lane quadrant = warp % 4) =====================
    const int quad = warp & 3;
    const int ehalf = (warp - 2) >> 2;                        // which half of the tile's columns this warp drains
    constexpr int kColsPerWarp = BN / (epi_warps(kEpi) / 4);
    const int c_lo = ehalf * kColsPerWarp, c_hi = c_lo + kColsPerWarp;
    int acc = 0; uint32_t acc_phase = 0;
    WorkItem w;
    for (int it = 0; ehalf < epi_warps(kEpi) / 4 && take_tile(it, w); ++it) {
      int g, mb, nb;
      tile_coords(w.tile, tile_m, p.num_n, g, mb, nb);
      if (kCta2) mb = 2 * mb + (int)rank;
      const int64_t m = (int64_t)mb * BM + quad * 32 + lane;
      const int64_t n0 = (int64_t)nb * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(quad * 32) << 16);
      const bool row_ok = m < p.M;
      // coalesced tile stores (store_tile_32x32): the warp's first row and how many of its 32 rows exist
      const int64_t wm0 = (int64_t)mb * BM + quad * 32;
      const int wrows = (int)max((int64_t)0, min((int64_t)32, p.M - wm0));
      const uint32_t stage_s = smem_u32(smem + kStages * Cfg::kStageBytes + 256) + (uint32_t)(warp - 2) * 2048u;

      if (kEpi == EPI_STORE) {
        const int64_t crow = (int64_t)g * p.c_group_stride + m * p.ldc;
        const bf16* rrow = p.R ? p.R + (int64_t)g * p.r_group_stride + m * p.ldr : nullptr;
        // stream-K partial tiles: column-major [BN][128] fp32 per (pair, CTA) so a warp's 32 rows are contiguous
        const int my_pair = first_tile;
        float* sk_out = nullptr;
        const float* sk_in = nullptr;
        // (layout [BN / 4][128 rows][4 cols]: one float4 per lane, 512 contiguous bytes per warp access)
        if (kExperiments && kCta2 && w.kind == 1)
          sk_out = p.sk_ws + ((size_t)(my_pair * 2 + (int)rank) * BN) * BM + (quad * 32 + lane) * 4;
        if (kExperiments && kCta2 && w.kind == 2) {
          sk_in = p.sk_ws + ((size_t)((my_pair - 1) * 2 + (int)rank) * BN) * BM + (quad * 32 + lane) * 4;
          const int* ready = p.sk_flags + ((my_pair - 1) * 2 + (int)rank) * 2;
          int seen;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ready) : "memory");
          } while (seen < 8);
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          if (sk_out) {                               // head part of a cut tile: park the raw accumulator, nothing else
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<uint4*>(sk_out + (size_t)(c + i) * BM) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            continue;
          }
          if (sk_in) {
            float4 part[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) part[i] = *reinterpret_cast<const float4*>(sk_in + (size_t)(c + 4 * i) * BM);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i + 0] = __float_as_uint(__uint_as_float(v[4 * i + 0]) + part[i].x);
              v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + part[i].y);
              v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + part[i].z);
              v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + part[i].w);
            }
          }
          const int64_t n = n0 + c;
          if (n >= p.N) continue;
          if (!p.c_f32 && n + 32 <= p.N && (p.ldc & 7) == 0 && (!rrow || (p.ldr & 7) == 0)) {
            // warp-uniform fast path: full 32-column bf16 chunk, written as whole 64-byte row segments
            uint32_t w[16];
            bf16* cp = reinterpret_cast<bf16*>(p.C) + crow + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f8[i] = __uint_as_float(v[q * 8 + i]) * p.alpha;
              if (rrow && row_ok) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(rrow + n + q * 8);
                f8[0] += bf16_lo(r4.x); f8[1] += bf16_hi(r4.x); f8[2] += bf16_lo(r4.y); f8[3] += bf16_hi(r4.y);
                f8[4] += bf16_lo(r4.z); f8[5] += bf16_hi(r4.z); f8[6] += bf16_lo(r4.w); f8[7] += bf16_hi(r4.w);
              }
              if (p.accumulate && row_ok) {
                const uint4 old = *reinterpret_cast<const uint4*>(cp + q * 8);
                f8[0] += bf16_lo(old.x); f8[1] += bf16_hi(old.x); f8[2] += bf16_lo(old.y); f8[3] += bf16_hi(old.y);
                f8[4] += bf16_lo(old.z); f8[5] += bf16_hi(old.z); f8[6] += bf16_lo(old.w); f8[7] += bf16_hi(old.w);
              }
              w[q * 4 + 0] = pack_bf16(f8[0], f8[1]); w[q * 4 + 1] = pack_bf16(f8[2], f8[3]);
              w[q * 4 + 2] = pack_bf16(f8[4], f8[5]); w[q * 4 + 3] = pack_bf16(f8[6], f8[7]);
              if (p.rope_cache && n + q * 8 < p.rope_cols && row_ok)
                rope_rotate8(w + q * 4, p.rope_cache + (int64_t)(p.rope_pos ? p.rope_pos[m] : m % p.rope_seq) * p.rope_hd,
                             (int)((n + q * 8) % p.rope_hd) >> 1, 1.f);
            }
            store_tile_32x32(stage_s, reinterpret_cast<bf16*>(p.C) + (int64_t)g * p.c_group_stride + wm0 * p.ldc + n,
                             p.ldc, lane, w, wrows);
            continue;
          }
          if (p.c_f32 && !p.accumulate && n + 32 <= p.N && (p.ldc & 3) == 0 &&
              (!rrow || (p.ldr & (p.r_f32 ? 3 : 7)) == 0)) {
            // warp-uniform fp32 output (the fp32 residual stream: o-proj / down-proj): alpha, residual (fp32 or bf16),
            // then two 32 x 16 fp32 half tiles through the staging tile (64 B per row, like the bf16 32 x 32 tile)
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
            if (rrow && row_ok) {
              if (p.r_f32) {
                const float* r32 = reinterpret_cast<const float*>(p.R) + (int64_t)g * p.r_group_stride + m * p.ldr + n;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 r4 = *reinterpret_cast<const float4*>(r32 + q * 4);
                  f[q * 4 + 0] += r4.x; f[q * 4 + 1] += r4.y; f[q * 4 + 2] += r4.z; f[q * 4 + 3] += r4.w;
                }
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 r4 = *reinterpret_cast<const uint4*>(rrow + n + q * 8);
                  f[q * 8 + 0] += bf16_lo(r4.x); f[q * 8 + 1] += bf16_hi(r4.x);
                  f[q * 8 + 2] += bf16_lo(r4.y); f[q * 8 + 3] += bf16_hi(r4.y);
                  f[q * 8 + 4] += bf16_lo(r4.z); f[q * 8 + 5] += bf16_hi(r4.z);
                  f[q * 8 + 6] += bf16_lo(r4.w); f[q * 8 + 7] += bf16_hi(r4.w);
                }
              }
            }
            float* c32 = reinterpret_cast<float*>(p.C) + (int64_t)g * p.c_group_stride + wm0 * p.ldc + n;
            store_tile_32x32(stage_s, reinterpret_cast<bf16*>(c32), 2 * p.ldc, lane,
                             reinterpret_cast<const uint32_t*>(f), wrows);
            store_tile_32x32(stage_s, reinterpret_cast<bf16*>(c32 + 16), 2 * p.ldc, lane,
                             reinterpret_cast<const uint32_t*>(f) + 16, wrows);
            continue;
          }
          if (!row_ok) continue;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
          const bool full_vec = (n + 32 <= p.N);
          if (rrow && p.r_f32) {
            const float* r32 = reinterpret_cast<const float*>(p.R) + (int64_t)g * p.r_group_stride + m * p.ldr + n;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n + i < p.N) f[i] += r32[i];
          } else if (rrow) {
            if (full_vec && (p.ldr % 8 == 0)) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(rrow + n + q * 8);
                f[q * 8 + 0] += bf16_lo(r4.x); f[q * 8 + 1] += bf16_hi(r4.x);
                f[q * 8 + 2] += bf16_lo(r4.y); f[q * 8 + 3] += bf16_hi(r4.y);
                f[q * 8 + 4] += bf16_lo(r4.z); f[q * 8 + 5] += bf16_hi(r4.z);
                f[q * 8 + 6] += bf16_lo(r4.w); f[q * 8 + 7] += bf16_hi(r4.w);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.N) f[i] += __bfloat162float(rrow[n + i]);
            }
          }
          if (p.c_f32) {
            float* cp = reinterpret_cast<float*>(p.C) + crow + n;
            if (full_vec && (p.ldc % 4 == 0)) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 o = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
                if (p.accumulate) {
                  const float4 old = *reinterpret_cast<const float4*>(cp + q * 4);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(cp + q * 4) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.N) cp[i] = p.accumulate ? cp[i] + f[i] : f[i];
            }
          } else {
            bf16* cp = reinterpret_cast<bf16*>(p.C) + crow + n;
            if (full_vec && (p.ldc % 8 == 0)) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (p.accumulate) {
                  const uint4 old = *reinterpret_cast<const uint4*>(cp + q * 8);
                  f[q * 8 + 0] += bf16_lo(old.x); f[q * 8 + 1] += bf16_hi(old.x);
                  f[q * 8 + 2] += bf16_lo(old.y); f[q * 8 + 3] += bf16_hi(old.y);
                  f[q * 8 + 4] += bf16_lo(old.z); f[q * 8 + 5] += bf16_hi(old.z);
                  f[q * 8 + 6] += bf16_lo(old.w); f[q * 8 + 7] += bf16_hi(old.w);
                }
                uint4 o;
                o.x = pack_bf16(f[q * 8 + 0], f[q * 8 + 1]); o.y = pack_bf16(f[q * 8 + 2], f[q * 8 + 3]);
                o.z = pack_bf16(f[q * 8 + 4], f[q * 8 + 5]); o.w = pack_bf16(f[q * 8 + 6], f[q * 8 + 7]);
                *reinterpret_cast<uint4*>(cp + q * 8) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.N) {
                  float o = f[i];
                  if (p.accumulate) o += __bfloat162float(cp[i]);
                  cp[i] = __float2bfloat16_rn(o);
                }
            }
          }
        }
        if (sk_out) {                                 // publish the partial: 8 warp arrivals per CTA
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(p.sk_flags + (my_pair * 2 + (int)rank) * 2, 1);
        }
        if (sk_in) {                                  // the last of the 8 reader warps re-arms the flags
          __syncwarp();
          if (lane == 0) {
            int* fl = p.sk_flags + ((my_pair - 1) * 2 + (int)rank) * 2;
            if (atomicAdd(fl + 1, 1) == 7) {
              fl[1] = 0;
              __threadfence();
              asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(fl), "r"(0) : "memory");
            }
          }
        }
      } else if (kEpi == EPI_SWIGLU_FWD) {
        bf16* gu0 = reinterpret_cast<bf16*>(p.C) + wm0 * p.ldc;      // the warp's first row
        bf16* act0 = reinterpret_cast<bf16*>(p.C2) + wm0 * p.ldc2;
        const int64_t c0 = (int64_t)nb * (BN / 2);
#pragma unroll 1
        for (int c = c_lo / 2; c < c_hi / 2; c += 32) {
          uint32_t vg[32], vu[32];
          __syncwarp();
          tmem_ld32(t_addr + c, vg);
          tmem_ld32(t_addr + BN / 2 + c, vu);
          tmem_ld_wait();
          const int64_t n = c0 + c;
          if (n >= p.inter) continue;                  // inter is a multiple of 128: chunks are all-or-nothing
          uint32_t pg[16], pu[16], pa[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            // the linear outputs are rounded to bf16 first (what the unfused path stores and re-reads)
            pg[e] = pack_bf16(__uint_as_float(vg[2 * e]), __uint_as_float(vg[2 * e + 1]));
            pu[e] = pack_bf16(__uint_as_float(vu[2 * e]), __uint_as_float(vu[2 * e + 1]));
            const float g0 = bf16_lo(pg[e]), g1 = bf16_hi(pg[e]);
            const uint32_t ps = pack_bf16(g0 * fast_sigmoid(g0), g1 * fast_sigmoid(g1));   // bf16(silu(gate))
            pa[e] = pack_bf16(bf16_lo(ps) * bf16_lo(pu[e]), bf16_hi(ps) * bf16_hi(pu[e]));
          }
          store_tile_32x32(stage_s, gu0 + n, p.ldc, lane, pg, wrows);
          store_tile_32x32(stage_s, gu0 + p.inter + n, p.ldc, lane, pu, wrows);
          store_tile_32x32(stage_s, act0 + n, p.ldc2, lane, pa, wrows);
        }
      } else if (kEpi == EPI_SWIGLU_BWD) {
        const bf16* g0p = p.R + wm0 * p.ldr;                          // gate|up, the warp's first row
        bf16* d0 = reinterpret_cast<bf16*>(p.C) + wm0 * p.ldc;        // dgate|dup, the warp's first row
        // gate/up of the NEXT 32-column chunk are fetched (whole 64-byte row segments, four rows x two tensors in flight
        // per lane) while this one is computed; they reach the row-per-lane layout through the staging tile
        uint4 gn[4], un[4];
        auto fetch = [&](int c) {
          const int64_t n = n0 + c;
          const bool chunk_ok = n < p.N && c < c_hi;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2);
            const bf16* src = g0p + (int64_t)row * p.ldr + n + (lane & 3) * 8;
            const bool ok = chunk_ok && row < wrows;
            gn[i] = ok ? ld_nc16(src) : make_uint4(0, 0, 0, 0);
            un[i] = ok ? ld_nc16(src + p.inter) : make_uint4(0, 0, 0, 0);
          }
        };
        fetch(c_lo);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          const int64_t n = n0 + c;
          uint32_t gw[16], uw[16];
          rows_from_coalesced(stage_s, lane, gn, gw);
          rows_from_coalesced(stage_s, lane, un, uw);
          fetch(c + 32);
          if (n >= p.N) continue;                      // inter is a multiple of 128: chunks are all-or-nothing
          uint32_t wdg[16], wdu[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            // dact rounded to bf16 first (what the unfused path stores and re-reads)
            const uint32_t dp = pack_bf16(__uint_as_float(v[2 * e]) * p.alpha, __uint_as_float(v[2 * e + 1]) * p.alpha);
            const float d0f = bf16_lo(dp), d1f = bf16_hi(dp);
            const float g0 = bf16_lo(gw[e]), g1 = bf16_hi(gw[e]);
            const float s0 = fast_sigmoid(g0), s1 = fast_sigmoid(g1);
            wdu[e] = pack_bf16(d0f * g0 * s0, d1f * g1 * s1);
            wdg[e] = pack_bf16(d0f * bf16_lo(uw[e]) * s0 * fmaf(g0, 1.f - s0, 1.f),
                               d1f * bf16_hi(uw[e]) * s1 * fmaf(g1, 1.f - s1, 1.f));
          }
          store_tile_32x32(stage_s, d0 + n, p.ldc, lane, wdg, wrows);
          store_tile_32x32(stage_s, d0 + p.inter + n, p.ldc, lane, wdu, wrows);
        }
      } else if (kEpi == EPI_CE_PARTIAL) {
        // online softmax over this tile's columns; the row is thread-local (lane == TMEM lane == row)
        int64_t tgt = -1;
        if (row_ok) tgt = p.targets[(int64_t)g * p.tgt_group_stride + m * p.tgt_row_stride];
        float mx = -INFINITY, se = 0.f, tl = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          const int64_t n = n0 + c;
          if (n >= p.N) continue;
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = (n + i < p.N) ? __uint_as_float(v[i]) : -INFINITY;
            v[i] = __float_as_uint(x);
            cm = fmaxf(cm, x);
            if (n + i == tgt) tl = x;
          }
          const float nm = fmaxf(mx, cm);
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) s += __expf(__uint_as_float(v[i]) - nm);
          se = se * __expf(mx - nm) + s;
          mx = nm;
        }
        if (row_ok) p.ce_part[((int64_t)g * p.M + m) * p.num_n + nb] = make_float4(mx, se, tl, 0.f);
      } else {  // EPI_CE_DLOGITS: bf16 gscale * (softmax - onehot), zero beyond N up to ldc
        int64_t tgt = -1;
        float L = 0.f;
        float gs = p.gscale;
        if (p.gscale_dev) gs *= *p.gscale_dev;
        if (row_ok) {
          tgt = p.targets[(int64_t)g * p.tgt_group_stride + m * p.tgt_row_stride];
          L = p.lse[(int64_t)g * p.M + m];
          if (tgt < 0 || tgt >= p.N) gs = 0.f;  // ignored row
        }
        bf16* crow = reinterpret_cast<bf16*>(p.C) + (int64_t)g * p.c_group_stride + m * p.ldc;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          const int64_t n = n0 + c;
          if (!row_ok || n >= p.ldc) continue;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pr = (n + i < p.N) ? __expf(__uint_as_float(v[i]) - L) : 0.f;
            f[i] = gs * (pr - ((n + i == tgt) ? 1.f : 0.f));
          }
          if (n + 32 <= p.ldc) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_bf16(f[q * 8 + 0], f[q * 8 + 1]); o.y = pack_bf16(f[q * 8 + 2], f[q * 8 + 3]);
              o.z = pack_bf16(f[q * 8 + 4], f[q * 8 + 5]); o.w = pack_bf16(f[q * 8 + 6], f[q * 8 + 7]);
              *reinterpret_cast<uint4*>(crow + n + q * 8) = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n + i < p.ldc) crow[n + i] = __float2bfloat16_rn(f[i]);
          }
        }
      }
      // this warp has drained its quadrant of the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta2 && rank != 0) mbar_arrive_cluster(mapa(smem_u32(&tempty[acc]), 0));
        else mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCta2) cluster_sync();      // no CTA of the pair exits (or frees TMEM) while the other may still touch it
  if (warp == 1) {
    tc_fence_after();
    if (kCta2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t groups,
                     uint64_t row_stride_elems, uint64_t group_stride_elems, uint32_t box_rows) {
  EncodeFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return CSM_ERR_CUDA; }
  cuuint64_t dims[3] = {inner, rows, groups};
  cuuint64_t strides[2] = {row_stride_elems * 2, group_stride_elems * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu groups=%llu row_stride=%llu group_stride=%llu",
              (int)r, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)groups,
              (unsigned long long)row_stride_elems, (unsigned long long)group_stride_elems);
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}

struct GemmTcOperands {
  const void* A; const void* B; const void* A2; const void* B2;
  int64_t lda, ldb, lda2, ldb2, K2;
  int64_t a_group_stride, b_group_stride;
  int transA, transB;
};

// Stream-K scratch, registered once per process by the host side (csm_gemm_set_streamk_workspace): flags first, then
// the fp32 partial tiles.  One buffer per device, so stream-K GEMMs must not run concurrently on two streams — on this
// path every GEMM is issued on the step's main stream (only NCCL uses the side stream).
static float* g_sk_ws = nullptr;
static int* g_sk_flags = nullptr;
// Measured (tools/bench_streamk.py): no gain on B200 — 4096x2048x8192: 94.2 us whole tiles vs 96.3 us stream-K;
// x16384: 176.9 vs 174.8; x3072: 60.4 vs 66.5.  The GPU is power-capped, so the pairs that sit out the last partial
// wave are not lost throughput (the busy ones clock higher).  Correct and tested, therefore kept, but OFF by default.
static std::atomic<int> g_sk_mode{0};    // 0 off (default), 1 cut tiles when the last wave is badly filled
static std::atomic<int> g_dyn_mode{0};   // 1: tiles after a CTA's first are drawn from a global counter; 0 (default): static
                                         // round-robin (~1 us per launch cheaper when the GEMM has the GPU to itself)
constexpr int kTileCtrOffset = 512;      // ints into the flag area (the stream-K flags use the first 320)
constexpr int kTileCtrSlots = 32;
void gemm_tc_set_dynamic_tiles(int m) { g_dyn_mode.store(m); }
constexpr size_t kSkFlagBytes = 4096;           // stream-K flags + tile counters
constexpr int kSkMaxPairs = 80;
size_t gemm_tc_streamk_workspace_bytes() { return kSkFlagBytes + (size_t)kSkMaxPairs * 2 * 256 * BM * sizeof(float); }
void gemm_tc_set_streamk_workspace(void* ptr, size_t bytes) {
  if (ptr && bytes >= gemm_tc_streamk_workspace_bytes()) {
    g_sk_flags = reinterpret_cast<int*>(ptr);
    g_sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ptr) + kSkFlagBytes);
  } else {
    g_sk_flags = nullptr; g_sk_ws = nullptr;
  }
}
void gemm_tc_set_streamk_mode(int m) { g_sk_mode.store(m); }
int gemm_tc_experiments_compiled() { return kExperiments ? 1 : 0; }
static std::atomic<int> g_tail_mode{0};  // 0 off (default), 1 narrow MMAs on ragged last column tiles (bit-identical, no gain measured)
void gemm_tc_set_narrow_tail_mode(int m) { g_tail_mode.store(m); }

// how many CTA pairs can be co-resident (a pair needs two SMs of one TPC); every pair instantiation has the same
// block size and shared-memory footprint, so one query serves all
static int pair_capacity();

template <bool TA, bool TB, int BN, int EPI, bool CTA2>
static int launch_inst(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2,
                       const GemmTcParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CTA2>;
  auto kern = gemm_tc_kernel<TA, TB, BN, EPI, CTA2>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
    configured = true;
  }
  const int max_pairs = CTA2 ? pair_capacity() : 0;
  if (!CTA2) {
    const int tiles = p.groups * p.num_m * p.num_n;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    cudaError_t e = launch_k(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, st, 1, a, b, a2, b2, p);
    if (e != cudaSuccess) { set_error("gemm_tc: launch failed: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
  } else {
    const int tiles = p.groups * ((p.num_m + 1) / 2) * p.num_n;
    const int avail = max_pairs < num_sms() / 2 ? max_pairs : num_sms() / 2;   // num_sms() honours reserved SMs
    const int pairs = tiles < avail ? tiles : avail;
    cudaError_t e = launch_k(kern, dim3(2 * pairs), dim3(kGemmThreads), Cfg::kSmemBytes, st, 2, a, b, a2, b2, p);
    if (e != cudaSuccess) { set_error("gemm_tc (CTA pair): launch failed: %s", cudaGetErrorString(e)); return CSM_ERR_CUDA; }
  }
  CSM_CHECK_LAUNCH("gemm_tc");
  return CSM_OK;
}

static int pair_capacity() {
  static int cap = 0;
  if (cap == 0) {
    // the persistent grid must not exceed the co-resident cluster count, or a late pair would start only after an
    // early one has finished its whole share (and a stream-K finisher could wait for a pair that is not running)
    using Cfg = GemmCfg<256, true>;
    auto kern = gemm_tc_kernel<false, false, 256, EPI_STORE, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(148); q.blockDim = dim3(kGemmThreads); q.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at; q.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &q);
    if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = 70; }
    cap = n;
  }
  return cap;
}

template <int BN, int EPI, bool CTA2>
static int launch_major(const GemmTcOperands& o, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2,
                        const CUtensorMap& b2, const GemmTcParams& p, cudaStream_t st) {
  if (!o.transA && !o.transB) return launch_inst<false, false, BN, EPI, CTA2>(a, b, a2, b2, p, st);
  if (!o.transA && o.transB) return launch_inst<false, true, BN, EPI, CTA2>(a, b, a2, b2, p, st);
  if (o.transA && !o.transB) return launch_inst<true, false, BN, EPI, CTA2>(a, b, a2, b2, p, st);
  return launch_inst<true, true, BN, EPI, CTA2>(a, b, a2, b2, p, st);
}

static std::atomic<int> g_cta2_mode{-1};   // -1 auto, 0 never, 1 whenever legal (test hook)
void gemm_tc_set_cta_pair_mode(int m) { g_cta2_mode.store(m); }

// BN = 256 unless that leaves SMs idle (or N is narrow) and 128 would fill them better
static int choose_bn(int groups, int64_t M, int64_t N) {
  const int64_t t256 = (int64_t)groups * ((M + BM - 1) / BM) * ((N + 255) / 256);
  return (t256 < num_sms() || N <= 128) ? 128 : 256;
}

// Tile width and CTA-pair decision for one launch (also used by the CE drivers, whose combine kernel must know how
// many column tiles the partials have).
static int gemm_tc_tiling(int epi, int groups, int64_t M, int64_t N, int transA, int transB, bool* pair) {
  const int mode = g_cta2_mode.load();
  const int num_m = (int)((M + BM - 1) / BM);
  // CTA-pair mode (256 x 256 super-tiles) for the GEMMs that can keep at least ~2/3 of the 74 pairs busy; the test
  // hook (mode 1) turns it on for every shape with two m-blocks and more than one 128-column tile
  const int64_t super256 = (int64_t)groups * ((num_m + 1) / 2) * ((N + 255) / 256);
  // (the fused-CE epilogues take K-major operands only: activations [M, K] and the [V, K] head)
  const bool ce_pair_ok = (epi == EPI_CE_PARTIAL || epi == EPI_CE_DLOGITS) && !transA && !transB;
  bool cta2 = (epi == EPI_STORE || epi == EPI_SWIGLU_BWD || ce_pair_ok) && num_m >= 2 && N > 128 && mode != 0 &&
              (mode == 1 || super256 >= 48);
  if (epi == EPI_SWIGLU_FWD) cta2 = true;                 // callers check gemm_tc_swiglu_supported()
  if (pair) *pair = cta2;
  return (cta2 || epi == EPI_SWIGLU_BWD) ? 256 : choose_bn(groups, M, N);
}

// Shared driver for the plain GEMM and the CE epilogues.
int gemm_tc_run(const GemmTcOperands& o, GemmTcParams p, int epi, cudaStream_t st) {
  p.num_m = (int)((p.M + BM - 1) / BM);
  bool cta2 = false;
  const int bn = gemm_tc_tiling(epi, p.groups, p.M, p.N, o.transA, o.transB, &cta2);
  if (kExperiments && cta2 && epi == EPI_STORE && p.groups == 1 && g_sk_ws && g_sk_mode.load() == 1) {
    // stream-K when the tile count leaves the last wave badly filled (e.g. 128 super-tiles on 74 pairs = 1.73 waves)
    const int cap = pair_capacity();
    const int P = cap < num_sms() / 2 ? cap : num_sms() / 2;
    const int64_t T = (int64_t)((p.num_m + 1) / 2) * ((p.N + 255) / 256);
    const int64_t kb = (p.K + BK - 1) / BK + ((o.A2 && o.K2 > 0) ? (o.K2 + BK - 1) / BK : 0);
    const int64_t waves = (T + P - 1) / P;
    if (T > P && P <= kSkMaxPairs && kb >= 8 && (double)T / (double)(waves * P) < 0.93) {
      p.streamk = 1; p.sk_ws = g_sk_ws; p.sk_flags = g_sk_flags;
    }
  }
  p.narrow_tail = (kExperiments && g_tail_mode.load() == 1 && !p.streamk) ? 1 : 0;
  // dynamic tile scheduler: one of 32 self-resetting counter pairs in the registered scratch (launches on one stream are
  // serialised, the rotation only separates launches that might overlap on different streams)
  p.tile_ctr = nullptr;
  if (g_sk_flags && !p.streamk && g_dyn_mode.load() != 0) {
    static std::atomic<unsigned> seq{0};
    p.tile_ctr = g_sk_flags + kTileCtrOffset + 2 * (int)(seq.fetch_add(1) % kTileCtrSlots);
  }
  const uint32_t b_box = cta2 ? (uint32_t)bn / 2 : (uint32_t)bn;
  p.num_n = epi == EPI_SWIGLU_FWD ? (int)((p.N + 127) / 128) : (int)((p.N + bn - 1) / bn);
  const uint64_t b_rows = epi == EPI_SWIGLU_FWD ? 2 * (uint64_t)p.N : (uint64_t)p.N;   // packed gate|up weight
  p.k_blocks = (int)((p.K + BK - 1) / BK);
  p.has_tail = (o.A2 && o.K2 > 0) ? (int)((o.K2 + BK - 1) / BK) : 0;
  CUtensorMap ta, tb, ta2, tb2;
  int rc;
  const uint64_t G = (uint64_t)p.groups;
  const uint64_t ags = p.groups > 1 ? (uint64_t)o.a_group_stride : (uint64_t)8;
  const uint64_t bgs = p.groups > 1 ? (uint64_t)o.b_group_stride : (uint64_t)8;
  // K-major operand: {K, rows}; MN-major: {rows, K}
  rc = o.transA ? encode_tmap_bf16(&ta, o.A, p.M, p.K, G, o.lda, ags, 64)
                : encode_tmap_bf16(&ta, o.A, p.K, p.M, G, o.lda, ags, BM);
  if (rc) return rc;
  rc = o.transB ? encode_tmap_bf16(&tb, o.B, p.N, p.K, G, o.ldb, bgs, 64)
                : encode_tmap_bf16(&tb, o.B, p.K, b_rows, G, o.ldb, bgs, b_box);
  if (rc) return rc;
  if (p.has_tail) {
    rc = o.transA ? encode_tmap_bf16(&ta2, o.A2, p.M, o.K2, 1, o.lda2, 8, 64)
                  : encode_tmap_bf16(&ta2, o.A2, o.K2, p.M, 1, o.lda2, 8, BM);
    if (rc) return rc;
    rc = o.transB ? encode_tmap_bf16(&tb2, o.B2, p.N, o.K2, 1, o.ldb2, 8, 64)
                  : encode_tmap_bf16(&tb2, o.B2, o.K2, b_rows, 1, o.ldb2, 8, b_box);
    if (rc) return rc;
  } else {
    ta2 = ta; tb2 = tb;
  }
#define DISPATCH(BN_, EPI_) return launch_major<BN_, EPI_, false>(o, ta, tb, ta2, tb2, p, st)
  if (epi == EPI_SWIGLU_FWD) return launch_inst<false, false, 256, EPI_SWIGLU_FWD, true>(ta, tb, ta2, tb2, p, st);
  if (epi == EPI_SWIGLU_BWD) {
    if (cta2) return launch_inst<false, true, 256, EPI_SWIGLU_BWD, true>(ta, tb, ta2, tb2, p, st);
    return launch_inst<false, true, 256, EPI_SWIGLU_BWD, false>(ta, tb, ta2, tb2, p, st);
  }
  if (cta2 && epi == EPI_CE_PARTIAL) return launch_inst<false, false, 256, EPI_CE_PARTIAL, true>(ta, tb, ta2, tb2, p, st);
  if (cta2 && epi == EPI_CE_DLOGITS) return launch_inst<false, false, 256, EPI_CE_DLOGITS, true>(ta, tb, ta2, tb2, p, st);
  if (cta2) return launch_major<256, EPI_STORE, true>(o, ta, tb, ta2, tb2, p, st);
  if (epi == EPI_STORE) { if (bn == 256) DISPATCH(256, EPI_STORE); else DISPATCH(128, EPI_STORE); }
  if (epi == EPI_CE_PARTIAL) { if (bn == 256) DISPATCH(256, EPI_CE_PARTIAL); else DISPATCH(128, EPI_CE_PARTIAL); }
  if (bn == 256) DISPATCH(256, EPI_CE_DLOGITS); else DISPATCH(128, EPI_CE_DLOGITS);
#undef DISPATCH
}

static bool tma_ok(const void* p, int64_t ld) { return aligned16(p) && ld > 0 && (ld % 8) == 0; }

bool gemm_tc_supported(const void* A, const void* B, const void* C, const void* R, int64_t M, int64_t N, int64_t K,
                       int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                       const void* A2, const void* B2, int64_t K2, int64_t lda2, int64_t ldb2) {
  (void)C; (void)R; (void)ldc; (void)ldr; (void)c_dtype; (void)transA; (void)transB;
  if (M < 1 || N < 1 || K < 16) return false;
  // tiny problems (the tiny test model) are the scalar kernel's job; a small output with a long reduction
  // (LoRA dA / dB: [r x in] = dt^T x over thousands of rows) still belongs on the tensor cores
  if (K < 64 || (double)M * (double)N * (double)K < (double)(1 << 21)) return false;
  if (!tma_ok(A, lda) || !tma_ok(B, ldb)) return false;
  if (A2 && (!tma_ok(A2, lda2) || !tma_ok(B2, ldb2) || K2 < 1 || K2 > 256)) return false;
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return false;
  return true;
}

int gemm_tc_launch_rope(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                        int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                        int accumulate, float alpha, const void* A2, const void* B2, int64_t K2, int64_t lda2,
                        int64_t ldb2, const float* rope_cache, int rope_seq, int rope_cols, int rope_hd,
                        const int32_t* rope_pos, cudaStream_t stream) {
  GemmTcOperands o{A, B, A2, B2, lda, ldb, lda2, ldb2, K2, 0, 0, transA, transB};
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = K; p.groups = 1;
  p.C = C; p.R = (const bf16*)R; p.ldc = ldc; p.ldr = ldr;
  p.c_f32 = (c_dtype & CSM_DT_F32) != 0; p.r_f32 = (c_dtype & CSM_DT_RES_F32) != 0;
  p.accumulate = accumulate; p.alpha = alpha;
  p.rope_cache = rope_cache; p.rope_seq = rope_seq; p.rope_cols = rope_cols; p.rope_hd = rope_hd;
  p.rope_pos = rope_pos;
  return gemm_tc_run(o, p, EPI_STORE, stream);
}

int gemm_tc_launch(const void* A, const void* B, void* C, const void* R, int64_t M, int64_t N, int64_t K,
                   int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int transA, int transB, int c_dtype,
                   int accumulate, float alpha, const void* A2, const void* B2, int64_t K2, int64_t lda2,
                   int64_t ldb2, cudaStream_t stream) {
  return gemm_tc_launch_rope(A, B, C, R, M, N, K, lda, ldb, ldc, ldr, transA, transB, c_dtype, accumulate, alpha, A2, B2,
                             K2, lda2, ldb2, nullptr, 0, 0, 0, nullptr, stream);
}

// ------------------------------------------------------------------------------------------- fused SwiGLU MLP
// forward : gate|up = x W13^T (+ LoRA tail) and act = silu(gate) * up in one launch (CTA-pair tiles of 128 act columns)
// backward: dgate|dup from dact = dy W2 (+ LoRA tail) in the dgrad GEMM's epilogue
bool gemm_tc_swiglu_supported(int64_t M, int64_t inter, int64_t K) {
  if (g_cta2_mode.load() == 0) return false;
  if (M < 2 * BM || inter < 256 || inter % 128 != 0 || K < 64 || K % 8 != 0) return false;
  if (g_cta2_mode.load() == 1) return true;                         // test hook: every legal shape
  return ((M + 2 * BM - 1) / (2 * BM)) * (inter / 128) >= 48;       // enough super-tiles for the 74 CTA pairs
}

int gemm_tc_swiglu_fwd(const void* X, const void* W13, void* GU, void* ACT, int64_t M, int64_t inter, int64_t K,
                       int64_t ldx, int64_t ldw, int64_t ldgu, int64_t ldact, const void* A2, const void* B2, int64_t K2,
                       int64_t lda2, int64_t ldb2, cudaStream_t st) {
  GemmTcOperands o{X, W13, A2, B2, ldx, ldw, lda2, ldb2, K2, 0, 0, 0, 0};
  GemmTcParams p{};
  p.M = M; p.N = inter; p.K = K; p.groups = 1;
  p.C = GU; p.ldc = ldgu; p.C2 = ACT; p.ldc2 = ldact; p.inter = inter; p.alpha = 1.f;
  return gemm_tc_run(o, p, EPI_SWIGLU_FWD, st);
}

int gemm_tc_swiglu_bwd(const void* DY, const void* W2, const void* GU, void* DGU, int64_t M, int64_t inter, int64_t K,
                       int64_t lddy, int64_t ldw, int64_t ldgu, int64_t lddgu, const void* A2, const void* B2,
                       int64_t K2, int64_t lda2, int64_t ldb2, cudaStream_t st) {
  // dact[M, inter] = dy[M, K] @ W2[K, inter]  (W2 is nn.Linear [out=K, in=inter]: the MN-major B operand)
  GemmTcOperands o{DY, W2, A2, B2, lddy, ldw, lda2, ldb2, K2, 0, 0, 0, 1};
  GemmTcParams p{};
  p.M = M; p.N = inter; p.K = K; p.groups = 1;
  p.C = DGU; p.ldc = lddgu; p.R = (const bf16*)GU; p.ldr = ldgu; p.inter = inter; p.alpha = 1.f;
  return gemm_tc_run(o, p, EPI_SWIGLU_BWD, st);
}

// ------------------------------------------------------------------------------------------- fused CE
int ce_combine_launch(const void* part, int nt, float* loss, float* lse, int64_t rows, cudaStream_t st);

bool linear_ce_tc_supported(int64_t M, int64_t V, int64_t K, int64_t ldh, int64_t ldw, int transW, const void* H,
                            const void* W) {
  if (transW) return false;  // [K,V] rows of V=2051 bf16 are not 16-byte aligned: callers pass the [V,K] shadow
  if (M < 1 || V < 64 || K < 64) return false;
  return tma_ok(H, ldh) && tma_ok(W, ldw);
}

size_t linear_ce_tc_workspace(int64_t M, int64_t V, int groups) {
  const int64_t nt = (V + 127) / 128;  // upper bound on N tiles (BN >= 128)
  return (size_t)groups * M * nt * sizeof(float4) + 256;
}

int linear_ce_tc_fwd(const void* H, const void* W, const int64_t* targets, float* loss_rows, float* lse, int64_t M,
                     int64_t V, int64_t K, int groups, int64_t ldh, int64_t hgs, int64_t ldw, int64_t wgs,
                     int transW, int64_t trs, int64_t tgs, void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)transW; (void)ws_bytes;
  GemmTcOperands o{H, W, nullptr, nullptr, ldh, ldw, 0, 0, 0, hgs, wgs, 0, 0};
  GemmTcParams p{};
  p.M = M; p.N = V; p.K = K; p.groups = groups;
  p.alpha = 1.f;
  p.targets = targets; p.tgt_row_stride = trs; p.tgt_group_stride = tgs;
  p.ce_part = reinterpret_cast<float4*>(ws);
  // (Merging the partials in the epilogue of each row group's last column tile — arrival counters, no second launch —
  // was built and measured in round 2: bit-identical, but SLOWER: 60.4 vs 50.2 us at N_sel = 232, 156.6 vs 107.5 us at
  // 1024; the fence + atomic per tile and the serialised merge stall the epilogue warps that gate the next tile's
  // MMAs.  The separate 5 us combine kernel stays.)
  int rc = gemm_tc_run(o, p, EPI_CE_PARTIAL, st);
  if (rc) return rc;
  const int bn = gemm_tc_tiling(EPI_CE_PARTIAL, groups, M, V, 0, 0, nullptr);  // the choice gemm_tc_run made
  const int nt = (int)((V + bn - 1) / bn);
  return ce_combine_launch(ws, nt, loss_rows, lse, (int64_t)groups * M, st);
}

int linear_ce_tc_bwd_dlogits(const void* H, const void* W, const int64_t* targets, const float* lse,
                             float grad_scale, const float* grad_scale_dev, void* dlogits, int64_t ldd, int64_t M,
                             int64_t V, int64_t K, int groups, int64_t ldh, int64_t hgs, int64_t ldw, int64_t wgs,
                             int64_t trs, int64_t tgs, cudaStream_t st) {
  // all heads in ONE launch: dlogits [groups, M, ldd]
  GemmTcOperands o{H, W, nullptr, nullptr, ldh, ldw, 0, 0, 0, hgs, wgs, 0, 0};
  GemmTcParams p{};
  p.M = M; p.N = V; p.K = K; p.groups = groups;
  p.C = dlogits; p.ldc = ldd; p.c_group_stride = M * ldd;
  p.targets = targets; p.tgt_row_stride = trs; p.tgt_group_stride = tgs;
  p.lse = lse; p.gscale = grad_scale; p.gscale_dev = grad_scale_dev;
  return gemm_tc_run(o, p, EPI_CE_DLOGITS, st);
}

// ------------------------------------------------------------------------------------------- split-reduction GEMM
// Skinny products (LoRA: t = x A^T [M x 16], dA = dt^T x [16 x in], ...) have a handful of output tiles and a long
// reduction: 16-32 CTAs each streaming its operand at one SM's share of HBM (~1.3 TB/s in total).  Splitting the
// reduction into `splits` groups (the kernel's `groups` dimension: group g covers K range [g*K/splits, (g+1)*K/splits))
// puts 96-128 CTAs on the machine; the fp32 partials [splits, M, N] are summed by a small kernel.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, bf16* __restrict__ out, int64_t M, int64_t N, int64_t ldc,
                     int splits, float alpha) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int g = 0; g < splits; ++g) acc += part[(int64_t)g * total + i];
    out[(i / N) * ldc + (i % N)] = __float2bfloat16_rn(acc * alpha);
  }
}

bool gemm_tc_splitk_supported(int64_t M, int64_t N, int64_t K, int splits) {
  return splits >= 2 && splits <= 16 && K % splits == 0 && (K / splits) % 64 == 0 && M >= 1 && N >= 1;
}

int gemm_tc_splitk(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                   int64_t ldc, int transA, int transB, float alpha, int splits, float* ws, cudaStream_t st) {
  const int64_t Kg = K / splits;
  // K-major operand: the group advances along the inner (K) dimension; MN-major: along the rows (K) dimension
  GemmTcOperands o{A, B, nullptr, nullptr, lda, ldb, 0, 0, 0, transA ? Kg * lda : Kg, transB ? Kg * ldb : Kg, transA,
                   transB};
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = Kg; p.groups = splits;
  p.C = ws; p.ldc = N; p.c_group_stride = M * N;
  p.c_f32 = 1; p.accumulate = 0; p.alpha = 1.f;
  int rc = gemm_tc_run(o, p, EPI_STORE, st);
  if (rc) return rc;
  const int64_t total = M * N;
  const unsigned grid = (unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  if (launch_k(splitk_reduce_kernel, dim3(grid), dim3(256), 0, st, 1, (const float*)ws, (bf16*)C, M, N, ldc, splits,
               alpha) != cudaSuccess) { set_error("splitk_reduce: launch failed"); return CSM_ERR_CUDA; }
  CSM_CHECK_LAUNCH("splitk_reduce");
  return CSM_OK;
}

// groups independent GEMMs in one launch (3-D tensor maps): C_g = op(A_g) op(B_g)
int gemm_tc_grouped(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int groups, int64_t lda,
                    int64_t ags, int64_t ldb, int64_t bgs, int64_t ldc, int64_t cgs, int transA, int transB,
                    int accumulate, cudaStream_t st) {
  GemmTcOperands o{A, B, nullptr, nullptr, lda, ldb, 0, 0, 0, ags, bgs, transA, transB};
  GemmTcParams p{};
  p.M = M; p.N = N; p.K = K; p.groups = groups;
  p.C = C; p.ldc = ldc; p.c_group_stride = cgs;
  p.c_f32 = 0; p.accumulate = accumulate; p.alpha = 1.f;
  return gemm_tc_run(o, p, EPI_STORE, st);
}

}  // namespace csm
