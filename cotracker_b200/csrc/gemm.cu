// gemm.cu -- split-bf16x3 linear layers on the 5th-gen tensor cores.
//
// Persistent warp-specialised kernel, one CTA per SM:
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 3-stage ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (3 MMAs per k16 step: hi*hi, lo*hi, hi*lo)
//   warps 2..17 : epilogue       (tcgen05.ld -> bias / row-bias / GELU / residual -> fp32 and/or split-bf16 stores)
// Two 128-column fp32 accumulators in TMEM are double buffered so the epilogue of tile i overlaps the
// main loop of tile i+1.  Tiles are 128 x 128; consecutive tile ids share the X (activation) tile so the
// big operand is read from HBM once and hit in L2 by the CTAs working on its other N-tiles.
//
// A second, deliberately simple SIMT kernel computes the same contraction from the same split operands
// (reconstructing hi+lo in fp32); tests use it to cross-check the tensor-core path on the GPU.
#include "gemm.cuh"

namespace ct3 {
namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int ACC = 2;
constexpr int TILE_A = BM * BK * 2;                    // 16 KiB (one 16-bit plane)
constexpr int TILE_B = BN * BK * 2;                    // 16 KiB
constexpr int EPI_WARPS = 16;                          // four warps per TMEM lane quarter, a quarter of the columns each
constexpr int CW = 16;                                 // epilogue chunk width (columns)
constexpr int STG_WORDS = 32 * CW;                     // per-warp transpose buffer (rotated rows: conflict-free both ways)
constexpr int MAX_STAGES = 6;
// Per products-per-FLOP variant: which operand planes a stage holds and how deep the ring is (always 192 KiB).
//   3: A hi|lo, W hi|lo (64 KiB x 3)   2: A hi, W hi|lo (48 KiB x 4)   1: A hi, W hi (32 KiB x 6)
template <int NPROD>
struct Cfg {
  // NPROD == 4 / 5: 3 products + the LayerNorm-producer epilogue (raw split rows + partial statistics) / the
  // LayerNorm-consumer epilogue -- separate instantiations so that the default kernels do not carry their registers
  // (compiled into every kernel they cost each GEMM 20-64 bytes of spills and ~7 % of its time)
  static constexpr int AP = NPROD >= 3 ? 2 : 1;
  static constexpr int BP = NPROD >= 2 ? 2 : 1;
  static constexpr int STAGE_BYTES = AP * TILE_A + BP * TILE_B;
  static constexpr int STAGES = NPROD >= 3 ? 3 : (NPROD == 2 ? 4 : 6);
  static constexpr int OFF_B = AP * TILE_A;
};
constexpr int OFF_STG = 3 * 65536;                     // = STAGES * STAGE_BYTES of every variant
static_assert(Cfg<1>::STAGES * Cfg<1>::STAGE_BYTES == OFF_STG && Cfg<2>::STAGES * Cfg<2>::STAGE_BYTES == OFF_STG &&
              Cfg<3>::STAGES * Cfg<3>::STAGE_BYTES == OFF_STG && Cfg<4>::STAGES * Cfg<4>::STAGE_BYTES == OFF_STG &&
              Cfg<5>::STAGES * Cfg<5>::STAGE_BYTES == OFF_STG, "ring size");
constexpr int OFF_BAR = OFF_STG + EPI_WARPS * STG_WORDS * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 /*barriers*/ + 1024 /*align slack*/;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr uint32_t TMEM_COLS = ACC * BN;               // 256 columns (power of two)

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return gelu_erf(x);
  if (act == 2) return gelu_tanh(x);
  return x;
}

// Epilogue of one 32-row x 16-column chunk by one warp.  Phase 1 (lane = row, straight out of tcgen05.ld): bias /
// row-bias / activation, then the 16 output words of the row (16 fp32, or 8 packed-hi | 8 packed-lo bf16 pairs) go to
// a per-warp staging buffer whose rows are rotated by row/2 words (conflict-free for both access patterns without
// padding).  Phase 2 (lane = (row in a group of 8, 16-byte column group)): every warp instruction moves 8 rows x 64 B
// of global memory in whole 32-byte sectors.
__device__ __forceinline__ int stg_idx(int row, int word) { return row * CW + ((word + (row >> 1)) & (CW - 1)); }

template <bool PROD, bool CONS>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& e, int M, int N, int row0, int col0, int lane,
                                               float (&v)[CW], uint32_t* stg, float ln_mean = 0.f, float ln_rstd = 1.f) {
  const int row = row0 + lane;
  if constexpr (CONS) {   // LayerNorm of the A rows applied after the contraction (see GemmEpilogue)
    const float4* w4 = reinterpret_cast<const float4*>(e.ln_wsum + col0);
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 ws = __ldg(w4 + i);
      v[4 * i + 0] = ln_rstd * (v[4 * i + 0] - ln_mean * ws.x); v[4 * i + 1] = ln_rstd * (v[4 * i + 1] - ln_mean * ws.y);
      v[4 * i + 2] = ln_rstd * (v[4 * i + 2] - ln_mean * ws.z); v[4 * i + 3] = ln_rstd * (v[4 * i + 3] - ln_mean * ws.w);
    }
  }
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (e.row_bias && row < M) {
    const float4* b4 = reinterpret_cast<const float4*>(e.row_bias + (int64_t)(row % e.row_mod) * N + col0);
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (e.act != 0) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = apply_act(v[i], e.act);
  }
  const int rsub = lane >> 2, q = lane & 3;   // phase-2 mapping: rows rsub + 8k, 16-byte group q of the 64-byte row
  if (e.out_f32) {
#pragma unroll
    for (int c = 0; c < CW; ++c) stg[stg_idx(lane, c)] = __float_as_uint(v[c]);
    __syncwarp();
    float* base = e.out_f32 + (int64_t)(row0 + rsub) * e.ld_f32 + col0 + 4 * q;
    const int64_t rstep = 8 * e.ld_f32;
    float4 x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 8 * k + rsub;
      x[k] = make_float4(__uint_as_float(stg[stg_idx(r, 4 * q + 0)]), __uint_as_float(stg[stg_idx(r, 4 * q + 1)]),
                         __uint_as_float(stg[stg_idx(r, 4 * q + 2)]), __uint_as_float(stg[stg_idx(r, 4 * q + 3)]));
    }
    if (e.residual) {
      float4 r4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)   // all residual rows in flight at once
        r4[k] = (row0 + 8 * k + rsub < M) ? *reinterpret_cast<const float4*>(base + k * rstep) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) { x[k].x += r4[k].x; x[k].y += r4[k].y; x[k].z += r4[k].z; x[k].w += r4[k].w; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (row0 + 8 * k + rsub < M) *reinterpret_cast<float4*>(base + k * rstep) = x[k];
    if constexpr (PROD) {
      // the final rows once more as a split-bf16 operand + the partial LayerNorm statistics of this 16-column chunk
      const int chunk = col0 >> 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t r = row0 + 8 * k + rsub;
        float s1 = x[k].x + x[k].y + x[k].z + x[k].w;
        float s2 = x[k].x * x[k].x + x[k].y * x[k].y + x[k].z * x[k].z + x[k].w * x[k].w;
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        if (r < M) {
          uint32_t h0, l0, h1, l1;
          split2(x[k].x, x[k].y, h0, l0);
          split2(x[k].z, x[k].w, h1, l1);
          __nv_bfloat16* o = e.raw_split + r * (2 * kLnDim) + col0 + 4 * q;
          *reinterpret_cast<uint2*>(o) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(o + kLnDim) = make_uint2(l0, l1);
          if (q == 0) *reinterpret_cast<float2*>(e.stat_part + (r * kLnParts + chunk) * 2) = make_float2(s1, s2);
        }
      }
    }
    __syncwarp();
  }
  if (e.out_split) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) {
      uint32_t hi, lo;
      split2(v[2 * i], v[2 * i + 1], hi, lo);
      stg[stg_idx(lane, i)] = hi;              // words 0..7 : hi plane of this row (16 bf16)
      stg[stg_idx(lane, CW / 2 + i)] = lo;     // words 8..15: lo plane
    }
    __syncwarp();
    // output offset of this lane's row in 16-byte units (every offset is a multiple of 16 elements); one division
    // per chunk, then one 32-bit shuffle per stored row group
    uint32_t off16_lane = 0xffffffffu;
    if (row < M) {
      const int orow = row / e.row_group;
      off16_lane = (uint32_t)(((long long)orow * e.ld_split + (long long)(row % e.row_group) * N + col0) >> 3);
    }
    uint4* plane = reinterpret_cast<uint4*>(e.out_split + (q >= 2 ? e.lo_off : 0)) + (q & 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 8 * k + rsub;
      const uint32_t o16 = __shfl_sync(0xffffffffu, off16_lane, r);
      const uint4 w4 = make_uint4(stg[stg_idx(r, 4 * q + 0)], stg[stg_idx(r, 4 * q + 1)], stg[stg_idx(r, 4 * q + 2)],
                                  stg[stg_idx(r, 4 * q + 3)]);
      if (o16 != 0xffffffffu) plane[o16] = w4;
    }
    __syncwarp();
  }
}

template <int NPROD>
__global__ void __launch_bounds__(THREADS, 1)
gemm_split3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int M,
                      int N, int Kpad, int fp16, GemmEpilogue epi) {
  using C = Cfg<NPROD>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int num_mt = (M + BM - 1) / BM;
  const int num_nt = N / BN;
  const int num_tiles = num_mt * num_nt;
  const int num_kb = Kpad / BK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / num_nt, nt = tile % num_nt;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* s = smem + stage * STAGE_BYTES;
          tma_load_2d(s, &tmX, kb * BK, mt * BM, &full_bar[stage]);
          if (C::AP == 2) tma_load_2d(s + TILE_A, &tmX, Kpad + kb * BK, mt * BM, &full_bar[stage]);
          tma_load_2d(s + C::OFF_B, &tmW, kb * BK, nt * BN, &full_bar[stage]);
          if (C::BP == 2) tma_load_2d(s + C::OFF_B + TILE_B, &tmW, Kpad + kb * BK, nt * BN, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_16(BM, BN, fp16 != 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);          // TMA bytes have landed
          tc_fence_after_sync();
          const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t a_hi = s, a_lo = s + TILE_A, b_hi = s + C::OFF_B, b_lo = b_hi + TILE_B;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint32_t koff = kk * 32;  // 16 elements = 32 bytes inside the 128-byte swizzle atom
            const uint64_t dah = umma_desc_sw128(a_hi + koff), dbh = umma_desc_sw128(b_hi + koff);
            const uint32_t first = (kb | kk) != 0 ? 1u : 0u;
            if (NPROD >= 3) {   // small terms first
              umma_bf16(d_tmem, umma_desc_sw128(a_lo + koff), dbh, idesc, first);
              umma_bf16(d_tmem, dah, umma_desc_sw128(b_lo + koff), idesc, 1u);
              umma_bf16(d_tmem, dah, dbh, idesc, 1u);
            } else if (NPROD == 2) {
              umma_bf16(d_tmem, dah, umma_desc_sw128(b_lo + koff), idesc, first);
              umma_bf16(d_tmem, dah, dbh, idesc, 1u);
            } else {
              umma_bf16(d_tmem, dah, dbh, idesc, first);
            }
          }
          umma_commit(&empty_bar[stage]);              // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);                  // accumulator complete -> epilogue
        if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access (warp id % 4)
    constexpr int CH = BN / CW / (EPI_WARPS / 4);      // 16-column chunks per warp
    const int chunk0 = ((warp - 2) >> 2) * CH;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / num_nt, nt = tile % num_nt;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const int row0 = mt * BM + quarter * 32;
      uint32_t* stg = reinterpret_cast<uint32_t*>(smem + OFF_STG) + (warp - 2) * STG_WORDS;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if constexpr (NPROD == 5) {
        if (row0 + lane < M) ln_row_stats(epi.ln_part + (int64_t)(row0 + lane) * kLnParts * 2, epi.ln_eps, ln_mean, ln_rstd);
      }
#pragma unroll 1
      for (int chunk = chunk0; chunk < chunk0 + CH; ++chunk) {
        float v[CW];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + chunk * CW);
        tmem_ld16(taddr, v);
        epilogue_chunk<NPROD == 4, NPROD == 5>(epi, M, N, row0, nt * BN + chunk * CW, lane, v, stg, ln_mean, ln_rstd);
      }
      tc_fence_before_sync();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2).  With 128x128 tiles the tensor pipe stalls at ~55 %: every SM must
// ingest 64 KiB of operands per 768 MMA cycles (83 B/cycle) while the L2->SM path delivers ~45 B/cycle (measured:
// 12.7 TB/s chip-wide, profiles/).  A pair of SMs works on a 256 x BN tile instead: each CTA loads only ITS 128 rows
// of X and ITS half (BN/2 rows) of W, one leader thread issues M=256 MMAs that read both shared memories and write
// both TMEMs, so the bytes ingested per FLOP halve.  Both CTAs run the same producer / epilogue code on their own
// 128 rows; barriers: full (leader, tx from both CTAs), empty + tmem_full (multicast commit to both),
// tmem_empty (leader, remote arrives from the peer's epilogue warps).
template <int BNP, int NPROD>
__global__ void __launch_bounds__(THREADS, 1)
gemm_split3_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int M,
                        int N, int Kpad, int fp16, GemmEpilogue epi) {
  using C = Cfg<NPROD>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int TILE_BH = (BNP / 2) * BK * 2;                 // this CTA's half of the W tile, one plane
  constexpr uint32_t TX_BYTES = 2u * ((uint32_t)C::AP * TILE_A + (uint32_t)C::BP * TILE_BH);   // both CTAs, all planes
  constexpr uint32_t TCOLS = 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TCOLS);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                            // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int num_mp = ((M + BM - 1) / BM + 1) / 2;   // pairs of M-tiles
  const int num_nt = N / BNP;
  const int num_tiles = num_mp * num_nt;
  const int num_kb = Kpad / BK;
  const int first = blockIdx.x >> 1, step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < num_tiles; tile += step) {
        const int mt = (tile / num_nt) * 2 + (int)rank, nt = tile % num_nt;
        const int wrow = nt * BNP + (int)rank * (BNP / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], TX_BYTES);
          uint8_t* s = smem + stage * STAGE_BYTES;
          tma_load_2d_2sm(s, &tmX, kb * BK, mt * BM, &full_bar[stage]);
          if (C::AP == 2) tma_load_2d_2sm(s + TILE_A, &tmX, Kpad + kb * BK, mt * BM, &full_bar[stage]);
          tma_load_2d_2sm(s + C::OFF_B, &tmW, kb * BK, wrow, &full_bar[stage]);
          if (C::BP == 2) tma_load_2d_2sm(s + C::OFF_B + TILE_B, &tmW, Kpad + kb * BK, wrow, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_16(256, BNP, fp16 != 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = first; tile < num_tiles; tile += step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);   // both epilogues have drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNP);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);          // both CTAs' TMA bytes have landed
          tc_fence_after_sync();
          const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t a_hi = s, a_lo = s + TILE_A, b_hi = s + C::OFF_B, b_lo = b_hi + TILE_B;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint32_t koff = kk * 32;
            const uint64_t dah = umma_desc_sw128(a_hi + koff), dbh = umma_desc_sw128(b_hi + koff);
            const uint32_t first = (kb | kk) != 0 ? 1u : 0u;
            if (NPROD >= 3) {
              umma_bf16_2sm(d_tmem, umma_desc_sw128(a_lo + koff), dbh, idesc, first);
              umma_bf16_2sm(d_tmem, dah, umma_desc_sw128(b_lo + koff), idesc, 1u);
              umma_bf16_2sm(d_tmem, dah, dbh, idesc, 1u);
            } else if (NPROD == 2) {
              umma_bf16_2sm(d_tmem, dah, umma_desc_sw128(b_lo + koff), idesc, first);
              umma_bf16_2sm(d_tmem, dah, dbh, idesc, 1u);
            } else {
              umma_bf16_2sm(d_tmem, dah, dbh, idesc, first);
            }
          }
          umma_commit_2sm(&empty_bar[stage]);          // frees the stage in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tfull_bar[acc]);              // accumulator complete -> both epilogues
        if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int quarter = warp & 3;
    constexpr int CH = BNP / CW / (EPI_WARPS / 4);     // 16-column chunks per warp
    const int chunk0 = ((warp - 2) >> 2) * CH;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first; tile < num_tiles; tile += step) {
      const int mt = (tile / num_nt) * 2 + (int)rank, nt = tile % num_nt;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const int row0 = mt * BM + quarter * 32;
      uint32_t* stg = reinterpret_cast<uint32_t*>(smem + OFF_STG) + (warp - 2) * STG_WORDS;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if constexpr (NPROD == 5) {
        if (row0 + lane < M) ln_row_stats(epi.ln_part + (int64_t)(row0 + lane) * kLnParts * 2, epi.ln_eps, ln_mean, ln_rstd);
      }
#pragma unroll 1
      for (int chunk = chunk0; chunk < chunk0 + CH; ++chunk) {
        float v[CW];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BNP + chunk * CW);
        tmem_ld16(taddr, v);
        epilogue_chunk<NPROD == 4, NPROD == 5>(epi, M, N, row0, nt * BNP + chunk * CW, lane, v, stg, ln_mean, ln_rstd);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty_bar[acc], 0);   // leader's barrier (local for the leader itself)
      if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                            // the pair retires together
  if (warp == 1) tmem_dealloc_2sm(tmem_base, TCOLS);
}

// ------------------------------------------------------------------------------------------------
// Fused  q|k|v projection + per-track time attention  (Attention.forward inside the time AttnBlock,
// blocks.py:379-398, 426-432; cotracker.py:494-495):
//     att[row, h*48 .. h*48+47] = softmax( q_h k_h^T * 48^-1/2 ) v_h      over the T frames of the row's track.
// GEMM: X = LN(tokens) [M, 768 split] x Wqkv^T, the 1152 weight rows regrouped per head as [q_h(48) | k_h(48) | v_h(48)],
// so one 128 x 144 output tile holds everything head h needs for the tracks of a row tile.  Token rows are
// track-major (row = n*T + t): a tile starts at row mt*R with R = floor(128 / T) * T, i.e. it owns whole tracks
// (the MMA still multiplies 128 rows; the rows past R belong to the next tile and are ignored).
// cta_group::2 pairs as in gemm_split3_pair_kernel (256-row MMA, each CTA loads its 128 rows of X and 72 of the 144
// weight rows).  Epilogue, per CTA on its own 128 rows: 2 independent groups of 4 warps (TMEM lane quarters) take
// alternate tiles (group g owns accumulator g and its own K/V buffer), so two tiles are in their epilogue at once.
// Per tile a thread (= row) reads q (+ bias) into registers, writes k and v (+ bias) to shared [row][k 48 | v 48] fp32,
// releases the accumulator, and -- after the group's named barrier -- runs exact fp32 online-softmax attention of its
// row against the T key rows of its track, then writes the 48 outputs as split bf16 straight into the
// out-projection's operand buffer.
// The fp32 q|k|v tensor (4.6 KB per token) never reaches HBM and the separate attention launch disappears.
namespace qa {
constexpr int BNQ = 144;                               // q|k|v of one head
constexpr int TILE_BQ = (BNQ / 2) * BK * 2;            // 9216 B: this CTA's 72 weight rows, one plane
constexpr int STAGE = 2 * TILE_A + 2 * TILE_BQ;        // 51200 B
constexpr int NSTAGE = 2;                              // the kernel is epilogue-bound: a shallow ring buys the second K/V buffer
constexpr int KV_LD = 100;                             // floats per row of a K/V buffer (96 + 4: conflict-free float4)
constexpr int KV_BYTES = BM * KV_LD * 4;               // 51200 per epilogue group
constexpr int OFF_KV = NSTAGE * STAGE;                 // 102400
constexpr int OFF_BARQ = OFF_KV + 2 * KV_BYTES;        // + 102400
constexpr int SMEM = OFF_BARQ + 256 + 1024;
constexpr int EPIW = 8;
constexpr int NTHREADS = (2 + EPIW) * 32;              // 320 threads (10 warps): up to 168 registers per thread
constexpr int ACC_STRIDE = 256;                        // TMEM columns between the two accumulators
static_assert(SMEM <= 232448, "shared memory budget");
}  // namespace qa

__global__ void __launch_bounds__(qa::NTHREADS, 1)
gemm_qkv_time_attn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int M,
                          int Kpad, int T, int R, float scale_log2e, GemmEpilogue epi) {
  using namespace qa;
  constexpr uint32_t TX_BYTES = 2u * (2u * TILE_A + 2u * TILE_BQ);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BARQ);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;
  uint64_t* tempty_bar = tfull_bar + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC);
  float* kvs = reinterpret_cast<float*>(smem + OFF_KV);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * 4);        // one epilogue group (4 warps) per CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int num_mt = (M + R - 1) / R;
  const int num_mp = (num_mt + 1) / 2;
  const int num_tiles = num_mp * kHeads;       // consecutive tiles = the 8 heads of one row-tile pair (X stays in L2)
  const int num_kb = Kpad / BK;
  const int first = blockIdx.x >> 1, step = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < num_tiles; tile += step) {
        const int mt = (tile / kHeads) * 2 + (int)rank, h = tile % kHeads;
        const int wrow = h * BNQ + (int)rank * (BNQ / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], TX_BYTES);
          uint8_t* s = smem + stage * STAGE;
          tma_load_2d_2sm(s, &tmX, kb * BK, mt * R, &full_bar[stage]);
          tma_load_2d_2sm(s + TILE_A, &tmX, Kpad + kb * BK, mt * R, &full_bar[stage]);
          tma_load_2d_2sm(s + 2 * TILE_A, &tmW, kb * BK, wrow, &full_bar[stage]);
          tma_load_2d_2sm(s + 2 * TILE_A + TILE_BQ, &tmW, Kpad + kb * BK, wrow, &full_bar[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BNQ);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = first; tile < num_tiles; tile += step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t s = smem_u32(smem + stage * STAGE);
          const uint32_t a_hi = s, a_lo = s + TILE_A, b_hi = s + 2 * TILE_A, b_lo = b_hi + TILE_BQ;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint32_t koff = kk * 32;
            const uint64_t dah = umma_desc_sw128(a_hi + koff), dbh = umma_desc_sw128(b_hi + koff);
            umma_bf16_2sm(d_tmem, umma_desc_sw128(a_lo + koff), dbh, idesc, (kb | kk) != 0 ? 1u : 0u);
            umma_bf16_2sm(d_tmem, dah, umma_desc_sw128(b_lo + koff), idesc, 1u);
            umma_bf16_2sm(d_tmem, dah, dbh, idesc, 1u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tfull_bar[acc]);
        if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;             // tiles alternate between the two groups; group g <-> accumulator g
    const int r = quarter * 32 + lane;             // row of the tile = TMEM lane
    const int tracks = R / T;
    const int jtrack = min(r / T, tracks - 1);     // rows past R are computed on a clamped track and never stored
    float* kvg = kvs + group * (KV_BYTES / 4);
    const float* kbase = kvg + (int64_t)jtrack * T * KV_LD;
    const int bar_id = 1 + group;
    int it = 0;
    for (int tile = first; tile < num_tiles; tile += step, ++it) {
      if ((it & 1) != group) continue;
      const int acc = group;
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const int mt = (tile / kHeads) * 2 + (int)rank, h = tile % kHeads;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
      float ln_mean = 0.f, ln_rstd = 1.f;
      {
        const int64_t grow_ln = (int64_t)mt * R + r;
        if (epi.ln_part && grow_ln < M) ln_row_stats(epi.ln_part + grow_ln * kLnParts * 2, epi.ln_eps, ln_mean, ln_rstd);
      }
      float xq[kDh];
      // k, v -> shared memory first, q (kept in registers for the attention) last
#pragma unroll
      for (int pi = 0; pi < 3; ++pi) {
        const int part = (pi + 1) % 3;   // 1 = k, 2 = v, 0 = q
        float x[kDh];
        const int col0 = kDh * part;
#pragma unroll
        for (int c = 0; c < kDh / 16; ++c) {
          float v[16];
          tmem_ld16(tlane + (uint32_t)(col0 + 16 * c), v);
          const float4* b4 = reinterpret_cast<const float4*>(epi.bias + h * BNQ + col0 + 16 * c);
          if (epi.ln_part) {
            const float4* w4 = reinterpret_cast<const float4*>(epi.ln_wsum + h * BNQ + col0 + 16 * c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 ws = __ldg(w4 + i);
              v[4 * i + 0] = ln_rstd * (v[4 * i + 0] - ln_mean * ws.x); v[4 * i + 1] = ln_rstd * (v[4 * i + 1] - ln_mean * ws.y);
              v[4 * i + 2] = ln_rstd * (v[4 * i + 2] - ln_mean * ws.z); v[4 * i + 3] = ln_rstd * (v[4 * i + 3] - ln_mean * ws.w);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b = __ldg(b4 + i);
            x[16 * c + 4 * i + 0] = v[4 * i + 0] + b.x; x[16 * c + 4 * i + 1] = v[4 * i + 1] + b.y;
            x[16 * c + 4 * i + 2] = v[4 * i + 2] + b.z; x[16 * c + 4 * i + 3] = v[4 * i + 3] + b.w;
          }
        }
        if (part == 0) {
#pragma unroll
          for (int i = 0; i < kDh; ++i) xq[i] = x[i];
        } else {
          float4* dst = reinterpret_cast<float4*>(kvg + r * KV_LD + (part - 1) * kDh);
#pragma unroll
          for (int i = 0; i < kDh / 4; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty_bar[acc], 0);   // accumulator drained (leader's barrier)
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // K/V of this tile are in shared memory
      {
        float m = -INFINITY, l = 0.f;
        float o[kDh];
#pragma unroll
        for (int i = 0; i < kDh; ++i) o[i] = 0.f;
        for (int t0 = 0; t0 < T; t0 += 8) {
          float sc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sc[i] = 0.f;
#pragma unroll
          for (int d4 = 0; d4 < kDh / 4; ++d4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 kk = *reinterpret_cast<const float4*>(kbase + min(t0 + i, T - 1) * KV_LD + 4 * d4);
              sc[i] = fmaf(xq[4 * d4 + 0], kk.x, sc[i]);
              sc[i] = fmaf(xq[4 * d4 + 1], kk.y, sc[i]);
              sc[i] = fmaf(xq[4 * d4 + 2], kk.z, sc[i]);
              sc[i] = fmaf(xq[4 * d4 + 3], kk.w, sc[i]);
            }
          }
          float mnew = m;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sc[i] = (t0 + i < T) ? sc[i] * scale_log2e : -INFINITY;
            mnew = fmaxf(mnew, sc[i]);
          }
          const float corr = exp2f(m - mnew);          // exp2(-inf) = 0 on the first chunk
          l *= corr;
#pragma unroll
          for (int i = 0; i < kDh; ++i) o[i] *= corr;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float pi = exp2f(sc[i] - mnew);      // 0 for the masked tail
            l += pi;
            const float* vr = kbase + min(t0 + i, T - 1) * KV_LD + kDh;
#pragma unroll
            for (int d4 = 0; d4 < kDh / 4; ++d4) {
              const float4 vv = *reinterpret_cast<const float4*>(vr + 4 * d4);
              o[4 * d4 + 0] = fmaf(pi, vv.x, o[4 * d4 + 0]);
              o[4 * d4 + 1] = fmaf(pi, vv.y, o[4 * d4 + 1]);
              o[4 * d4 + 2] = fmaf(pi, vv.z, o[4 * d4 + 2]);
              o[4 * d4 + 3] = fmaf(pi, vv.w, o[4 * d4 + 3]);
            }
          }
          m = mnew;
        }
        const int64_t grow = (int64_t)mt * R + r;
        if (r < R && grow < M) {
          const float inv = 1.0f / l;
          uint4* ph = reinterpret_cast<uint4*>(epi.out_split + grow * epi.ld_split + h * kDh);
          uint4* pl = reinterpret_cast<uint4*>(epi.out_split + grow * epi.ld_split + epi.lo_off + h * kDh);
#pragma unroll
          for (int i = 0; i < kDh / 8; ++i) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split2(o[8 * i + 2 * j] * inv, o[8 * i + 2 * j + 1] * inv, hw[j], lw[j]);
            ph[i] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            pl[i] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // K/V consumed: this group's next tile may overwrite
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// SIMT verification kernel: 64x64 tile, 256 threads, each 4x4 outputs; fp32 FMA on hi+lo.
__device__ __forceinline__ void epilogue_store1(const GemmEpilogue& e, int N, int row, int col, float v) {
  if (e.ln_part) {
    float mean, rstd;
    ln_row_stats(e.ln_part + (int64_t)row * kLnParts * 2, e.ln_eps, mean, rstd);
    v = rstd * (v - mean * e.ln_wsum[col]);
  }
  if (e.bias) v += e.bias[col];
  if (e.row_bias) v += e.row_bias[(int64_t)(row % e.row_mod) * N + col];
  v = apply_act(v, e.act);
  if (e.out_f32) {
    float* o = e.out_f32 + (int64_t)row * e.ld_f32 + col;
    if (e.residual) v += *o;
    *o = v;
  }
  if (e.out_split) {
    const int orow = row / e.row_group;
    const int ocol = (row % e.row_group) * N + col;
    bf16pair p = split_bf16(v);
    __nv_bfloat16* hp = e.out_split + (int64_t)orow * e.ld_split + ocol;
    hp[0] = p.hi;
    hp[e.lo_off] = p.lo;
  }
}

__global__ void __launch_bounds__(256)
gemm_split3_simt_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ W, int M, int N,
                        int Kpad, int64_t x_ld, int products, int fp16, GemmEpilogue epi) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  const int64_t ld = 2 * (int64_t)Kpad;
  // the planes the tensor-core path multiplies: x_hi (+ x_lo when 3 products), w_hi (+ w_lo when >= 2 products)
  auto val = [fp16](const __nv_bfloat16* p) {
    return fp16 ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*p);
  };
  for (int k0 = 0; k0 < Kpad; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      const int gm = m0 + r, gn = n0 + r;
      float a = 0.f, b = 0.f;
      if (gm < M) a = val(X + gm * x_ld + k0 + k) + (products == 3 ? val(X + gm * x_ld + Kpad + k0 + k) : 0.f);
      if (gn < N) b = val(W + gn * ld + k0 + k) + (products >= 2 ? val(W + gn * ld + Kpad + k0 + k) : 0.f);
      As[k][r] = a;
      Bs[k][r] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (r < M && c < N) epilogue_store1(epi, N, r, c, acc[i][j]);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static const EncodeTiledFn fn = [] {   // C++11 thread-safe one-time initialisation
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

// 2-D bf16 tensor [rows, cols] row-major, box = 64 cols (128 B) x 128 rows, 128-byte swizzle, OOB -> 0
bool make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows = 128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

bool encode_tensor_map(CUtensorMap* m, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn || rank < 1 || rank > 5) return false;
  cuuint64_t d[5], st[5];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  return fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

template <int NPROD>
cudaError_t set_attrs() {
  cudaError_t e = cudaFuncSetAttribute(gemm_split3_tc_kernel<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_split3_pair_kernel<256, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_split3_pair_kernel<192, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  return e;
}

template <int NPROD>
cudaError_t launch_tc(const GemmProblem& p, const CUtensorMap& tmX, const CUtensorMap& tmW, int bnp, int num_mt,
                      int num_sms, cudaStream_t stream) {
  if (bnp == 0) {
    const int num_tiles = num_mt * (p.N / BN);
    const int grid = num_tiles < num_sms ? num_tiles : num_sms;
    gemm_split3_tc_kernel<NPROD><<<grid, THREADS, SMEM_BYTES, stream>>>(tmX, tmW, p.M, p.N, p.Kpad, p.fp16, p.epi);
    return cudaGetLastError();
  }
  const int groups = ((num_mt + 1) / 2) * (p.N / bnp);
  int pairs = num_sms / 2;
  if (pairs > groups) pairs = groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = (bnp == 256)
      ? cudaLaunchKernelEx(&cfg, gemm_split3_pair_kernel<256, NPROD>, tmX, tmW, p.M, p.N, p.Kpad, p.fp16, p.epi)
      : cudaLaunchKernelEx(&cfg, gemm_split3_pair_kernel<192, NPROD>, tmX, tmW, p.M, p.N, p.Kpad, p.fp16, p.epi);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace

bool qkv_time_attn_supported(int T) { return T >= 1 && T <= BM; }

int gemm_qkv_time_attn_launch(const __nv_bfloat16* x_split, const __nv_bfloat16* w_heads, const float* bias_heads,
                              int M, int Kpad, int T, __nv_bfloat16* att_split, int64_t ld_split, int lo_off,
                              float scale, const float* ln_part, const float* ln_wsum, float ln_eps, int num_sms,
                              cudaStream_t stream, const char** err) {
  *err = nullptr;
  if (M <= 0 || Kpad <= 0 || (Kpad % BK) != 0 || !qkv_time_attn_supported(T) || (M % T) != 0) {
    *err = "qkv_time_attn: need M > 0, M % T == 0, 1 <= T <= 128, Kpad % 64 == 0";
    return (int)cudaErrorInvalidValue;
  }
  if (((reinterpret_cast<uintptr_t>(x_split) | reinterpret_cast<uintptr_t>(w_heads) |
        reinterpret_cast<uintptr_t>(att_split)) & 15) || (ld_split % 8) != 0 || (lo_off % 8) != 0) {
    *err = "qkv_time_attn: operands must be 16-byte aligned";
    return (int)cudaErrorInvalidValue;
  }
  const int R = (BM / T) * T;
  CUtensorMap tmX, tmW;
  if (!make_tmap(&tmX, x_split, (uint64_t)M, 2ull * Kpad) ||
      !make_tmap(&tmW, w_heads, (uint64_t)(kHeads * qa::BNQ), 2ull * Kpad, (uint32_t)(qa::BNQ / 2))) {
    *err = "qkv_time_attn: cuTensorMapEncodeTiled failed";
    return (int)cudaErrorInvalidValue;
  }
  static DeviceOnce attr;
  cudaError_t e = once_per_device(attr, [&] {
    return cudaFuncSetAttribute(gemm_qkv_time_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qa::SMEM);
  });
  if (e != cudaSuccess) { *err = "qkv_time_attn: cudaFuncSetAttribute failed"; return (int)e; }
  const int num_mt = (M + R - 1) / R;
  const int groups = ((num_mt + 1) / 2) * kHeads;
  int pairs = num_sms / 2;
  if (pairs > groups) pairs = groups;
  GemmEpilogue epi;
  epi.bias = bias_heads;
  epi.out_split = att_split;
  epi.ld_split = ld_split;
  epi.lo_off = lo_off;
  epi.ln_part = ln_part;
  epi.ln_wsum = ln_wsum;
  epi.ln_eps = ln_eps;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(qa::NTHREADS);
  cfg.dynamicSmemBytes = qa::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, gemm_qkv_time_attn_kernel, tmX, tmW, M, Kpad, T, R,
                         scale * 1.44269504088896340736f, epi);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

int gemm_launch(const GemmProblem& p, int impl, int num_sms, cudaStream_t stream, const char** err) {
  *err = nullptr;
  if (p.M <= 0 || p.N <= 0 || p.Kpad <= 0 || (p.N % BN) != 0 || (p.Kpad % BK) != 0) {
    *err = "gemm: need M>0, N % 128 == 0, Kpad % 64 == 0";
    return (int)cudaErrorInvalidValue;
  }
  if (p.products < 1 || p.products > 3) {
    *err = "gemm: products must be 1, 2 or 3";
    return (int)cudaErrorInvalidValue;
  }
  const int64_t x_ld = p.x_ld ? p.x_ld : 2 * (int64_t)p.Kpad;
  if (x_ld < (p.products == 3 ? 2 : 1) * (int64_t)p.Kpad || (x_ld % 8) != 0) {
    *err = "gemm: x_ld too small for the operand planes this product count reads";
    return (int)cudaErrorInvalidValue;
  }
  if ((reinterpret_cast<uintptr_t>(p.x_split) | reinterpret_cast<uintptr_t>(p.w_split)) & 15) {
    *err = "gemm: operands must be 16-byte aligned";
    return (int)cudaErrorInvalidValue;
  }
  if (p.epi.raw_split && (p.N != kLnDim || !p.epi.out_f32 || !p.epi.stat_part || impl == 1)) {
    *err = "gemm: raw_split/stat_part need the fp32 output path of the tensor-core kernels with N == 384";
    return (int)cudaErrorInvalidValue;
  }
  if (p.epi.ln_part && !p.epi.ln_wsum) {
    *err = "gemm: ln_part needs ln_wsum";
    return (int)cudaErrorInvalidValue;
  }
  if (impl == 1) {
    dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
    gemm_split3_simt_kernel<<<grid, 256, 0, stream>>>(p.x_split, p.w_split, p.M, p.N, p.Kpad, x_ld, p.products,
                                                      p.fp16, p.epi);
    return (int)cudaGetLastError();
  }
  const int num_mt = (p.M + BM - 1) / BM;
  // CTA pairs (cta_group::2) once there is enough work to fill the machine with 256-row tiles
  int bnp = 0;
  if (impl == 0 && num_mt >= 2 * num_sms) bnp = (p.N % 256 == 0) ? 256 : ((p.N % 192 == 0) ? 192 : 0);
  CUtensorMap tmX, tmW;
  if (!make_tmap(&tmX, p.x_split, (uint64_t)p.M, (uint64_t)x_ld) ||
      !make_tmap(&tmW, p.w_split, (uint64_t)p.N, 2ull * p.Kpad, bnp ? (uint32_t)(bnp / 2) : 128u)) {
    *err = "gemm: cuTensorMapEncodeTiled failed";
    return (int)cudaErrorInvalidValue;
  }
  static DeviceOnce attr_set;
  {
    cudaError_t e = once_per_device(attr_set, [&] {
      cudaError_t e = set_attrs<3>();
      if (e == cudaSuccess) e = set_attrs<4>();
      if (e == cudaSuccess) e = set_attrs<5>();
      if (e == cudaSuccess) e = set_attrs<2>();
      if (e == cudaSuccess) e = set_attrs<1>();
      return e;
    });
    if (e != cudaSuccess) { *err = "gemm: cudaFuncSetAttribute(max dynamic smem) failed"; return (int)e; }
  }
  if (p.epi.ln_part && (p.products != 3 || p.epi.raw_split)) {
    *err = "gemm: the LayerNorm-consumer epilogue needs 3 products and cannot be combined with raw_split";
    return (int)cudaErrorInvalidValue;
  }
  cudaError_t e = (p.products == 3 && p.epi.raw_split) ? launch_tc<4>(p, tmX, tmW, bnp, num_mt, num_sms, stream)
                : (p.products == 3 && p.epi.ln_part) ? launch_tc<5>(p, tmX, tmW, bnp, num_mt, num_sms, stream)
                : p.products == 3 ? launch_tc<3>(p, tmX, tmW, bnp, num_mt, num_sms, stream)
                : p.products == 2 ? launch_tc<2>(p, tmX, tmW, bnp, num_mt, num_sms, stream)
                                  : launch_tc<1>(p, tmX, tmW, bnp, num_mt, num_sms, stream);
  return (int)e;
}

}  // namespace ct3
