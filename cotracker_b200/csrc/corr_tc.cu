// corr_tc.cu -- fused bilinear sampling + 4-D correlation on the 5th-gen tensor cores (the production path of
// launch_corr_sample; corr.cu keeps the exact-fp32 SIMT version the tests cross-check against).
//
//   vol[(n,t,l)][(a*7+b)*49 + (i*7+j)] = < bilinear(F_l[t], cx/2^l + a-3, cy/2^l + b-3) , S_l[n, i*7+j, :] >
//   (get_correlation_feat + einsum, cotracker3_online.py:130-143, cotracker3_offline.py:144-156)
//
// Persistent, warp-specialised; work unit = (track n, level l), tile = two frames of that unit:
//   A tile  [128 x 128] : rows f*49 + a*7 + b (98 used) = sampled feature vectors, built IN SHARED MEMORY by the
//                         producer warps (separable 4-tap blend from the channels-last pyramid: one coalesced
//                         512-byte texel line per tap), stored split-bf16 in the 128B-swizzled K-major layout
//   B tile  [ 64 x 128] : the 49 support vectors of (n,l) (rows 49..63 zero), split-bf16, built once per unit
//   D       [128 x  64] : fp32 in TMEM, 3 tcgen05.mma per k16 step (lo*hi + hi*lo + hi*hi), double buffered
//   epilogue            : tcgen05.ld -> split-bf16 -> staging image of the complete 9728-byte volume rows
//                         ([hi(2432) | lo(2432)], K padding zero) -> fully coalesced 16-byte stores
// Warps: 0..13 producers (warp w owns frame w/7 and x-offset a = w%7 of the tile; all 16 tap loads of a column
// are issued before the first use), 14 = TMEM alloc + MMA issuer, 15..18 epilogue.
// Neither the sampled features (10 GB/iteration in the reference) nor an fp32 volume ever touch HBM.
#include "kernels.cuh"

namespace ct3 {
namespace {

constexpr int PW = 14;                    // producer warps: warp w -> (frame w/7 of the tile, x-offset a = w%7)
constexpr int MMA_WARP = 14;
constexpr int EPI_WARP0 = 15;             // warps 15..18 cover the four TMEM lane quarters (warp & 3)
constexpr int THREADS = 19 * 32;
constexpr int A_PART = 2 * 16384;         // one bf16 plane of A: 2 K-atoms x [128 rows x 128 B]
constexpr int A_STAGE = 2 * A_PART;       // hi + lo = 64 KiB
constexpr int S_PART = 2 * 8192;          // one plane of S: 2 K-atoms x [64 rows x 128 B]
constexpr int S_BYTES = 2 * S_PART;       // 32 KiB
constexpr int ROW_BYTES = 2 * kVolPad * 2;  // 9728: one volume row image [hi | lo]
constexpr int STG_BYTES = 2 * ROW_BYTES;  // two frames per tile
constexpr int OFF_S = 0;
constexpr int OFF_A = OFF_S + S_BYTES;
constexpr int OFF_STG = OFF_A + 2 * A_STAGE;
constexpr int OFF_BAR = OFF_STG + 2 * STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr uint32_t TMEM_COLS = 128;       // 2 accumulators x 64 columns

struct CorrTcArgs {
  const float* pyr;
  PyramidLayout lay;
  const float* support;        // [4][49, N, 128]
  const uint8_t* track_valid;  // [N] or null
  const float* coords;         // [T, N, 2]
  int T, N;
  __nv_bfloat16* vol;          // [N*T*4, 2*kVolPad]
};

// byte offset of (row r, 16-byte chunk c) inside one [rows x 128 B] swizzle-128B K-atom
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// lane -> where its 4 channels [4*lane, 4*lane+4) live: K-atom, 16B chunk, 8-byte half
struct LanePos { int atom, chunk, half; };
__device__ __forceinline__ LanePos lane_pos(int lane) { return {lane >> 4, (lane & 15) >> 1, lane & 1}; }

__device__ __forceinline__ void store_split4(uint8_t* plane_hi, uint8_t* plane_lo, int atom_bytes, int row,
                                             LanePos lp, float4 v) {
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  const uint32_t off = (uint32_t)(lp.atom * atom_bytes) + sw128(row, lp.chunk) + (uint32_t)(lp.half * 8);
  *reinterpret_cast<uint2*>(plane_hi + off) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(plane_lo + off) = make_uint2(l0, l1);
}

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float w) {
  const float u = 1.f - w;
  return make_float4(u * a.x + w * b.x, u * a.y + w * b.y, u * a.z + w * b.z, u * a.w + w * b.w);
}

__global__ void __launch_bounds__(THREADS, 1)
corr_sample_tc_kernel(CorrTcArgs g, int num_units) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* a_full = bars;        // [2] producers -> MMA
  uint64_t* a_empty = bars + 2;   // [2] MMA -> producers (tcgen05.commit)
  uint64_t* d_full = bars + 4;    // [2] MMA -> epilogue (tcgen05.commit)
  uint64_t* d_empty = bars + 6;   // [2] epilogue -> MMA
  uint64_t* s_full = bars + 8;    // producers -> MMA, once per unit
  uint64_t* s_empty = bars + 9;   // MMA -> producers, once per unit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_unit = (g.T + 1) / 2;

  // one-time: zero S (rows 49..63 stay zero forever) and the staging images (K padding stays zero)
  for (int i = threadIdx.x; i < S_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + OFF_S)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 2 * STG_BYTES / 16; i += THREADS)
    reinterpret_cast<uint4*>(smem + OFF_STG)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], PW);
      mbar_init(&a_empty[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 4);
    }
    mbar_init(s_full, PW);
    mbar_init(s_empty, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PW) {
    // ================================================================== producers
    const LanePos lp = lane_pos(lane);
    const int f = warp / 7, a = warp % 7;  // frame of the tile / x-offset index owned by this warp
    uint32_t it = 0;     // tile counter of this CTA
    uint32_t ui = 0;     // unit counter of this CTA
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++ui) {
      const int n = u / kL, l = u % kL;
      const int H = g.lay.h[l], W = g.lay.w[l];
      const float inv = 1.0f / (float)(1 << l);
      // ---- support tile (B operand), once per unit
      if (ui > 0) mbar_wait(s_empty, (ui - 1) & 1u);   // MMAs of the previous unit have retired
      {
        const bool valid = g.track_valid == nullptr || g.track_valid[n] != 0;
        uint8_t* s_hi = smem + OFF_S;
        uint8_t* s_lo = smem + OFF_S + S_PART;
        for (int p = warp; p < kP; p += PW) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) v = __ldg(reinterpret_cast<const float4*>(g.support + ((int64_t)l * kP * g.N + (int64_t)p * g.N + n) * kD) + lane);
          store_split4(s_hi, s_lo, 8192, p, lp, v);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_full);
      }
      // ---- A tiles: two frames each, one (frame, x-offset) column of 7 samples per warp
      for (int tp = 0; tp < tiles_per_unit; ++tp, ++it) {
        const int stage = it & 1;
        const int t = 2 * tp + f;
        // addresses and weights first (independent of the smem slot) ...
        float wy[7];
        int y0[7], yl[8];   // yl[0] = y0 of sample 0, yl[k+1] = y1 of sample k: the (<= 8) distinct rows of the column
        float4 hrow[8];
        const float* fm = g.pyr + g.lay.off[l] + (int64_t)min(t, g.T - 1) * H * W * kD;
        int x0 = 0, x1 = 0;
        float wx = 0.f;
        if (t < g.T) {
          const float cx = g.coords[((int64_t)t * g.N + n) * 2 + 0] * inv;
          const float cy = g.coords[((int64_t)t * g.N + n) * 2 + 1] * inv;
          const float x = fminf(fmaxf(cx + (float)(a - kR), 0.f), (float)(W - 1));
          const float xf = floorf(x);
          x0 = (int)xf;
          x1 = min(x0 + 1, W - 1);
          wx = x - xf;
#pragma unroll
          for (int b = 0; b < 7; ++b) {
            const float y = fminf(fmaxf(cy + (float)(b - kR), 0.f), (float)(H - 1));
            const float yf = floorf(y);
            y0[b] = (int)yf;
            wy[b] = y - yf;
            if (b == 0) yl[0] = y0[0];
            yl[b + 1] = min(y0[b] + 1, H - 1);
          }
          // ... then all 16 taps in flight at once (one L2 round trip per column instead of eight)
          float4 p0[8], p1[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            p0[k] = __ldg(reinterpret_cast<const float4*>(fm + ((int64_t)yl[k] * W + x0) * kD) + lane);
            p1[k] = __ldg(reinterpret_cast<const float4*>(fm + ((int64_t)yl[k] * W + x1) * kD) + lane);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) hrow[k] = lerp4(p0[k], p1[k], wx);
        }
        mbar_wait(&a_empty[stage], ((it >> 1) & 1u) ^ 1u);
        if (t < g.T) {
          uint8_t* a_hi = smem + OFF_A + stage * A_STAGE;
          uint8_t* a_lo = a_hi + A_PART;
#pragma unroll
          for (int b = 0; b < 7; ++b) {
            // rows are consecutive unless the sample was clamped at the low border (y0 stays at row yl[0]); a
            // floor() jump caused by fp32 rounding of cy + offset (probability ~1e-7) takes the direct path
            float4 h0;
            if (y0[b] == yl[b]) {
              h0 = hrow[b];
            } else if (y0[b] == yl[0]) {
              h0 = hrow[0];
            } else {
              h0 = lerp4(__ldg(reinterpret_cast<const float4*>(fm + ((int64_t)y0[b] * W + x0) * kD) + lane),
                         __ldg(reinterpret_cast<const float4*>(fm + ((int64_t)y0[b] * W + x1) * kD) + lane), wx);
            }
            store_split4(a_hi, a_lo, 16384, f * kP + a * 7 + b, lp, lerp4(h0, hrow[b + 1], wy[b]));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[stage]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      uint32_t it = 0, ui = 0;
      const uint32_t s_base = smem_u32(smem + OFF_S);
      for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++ui) {
        mbar_wait(s_full, ui & 1u);
        for (int tp = 0; tp < tiles_per_unit; ++tp, ++it) {
          const int stage = it & 1;
          const uint32_t ph = (it >> 1) & 1u;
          mbar_wait(&a_full[stage], ph);
          mbar_wait(&d_empty[stage], ph ^ 1u);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(smem + OFF_A + stage * A_STAGE);
          const uint32_t d_tmem = tmem_base + (uint32_t)(stage * 64);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t ao = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
            const uint32_t so = (uint32_t)((ks >> 2) * 8192 + (ks & 3) * 32);
            const uint64_t dah = umma_desc_sw128(a_base + ao), dal = umma_desc_sw128(a_base + A_PART + ao);
            const uint64_t dsh = umma_desc_sw128(s_base + so), dsl = umma_desc_sw128(s_base + S_PART + so);
            umma_bf16(d_tmem, dal, dsh, idesc, ks != 0 ? 1u : 0u);
            umma_bf16(d_tmem, dah, dsl, idesc, 1u);
            umma_bf16(d_tmem, dah, dsh, idesc, 1u);
          }
          umma_commit(&a_empty[stage]);
          umma_commit(&d_full[stage]);
        }
        umma_commit(s_empty);
      }
    }
  } else {
    // ================================================================== epilogue
    const int q = warp & 3;              // TMEM lane quarter (warps 8..11 -> 0..3)
    const int r = q * 32 + lane;         // D row
    const int f = r >= kP ? 1 : 0;
    const int rho = r - f * kP;          // a*7+b
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    uint32_t it = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int n = u / kL, l = u % kL;
      for (int tp = 0; tp < tiles_per_unit; ++tp, ++it) {
        const int stage = it & 1;
        mbar_wait(&d_full[stage], (it >> 1) & 1u);
        tc_fence_after_sync();
        uint8_t* stg = smem + OFF_STG + stage * STG_BYTES;
        const bool row_ok = r < 2 * kP && (2 * tp + f) < g.T;
        __nv_bfloat16* dst_hi = reinterpret_cast<__nv_bfloat16*>(stg + f * ROW_BYTES) + rho * kP;
        __nv_bfloat16* dst_lo = dst_hi + kVolPad;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * 64);
        float v[32];
        tmem_ld32(taddr, v);
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const bf16pair p = split_bf16(v[c]);
            dst_hi[c] = p.hi;
            dst_lo[c] = p.lo;
          }
        }
        tmem_ld32(taddr + 32, v);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty[stage]);  // accumulator drained (registers hold the rest)
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < kP - 32; ++c) {
            const bf16pair p = split_bf16(v[c]);
            dst_hi[32 + c] = p.hi;
            dst_lo[32 + c] = p.lo;
          }
        }
        // all 128 epilogue threads: staging complete -> coalesced copy-out of whole volume rows
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int idx = et; idx < 2 * (ROW_BYTES / 16); idx += 128) {
          const int ff = idx / (ROW_BYTES / 16), w16 = idx % (ROW_BYTES / 16);
          const int t = 2 * tp + ff;
          if (t < g.T) {
            uint4* grow = reinterpret_cast<uint4*>(g.vol + (((int64_t)n * g.T + t) * kL + l) * (2 * kVolPad));
            grow[w16] = reinterpret_cast<const uint4*>(stg + ff * ROW_BYTES)[w16];
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

cudaError_t launch_corr_sample_tc(const float* pyr, int H4, int W4, const float* support,
                                  const uint8_t* track_valid, const float* coords, int T, int N,
                                  __nv_bfloat16* vol_split, int num_sms, cudaStream_t s) {
  CorrTcArgs g;
  g.pyr = pyr;
  g.lay = pyramid_layout(T, H4, W4);
  g.support = support;
  g.track_valid = track_valid;
  g.coords = coords;
  g.T = T;
  g.N = N;
  g.vol = vol_split;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(corr_sample_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int num_units = N * kL;
  const int grid = num_units < num_sms ? num_units : num_sms;
  corr_sample_tc_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(g, num_units);
  return cudaGetLastError();
}

}  // namespace ct3
