// attention_p2v.cu -- point <- virtual cross attention on the 5th-gen tensor cores (tcgen05):
//     out[n, t, h*48 ..] = softmax( q_h[n,t] . K_h[t]^T * 48^-1/2 ) V_h[t]          over the 64 virtual tokens of frame t
// (Attention.forward, blocks.py:379-398, called from CrossAttnBlock as space_point2virtual_blocks, cotracker.py:515-517).
// One CTA = one frame t x 128 consecutive tracks; per head:
//   1. the 128 query rows (48 fp32 each, pre-multiplied by 48^-1/2 log2 e) and the 64 key rows are split into bf16
//      hi|lo planes and written as 128B-swizzled K-major operand tiles (head dim 48 zero-padded to the 64-element
//      swizzle row); V_h is written TRANSPOSED ([48 dims x 64 keys], K = keys) as the B operand of the second product
//   2. S = Q K^T : tcgen05.mma M=128, N=64, 3 k16 steps x 3 split products -> TMEM
//   3. softmax on the thread's own row (TMEM lane = query row: 64 scores in registers, exp2), P normalised, split,
//      written as the K-major A tile of the second product (64 keys = exactly one 128-byte row)
//   4. O = P V : M=128, N=48, 4 k16 steps x 3 split products -> TMEM -> registers -> split bf16 rows of the
//      out-projection's operand buffer
// Three CTAs are resident per SM (65 KiB of shared memory, 128 TMEM columns each) and hide each other's phase latency.
// Replaces the mma.sync kernel of attention_tc.cu for this pattern (190 us -> see profiles/ per call at N=6400, T=16).
#include "gemm.cuh"
#include "kernels.cuh"

namespace ct3 {
namespace {

constexpr int PV_THREADS = 128;
constexpr int PV_TILE_Q = 128 * 128;     // [128 rows x 128 B] one plane of Q / P
constexpr int PV_TILE_K = 64 * 128;      // [64 keys x 128 B] one plane of K
constexpr int PV_OFF_Q = 0;                              // hi | lo
constexpr int PV_OFF_K = PV_OFF_Q + 2 * PV_TILE_Q;       // hi | lo
constexpr int PV_OFF_V = PV_OFF_K + 2 * PV_TILE_K;       // hi | lo (each padded to 8 KiB for 1024-byte alignment)
constexpr int PV_OFF_P = PV_OFF_Q;                       // P reuses the Q tiles (S = Q K^T has retired when P is written)
constexpr int PV_OFF_BAR = PV_OFF_V + 2 * 8192;
constexpr int PV_SMEM = PV_OFF_BAR + 64 + 1024;
static_assert(PV_SMEM <= 75 * 1024, "three CTAs per SM");

__device__ __forceinline__ uint32_t swz(int r, int c16) { return (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(PV_THREADS, 3)
attn_p2v_tc_kernel(AttnParams p, int tiles_per_seq) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + PV_OFF_BAR);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);
  const int r = threadIdx.x, warp = r >> 5;
  const int t = blockIdx.x / tiles_per_seq, n0 = (blockIdx.x % tiles_per_seq) * 128;
  const int n = n0 + r;
  const bool valid = n < p.Lq;

  // zero every operand tile once: the head-dim padding (columns 48..63 of Q / K rows) is never written again
  for (int i = r; i < PV_OFF_BAR / 16; i += PV_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (r == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16);         // S: columns 0..63
  const uint32_t t_o = t_s + 64;                                          // O: columns 64..111
  const float qscale = p.scale * 1.44269504088896340736f;
  const float* qrow = p.q + ((int64_t)t * p.q_seq_stride + (int64_t)(valid ? n : 0) * p.q_tok_stride) * p.q_ld + p.q_col;
  __nv_bfloat16* orow = p.out + ((int64_t)t * p.q_seq_stride + (int64_t)n * p.q_tok_stride) * p.out_ld;
  // threads 0..63 stage key row r, threads 64..127 stage value row r - 64
  const int kvi = r & 63;
  const float* kvrow = p.kv + ((int64_t)t * p.k_seq_stride + (int64_t)kvi * p.k_tok_stride) * p.kv_ld + (r < 64 ? p.k_col : p.v_col);
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64), idesc_o = umma_idesc_bf16(128, 48);
  const uint32_t sQ = smem_u32(smem + PV_OFF_Q), sK = smem_u32(smem + PV_OFF_K), sV = smem_u32(smem + PV_OFF_V),
                 sP = smem_u32(smem + PV_OFF_P);

  for (int h = 0; h < kHeads; ++h) {
    const uint32_t ph = (uint32_t)(h & 1);
    // ---- 1. stage Q (this thread's row), K / V^T (this thread's key)
    {
      const float4* q4 = reinterpret_cast<const float4*>(qrow + h * kDh);
#pragma unroll
      for (int c = 0; c < 8; ++c) {                // 6 chunks of 8 elements = 16 bytes per plane + 2 chunks of zero padding
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (valid && c < kDh / 8) { a = __ldg(q4 + 2 * c); b = __ldg(q4 + 2 * c + 1); }
        uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
        split2(a.x * qscale, a.y * qscale, h0, l0);
        split2(a.z * qscale, a.w * qscale, h1, l1);
        split2(b.x * qscale, b.y * qscale, h2, l2);
        split2(b.z * qscale, b.w * qscale, h3, l3);
        *reinterpret_cast<uint4*>(smem + PV_OFF_Q + swz(r, c)) = make_uint4(h0, h1, h2, h3);
        *reinterpret_cast<uint4*>(smem + PV_OFF_Q + PV_TILE_Q + swz(r, c)) = make_uint4(l0, l1, l2, l3);
      }
      const float4* k4 = reinterpret_cast<const float4*>(kvrow + h * kDh);
      if (r < 64) {
#pragma unroll
        for (int c = 0; c < kDh / 8; ++c) {
          const float4 a = __ldg(k4 + 2 * c), b = __ldg(k4 + 2 * c + 1);
          uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
          split2(a.x, a.y, h0, l0);
          split2(a.z, a.w, h1, l1);
          split2(b.x, b.y, h2, l2);
          split2(b.z, b.w, h3, l3);
          *reinterpret_cast<uint4*>(smem + PV_OFF_K + swz(kvi, c)) = make_uint4(h0, h1, h2, h3);
          *reinterpret_cast<uint4*>(smem + PV_OFF_K + PV_TILE_K + swz(kvi, c)) = make_uint4(l0, l1, l2, l3);
        }
      } else {
        // V^T: element (dim d, key kvi) at row d, 16-byte chunk kvi / 8, position kvi % 8
#pragma unroll
        for (int c = 0; c < kDh / 4; ++c) {
          const float4 a = __ldg(k4 + c);
          const float vals[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = 4 * c + j;
            const bf16pair sp = split_bf16(vals[j]);
            const uint32_t off = swz(d, kvi >> 3) + (uint32_t)((kvi & 7) * 2);
            *reinterpret_cast<__nv_bfloat16*>(smem + PV_OFF_V + off) = sp.hi;
            *reinterpret_cast<__nv_bfloat16*>(smem + PV_OFF_V + 8192 + off) = sp.lo;
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    // ---- 2. S = Q K^T
    if (warp == 0 && elect_one()) {
      tc_fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < kDh / 16; ++kk) {
        const uint32_t ko = kk * 32;
        const uint64_t qh = umma_desc_sw128(sQ + ko), ql = umma_desc_sw128(sQ + PV_TILE_Q + ko);
        const uint64_t kh = umma_desc_sw128(sK + ko), kl = umma_desc_sw128(sK + PV_TILE_K + ko);
        umma_bf16(tmem_base, ql, kh, idesc_s, kk != 0 ? 1u : 0u);
        umma_bf16(tmem_base, qh, kl, idesc_s, 1u);
        umma_bf16(tmem_base, qh, kh, idesc_s, 1u);
      }
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after_sync();
    // ---- 3. softmax of this thread's row -> normalised P (split) as the A tile of the second product
    {
      float s[64];
      tmem_ld64(t_s, s);
      float m = s[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) m = fmaxf(m, s[i]);
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { s[i] = exp2f(s[i] - m); l += s[i]; }
      const float inv = 1.0f / l;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t hh[4], ll[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(s[8 * c + 2 * j] * inv, s[8 * c + 2 * j + 1] * inv, hh[j], ll[j]);
        *reinterpret_cast<uint4*>(smem + PV_OFF_P + swz(r, c)) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<uint4*>(smem + PV_OFF_P + PV_TILE_Q + swz(r, c)) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
      }
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- 4. O = P V
    if (warp == 0 && elect_one()) {
      tc_fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t ko = kk * 32;
        const uint64_t phd = umma_desc_sw128(sP + ko), pld = umma_desc_sw128(sP + PV_TILE_Q + ko);
        const uint64_t vh = umma_desc_sw128(sV + ko), vl = umma_desc_sw128(sV + 8192 + ko);
        umma_bf16(tmem_base + 64, pld, vh, idesc_o, kk != 0 ? 1u : 0u);
        umma_bf16(tmem_base + 64, phd, vl, idesc_o, 1u);
        umma_bf16(tmem_base + 64, phd, vh, idesc_o, 1u);
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after_sync();
    {
      float o[kDh];
#pragma unroll
      for (int c = 0; c < kDh / 16; ++c) {
        float v[16];
        tmem_ld16(t_o + 16 * c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[16 * c + j] = v[j];
      }
      if (valid) {
        uint4* oh = reinterpret_cast<uint4*>(orow + h * kDh);
        uint4* ol = reinterpret_cast<uint4*>(orow + p.lo_off + h * kDh);
#pragma unroll
        for (int c = 0; c < kDh / 8; ++c) {
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split2(o[8 * c + 2 * j], o[8 * c + 2 * j + 1], hh[j], ll[j]);
          oh[c] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          ol[c] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();   // every thread has read S and O of this head: the next head may overwrite the tiles and TMEM
  }
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace

bool attention_p2v_supported(const AttnParams& p) {
  return p.Lk == kV && p.Lq >= 1 && (p.q_ld % 4) == 0 && (p.kv_ld % 4) == 0 && (p.q_col % 4) == 0 && (p.k_col % 4) == 0 &&
         (p.v_col % 4) == 0 && (p.out_ld % 8) == 0 && (p.lo_off % 8) == 0;
}

cudaError_t launch_attention_p2v(const AttnParams& p, cudaStream_t s) {
  if (!attention_p2v_supported(p)) return cudaErrorInvalidValue;
  static DeviceOnce attr;
  cudaError_t e = once_per_device(attr, [&] {
    return cudaFuncSetAttribute(attn_p2v_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PV_SMEM);
  });
  if (e != cudaSuccess) return e;
  const int tiles = (p.Lq + 127) / 128;
  attn_p2v_tc_kernel<<<p.num_seq * tiles, PV_THREADS, PV_SMEM, s>>>(p, tiles);
  return cudaGetLastError();
}

}  // namespace ct3
