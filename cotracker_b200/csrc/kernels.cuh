// kernels.cuh -- launchers of the non-GEMM kernels of the update loop (definitions in *.cu).
#pragma once
#include "common.cuh"

namespace ct3 {

// ---- prep.cu : per-clip preparation -------------------------------------------------------------
struct PyramidLayout {
  int64_t off[kL];  // float offset of each level inside the pyramid buffer
  int h[kL], w[kL];
  int64_t total;
};
PyramidLayout pyramid_layout(int T, int H4, int W4);
cudaError_t launch_prepare_pyramid(const float* fmaps, int T, int H4, int W4, float* pyr, cudaStream_t s);
cudaError_t launch_sample_support(const float* pyr, int T, int H4, int W4, const int32_t* qframes,
                                  const float* qcoords, int N, const uint8_t* acc_mask, float* support,
                                  cudaStream_t s);

// pools levels 1..3 from an already normalised channels-last level 0 living in `pyr`
cudaError_t launch_pyramid_pools(int T, int H4, int W4, float* pyr, cudaStream_t s);

// ---- enc_tail.cu : conv2 -> InstanceNorm -> ReLU -> conv3 of the encoder on the GEMM engine --------
cudaError_t launch_im2col3x3_split(const float* in, int T, int C, int H, int W, int Kpad, __nv_bfloat16* out,
                                   cudaStream_t s);
// scratch: instnorm_scratch_bytes(T, C) bytes (fp64 partial sums of the two-stage reduction)
size_t instnorm_scratch_bytes(int T, int C);
cudaError_t launch_instnorm_stats(const float* y, int T, int HW, int C, float eps, float* stats, void* scratch,
                                  cudaStream_t s);
cudaError_t launch_instnorm_relu_split(const float* y, const float* stats, int64_t rows, int HW, int C,
                                       __nv_bfloat16* out, cudaStream_t s);
cudaError_t launch_l2norm_rows(const float* in, int64_t rows, float* out, cudaStream_t s);
cudaError_t launch_upsample_concat(const float* const src[4], const int c[4], const int h[4], const int w[4], int T,
                                   int H, int W, float* out, cudaStream_t s);

// ---- enc_front.cu : the CNN encoder's convolutions, channels-last, on the tensor cores --------------------
// conv1 7x7/2 pad 3 (3 -> 64), fp32 SIMT: frames [T,3,H,W] -> out [T,Ho,Wo,64] NHWC (+ bias)
cudaError_t launch_conv_stem(const float* frames, const float* w, const float* bias, int T, int H, int W, float* out,
                             cudaStream_t s);
// 3x3 stride-1 pad-1 implicit-GEMM convolution: x_split [T*H*W, 2*C] (hi C | lo C), w_split [Cout, 2*9*C] packed by
// launch_pack_conv (tap-major K), out fp32 NHWC [T,H,W,Cout] (+ bias).  C, Cout multiples of 64.
cudaError_t launch_conv3x3_tc(const __nv_bfloat16* x_split, const __nv_bfloat16* w_split, const float* bias, int T,
                              int H, int W, int C, int Cout, float* out, int num_sms, cudaStream_t s);
// [Cout, Cin, taps] fp32 -> split [Cout_pad, 2*taps*Cp] with K = tap*Cp + c, zero padded
cudaError_t launch_pack_conv(const float* w, int Cout, int Cin, int taps, int Cp, int Cout_pad, __nv_bfloat16* out,
                             cudaStream_t s);
// stride-2 gather (taps 9: 3x3 pad 1; taps 1: 1x1) of a split NHWC activation into GEMM rows [T*Ho*Wo, 2*taps*C]
cudaError_t launch_gather_s2(const __nv_bfloat16* x_split, int T, int H, int W, int C, int taps, __nv_bfloat16* out,
                             cudaStream_t s);
// InstanceNorm + ReLU (+ residual: mode 1 x fp32; mode 2 x = 1x1/2 conv output normalised with stats_d) on NHWC rows
cudaError_t launch_norm_act(const float* y, const float* stats, const float* x, const float* stats_d, int mode,
                            int64_t rows, int HW, int C, float* out_f32, __nv_bfloat16* out_split, cudaStream_t s);
// bilinear (align_corners) resize of 4 NHWC fp32 sources (c channels used, cs channel stride) + concat -> split [rows, 2*Cp]
cudaError_t launch_upsample_concat_split(const float* const src[4], const int c[4], const int cs[4], const int h[4],
                                         const int w[4], int T, int Cp, int H, int W, __nv_bfloat16* out, cudaStream_t s);

// ---- corr.cu : correlation sampling ---------------------------------------------------------------
// vol_split [N*T*4, 2*kVolPad] bf16, row (n*T+t)*4+level
// impl: 0 tensor cores (correlate-then-interpolate when pyr_split is given and every level is >= 8x8: corr_tc3.cu for
//         mode 2, corr_tc2.cu for modes 3 / 1; else corr_tc.cu),
//       1 exact-fp32 SIMT, 2 corr_tc.cu always, 3 like 0 but corr_tc2.cu for every mode (A/B of the two kernels)
// mode / vol16 apply to the corr_tc2.cu path only (corr_uses_patch_kernel): products per correlation FLOP (3|2|1;
// pyr_split must have been made with the same mode) and a single-fp16-plane volume [N*T*4, kVolPad] instead of the
// split one; the other kernels always compute in fp32 / bf16x3 and write the split volume.
bool corr_uses_patch_kernel(int impl, bool have_pyr_split, int T, int H4, int W4);
cudaError_t launch_corr_sample(const float* pyr, const __nv_bfloat16* pyr_split, int H4, int W4, const float* support,
                               const uint8_t* track_valid, const float* coords, int T, int N,
                               __nv_bfloat16* vol_split, int impl, int mode, int vol16, int num_sms, cudaStream_t s);

cudaError_t launch_corr_sample_tc(const float* pyr, int H4, int W4, const float* support,
                                  const uint8_t* track_valid, const float* coords, int T, int N,
                                  __nv_bfloat16* vol_split, int num_sms, cudaStream_t s);

// corr_tc2.cu: correlate-then-interpolate on a split-bf16 copy of the pyramid
//   pyr_split: per level at bf16 offset 2*off[l]: [plane hi|lo][T][H][W][128]   (same bytes as the fp32 pyramid)
bool corr_patch_supported(int T, int H4, int W4);
//   mode 3: [plane hi|lo][T][H][W][128] bf16;  mode 1/2: one fp16 plane [T][H][W][128] at the same level offset
cudaError_t launch_split_pyramid(const float* pyr, int T, int H4, int W4, __nv_bfloat16* pyr_split, int mode,
                                 cudaStream_t s);
cudaError_t launch_corr_patch_tc(const __nv_bfloat16* pyr_split, int H4, int W4, const float* support,
                                 const uint8_t* track_valid, const float* coords, int T, int N,
                                 __nv_bfloat16* vol_split, int mode, int vol16, int num_sms, cudaStream_t s);

// corr_tc3.cu: the production kernel -- same algorithm with the MMA transposed (supports = M side), one fp16 texel
// plane (pyr_split made with mode 1 or 2), supports split fp16 (mode 2) or one fp16 plane (one_product, mode 1)
cudaError_t launch_corr_patch_t(const __nv_bfloat16* pyr_half, int H4, int W4, const float* support,
                                const uint8_t* track_valid, const float* coords, int T, int N,
                                __nv_bfloat16* vol, int vol16, int one_product, int num_sms, cudaStream_t s);

// ---- tokens.cu : elementwise / row-wise pieces of the transformer ---------------------------------
cudaError_t launch_layernorm_split(const float* x, int rows, const float* gamma, const float* beta, float eps,
                                   __nv_bfloat16* out_split, cudaStream_t s);
// LayerNorm fold helpers: split copy + partial row statistics of fp32 token rows; rowsum / affine fold at pack time
cudaError_t launch_rowstats_split(const float* x, int rows, __nv_bfloat16* raw_split, float* stat_part, cudaStream_t s);
cudaError_t launch_rowsum(const float* w, int N, int K, float* out, cudaStream_t s);
cudaError_t launch_affine_fold(const float* w, const float* b, const float* gamma, const float* beta, int N, int K,
                               float* w2, float* b2, cudaStream_t s);
cudaError_t launch_build_x_small(const float* coords, const float* vis, const float* conf, int T, int N,
                                 __nv_bfloat16* x_split, cudaStream_t s);
cudaError_t launch_init_virtual(float* tokens, const float* virt, int T, int N, cudaStream_t s);
// delta_out == nullptr: in-place state update; else write delta [N,T,4] and leave the state alone
cudaError_t launch_heads(const float* tokens, const float* w4, const float* b4, float* coords, float* vis,
                         float* conf, float* delta_out, int T, int N, cudaStream_t s);
cudaError_t launch_row_bias(const float* time_emb, const float* w_in, int T, float* out, cudaStream_t s);
// fp32 [rows,K] -> split [rows, 2*Kpad]; perm_x: apply the X column permutation (x_src_col)
// fp16 != 0: the planes hold IEEE fp16 (hi = fp16(x), lo = fp16(x - hi)) instead of bf16
cudaError_t launch_split_rows(const float* x, int rows, int K, int Kpad, int perm_x, __nv_bfloat16* out,
                              int64_t dst_row_off, cudaStream_t s, int fp16 = 0);

// ---- attention.cu -------------------------------------------------------------------------------
struct AttnParams {
  const float* q;  int64_t q_ld;  int q_col;
  const float* kv; int64_t kv_ld; int k_col, v_col;
  __nv_bfloat16* out; int64_t out_ld; int lo_off;
  int num_seq, Lq, Lk;
  int64_t q_seq_stride, q_tok_stride;  // q/out row = s*q_seq_stride + i*q_tok_stride
  int64_t k_seq_stride, k_tok_stride;  // k/v row  = s*k_seq_stride + j*k_tok_stride
  float scale;
};
cudaError_t launch_attention(const AttnParams& p, cudaStream_t s);   // exact-fp32 SIMT (verification)

// ---- attention_p2v.cu : point <- virtual cross attention (Lk == 64 keys) on tcgen05 ----------------------------------
bool attention_p2v_supported(const AttnParams& p);
cudaError_t launch_attention_p2v(const AttnParams& p, cudaStream_t s);

// ---- attention_tc.cu : tensor-core (mma.sync split-bf16x3) production path ------------------------
constexpr int kAttnMaxSplits = 32;
size_t attention_partial_bytes(int num_seq, int Lq, int max_splits);
cudaError_t launch_attention_tc(const AttnParams& p, bool per_warp, float* part, int num_sms, cudaStream_t s);

}  // namespace ct3
