/*
 * ct3_b200.h -- C ABI of libct3_b200.so: the CoTracker3 iterative update loop
 * (correlation sampling + correlation MLP + EfficientUpdateFormer + delta heads)
 * as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary for the reference's inference hot path
 * (citations are file:line inside facebookresearch/co-tracker):
 *   cotracker/models/core/cotracker/cotracker3_offline.py:139-216   (offline loop body)
 *   cotracker/models/core/cotracker/cotracker3_online.py:187-263    (forward_window loop body)
 * and the per-clip preparation either side of it
 *   cotracker3_offline.py:92-127  (L2-normalise, avg-pool pyramid, support sampling).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     name ends in _host; the caller (PyTorch) owns every allocation, the library
 *     never allocates persistent device memory (scratch = caller workspace);
 *   - every entry point returns 0 on success, a negative CT3_E* code otherwise,
 *     never throws / exits; ct3_last_error() returns a thread-local message;
 *   - all work is enqueued on the given cudaStream_t and is asynchronous w.r.t.
 *     the host; B (batch) is 1 as in the reference (cotracker3_offline.py:135,141).
 *
 * Symbols (all extern "C"):  see the declarations below; tests/test_host_logic.py checks
 * that the built library exports each of them.
 */
#ifndef CT3_B200_H_
#define CT3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ct3_stream_t; /* == cudaStream_t */

enum {
  CT3_OK = 0,
  CT3_EINVAL = -1,   /* bad argument (shape, null pointer, alignment)          */
  CT3_ECUDA = -2,    /* a CUDA runtime/driver call failed (see ct3_last_error) */
  CT3_ENOSPC = -3,   /* workspace / packed buffer too small                    */
  CT3_EUNSUPPORTED = -4
};

/* Model constants fixed by the reference architecture
 * (cotracker3_online.py:43-92, cotracker.py:392-462). */
enum {
  CT3_LATENT = 128,     /* feature channels D                                  */
  CT3_LEVELS = 4,       /* corr_levels                                         */
  CT3_RADIUS = 3,       /* corr_radius -> 7x7 = 49 samples                     */
  CT3_P = 49,
  CT3_VOL = 2401,       /* 49*49 correlation volume per (t,n,level)            */
  CT3_VOL_PAD = 2432,   /* padded to a multiple of 64 for the tensor-core GEMM */
  CT3_HID = 384,        /* transformer width C                                 */
  CT3_HEADS = 8,
  CT3_DHEAD = 48,
  CT3_VIRT = 64,        /* virtual tracks                                      */
  CT3_XDIM = 1110,      /* transformer input width                             */
  CT3_XDIM_PAD = 1152,
  CT3_DEPTH = 3         /* time_depth == space_depth                           */
};

/* Order of the fp32 tensors handed to ct3_pack_weights (names are the
 * state-dict keys of the reference).  Per transformer
 * layer i in [0,3) the block tensors follow in the order listed by
 * ct3_weight_name(). */
int ct3_num_weight_tensors(void);                 /* how many pointers ct3_pack_weights expects     */
const char* ct3_weight_name(int index);           /* state-dict key of tensor #index (NULL if OOR)  */

int ct3_version(void);
const char* ct3_last_error(void);

/* Debug/verification options ("gemm", "corr", "attn": 0 = tensor-core path (default),
 * 1 = SIMT fp32 verification kernel used by the tests to cross-check; "corr" = 2 forces the
 * sample-then-correlate tensor-core kernel that otherwise only serves pyramids with a level below 8x8;
 * "attn" = 2 runs the point<-virtual attention on the mma.sync kernel instead of the tcgen05 one, for A/B). */
int ct3_set_option(const char* name, int value);
int ct3_get_option(const char* name, int* value);
/* Precision switches of the correlation branch ("prec.corr", "prec.fc1": tensor-core products per FLOP, 3 | 2 | 1;
 * DESIGN.md section 2).  Options and the live profiler are per HOST THREAD (thread_local), values are range-checked.
 * ct3_precision_info reports what actually runs for a (T, H4, W4) problem under the calling thread's options:
 * products of the 49x128x49 correlation contraction (cotracker3_offline.py:148-156), of corr_mlp.fc1
 * (blocks.py:61), and the bytes per element of the correlation volume (4 = split bf16 hi|lo, 2 = one fp16 plane). */
int ct3_precision_info(int T, int H4, int W4, int* corr_products, int* fc1_products, int* volume_bytes_per_element);
/* 1 when ct3_corr_sample / the update loop emit volume rows support-major (element k*49 + (a*7+b): corr_tc3.cu, the
 * default kernel) instead of the reference's sample-major (a*7+b)*49 + k (cotracker3_offline.py:148-156); the matching
 * column permutation of corr_mlp.fc1 is applied at pack time, so this only matters to callers of the stage API. */
int ct3_volume_is_support_major(int T, int H4, int W4, int* flag);

/* ---- one-time weight packing ------------------------------------------------
 * Replaces the nn.Module parameter storage read by cotracker3_online.py:73-92.
 * Splits every Linear weight into bf16 hi/lo planes ([out, 2*Kpad], K padded to
 * a multiple of 64), concatenates to_q|to_kv for self-attention blocks, permutes
 * the input_transform columns to the X layout documented in DESIGN.md. */
int ct3_packed_weights_bytes(size_t* out_bytes);
int ct3_pack_weights(const float* const* tensors_host_array_of_device_ptrs, int n_tensors,
                     void* packed, size_t packed_bytes, ct3_stream_t stream);

/* ---- per-clip preparation ---------------------------------------------------
 * ct3_prepare_pyramid: cotracker3_offline.py:92-117 (L2-normalise over channels,
 * 3x avg_pool2d(2,2)).  in: fnet output [T,128,H4,W4] fp32 channel-planar.
 * out: pyr = 4 levels, channels-last [T,Hl,Wl,128] fp32, concatenated; level
 * offsets (in floats) are returned by ct3_pyramid_layout. */
int ct3_pyramid_layout(int T, int H4, int W4, int64_t level_off[4], int level_h[4], int level_w[4],
                       int64_t* total_floats);
int ct3_prepare_pyramid(const float* fmaps, int T, int H4, int W4, float* pyr, ct3_stream_t stream);

/* ct3_sample_support: get_track_feat / sample_features5d
 * (cotracker3_online.py:113-128, model_utils.py:293-323) for all 4 levels.
 * queried_frames [N] int32 (already relative to the window, clamped into [0,T-1]),
 * queried_coords [N,2] fp32 in stride-4 feature units (x,y).
 * support out: [4][49,N,128] fp32 (the reference's [B,49,N,C] layout per level).
 * If accumulate_mask != NULL ([N] uint8) the sampled features are ADDED to
 * `support` where mask!=0 and nothing is written elsewhere (online accumulation,
 * cotracker3_online.py:433-434); otherwise they overwrite. */
int ct3_sample_support(const float* pyr, int T, int H4, int W4, const int32_t* queried_frames,
                       const float* queried_coords, int N, const uint8_t* accumulate_mask,
                       float* support, ct3_stream_t stream);

/* ---- encoder tail (partially): conv2 3x3 (416->256) -> InstanceNorm -> ReLU ->
 * conv3 1x1 (256->128) (BasicEncoder.forward tail, blocks.py:215-218) + L2-normalise + pyramid
 * (cotracker3_offline.py:92-117) on the GEMM engine.  cat: [T,416,H4,W4] fp32 channel-planar = the concatenated,
 * bilinearly resized stage outputs (blocks.py:210-215).  Output = the same channels-last pyramid as
 * ct3_prepare_pyramid.  Weights: conv2.weight [256,416,3,3], conv2.bias, conv3.weight [128,256,1,1], conv3.bias. */
/* bilinear (align_corners=True) resize of the 4 stage outputs [T,Cs,Hs,Ws] (fp32, planar) to H x W and channel concat
 * -> out [T, sum Cs, H, W] (blocks.py:202-215); src/channels/heights/widths are HOST arrays of 4 entries. */
int ct3_upsample_concat(const float* const* src, const int* channels, const int* heights, const int* widths, int T,
                        int H, int W, float* out, ct3_stream_t stream);
int ct3_enc_tail_packed_bytes(size_t* out_bytes);
int ct3_enc_tail_pack(const float* conv2_w, const float* conv2_b, const float* conv3_w, const float* conv3_b,
                      void* packed, size_t packed_bytes, ct3_stream_t stream);
int ct3_enc_tail_workspace_bytes(int T, int H4, int W4, size_t* out_bytes);
int ct3_enc_tail(const void* packed, const float* cat, int T, int H4, int W4, float* pyr, void* workspace,
                 size_t workspace_bytes, ct3_stream_t stream);

/* ---- the hot loop -----------------------------------------------------------
 * ct3_workspace_bytes: scratch needed by ct3_update_loop for a window of T frames of H4 x W4 feature maps and N
 * tracks (includes the split-bf16 copy of the pyramid the correlation kernel reads through TMA);
 * H4 = W4 = 0 sizes it for ct3_updateformer alone. */
int ct3_workspace_bytes(int T, int N, int H4, int W4, size_t* out_bytes);

/* ct3_update_loop: `iters` refinement iterations, in place on the state.
 *   packed   : ct3_pack_weights output
 *   pyr      : ct3_prepare_pyramid output (T frames of the window)
 *   support  : [4][49,N,128] fp32; track_valid (may be NULL) [N] uint8 zeroes the
 *              support of not-yet-queried tracks (cotracker3_online.py:493-496)
 *   coords   : [T,N,2] fp32 stride-4 feature units, in/out
 *   vis,conf : [T,N] fp32 logits, in/out
 *   time_emb : [T,1110] fp32 (buffer already interpolated to T,
 *              cotracker3_online.py:145-156)
 * On return coords/vis/conf hold the state after the last iteration (the caller
 * multiplies coords by the stride and applies sigmoid, cotracker3_offline.py:213-216). */
int ct3_update_loop(const void* packed, const float* pyr, int H4, int W4, const float* support,
                    const uint8_t* track_valid, float* coords, float* vis, float* conf,
                    const float* time_emb, int T, int N, int iters, void* workspace,
                    size_t workspace_bytes, ct3_stream_t stream);

/* ---- live profiler (bench.py roofline): CUDA events around every launch of the library, summed per
 * kernel category: 0 corr_sample, 1 gemm (tcgen05), 2 attention, 3 layernorm, 4 misc.
 * ct3_profile_enable(1) clears and starts recording; ct3_profile_read synchronises and sums. */
int ct3_profile_enable(int on);
int ct3_profile_read(double ms[7], int launches[7], double* gemm_flops);

/* ---- stage-level entry points (used by the parity tests and profiles) ------- */

/* Correlation sampling alone (get_correlation_feat + einsum,
 * cotracker3_online.py:130-143, cotracker3_offline.py:144-156) for all levels:
 * vol_split [N*T*4, 2*2432] bf16: row ((n*T+t)*4+level), hi plane cols [0,2432),
 * lo plane cols [2432,4864); value = hi+lo, cols 2401..2431 are zero. */
/* scratch: ct3_pyramid_layout's total * 4 bytes (256-byte aligned) for the split-bf16 pyramid copy of the
 * correlate-then-interpolate kernel (used when every level is >= 8x8 texels); NULL selects the
 * sample-then-correlate kernel.  With prec.corr < 3 the copy is ONE fp16 plane per level (the first half of each
 * level's region); the default kernel keeps its 4-byte work counter behind level 0's plane, zeroed by the call. */
int ct3_corr_sample(const float* pyr, int H4, int W4, const float* support,
                    const uint8_t* track_valid, const float* coords, int T, int N,
                    void* vol_split, void* scratch, size_t scratch_bytes, ct3_stream_t stream);

/* Generic split-bf16x3 linear layer  Y = act(X W^T + b)  (nn.Linear, blocks.py:61-67)
 *   x_split [M, 2*Kpad] bf16 (hi|lo), w_split [Nout, 2*Kpad] bf16, bias [Nout] fp32 or NULL
 *   act: 0 none, 1 GELU(erf), 2 GELU(tanh);  y fp32 [M, Nout] */
int ct3_linear(const void* x_split, const void* w_split, const float* bias, int M, int Nout,
               int Kpad, int act, float* y, ct3_stream_t stream);

/* Same with the precision switches of the GEMM engine: `products` tensor-core products per FLOP (3: split x split,
 * 2: x_hi x (w_hi + w_lo), 1: x_hi x w_hi) on bf16 (fp16 = 0) or IEEE-fp16 (fp16 = 1) planes. */
int ct3_linear_prec(const void* x_split, const void* w_split, const float* bias, int M, int Nout,
                    int Kpad, int act, int products, int fp16, float* y, ct3_stream_t stream);

/* fp32 [rows, K] -> split bf16 [rows, 2*Kpad] (zero padded); _fp16: the planes hold IEEE fp16 instead */
int ct3_split_rows(const float* x, int rows, int K, int Kpad, void* x_split, ct3_stream_t stream);
int ct3_split_rows_fp16(const float* x, int rows, int K, int Kpad, void* x_split, ct3_stream_t stream);

/* One EfficientUpdateFormer forward (cotracker.py:483-531) on an explicit token
 * input x [N, T, 1110] fp32 (time embedding already added, reference column order);
 * delta out [N, T, 4] fp32. */
int ct3_updateformer(const void* packed, const float* x, int T, int N, float* delta, void* workspace,
                     size_t workspace_bytes, ct3_stream_t stream);

/* ---- the whole CNN encoder (BasicEncoder.forward, blocks.py:190-219; normalise + pyramid,
 * cotracker3_offline.py:92-117) on the tensor-core engine, channels-last ------------------------------------
 * frames [T,3,H,W] fp32 already scaled to [-1,1] (cotracker3_offline.py:63) -> pyr (ct3_pyramid_layout(T, H/4, W/4)).
 * conv1 7x7/2 runs as fp32 SIMT, every other convolution as split-bf16x3 tcgen05 GEMMs (3x3 stride-1: implicit GEMM
 * over TMA-shifted NHWC boxes; strided ones: gather + GEMM); InstanceNorm statistics in fp64.
 * Weight tensors in the order of ct3_encoder_weight_name() (state-dict keys below `fnet.`). */
int ct3_encoder_num_weight_tensors(void);
const char* ct3_encoder_weight_name(int index);
int ct3_encoder_packed_bytes(size_t* out_bytes);
int ct3_encoder_pack(const float* const* tensors_host_array_of_device_ptrs, int n_tensors, void* packed,
                     size_t packed_bytes, ct3_stream_t stream);
int ct3_encoder_workspace_bytes(int T, int H, int W, size_t* out_bytes);
int ct3_encoder(const void* packed, const float* frames, int T, int H, int W, float* pyr, void* workspace,
                size_t workspace_bytes, ct3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CT3_B200_H_ */
