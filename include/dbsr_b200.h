/*
 * dbsr_b200.h -- C ABI of libdbsr_b200.so: hand-written sm_100a kernels for the DBSR burst forward pass.
 *
 * Drop-in boundary.  The reference (Tony-Tseng/deep-rawburst-sr) has no FFI of its own: its only native
 * code are CUDA-C strings launched through cupy with raw `data_ptr()` integers
 * (external/pwcnet/correlation/correlation.py:293-322).  This header is the compiled replacement for that
 * seam and for every library op on the hot path `DBSRNet.forward` (models/dbsr/dbsrnet.py:33-38).  Host
 * code (Python, `deep_rawburst_sr_b200/`) allocates all device memory with torch and passes raw pointers,
 * sizes and the current `cudaStream_t`; nothing in here allocates device memory or synchronises.
 *
 * Conventions
 *  - Every entry point returns 0 on success, non-zero on error; `dbsr_last_error()` returns the message of
 *    the last failing call on the calling thread.
 *  - Activations inside the path are NHWC ("channels-last") views described by `dbsr_nhwc_t`; a view may
 *    be a channel slice [c_off, c_off + c) of a wider buffer (c_pitch channels per pixel).  That is how
 *    the PWC-Net dense concatenations (models/alignment/pwcnet.py:171-177) are written in place.
 *  - The module seams of the reference are NCHW fp32; `dbsr_nchw_to_nhwc` / `dbsr_nhwc_to_nchw` convert.
 *  - `stream` is a `cudaStream_t` passed as `void*`.
 *  - No entry point falls back to the CPU: on a device that is not sm_100 `dbsr_device_check` fails and
 *    the Python layer refuses to run.
 */
#ifndef DBSR_B200_H_
#define DBSR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBSR_B200_VERSION 100

enum { DBSR_F32 = 0, DBSR_BF16 = 1 };
enum { DBSR_ACT_NONE = 0, DBSR_ACT_RELU = 1, DBSR_ACT_LRELU = 2 /* slope 0.1 */ };

/* NHWC activation view (element (n,y,x,ch) at data[((n*h + y)*w + x)*c_pitch + c_off + ch]). */
typedef struct dbsr_nhwc {
  void*   data;
  int32_t n, h, w;
  int32_t c;        /* channels in the view                          */
  int32_t c_off;    /* first channel of the view inside a pixel      */
  int32_t c_pitch;  /* channels per pixel of the underlying buffer   */
  int32_t dtype;    /* DBSR_F32 | DBSR_BF16                          */
} dbsr_nhwc_t;

/* -------------------------------------------------------------------------------------------------- */
/* probes                                                                                             */
/* -------------------------------------------------------------------------------------------------- */
int         dbsr_version(void);
const char* dbsr_last_error(void);
/* 0 iff `device` is a compute-capability 10.x part (B200); anything else is refused (no fallback).   */
int         dbsr_device_check(int device);

/* -------------------------------------------------------------------------------------------------- */
/* layout converters (module seams; NCHW fp32 <-> NHWC view)                                          */
/* -------------------------------------------------------------------------------------------------- */
int dbsr_nchw_to_nhwc(const float* src, const dbsr_nhwc_t* dst, void* stream);
int dbsr_nhwc_to_nchw(const dbsr_nhwc_t* src, float* dst, void* stream);
/* dst[p] = src[map(p)]: map(p) = p if group == 0, else (p / group) * src_group  (reference frame of the
 * burst replicated over its pairs; replaces `x_rgb[:, :1].repeat(...)`, models/dbsr/encoders.py:53).  */
int dbsr_copy_channels(const dbsr_nhwc_t* src, const dbsr_nhwc_t* dst, int32_t group, int32_t src_group,
                       int32_t src_first, void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* burst preparation: replaces encoders.py:52 (RGGB->RGB) + pwcnet.py:262-271 (bilinear resize to x64) */
/*   burst  [B*N, 4, H, W] fp32 NCHW                                                                  */
/*   enc_in [B*N, H, W, >=4]  packed RAW, channels-last, extra channels zero-filled                   */
/*   pwc_in [B*N, Hp, Wp, >=3] RGB resized to (Hp, Wp), extra channels zero-filled                    */
/* -------------------------------------------------------------------------------------------------- */
int dbsr_prep_burst(const float* burst, int32_t frames, int32_t H, int32_t W, const dbsr_nhwc_t* enc_in,
                    const dbsr_nhwc_t* pwc_in, void* stream);
/* The same for the bf16 tensor-core PWC-Net path with the space-to-depth of the extractor's first stride-2 convolution
 * (dbsr_space_to_depth2 above) folded in: the resized RGB image is written directly as
 *   pwc_s2d[n, Y, X, (2p + q) * 3 + c] = rgb[n, 2Y + p, 2X + q, c]   bf16, dense [frames, Hp/2, Wp/2, 12 (+4 zero pad)]
 * so the fp32 [Hp, Wp] image never exists in HBM.  Hp, Wp: the multiples of 64 of pwcnet.py:263-264.                 */
int dbsr_prep_burst_s2d(const float* burst, int32_t frames, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                        const dbsr_nhwc_t* enc_in, const dbsr_nhwc_t* pwc_s2d, void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* convolution, CUDA-core path (exact fp32 accumulate; any Cin/Cout/stride/dilation)                   */
/* replaces nn.Conv2d + activation + residual of models/layers/blocks.py:46-96 and pwcnet.py:49-204    */
/*   w: fp32 [k*k][Cin][Cout]  (tap-major, Cout contiguous) -- see dbsr_pack_conv_weight_direct         */
/*   y = act(conv(x) + bias (+ residual))                                                              */
/*   shuffle_r > 1: output channel co = (c*r + i)*r + j is written to pixel (r*y+i, r*x+j), channel c   */
/*   of the (r*H, r*W) output view (nn.PixelShuffle folded in, models/layers/upsampling.py:57).         */
/* -------------------------------------------------------------------------------------------------- */
typedef struct dbsr_conv {
  dbsr_nhwc_t x, y;
  dbsr_nhwc_t residual;      /* data == NULL: none; same geometry as y                              */
  const void* w;             /* packed weights (layout depends on the entry point)                  */
  const float* bias;         /* NULL: none                                                          */
  int32_t ksize;             /* 1 or 3                                                              */
  int32_t stride, dilation;  /* padding is dilation*(ksize-1)/2 ("same" for stride 1)               */
  int32_t act;
  int32_t shuffle_r;         /* 0/1: plain; 8: pixel-shuffle scatter                                */
  int32_t grid_limit;        /* dbsr_conv2d_tc*: cap of the persistent grid (CTAs); 0 = one per SM.  Per call, so
                                that concurrent engines / devices never share launch state.                */
  int32_t residual_group;    /* dbsr_conv2d_tc: 0 / 1: residual image i belongs to output image i; g > 1: output image i
                                takes residual image i / g (one map per burst broadcast over its g frames: the
                                per-burst term of the fusion weight predictor's first conv, merging.py:108-112)  */
  int32_t flags;             /* DBSR_CONV_* bits                                                     */
} dbsr_conv_t;
/* flags: w and bias are constants of the caller (uploaded before, never written by a kernel that precedes this launch on the
 * stream).  The tensor-core kernels then start fetching weights while the PRECEDING kernel is still running (programmatic
 * dependent launch: only the activation loads and the stores wait for it).  Leave it clear when a kernel of the same stream
 * produced the weights (the reference has no such layer; the parity tests pack weights right before a call).            */
enum { DBSR_CONV_STATIC_WEIGHTS = 1 };
int dbsr_conv2d_direct(const dbsr_conv_t* p, void* stream);

/* tcgen05 / TMEM implicit-GEMM path (bf16 operands, fp32 accumulate in tensor memory), stride 1.
 *   x: bf16 NHWC view, c_off and c_pitch multiples of 8 (16-byte rows for TMA); channels beyond x.c are
 *      zero-filled by TMA, so Cin needs no padding in memory.
 *   y: bf16 or fp32 view (any channel count / offset; aligned multiples of 16 channels take a vector path)
 *   residual: bf16 or fp32 view with the geometry of y.
 *   w: bf16 [k*k][cout_pad][kpad] K-major with (ck, kpad, n_tile, cout_pad) = dbsr_conv2d_tc_geometry(Cin, Cout);
 *      rows >= Cout and columns >= Cin are zero.  shuffle_r = 8: rows permuted to (i, j, c) order.
 *   Returns non-zero for shapes it does not cover; the Python engine decides per layer at plan time
 *   (dbsr_conv2d_tc_supported) and never falls back silently inside a call.                              */
int dbsr_conv2d_tc(const dbsr_conv_t* p, void* stream);
int dbsr_conv2d_tc_supported(const dbsr_conv_t* p);
/* p->grid_limit caps the persistent grid of THIS launch (0 = one CTA per SM).  The engine lowers it for the encoder conv
 * stack while PWC-Net runs on a second stream, so that the alignment kernels always find free SMs.  The library keeps no
 * mutable launch state: every entry point is re-entrant per device (host-side caches are per device).                  */
/* The same convolution with the decoder's 1x1 predictor + ReLU (models/dbsr/decoders.py:52,61: conv_block(post_conv_dim, 3,
 * 1, activation) after the last post-res block) folded into the epilogue: every epilogue thread owns all (<= 32) output
 * channels of its pixel, so  pred[n, k, y, x] = relu(pred_b[k] + sum_c pred_w[k][c] * act(conv(x) + bias (+ residual))[c])
 * is computed in fp32 registers and stored as fp32 NCHW; the map y itself is NOT written (p->y only gives the geometry).
 *   Needs Cout <= 32 on a map wider than 8 pixels, a residual (if any) that the kernel accumulates on the tensor core
 *   (bf16, Cout == 32), 1 <= pred_c <= 4.  pred_w: fp32 [pred_c][Cout], pred_b: fp32 [pred_c] -- HOST arrays: the <= 132
 *   values are copied into the kernel parameters (constant bank), the only non-device pointers of this ABI.
 *   pred_q14 = 1: `pred` is an int16 [n, pred_c, h, w] buffer and receives (int16)(min(value, 1) * 2^14) -- the 14-bit
 *   quantisation the reference's evaluation / result writers apply to the prediction (evaluation/burstsr/compute_score.py:
 *   110-111, and the save_results.py writers), which halves the device-to-host / gather bytes.                            */
int dbsr_conv2d_tc_predictor(const dbsr_conv_t* p, const float* pred_w, const float* pred_b, int32_t pred_c, void* pred,
                             int32_t pred_q14, void* stream);
/* Fused residual block of the decoder's high-resolution stage (models/layers/blocks.py:84-96 as used by the four
 * `post_res_layers` of ResPixShuffleConv, models/dbsr/decoders.py:45-50, 59):
 *     y = relu(x + conv2(relu(conv1(x) + b1)) + b2),  3x3 / stride 1 / zero padding, 32 -> 32 channels
 * in ONE launch: the intermediate map stays in shared memory (bf16, rounded exactly where the two-launch path rounds it),
 * so the block reads x once and writes y once.  Results are bit-identical to dbsr_conv2d_tc(conv1) + dbsr_conv2d_tc(conv2,
 * residual = x).
 *   x, y: bf16 NHWC views of 32 channels (c_off, c_pitch multiples of 8, 16-byte aligned), same geometry, NOT aliased.
 *   w1, w2: bf16 [9][32][32] as dbsr_conv2d_tc packs a 32 -> 32 3x3 kernel; b1, b2: fp32 [32], 16-byte aligned.
 *   pred != NULL: the decoder's 1x1 predictor + ReLU (decoders.py:52, 61) is applied in the epilogue as in
 *   dbsr_conv2d_tc_predictor (pred_w [pred_c][32], pred_b [pred_c]: HOST arrays; pred: fp32 or, with pred_q14, int16
 *   [n, pred_c, h, w]); y is then not written and may be empty.                                                       */
typedef struct dbsr_resblock {
  dbsr_nhwc_t x, y;
  const void*  w1;
  const float* b1;
  const void*  w2;
  const float* b2;
  const float* pred_w;
  const float* pred_b;
  void*        pred;
  int32_t      pred_c;
  int32_t      pred_q14;
  int32_t      grid_limit;   /* cap of the persistent grid (CTAs); 0 = one per SM */
  int32_t      flags;        /* DBSR_CONV_STATIC_WEIGHTS: w1, b1, w2, b2 are constants of the caller */
} dbsr_resblock_t;
int dbsr_resblock32_tc(const dbsr_resblock_t* p, void* stream);
int dbsr_resblock32_tc_supported(const dbsr_resblock_t* p);
/* the same quantisation as a stand-alone pass for the paths without the fused epilogue: dst[i] = (int16)(clamp(src[i],0,1)*2^14) */
int dbsr_quantize_q14(const float* src, int16_t* dst, int64_t count, void* stream);
/* tiling chosen for (Cin, Cout): K chunk (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B), padded K, UMMA N, padded Cout */
int dbsr_conv2d_tc_geometry(int32_t cin, int32_t cout, int32_t* ck, int32_t* kpad, int32_t* n_tile,
                            int32_t* cout_pad);

/* y[n, Y, X, (p*2 + q)*C + c] = x[n, 2Y + p, 2X + q, c], zero beyond the image; dtype converted to y's.
 *   Turns the stride-2 3x3 convs of the PWC-Net extractor (pwcnet.py:49-97) into stride-1 3x3 convs over 4C channels
 *   that dbsr_conv2d_tc covers (weights repacked on the host: taps (ky', kx') in {0,1}^2, the other five are zero). */
int dbsr_space_to_depth2(const dbsr_nhwc_t* x, const dbsr_nhwc_t* y, void* stream);

/* ConvTranspose2d(k=4, s=2, p=1), Cout = 2 (pwcnet.py:119-120 netUpflow / netUpfeat).
 *   w: fp32 [4][4][2][Cin]; y, y2: [n, 2h, 2w, 2] views (y2 optional second destination, data NULL ok) */
int dbsr_deconv4x4s2(const dbsr_nhwc_t* x, const float* w, const float* bias, const dbsr_nhwc_t* y,
                     const dbsr_nhwc_t* y2, void* stream);
/* The same transposed convolution split for wide inputs (netUpfeat, Cin = 529 .. 565): the channel contraction runs
 * as a 1x1 convolution Cin -> 32 (dbsr_conv2d_tc / dbsr_conv2d_direct, output channel (ky*4 + kx)*2 + oc) and this
 * entry scatters the taps: y_t[n, oy, ox, oc] = bias_t[oc] + sum of the <= 4 taps that land on (oy, ox).
 *   taps: [n, h, w, 32];  y_t: [n, 2h, 2w, 2].
 *   flow != NULL (data != NULL): additionally netUpflow (pwcnet.py:119) of the 2-channel flow [n, h, w, 2] with
 *   wf fp32 [4][4][2][2] ([ky][kx][oc][ic]) and bias_f, written to y_f and (optional) y_f2.                  */
int dbsr_deconv_col2im(const dbsr_nhwc_t* taps, const float* bias_t, const dbsr_nhwc_t* y_t, const dbsr_nhwc_t* flow,
                       const float* wf, const float* bias_f, const dbsr_nhwc_t* y_f, const dbsr_nhwc_t* y_f2,
                       void* stream);
/* The decoders' flow heads (netSix: 3x3, Cin ~ 600 -> 2, pwcnet.py:150) in the same split form: the channel contraction runs
 * as a 1x1 convolution Cin -> 18 planes, ftaps[n, y, x, (ky*3 + kx)*2 + oc] = sum_ic x[n, y, x, ic] * w[oc, ic, ky, kx] (ONE
 * N = 32 tensor-core MMA per K step instead of nine N = 16 ones; it shares its launch with netUpfeat's 32 planes, which read
 * the same concat buffer), and
 *   dbsr_flow_from_taps       y[n, y, x, oc] = bias[oc] + sum of the taps inside the map  (the level-2 flow, fp32 [n, h, w, 2]);
 *   dbsr_deconv_col2im_ftaps  dbsr_deconv_col2im with the coarser flow given as its 18 tap planes `ftaps` + `bias6` instead of
 *                             a map (`flow` NULL): netUpflow sums the taps on the fly, the coarser flow map is never written. */
int dbsr_flow_from_taps(const dbsr_nhwc_t* ftaps, const float* bias, const dbsr_nhwc_t* y, void* stream);
int dbsr_deconv_col2im_ftaps(const dbsr_nhwc_t* taps, const float* bias_t, const dbsr_nhwc_t* y_t, const dbsr_nhwc_t* flow,
                             const dbsr_nhwc_t* ftaps, const float* bias6, const float* wf, const float* bias_f,
                             const dbsr_nhwc_t* y_f, const dbsr_nhwc_t* y_f2, void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* PWC-Net cost volume: replaces correlation.FunctionCorrelation (correlation.py:280-330, 3 launches +  */
/* 2 padded temporaries), the LeakyReLU after it (pwcnet.py:161,169) and, when flow != NULL, backwarp   */
/* (pwcnet.py:16-38) of the second feature map fused into the halo staging.                            */
/*   out[p, 9*(dy+4)+(dx+4), y, x] = act( (1/C) sum_c f1[i1(p),c,y,x] * f2w[i2(p),c,y+dy,x+dx] )        */
/*   pair -> image mapping: group == 0: i1 = i2 = p; else burst b = p / group, i1 = b*(group+1),        */
/*   i2 = b*(group+1) + 1 + p % group  (frame 0 of each burst is the reference).                        */
/*   flow: [pairs, h, w, 2] fp32 view or data NULL; flow_scale = fltBackwarp (pwcnet.py:121)            */
/* -------------------------------------------------------------------------------------------------- */
int dbsr_corr81(const dbsr_nhwc_t* f1, const dbsr_nhwc_t* f2, const dbsr_nhwc_t* flow, float flow_scale,
                const dbsr_nhwc_t* out, int32_t pairs, int32_t group, int32_t act, int32_t mode, void* stream);
/* mode (per call): DBSR_CORR_AUTO: bf16 maps with C in {32, 64, 96, 128} larger than 8x8 run the banded product on the tensor
 * cores (warp-level mma.sync, bf16 operands, fp32 accumulate; a fused backwarp rounds the warped map to bf16);
 * DBSR_CORR_CUDA_CORES keeps every shape on the CUDA-core kernels (fp32 arithmetic on the bf16 inputs): A/B + tests.  */
enum { DBSR_CORR_AUTO = 0, DBSR_CORR_CUDA_CORES = 1 };
/* The same call with the first map also written into a second view: f1_copy[p] = f1[i1(p)], [pairs, h, w, C] of f1's dtype --
 * the `tenFirst` slice of the decoder's concat buffer (torch.cat([tenVolume, tenFirst, tenFlow, tenFeat], 1), pwcnet.py:173).
 * The tensor-core and small-map kernels store it from the tile they stage anyway (one launch less per pyramid level); the
 * other paths run dbsr_copy_channels on the same stream.  f1_copy == NULL or data == NULL: dbsr_corr81.                  */
int dbsr_corr81_copy(const dbsr_nhwc_t* f1, const dbsr_nhwc_t* f2, const dbsr_nhwc_t* flow, float flow_scale,
                     const dbsr_nhwc_t* out, const dbsr_nhwc_t* f1_copy, int32_t pairs, int32_t group, int32_t act,
                     int32_t mode, void* stream);

/* flow head: replaces pwcnet.py:274-279.  flow4 [P, h4, w4, 2] fp32 -> offsets [P, 2, H, W] fp32 NCHW  */
/*   offsets = 20 * bilinear_resize(flow4 -> (H, W)) * (W/Wp, H/Hp)                                     */
int dbsr_flow_head(const dbsr_nhwc_t* flow4, float* offsets, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                   void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* fusion                                                                                              */
/* -------------------------------------------------------------------------------------------------- */
/* warp (models/layers/warp.py:19-46): out[f] = bilinear(feat[f], (x + fx, y + fy)), zeros outside.
 *   offsets: [P, 2, H, W] fp32 NCHW.  frames == 0: plain mode, image f uses offsets[f].
 *   frames = N > 0: burst mode over B*N images: frame 0 of each burst is copied (the reference frame,
 *   models/dbsr/merging.py:72 `cat(ref_feat, oth_feat)`), frame n > 0 uses offsets[b*(N-1) + n-1].       */
int dbsr_warp(const dbsr_nhwc_t* feat, const float* offsets, const dbsr_nhwc_t* out, int32_t frames, void* stream);
/* merging.py:91-105: [B*N, H, W, >=2] <- (frame 0: zeros; others: offsets mod 1.0, floor-mod)          */
int dbsr_offsets_mod(const float* offsets, const dbsr_nhwc_t* out, int32_t bursts, int32_t frames,
                     float modulo, void* stream);
/* merging.py:79-89: wp_in[:, 0:C] = proj[b, 0]; wp_in[:, C:2C] = proj[b, n] - proj[b, 0]               */
int dbsr_build_wp_input(const dbsr_nhwc_t* proj, const dbsr_nhwc_t* wp_in, int32_t frames, void* stream);
/* merging.py:72-89 with the warp of encoders.py:80 folded in.  The 1x1 projection commutes with the bilinear warp,
 * so q = W_p * feat is computed on the UNWARPED embeddings (conv2d, no bias / activation) and this kernel produces
 *   p_n = relu(warp(q_n, offsets) + bias)  (frame 0 unwarped);  wp_in[:, 0:C] = p_0,  wp_in[:, C:2C] = p_n - p_0.
 * offsets == NULL: q is already aligned (WeightedSum called on pre-warped embeddings).                           */
int dbsr_warp_proj(const dbsr_nhwc_t* q, const float* bias, const float* offsets, const dbsr_nhwc_t* wp_in,
                   int32_t frames, void* stream);
/* The same with the weight predictor's first convolution SPLIT into a per-frame and a per-burst part:
 *   W [p_0 | p_n - p_0 | e_n] = W_d p_n + W_o e_n + (W_b - W_d) p_0      (merging.py:108-112: the last term is per burst)
 * so this form writes wp_in[:, 0:C] = p_n for every frame (frame 0: p_0) and p0[b] = p_0 of every burst ([B, H, W, C]); the
 * engine convolves p0 once per burst and adds the result as a broadcast residual (dbsr_conv_t.residual_group).          */
int dbsr_warp_proj_split(const dbsr_nhwc_t* q, const float* bias, const float* offsets, const dbsr_nhwc_t* wp_in,
                         const dbsr_nhwc_t* p0, int32_t frames, void* stream);
/* merging.py:117-124 fused with the warp: fused[b] = sum_n softmax_n(logits[b,n]) * A[b,n] where
 *   A[b,0] = feat[b*N], A[b,n>0] = bilinear(feat[b*N+n], (x,y) + offsets[b*(N-1)+n-1]) gathered on the fly
 *   (offsets == NULL: `feat` already holds the aligned maps).  weights_out (optional, may be NULL):
 *   [B, N, C, H, W] fp32 NCHW -- the reference's `fusion_weights`.                                     */
int dbsr_softmax_wsum(const dbsr_nhwc_t* feat, const dbsr_nhwc_t* logits, const float* offsets,
                      const dbsr_nhwc_t* fused, float* weights_out, int32_t frames, void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* decoder tail                                                                                        */
/* -------------------------------------------------------------------------------------------------- */
/* per-channel 3x3 blur, zero padding (upsampling.py:59-65); k: 9 floats on the host                    */
int dbsr_blur3x3(const dbsr_nhwc_t* x, const dbsr_nhwc_t* y, const float* k9, void* stream);
/* predictor (decoders.py:52,60): pred[n, co, y, x] = relu(b[co] + sum_c w[co][c] x[n,y,x,c]), NCHW fp32 */
int dbsr_predictor(const dbsr_nhwc_t* x, const float* w, const float* bias, int32_t cout, float* pred,
                   void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* evaluation metrics: the multi-GPU run reduces scalars instead of gathering images  */
/* -------------------------------------------------------------------------------------------------- */
/* SSIM (models/loss/msssim.py:22-74 `ssim`, called by msssim.SSIM / MSSSIM and by image_quality_v2.SSIM :104-136).
 *   img1, img2: fp32 NCHW [n, c, h, w], contiguous.  `crop` = boundary_ignore (image_quality_v2.py:112-114), applied as
 *   index arithmetic.  window1d: `window` (1..11) HOST floats = msssim.gaussian(window, 1.5) (:10-12); the 2-D window of
 *   create_window (:15-19) is its outer product and is applied separably, without padding (:44-45, padd = 0).
 *   val_range > 0: L = val_range; <= 0: L is derived on the device from max / min of the cropped img1 (:24-35), no host
 *   synchronisation.  stats: fp32 [n][2] = per-image mean of the ssim map and of the contrast term v1 / v2 (:59).
 *   ssim_map: optional fp32 [n, c, h - 2 crop - window + 1, w - 2 crop - window + 1] (`spatial_out`), or NULL.
 *   workspace: dbsr_ssim_workspace_floats(...) floats of device memory (per-CTA partial sums: the reduction order is
 *   fixed, results are bit-identical from run to run).
 *   valid: optional [n, 1, h, w] bytes (non-zero = valid pixel), the mask of image_quality_v2.SSIM.forward (:127-131: cropped
 *   like the images, then by 5 on each side -- window 11 only); with a mask stats[n] = (sum ssim * valid, sum valid) over the
 *   positions of all channels, un-normalised, so that the caller forms `sum / (count + eps)` per image or per batch.   */
int dbsr_ssim_workspace_floats(int32_t n, int32_t c, int32_t h, int32_t w, int32_t crop, int32_t window);
int dbsr_ssim(const float* img1, const float* img2, int32_t n, int32_t c, int32_t h, int32_t w, int32_t crop,
              const float* window1d, int32_t window, float val_range, const uint8_t* valid, float* workspace, float* stats,
              float* ssim_map, void* stream);
/* F.avg_pool2d(img, (2, 2)) of both images between two MS-SSIM levels (msssim.py:88-89): [planes, h, w] -> [planes, h/2, w/2] */
int dbsr_avgpool2_pair(const float* img1, const float* img2, float* out1, float* out2, int32_t planes, int32_t h, int32_t w,
                       void* stream);
/* per-image mean squared error over the interior (image_quality_v2.py:47-66 with metric 'l2', valid = None), the quantity
 * PSNR.psnr (:75-86) takes the log of.  pred, gt: fp32 NCHW [n, c, h, w]; mse: fp32 [n]; workspace: dbsr_mse_workspace_floats(n).
 * valid (optional, [n, 1, h, w] bytes): mse becomes fp32 [n][2] = (sum valid * err^2, sum valid over pixels and channels), the two
 * sums of the masked form `(err * valid).sum() / (valid.sum() * elem_ratio + eps)` (:60-64).                                */
int dbsr_mse_workspace_floats(int32_t n);
int dbsr_mse_per_image(const float* pred, const float* gt, const uint8_t* valid, int32_t n, int32_t c, int32_t h, int32_t w,
                       int32_t crop, float* workspace, float* mse, void* stream);

/* -------------------------------------------------------------------------------------------------- */
/* synthetic burst generation, inverse camera pipeline (the parameter-deterministic part)  */
/* -------------------------------------------------------------------------------------------------- */
/* data/synthetic_burst_generation.py:59-79 in one pass: invert_smoothstep (camera_pipeline.py:78-81, if `smoothstep`) ->
 * gamma_expansion (:84-87, if `gamma`) -> apply_ccm(rgb2cam) (:96-107) -> safe_invert_gains (:121-136) -> clamp(0, 1).
 *   image, out: fp32 [batch, 3, h, w]; rgb2cam9: 9 HOST floats (row-major); gains3: 3 HOST floats =
 *   [1 / red_gain, 1, 1 / blue_gain] / rgb_gain as the reference forms them in fp32 (:125).                         */
int dbsr_unprocess_rgb(const float* image, float* out, int32_t batch, int32_t h, int32_t w, const float* rgb2cam9,
                       const float* gains3, int32_t smoothstep, int32_t gamma, void* stream);
/* synthetic_burst_generation.py:88-99: mosaic (camera_pipeline.py:139-150, 'rggb') -> add_noise (:178-183) -> clamp(0, 1).
 *   rgb: fp32 [n, 3, h, w] (h, w even); raw: fp32 [n, 4, h/2, w/2]; noise: standard-normal fp32 of raw's shape drawn by the
 *   caller (the reference draws it on the host inside add_noise), or NULL for no noise.                              */
int dbsr_mosaic_noise(const float* rgb, const float* noise, float* raw, int32_t n, int32_t h, int32_t w, float shot_noise,
                      float read_noise, void* stream);
/* single2lrburst (data/synthetic_burst_generation.py:131-246) for given frame transforms, fused: (image * 255).astype(uint8)
 * -> cv2.warpAffine (8-bit, INTER_LINEAR, BORDER_CONSTANT, :211-212) -> border crop (:220-224) -> cv2.resize by 1 / factor (8-bit,
 * INTER_LINEAR, :227-228) -> / 255 (:236); BIT-EXACT against OpenCV's fixed-point arithmetic (restated in oracle/lrburst_oracle.py).
 *   image: fp32 [3, H, W]; inverse_maps: DEVICE doubles [n][6], the inverse of each 2x3 forward matrix as cv::warpAffine forms it;
 *   position_maps: DEVICE floats [n][6], rows 0..1 of the fp32 inverse of the 3x3 matrix (:214-215), needed for `flow`;
 *   burst: fp32 [n, 3, h, w], h = (H - 2 crop) / factor (must divide); flow (optional): fp32 [n, 2, h, w] flow vectors to frame 0
 *   (:229-246); normalize = 1: the image is in [0, 1] (the reference tests image.max() < 2, :154).                       */
int dbsr_single2lrburst(const float* image, int32_t H, int32_t W, const double* inverse_maps, const float* position_maps,
                        int32_t n, int32_t factor, int32_t border_crop, int32_t normalize, float* burst, float* flow, void* stream);

/* Batched forms of the three generator kernels (one launch for `batch` bursts with their OWN random parameters, all in device
 * memory, so that a generator front end samples the parameters of many bursts on the host and uploads them once):
 *   params12: fp32 [batch][12] = rgb2cam (9, row-major) + gains (3) per image;
 *   images [batch, 3, H, W]; inverse_maps [batch * frames_per_image][6] / position_maps likewise: frame k warps image
 *   k / frames_per_image and its flow is relative to the first frame of that image; burst [batch * frames_per_image, 3, h, w];
 *   levels: fp32 [n / frames_per_burst][2] = (shot, read) noise level per burst.  Same arithmetic as the single forms.     */
int dbsr_unprocess_rgb_batch(const float* image, float* out, int32_t batch, int32_t h, int32_t w, const float* params12,
                             int32_t smoothstep, int32_t gamma, void* stream);
int dbsr_single2lrburst_batch(const float* images, int32_t batch, int32_t H, int32_t W, const double* inverse_maps,
                              const float* position_maps, int32_t frames_per_image, int32_t factor, int32_t border_crop,
                              int32_t normalize, float* burst, float* flow, void* stream);
int dbsr_mosaic_noise_batch(const float* rgb, const float* noise, float* raw, int32_t n, int32_t h, int32_t w, const float* levels,
                            int32_t frames_per_burst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DBSR_B200_H_ */
