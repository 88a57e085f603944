/*
 * vunet_b200 — C ABI of the B200-native VUnet hot path.
 *
 * Every entry point takes plain device pointers + sizes and a CUDA stream (as void*), launches
 * hand-written sm_100a kernels asynchronously on that stream and returns
 *      0            success
 *     <0            bad argument (see VUB_E_*; vub_last_error() has a message)
 *     >0            a cudaError_t / CUresult raised by the launch
 * Nothing throws across the boundary, the caller owns every buffer, and there is no CPU fallback:
 * without a CUDA device the launchers return an error.
 *
 * Each function cites the reference code (CompVis/behavior-driven-video-synthesis) it replaces.
 * The Python host (`behavior-driven-video-synthesis_b200/`) binds these with ctypes and registers
 * them as torch ops; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Tensor conventions: activations inside the path are NHWC bf16 ("raw" = the tensor as produced,
 * "elu" = ELU(tensor), stored side by side because every consumer wants one or the other, see
 * SURVEY.md App. B); tensors crossing the module boundary are NCHW fp32 like the reference's.
 */
#ifndef VUNET_B200_H_
#define VUNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VUB_ABI_VERSION 5

#define VUB_OK 0
#define VUB_E_NULL (-1)        /* required pointer is NULL */
#define VUB_E_SHAPE (-2)       /* unsupported shape (non power-of-two extent, bad channel count) */
#define VUB_E_FLAGS (-3)       /* inconsistent flag / output combination */
#define VUB_E_ALIGN (-4)       /* pointer not 16-byte aligned */
#define VUB_E_NODRIVER (-5)    /* CUDA driver entry point unavailable (no GPU / no driver) */
#define VUB_E_COMM (-6)        /* NCCL unavailable or returned an error (vub_last_error() has its message) */

/* vub_conv_desc.flags */
#define VUB_CONV_D2S 1u        /* store through DepthToSpace(2) (lib/modules.py:24-34, DCR order) */
#define VUB_CONV_F32_NHWC 2u   /* out_f32 is [B,Ho,Wo,cout] fp32 */
#define VUB_CONV_F32_NCHW 4u   /* out_f32 is [B,cout_valid,Ho,Wo] fp32 */
#define VUB_CONV_ELU_IN0 8u    /* x[0] holds RAW activations: apply ELU while loading (VunetRNB's ELU(x), */
#define VUB_CONV_ELU_IN1 16u   /* lib/modules.py:225-231); same for x[1].  Only where vub_conv_fwd_elu_in_ok() */
#define VUB_CONV_U8_HWC 32u    /* out_f32 points to a uint8 HWC image [B,Ho,Wo,cout_valid]: the reference's video-frame
                                  conversion (scale_img(y) * 255).astype(uint8) (lib/utils.py:658-668,
                                  data/data_conversions_3d.py:1166-1173) fused into out_conv's epilogue.  Only where
                                  vub_conv_fwd_u8_out_ok() */

#define VUB_CONV_RELU 64u      /* the "elu" output (and, in vub_conv_dgrad, the mask) uses ReLU instead of ELU: the frozen
                                  VGG19 feature convs of the perceptual loss (models/imagenet_pretrained.py:46-59,
                                  lib/losses.py:81-119) */

#define VUB_CONV_LRELU 128u    /* the "elu" output holds LeakyReLU(y, 0.01) / tanh(y): the coupling MLPs of the behavior net's */
#define VUB_CONV_TANH 256u     /* latent flow (lib/modules.py:236-257 BasicFullyConnectedNet).  Per-tap kernel only */

/* ABI version of the loaded library (== VUB_ABI_VERSION it was built with). */
int vub_abi_version(void);
/* Message for the last non-zero return on this thread ("" if none). */
const char* vub_last_error(void);

/*
 * Weight-normalised convolution, forward:  NormConv2d.forward (lib/modules.py:140-145) with the
 * callers' surrounding elementwise work fused in:
 *     u      = conv2d(cat[x0, x1], v_hat)            v_hat = v / ||v||  (packed bf16, see vub_wn_prep)
 *     y      = scale * u + shift (+ residual)        scale = gamma*g, shift = gamma*bias + beta
 *     out_raw = bf16(y), out_elu = bf16(ELU(y))      (VunetRNB's act_fn, lib/modules.py:225,229)
 * `x0`/`x1` are the two halves of VunetRNB's torch.cat (lib/modules.py:227) — the concat is never
 * materialised.  ksize 3 => padding 1, ksize 1 => padding 0 (the only combinations the VUnet uses:
 * lib/modules.py:153,170,203,214; models/vunets.py:117,288,523,539).  stride 2 = Downsample
 * (lib/modules.py:148-161).  VUB_CONV_D2S = Upsample's DepthToSpace (lib/modules.py:164-182).
 * Implicit GEMM on tcgen05: M = 128 output pixels, N = cout tile, K = (source, tap, channel).
 */
typedef struct vub_conv_desc {
  const void* x[2];      /* NHWC bf16 [batch, hin, win, c_src]; x[1] NULL when nsrc == 1 */
  const void* w;         /* packed bf16 [cout_pad, nsrc*ksize*ksize*c_src], K order (src, tap, c) */
  const float* scale;    /* [cout_pad] */
  const float* shift;    /* [cout_pad] */
  const void* residual;  /* NHWC bf16 [batch, hout, wout, cout] or NULL */
  void* out_raw;         /* NHWC bf16 or NULL */
  void* out_elu;         /* NHWC bf16 or NULL */
  void* out_f32;         /* fp32, layout per flags, or NULL */
  int32_t nsrc;          /* 1 or 2 */
  int32_t c_src;         /* channels per source: 32, 64 or 128; 256 / 512 (VunetOrg's 4-group prior on the two smallest
                            latent maps: models/vunets.py:655-676) with stride 1 and no D2S; any multiple of 64 up to
                            2048 for a "linear layer" (ksize 1 on a 1x1 map: nn.Linear / nn.LSTMCell of the behavior net) */
  int32_t batch, hin, win;
  int32_t cout;          /* GEMM N = rows of w (padded to 16/32/64/128/256/512; linear layers: any multiple of 128 <= 4096) */
  int32_t cout_valid;    /* channels actually stored (== cout except for out_conv) */
  int32_t ksize;         /* 1 or 3 */
  int32_t stride;        /* 1 or 2 */
  uint32_t flags;
  int32_t stages;        /* smem pipeline depth, 0 = library default */
  const void* w_tiled;   /* optional tile-major copy of w (vub_wn_entry.w_tiled, c_src % 64 == 0), or NULL */
} vub_conv_desc;

int vub_conv_fwd(const vub_conv_desc* d, void* stream);
/* 1 when a conv of this geometry can store VUB_CONV_U8_HWC frames (the row-streaming kernel takes it), else 0 */
int vub_conv_fwd_u8_out_ok(int32_t nsrc, int32_t c_src, int32_t cout, int32_t ksize, int32_t stride, int32_t hin,
                           int32_t win);

/*
 * Pure host query: 1 when vub_conv_fwd accepts VUB_CONV_ELU_IN0/1 for a conv of this geometry (it is served by
 * the row-streaming kernel, which can apply ELU to its input rows in shared memory), else 0.  Lets the caller
 * decide whether the producer of a tensor has to store an ELU copy at all.
 */
int vub_conv_fwd_elu_in_ok(int32_t nsrc, int32_t c_src, int32_t cout, int32_t ksize, int32_t stride, int32_t hin,
                           int32_t win);

/*
 * Weight-norm preparation for many layers in one launch: replaces the per-forward
 * `torch.nn.utils.weight_norm` hook of every NormConv2d (lib/modules.py:135-138) and folds
 * gamma/beta/bias (lib/modules.py:144) into one fp32 scale/shift pair per output channel.
 * `table` is a DEVICE array of n_entries entries; entry e covers rows
 * [row_start, row_start + cout) of the launch grid; total_rows = sum of cout.
 */
typedef struct vub_wn_entry {
  const float* v;      /* conv.weight_v [cout, cin, k, k] */
  const float* g;      /* conv.weight_g [cout] */
  const float* gamma;  /* gamma [cout] */
  const float* beta;   /* beta [cout] */
  const float* bias;   /* conv.bias [cout] */
  void* w_packed;      /* bf16, first row of this layer in its packed matrix (row pitch ktot) */
  float* scale;        /* [cout] */
  float* shift;        /* [cout] */
  int32_t cout, cin, ktaps, c_src; /* cin = nsrc * c_src; ktaps = k*k */
  int32_t ktot;        /* row pitch of the packed matrix in elements (= cin*ktaps) */
  int32_t row_start;
  /* data-gradient weights (training): wd[(src*c_src + c)][tap*wd_cout + wd_col0 + co] =
   * bf16(scale[co] * v_hat[co, src, tap, c]) — the matrix vub_conv_dgrad contracts dy with.  NULL = skip. */
  void* wd;
  int32_t wd_cout;     /* K extent per tap of wd (output channels of the whole slot, padded to 32) */
  int32_t wd_col0;     /* first column of this conv inside a tap block (merged convs) */
  /* fp32 validation path: w32[(src*ktaps + tap)*c_src + c][wd_col0 + co] = v_hat (row pitch wd_cout floats).
   * NULL = skip.  w_packed may be NULL when only w32 is wanted. */
  float* w32;
  /* Tile-major copies for the tensor-core-bound layers (c_src a multiple of 64; NULL = skip): the same values as
   * w_packed / wd, stored as the exact shared-memory image of a SWIZZLE_128B operand tile so that a kernel can fetch
   * a weight tile with ONE 1-D bulk copy (measured 73 B/clk/SM, and it overlaps a tensor box; a 128-row tensor box
   * costs 384 cycles = 42.7 B/clk/SM: profiles/r01_exp_bulk.log):
   *   w_tiled [K/64][tiled_rows][64]   element (row = wd_col0 + co, k):  block k/64, 16-byte chunk ((k%64)/8) ^ (row%8)
   *   wd_tiled[ktaps*wd_cout/64][cin][64]  element (row = src*c_src + c, k = tap*wd_cout + wd_col0 + co): likewise */
  void* w_tiled;
  void* wd_tiled;
  int32_t tiled_rows;  /* rows of the whole slot in w_tiled (= padded output channels) */
  int32_t pad_;
} vub_wn_entry;

int vub_wn_prep(const vub_wn_entry* table_dev, int32_t n_entries, int32_t total_rows, void* stream);

/*
 * First 1x1 NormConv2d of EncUp / DecUp (models/vunets.py:117-119,137 and 230-232,250) fused with the
 * layout/dtype change: x NCHW fp32 [batch, cx, h, w] -> NHWC bf16 [batch, h, w, cout] raw and ELU.
 * K = cx (3 or 30) is not a tensor-core shape: CUDA-core kernel, HBM bound.  Takes the layer's raw
 * parameters (weight-norm done in-kernel).  cout must be 32 or 64.
 */
int vub_stem_fwd(const float* x_nchw, const float* v, const float* g, const float* gamma,
                 const float* beta, const float* bias, int32_t batch, int32_t cx, int32_t h,
                 int32_t w, int32_t cout, void* out_raw, void* out_elu, void* stream);
/*
 * Same layer fed with the uint8 HWC image [batch, h, w, cx] the reference holds BEFORE its `ToTensor()` +
 * `x * 2 - 1` transforms (data/__init__.py:15-43; the rendering loop's stickman: data/data_conversions_3d.py:1154-1163).
 * The kernel applies that conversion itself (fp32 (u / 255) * 2 - 1, IEEE division): bit-identical outputs to
 * vub_stem_fwd on the converted tensor, a quarter of the input bytes over PCIe and HBM.
 */
int vub_stem_fwd_u8(const uint8_t* x_nhwc_u8, const float* v, const float* g, const float* gamma,
                    const float* beta, const float* bias, int32_t batch, int32_t cx, int32_t h,
                    int32_t w, int32_t cout, void* out_raw, void* out_elu, void* stream);

/*
 * Latent head: EncDownAlter's squash + reparametrize (models/vunets.py:575,594-597) and the KL
 * summand of lib/losses.py:283-291, from the fused (mu || logit) conv output.
 *   ml_nhwc  fp32 [batch, h, w, 2c]  (cols [0,c) = make_latent_params, [c,2c) = make_logstds)
 *   eps_nchw fp32 [batch, c, h, w]   (torch.randn_like draw, reference order)
 *   logstd = sigmoid(logit); z = mu + exp(logstd) * eps
 * Writes mu/logstd/z NCHW fp32 (returned to the caller), ELU(z) and ELU(mu) NHWC bf16 (next convs'
 * operands; either may be NULL) and one partial sum of (-logstd + 0.5*(exp(2 logstd) + mu^2)) per
 * block into kl_partial[0 .. vub_latent_num_partials(...)).
 */
int32_t vub_latent_num_partials(int32_t batch, int32_t c, int32_t h, int32_t w);
int vub_latent_fwd(const float* ml_nhwc, const float* eps_nchw, int32_t batch, int32_t c, int32_t h,
                   int32_t w, float* mu_nchw, float* logstd_nchw, float* z_nchw, void* elu_z_nhwc,
                   void* elu_mu_nhwc, float* kl_partial, void* stream);

/*
 * Layout converters at the module boundary (the reference is NCHW fp32 throughout):
 * NCHW fp32 -> NHWC bf16 raw and/or ELU (DecDownAlter's prior sample, models/vunets.py:352-354),
 * and NHWC bf16 -> NCHW fp32 for callers that want reference-layout activations.
 */
int vub_nchw_f32_to_nhwc_bf16(const float* src, int32_t batch, int32_t c, int32_t h, int32_t w,
                              void* dst_raw, void* dst_elu, void* stream);
int vub_nhwc_bf16_to_nchw_f32(const void* src, int32_t batch, int32_t c, int32_t h, int32_t w,
                              float* dst, void* stream);

/*
 * Frame epilogue of the video-rendering loop: (scale_img(rgb) * 255).permute(0,2,3,1).astype(uint8)
 * (lib/utils.py:658-668, data/data_conversions_3d.py:1166-1173) — NCHW fp32 [batch, 3, h, w] in [-1, 1] ->
 * NHWC uint8 [batch, h, w, 3], bit-exact with the reference expression.  Quarters the device->host bytes.
 */
int vub_frames_to_uint8(const float* img_nchw, int32_t batch, int32_t h, int32_t w, uint8_t* out_nhwc, void* stream);

/* Same converter with the channel axis zero-padded to c_pad (multiple of 8) — the image gradient
 * d(loss)/d(imgs) [batch, 3, h, w] entering the backward pass of dd.out_conv (models/vunets.py:411). */
int vub_nchw_f32_to_nhwc_bf16_pad(const float* src, int32_t batch, int32_t c, int32_t h, int32_t w,
                                  int32_t c_pad, void* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 validation path (BASELINE.json north_star: relative L2 <= 1e-4 against the reference in an
 * fp32 validation mode).  Same fused conv as vub_conv_fwd with fp32 NHWC activations, fp32 weights
 * (vub_wn_entry.w32) and fp32 FMA accumulation on CUDA cores; any channel counts (the stems included).
 * Output layout: NHWC [batch, hout, wout, cout]; VUB_CONV_D2S: shuffled NHWC; VUB_CONV_F32_NCHW: NCHW
 * (the image).  A checker for the fused schedule, not a fast path.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vub_conv_f32_desc {
  const float* x[2];       /* NHWC fp32 [batch, hin, win, c_src] */
  const float* w;          /* [nsrc*ksize*ksize*c_src][w_pitch] fp32, 16-byte aligned */
  const float* scale;
  const float* shift;
  const float* residual;   /* same layout as the output, or NULL */
  float* out_raw;
  float* out_elu;
  int32_t nsrc, c_src, batch, hin, win, cout, w_pitch, ksize, stride;
  uint32_t flags;
} vub_conv_f32_desc;

int vub_conv_fwd_f32(const vub_conv_f32_desc* d, void* stream);
int vub_nchw_f32_to_nhwc_f32(const float* src, int32_t batch, int32_t c, int32_t h, int32_t w, float* dst_raw,
                             float* dst_elu, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training half of the path.  The reference gets these from autograd (convolution_backward, ELU /
 * add / cat / sigmoid / exp backward kernels, aten::_weight_norm_interface_backward); here each conv
 * needs exactly two tensor-core launches (data gradient, weight correlation) plus one reduction
 * launch per step for every layer's weight-norm / affine parameter gradients.
 * ---------------------------------------------------------------------------------------------- */

/*
 * Data gradient of a NormConv2d (autograd of lib/modules.py:143-144 w.r.t. x) with the backward of the
 * surrounding VunetRNB elementwise ops fused into the epilogue:
 *     dx_s = ELU'(x_s) * conv_transpose(dy, scale * v_hat)_s  + addend_s + addend2_s
 * ELU' is rebuilt from the stored ELU(x_s) copy (e > 0 ? 1 : e + 1; mask_elu[s] NULL = the conv read
 * the raw tensor); addend / addend2 carry gradient contributions of other consumers (the residual
 * identity path, skip connections) and may alias dx_s (in-place accumulation).
 * dy: NHWC bf16 [batch, hout, wout, cout_k]; with VUB_CONV_D2S the gradient of the SHUFFLED tensor
 * [batch, 2*hin, 2*win, cout_k/4].  stride 2 (Downsample) is computed per input-pixel parity class.
 */
typedef struct vub_dgrad_desc {
  const void* dy;
  const void* wd;            /* [nsrc*c_src, ksize*ksize*cout_k] bf16 (vub_wn_entry.wd) */
  void* dx[2];               /* NHWC bf16 [batch, hin, win, c_src] per source; c_src 32 / 64 / 128, or 256 / 512 (stride 1, no
                                D2S: one launch per 128-channel slice of every source) */
  const void* mask_elu[2];
  const void* addend[2];
  const void* addend2[2];
  int32_t nsrc, c_src;
  int32_t batch, hin, win;   /* extents of the conv INPUT */
  int32_t cout_k;            /* K extent per tap = output channels padded to a multiple of 32 */
  int32_t ksize, stride;
  uint32_t flags;            /* VUB_CONV_D2S */
  int32_t pad_;
  const void* wd_tiled;      /* optional tile-major copy of wd (vub_wn_entry.wd_tiled, cout_k % 64 == 0), or NULL */
} vub_dgrad_desc;

int vub_conv_dgrad(const vub_dgrad_desc* d, void* stream);

/*
 * Weight correlation of a NormConv2d (the raw part of convolution_backward w.r.t. the weight):
 *     gt[split][co][(src, tap, c)] = sum_{pixels of the split} dy[pix, co] * x_src[pix + tap offset, c]
 *     dy_sum[split][co]            = sum_{pixels of the split} dy[pix, co]
 * on tcgen05 with both operands MN-major (pixels are the contraction index), fp32.  The pixel range is split
 * over CTAs; every split WRITES its own slice (plain stores: no atomics, bit-reproducible, nothing to zero) and
 * vub_wn_bwd sums the slices.  vub_conv_wgrad_splits returns the split count that fills the GPU for a given
 * shape (pointers ignored); the caller sizes gt / dy_sum for the n_split it passes.
 * x[s] is the tensor the forward conv read (the ELU'd copy inside a VunetRNB).
 */
typedef struct vub_wgrad_desc {
  const void* dy;            /* as in vub_dgrad_desc */
  const void* x[2];          /* NHWC bf16 [batch, hin, win, c_src]; c_src 32 / 64 / 128, or 256 / 512 (stride 1, no D2S: one
                                launch per 128-channel slice of every source, each writing its own rows of gt) */
  float* gt;                 /* [n_split][cout_k][gt_pitch] (slices split_stride floats apart) */
  float* dy_sum;             /* [n_split][cout_k] (slices sum_stride floats apart) */
  int64_t split_stride;
  int32_t sum_stride;
  int32_t n_split;           /* 1 .. min(#64-pixel tiles, 1024) */
  int32_t nsrc, c_src;
  int32_t batch, hin, win;
  int32_t cout_k;            /* channels of dy (multiple of 32; out_conv: 32 with 3 valid) */
  int32_t gt_pitch;          /* row pitch of gt in floats: >= nsrc*ksize*ksize*c_src */
  int32_t ksize, stride;
  uint32_t flags;            /* VUB_CONV_D2S */
} vub_wgrad_desc;

int32_t vub_conv_wgrad_splits(const vub_wgrad_desc* d);
int vub_conv_wgrad(const vub_wgrad_desc* d, void* stream);

/*
 * Parameter gradients of every NormConv2d from its weight correlation, one launch for all layers
 * (replaces aten::_weight_norm_interface_backward + the mul/add backward of lib/modules.py:144).
 * With u = conv(x, v_hat), out = s*u + t, s = gamma*g, t = gamma*bias + beta, G = gt[:, co], S = dy_sum[co]:
 *     d_beta = S, d_bias = gamma*S, d_g = gamma*(v_hat.G), d_gamma = g*(v_hat.G) + bias*S,
 *     d_v = (gamma*g/||v||) * (G - v_hat*(v_hat.G))
 * accumulate != 0 adds to the gradient buffers instead of overwriting them.
 */
typedef struct vub_wnb_entry {
  const float* v;
  const float* g;
  const float* gamma;
  const float* bias;
  const float* gt;           /* partial correlations: element (split, co, k) at
                                gt[split*split_stride + (gt_col0 + co)*gt_pitch + k] */
  const float* dy_sum;       /* element (split, co) at dy_sum[split*sum_stride + gt_col0 + co] */
  float* d_v;
  float* d_g;
  float* d_gamma;
  float* d_beta;
  float* d_bias;
  int32_t cout, cin, ktaps, c_src;
  int32_t gt_pitch, gt_col0;
  int32_t row_start;
  int32_t n_split;           /* slices to sum (>= 1) */
  int64_t split_stride;
  int32_t sum_stride;
  int32_t pad;
} vub_wnb_entry;

/* max_len: the largest cin * ktaps among the entries (the kernel stages one weight row of gradients in shared
 * memory); <= 0 means 2304, the largest VunetAlter layer; at most 9216 (VunetOrg's 1024 x 3 x 3 prior convs). */
int vub_wn_bwd(const vub_wnb_entry* table_dev, int32_t n_entries, int32_t total_rows, int32_t accumulate,
               int32_t max_len, void* stream);

/*
 * Fused optimizer step: torch.optim.Adam (no weight decay, no amsgrad; the reference uses
 * betas = (0.5, 0.9), experiments/shape_and_pose_net.py:237-246) on beta / gamma / bias / weight_g / weight_v of
 * every listed NormConv2d, fused with the weight-norm preparation (vub_wn_prep) of the UPDATED weights for the
 * next step: one pass over the parameters, one launch.  `wn` describes the conv exactly like vub_wn_prep's entry
 * (prepared outputs may be NULL, e.g. for the stems); gradients and moment buffers have the parameters' shapes;
 * `step` is the 1-based step count (bias correction); `lr` is per entry (parameter groups).
 */
typedef struct vub_adam_entry {
  vub_wn_entry wn;
  const float *d_v, *d_g, *d_gamma, *d_beta, *d_bias;
  float *m_v, *m_g, *m_gamma, *m_beta, *m_bias;   /* exp_avg */
  float *s_v, *s_g, *s_gamma, *s_beta, *s_bias;   /* exp_avg_sq */
  float lr;
  int32_t pad;
} vub_adam_entry;

int vub_adam_wn_step(const vub_adam_entry* table_dev, int32_t n_entries, int32_t total_rows, float beta1, float beta2,
                     float eps, int32_t step, void* stream);
/* Same with the step count in device memory: increments *step_dev, then uses it for the bias corrections — the
 * launch arguments no longer change from step to step, so a captured CUDA graph of the training step can be
 * replayed (training.GraphedTrainStep). */
int vub_adam_wn_step_dev(const vub_adam_entry* table_dev, int32_t n_entries, int32_t total_rows, float beta1,
                         float beta2, float eps, int32_t* step_dev, void* stream);

/*
 * Backward of the stem (first 1x1 NormConv2d on the NCHW fp32 input, vub_stem_fwd): correlation of the
 * output gradient with the input, gt[co][ci] += sum dy[pix, co]*x[pix, ci], dy_sum[co] += sum dy.
 * The stem needs no data gradient (its input is the image).  gt pitch = cx.
 */
int vub_stem_bwd(const float* x_nchw, const void* dy_nhwc, int32_t batch, int32_t cx, int32_t h, int32_t w,
                 int32_t cout, float* gt, float* dy_sum, void* stream);

/*
 * Backward of the latent head (vub_latent_fwd): z = mu + exp(ls)*eps, ls = sigmoid(logit).
 *   d_mu_tot = d_mu + dz,  d_ls_tot = d_ls + dz*eps*exp(ls),  d_logit = d_ls_tot*ls*(1-ls)
 * d_mu / d_ls: NCHW fp32 gradients w.r.t. the returned means / logstds (NULL = zero); dz: NHWC bf16
 * gradient w.r.t. z from the decoder (NULL = zero).  Writes d_ml NHWC bf16 [batch, h, w, 2c], the
 * gradient of the fused (mu || logit) conv output.
 */
int vub_latent_bwd(const float* mu_nchw, const float* logstd_nchw, const float* eps_nchw, const float* d_mu_nchw,
                   const float* d_ls_nchw, const void* dz_nhwc, int32_t batch, int32_t c, int32_t h, int32_t w,
                   void* d_ml_nhwc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 validation mode, training half: the same backward schedule with fp32 activations / gradients and fp32
 * v / ||v|| matrices on CUDA cores (north_star: gradients within 1e-4 of the reference "in an fp32-accumulate
 * validation mode").  The weight-norm / affine parameter gradients come from the SAME vub_wn_bwd launch as in the
 * bf16 path (its inputs were fp32 all along).  Checkers, not fast paths.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vub_dgrad_f32_desc {
  const float* dy;           /* NHWC fp32 [batch, hout, wout, cout]; VUB_CONV_D2S: [batch, 2*hin, 2*win, cout/4] */
  const float* w;            /* [nsrc*ksize*ksize*c_src][w_pitch] fp32 v / ||v|| (the forward matrix) */
  const float* scale;        /* gamma * g per output channel */
  float* dx[2];              /* NHWC fp32 [batch, hin, win, c_src] per source */
  const float* mask_elu[2];  /* stored ELU(x_s) or NULL */
  const float* addend[2];    /* may alias dx */
  const float* addend2[2];
  int32_t nsrc, c_src, batch, hin, win, cout, w_pitch, ksize, stride;
  uint32_t flags;            /* VUB_CONV_D2S */
} vub_dgrad_f32_desc;
int vub_conv_dgrad_f32(const vub_dgrad_f32_desc* d, void* stream);

typedef struct vub_wgrad_f32_desc {
  const float* dy;
  const float* x[2];
  float* gt;                 /* [n_split][cout rows of gt_pitch floats]: partial correlations, K order (src, tap, ch) */
  float* dy_sum;             /* [n_split][cout] */
  int32_t nsrc, c_src, batch, hin, win, cout, ksize, stride;
  uint32_t flags;            /* VUB_CONV_D2S */
  int32_t n_split, gt_pitch;
  int64_t split_stride, sum_stride;  /* floats between consecutive splits of gt / dy_sum */
} vub_wgrad_f32_desc;
int vub_conv_wgrad_f32(const vub_wgrad_f32_desc* d, void* stream);

int vub_latent_bwd_f32(const float* logstd_nchw, const float* eps_nchw, const float* d_mu_nchw, const float* d_ls_nchw,
                       const float* dz_nhwc, int32_t batch, int32_t c, int32_t h, int32_t w, float* d_ml_nhwc,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rendering front-end: batched stickman rasteriser, the GPU counterpart of `make_joint_img`
 * (lib/utils.py:325-513: cv2.fillPoly for the body polygon, cv2.line with thickness 1 for the limbs, three uint8
 * planes) as called per frame by the rendering loop (data/data_conversions_3d.py:1154-1160).  Output: uint8 HWC
 * images [batch, height, width, 3] — exactly what vub_stem_fwd_u8 consumes.  Bit-exact with the reference function
 * (tests/golden/stickman.npz) and with cv2 for joints anywhere inside or outside the image.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vub_stickman_model {
  int32_t n_body, body[8];                 /* joint_model.body (polygon, > 2 joints to be drawn) */
  int32_t n_right, right[16][2];           /* joint_model.right_lines */
  int32_t n_left, left[16][2];             /* joint_model.left_lines */
  int32_t n_head, head[8][2];              /* joint_model.head_lines (empty: throat from the shoulder midpoint) */
  int32_t n_face, face[16][2];             /* joint_model.face (drawn when shorter than the throat) */
  int32_t rshoulder, lshoulder, headup;
  int32_t colored;                         /* 0: line_colors=None (right: plane 1 / 255, left: plane 0 / 255, head: planes 0+1 / 127) */
  int32_t right_chan[16], right_col[16];   /* colored: plane and value of line_colors[0][i] */
  int32_t left_chan[16], left_col[16];     /* line_colors[1][i] */
  int32_t head_chan[16], head_col[16];     /* line_colors[2][i] (the face loop indexes the same table from 0) */
} vub_stickman_model;
int vub_stickman_render(const double* joints_xy /* [batch][n_joints][2] */, int32_t batch, int32_t n_joints,
                        const vub_stickman_model* model, int32_t height, int32_t width, uint8_t* out_hwc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VunetOrg family: latent bookkeeping of `latent_sample` (models/vunets.py:151-156) and of the
 * 4-group autoregressive prior in DecDown.forward (704-760) with its SpaceToDepth / DepthToSpace regrouping
 * (lib/modules.py:11-34).  In NHWC, group (gi, gj) of SpaceToDepth(x) is the parity sub-grid x[:, gi::2, gj::2, :].
 * ---------------------------------------------------------------------------------------------- */
/*
 * z = p + eps for the parameter map `p_nhwc` [batch, h, w, c] fp32 (the output of make_latent_params / auto_lp) and
 * noise `eps_nchw` [batch, c, h, w] (NULL: z = p, the posterior mean fed by `transfer`).  sub == 1: the map is the
 * whole latent; sub == 2: it is prior group (gi, gj) and every output marked "merged" is the (2h x 2w) map of
 * `__merge_groups` (DepthToSpace of the concatenated groups), of which this call fills one parity sub-grid.
 * Outputs, each optional: mean_nchw / z_nchw fp32 [batch, c, h*sub, w*sub] (merged; what the module returns),
 * z_raw_nhwc / z_elu_nhwc [batch, h*sub, w*sub, c] (merged; the activations the next convs read), grp_elu_nhwc
 * [batch, h, w, c] = ELU(z) of this map alone (the `feedback` of test_forward).  Activations are bf16, or fp32 when
 * f32_acts != 0 (fp32 validation mode).
 */
int vub_latent_sample(const float* p_nhwc, const float* eps_nchw, int32_t batch, int32_t c, int32_t h, int32_t w,
                      int32_t sub, int32_t gi, int32_t gj, float* mean_nchw, float* z_nchw, void* z_raw_nhwc,
                      void* z_elu_nhwc, void* grp_elu_nhwc, int32_t f32_acts, void* stream);
/*
 * Backward of vub_latent_sample (what autograd does for `mean + eps`, torch.split / DepthToSpace in the reference):
 * d_p_nhwc [batch, h, w, c] = d_mean_nchw (gradient of the returned, merged mean; NULL = none) + dz_nhwc (gradient of
 * the merged NHWC activation z; NULL = none), both gathered from this map's parity positions (sub, gi, gj as above).
 * d_p is what vub_conv_dgrad / vub_conv_wgrad of the parameter conv take as dy.
 */
int vub_latent_sample_bwd(const float* d_mean_nchw, const void* dz_nhwc, int32_t batch, int32_t c, int32_t h, int32_t w,
                          int32_t sub, int32_t gi, int32_t gj, void* d_p_nhwc, int32_t f32_acts, void* stream);
/*
 * SpaceToDepth(2) of an NHWC activation `src` [batch, h, w, c] (lib/modules.py:11-21):
 *   group_major == 0: dst [batch, h/2, w/2, 4c], channel (gi*2 + gj)*c + ch   (the reference's channel order)
 *   group_major == 1: dst [4][batch, h/2, w/2, c]   (`__split_groups`: torch.split of the above, each group compact)
 * dst_copy receives the values, dst_elu ELU(values); either may be NULL.  inverse != 0 runs the index map backwards
 * (DepthToSpace, lib/modules.py:24-34: `src` is the grouped tensor, dst the [batch, h, w, c] one); accumulate != 0
 * adds to dst_copy instead of overwriting it (training: the backward of a regrouping is the inverse regrouping of
 * the gradient, summed with the tensor's other consumers' contributions).
 */
int vub_space_to_depth(const void* src, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t group_major,
                       int32_t inverse, int32_t accumulate, void* dst_copy, void* dst_elu, int32_t f32_acts,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * VGG19 perceptual loss: `vgg_loss` (lib/losses.py:81-119) over `PerceptualVGG`
 * (models/imagenet_pretrained.py:8-61).  The 3x3 conv + ReLU layers of torchvision's vgg19.features[0:32] run on
 * vub_conv_fwd / vub_conv_dgrad with VUB_CONV_RELU (frozen weights: no weight gradient); these are the memory-bound
 * pieces around them.  Feature maps are NHWC bf16 — or fp32 when f32_acts != 0: the fp32 validation mode, whose convs
 * are vub_conv_fwd_f32 / vub_conv_dgrad_f32 with VUB_CONV_RELU.
 * ---------------------------------------------------------------------------------------------- */
/* x = (x + 1) / 2; x = (x - mean) / std (imagenet_pretrained.py:46-47) of an NCHW fp32 image batch [batch, 3, h, w]:
 * out_nhwc32 = NHWC bf16 padded to 32 channels (conv1_1's input; NULL = skip), out_nchw = the fp32 tensor the module
 * returns as "input" (NULL = skip). */
int vub_vgg_normalize(const float* x_nchw, int32_t batch, int32_t h, int32_t w, void* out_nhwc32, float* out_nchw,
                      int32_t f32_acts, void* stream);
/* "input" tap of vgg_loss: partial[i] = per-block sums of |xn(pred) - xn(target)| (vub_vgg_input_l1_partials blocks;
 * the caller sums them and divides by 3*batch*h*w). */
int32_t vub_vgg_input_l1_partials(int32_t batch, int32_t h, int32_t w);
int vub_vgg_input_l1(const float* pred_nchw, const float* target_nchw, int32_t batch, int32_t h, int32_t w,
                     float* partial, void* stream);
/* Gradient w.r.t. pred: d_pred = (d_xn (gradient of conv1_1's input, NHWC bf16 x 32, or NULL) + g0 * scale0 *
 * sign(xn(pred) - xn(target))) / (2 std_c); g0 = upstream gradient of the "input" loss term (device scalar, NULL = 0),
 * scale0 = loss_weights[0] / (3*batch*h*w). */
int vub_vgg_input_bwd(const float* pred_nchw, const float* target_nchw, const void* d_xn_nhwc32, const float* g0_dev,
                      float scale0, int32_t batch, int32_t h, int32_t w, float* d_pred_nchw, int32_t f32_acts,
                      void* stream);
/* nn.MaxPool2d(kernel_size=2, stride=2) on [batch, h, w, c] -> [batch, h/2, w/2, c]. */
int vub_maxpool2_fwd(const void* x_nhwc, int32_t batch, int32_t h, int32_t w, int32_t c, void* y_nhwc, int32_t f32_acts,
                     void* stream);
/* Its backward fused with the ReLU in front of it and the tap gradient: dx = (x is the first maximum of its window ?
 * dy : 0) * (x > 0) + addend, x = the ReLU output that was pooled [batch, h, w, c], dy [batch, h/2, w/2, c]. */
int vub_maxpool2_relu_bwd(const void* x_nhwc, const void* dy_nhwc, const void* addend_nhwc, int32_t batch, int32_t h,
                          int32_t w, int32_t c, void* dx_nhwc, int32_t f32_acts, void* stream);
/* torch.mean(torch.abs(tf - pf)) of a feature tap: per-block partial sums of |pf - tf| (vub_l1_num_partials(numel)
 * of them), and the gradient w.r.t. the pre-activation of pf = ReLU(u): g * scale * sign(pf - tf) * (pf > 0) with the
 * upstream gradient g in device memory and scale = loss_weights[i] / numel. */
int32_t vub_l1_num_partials(int64_t numel);
int vub_l1_partial(const void* pf, const void* tf, int64_t numel, float* partial, int32_t f32_acts, void* stream);
int vub_l1_relu_grad(const void* pf, const void* tf, int64_t numel, const float* g_dev, float scale, void* out,
                     int32_t f32_acts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Behavior-net sampling (BASELINE.json configs[3]): b = flow.reverse(z) through the 15-block latent
 * flow (models/flow/simple_flow.py:136-180 UnsupervisedTransformer2 -> models/flow/blocks.py:95-128, 531-559), then
 * ResidualBehaviorNet.generate_seq (models/pose_behavior_rnn.py:603-626) with the residual LSTM decoder (463-506).
 * Every matrix product (coupling MLPs, LSTM gates, n_out) is a "linear layer" launch of vub_conv_fwd (ksize 1 on a
 * 1x1 map, VUB_CONV_LRELU epilogues); these are the elementwise pieces between them, on fp32 rows [batch, d].
 * ---------------------------------------------------------------------------------------------- */
/* y[b, j] = x[b, idx[j]] / scale[idx[j]] - loc[idx[j]]: Shuffle.reverse (blocks.py:700-704, idx = backward_shuffle_idx;
 * NULL = identity) fused with the ActNorm.reverse (lib/modules.py:318-330) of the block processed before it (scale /
 * loc NULL = none).  y_bf16 (optional): compact bf16 copy [batch, bf_cols] of columns [bf_col0, bf_col0 + bf_cols). */
int vub_gather_affine(const float* x, const int32_t* idx, const float* scale, const float* loc, int32_t batch, int32_t d,
                      float* y, void* y_bf16, int32_t bf_col0, int32_t bf_cols, void* stream);
/* Reverse affine coupling (blocks.py:311-318): state_out[:, col0 : col0 + d2] = (state_in[:, col0 : col0 + d2] - t) *
 * exp(-tanh(s_raw)), s_raw / t = fp32 [batch, d2] outputs of the two MLPs' last linear layers (the s-net's nn.Tanh is
 * applied here); out_bf16: compact bf16 copy [batch, d2] or NULL.  state_in may equal state_out. */
int vub_couple_inv(const float* state_in, const float* s_raw, const float* t, int32_t batch, int32_t d, int32_t d2,
                   int32_t col0, float* state_out, void* out_bf16, void* stream);
/* nn.LSTMCell from the gate pre-activations [batch, 4*hdim] (chunks i, f, g, o): c updated in place (fp32), h' written
 * as bf16 compact [batch, hdim] and into columns [0, hdim) of the next step's gate input [batch, cat_pitch]. */
int vub_lstm_cell(const float* gates, float* c, int32_t batch, int32_t hdim, void* h_lin_bf16, void* h_cat_bf16,
                  int32_t cat_pitch, float* h_f32, void* stream);   /* h_lin_bf16 / h_f32 optional */
/* The step input of BEncoder's nn.LSTM over a given sequence (pose_behavior_rnn.py:177-178): x_t [batch, n] fp32 with
 * row pitch x_pitch floats (x[:, t] of a [batch, T, n] tensor) as bf16, zero-padded, into columns [hdim, hdim + xpad). */
int vub_lstm_input(const float* x_t, int64_t x_pitch, int32_t batch, int32_t n, int32_t xpad, void* cat_bf16,
                   int32_t cat_pitch, int32_t hdim, void* stream);
/* ResidualRNNDecoder.forward's `out + res, res` (pose_behavior_rnn.py:487-506) for step t: x <- x + v (v fp32
 * [batch, vpitch], first n columns), xs[:, t] = new x, cs[:, t] = old x ([batch, steps, n] fp32), bf16 copy of the new
 * x (zero-padded to xpad) into columns [hdim, hdim + xpad) of the gate input. */
int vub_decoder_step(float* x, const float* v, int32_t batch, int32_t n, int32_t vpitch, int32_t xpad, int32_t t,
                     int32_t steps, float* xs, float* cs, void* x_cat_bf16, int32_t cat_pitch, int32_t hdim, void* stream);
/* generate_seq's start: hidden = cell = b (pose_behavior_rnn.py:612-614), x = x_pose[:, start_frame]. */
int vub_decoder_init(const float* b_lat, const float* x0, int32_t batch, int32_t hdim, int32_t n, int32_t xpad, float* c,
                     float* x, void* cat_bf16, int32_t cat_pitch, void* stream);

/*
 * Data-parallel gradient exchange (SURVEY.md 8b `allreduce_bucket(flat, ncclComm, stream)`): in-place all-reduce of
 * one flat fp32 gradient bucket (average != 0: ncclAvg, i.e. the 1/world scaling happens inside the collective),
 * enqueued on `stream`; capturable in a CUDA graph.  Replaces nn.DataParallel's per-step gather / reduce
 * (experiments/shape_and_pose_net.py:213-214).  `nccl_comm` is an ncclComm_t — from vub_nccl_comm_init or any other
 * owner; NCCL is bound at run time (the copy the process already loaded, else the system libnccl.so.2).
 */
int vub_allreduce_bucket(void* flat, int64_t n_floats, void* nccl_comm, int32_t average, void* stream);
/* Communicator plumbing for hosts that do not own one: rank 0 creates the 128-byte id and ships it to the other
 * ranks by any means (torch.distributed's store in training.BucketReducer). */
int32_t vub_nccl_version(void);  /* NCCL_VERSION_CODE of the bound library, 0 if none */
int vub_nccl_unique_id(void* id128);
int vub_nccl_comm_init(void** comm, int32_t n_ranks, const void* id128, int32_t rank);
int vub_nccl_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* VUNET_B200_H_ */
