/*
 * camvid_b200.h -- C ABI of libcamvid_b200.so: the sm_100a (B200) kernels behind the UNet / SegNet
 * training + eval hot path of weiaicunzai/pytorch-camvid.
 *
 * The reference has no FFI layer: its hot path is a chain of ATen calls made from nn.Modules.  Each entry point
 * below therefore cites the reference call site (file:line in the reference tree) whose ATen op it replaces;
 * INTEGRATION.md shows the ctypes / torch.library binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless the name ends in `_host`. The library never allocates, frees or
 *     retains caller memory; workspaces are passed in (sizes from the cvb_*_workspace_bytes helpers).
 *   - Activations are NHWC bf16 "views": channel stride 1, explicit element strides for N, H, W so that a view can
 *     be a channel slice of a concat buffer or a 44-row window of a 45-row buffer (UNet F.pad, models/unet.py:120-123).
 *     View base pointers must be 16-byte aligned, C a multiple of 8, strides multiples of 8 elements.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - Return value: 0 on success, negative cvb_status on error; cvb_last_error() returns a thread-local message.
 *     Nothing throws across this boundary. All functions are re-entrant (forward is called from the Python main
 *     thread, backward from autograd's per-device worker thread).
 */
#ifndef CAMVID_B200_H_
#define CAMVID_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_ABI_VERSION 4

#if defined(__GNUC__)
#define CVB_API __attribute__((visibility("default")))
#else
#define CVB_API
#endif

typedef enum {
  CVB_OK = 0,
  CVB_ERR_INVALID_ARG = -1,   /* shape / alignment / null pointer */
  CVB_ERR_UNSUPPORTED = -2,   /* valid request the kernels do not cover (e.g. C not multiple of 64 for the GEMMs) */
  CVB_ERR_CUDA = -3,          /* a CUDA runtime / driver call failed */
  CVB_ERR_NO_DEVICE = -4      /* no sm_100 device */
} cvb_status;

/* NHWC bf16 view. Element (n,h,w,c) lives at ptr[n*sn + h*sh + w*sw + c]. */
typedef struct {
  void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
} cvb_view;

CVB_API const char* cvb_last_error(void);
CVB_API int cvb_abi_version(void);
/* Number of SMs of the current device (grid sizing); <0 on error. */
CVB_API int cvb_sm_count(void);
/* Releases library-owned host state (nothing on the device is owned). */
CVB_API void cvb_shutdown(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Layout conversion at the nn.Module boundary (reference tensors are NCHW fp32: train.py:126-128).
 * ------------------------------------------------------------------------------------------------------------- */
/* NCHW fp32 [n,c_src,h,w] -> NHWC bf16 view; channels c_src..dst.c-1 of the view are written as zero. */
CVB_API int cvb_nchw_f32_to_nhwc_bf16(const float* src, int c_src, cvb_view dst, void* stream);
/* NHWC bf16 view (first c_dst channels) -> NCHW fp32 [n,c_dst,h,w]. */
CVB_API int cvb_nhwc_bf16_to_nchw_f32(cvb_view src, float* dst, int c_dst, void* stream);
/* First-layer im2col (models/unet.py:41 / models/segnet.py:24 have Cin=3): NCHW fp32 [n,c_src,h,w] ->
 * NHWC bf16 view with channel k = ci*9 + r*3 + s holding x[n, ci, h+r-1, w+s-1] (zero outside), k >= 9*c_src zero.
 * Turns the K=27 convolution into a single-tap GEMM whose weight matrix is the OIHW tensor flattened. */
CVB_API int cvb_im2col3x3_nchw_f32(const float* src, int c_src, cvb_view dst, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * 3x3 / stride 1 / pad 1 convolution as implicit GEMM on tcgen05 (nn.Conv2d(.,.,3,padding=1): models/unet.py:11,
 * models/segnet.py:8).  bf16 operands, fp32 accumulation in TMEM, TMA-fed.
 * ------------------------------------------------------------------------------------------------------------- */
/* Packs OIHW fp32 weights [cout,cin,3,3] into the bf16 GEMM-B matrix of the forward conv:
 * dst[co][tap][ci] with co padded to cout_pad, ci to cin_pad (zeros), tap = r*3+s.   taps==1: dst[co][k], k=cin*9
 * flattened OIHW padded to cin_pad (pairs with cvb_im2col3x3_nchw_f32). */
CVB_API int cvb_pack_weights_fprop(const float* w, int cout, int cin, int taps, int cout_pad, int cin_pad, void* dst,
                           void* stream);
/* Packs the same weights for the data-gradient conv (aten::convolution_backward input grad):
 * dst[ci][tap'][co] = w[co][ci][2-r][2-s], tap' = r*3+s, ci padded to cin_pad, co to cout_pad. */
CVB_API int cvb_pack_weights_dgrad(const float* w, int cout, int cin, int cout_pad, int cin_pad, void* dst, void* stream);

/* All 3x3 layers of a network in ONE launch (the optimizer touches every weight every step): for each entry reads the
 * OIHW fp32 tensor once and writes both GEMM operands, dst_fprop as cvb_pack_weights_fprop(taps=9) and dst_dgrad as
 * cvb_pack_weights_dgrad (dst_dgrad may be 0). `table` is a DEVICE array of `count` entries. */
typedef struct {
  const float* w;      /* [cout][cin][3][3] fp32 */
  void* dst_fprop;     /* bf16 [cout_pad][9][cin_pad] */
  void* dst_dgrad;     /* bf16 [cin_pad][9][cout_pad] or NULL */
  int64_t cout, cin, cout_pad, cin_pad;
  int64_t reserved;
} cvb_pack_entry;
CVB_API int cvb_pack_weights_batch(const cvb_pack_entry* table, int count, int max_cout_pad, int max_cin_pad, void* stream);

/* Epilogue selection for cvb_conv3x3_fprop. */
typedef struct {
  /* train mode: per-CTA partial sums for BatchNorm batch statistics, fp32 [cvb_conv_stat_rows()][2][y.c]
   * (sum, sum of squares over valid pixels). NULL = not produced. */
  float* stat_partials;
  /* eval mode (BatchNorm folded, models/unet.py:12 with running stats): y = relu?(acc*scale[c] + shift[c]).
   * NULL = store the raw accumulator. */
  const float* scale;
  const float* shift;
  int32_t relu;
  /* data-gradient mode (ABI 2): the output of this call is da, the gradient w.r.t. the activation of the block that
   * precedes it, and that block's BatchNorm+ReLU backward (aten::threshold_backward + native_batch_norm_backward,
   * train.py:131) starts with a reduction over exactly this tensor. With bwd_partials != NULL the kernel emits it from
   * its epilogue: per-CTA partial sums of g = da*[bwd_y*bwd_scale[c]+bwd_shift[c] > 0] and g*bwd_y over the valid pixels,
   * fp32 [cvb_conv_stat_rows()][2][y.c], from the bf16-rounded da it stores -- what cvb_bn_relu_bwd_reduce would read
   * back from memory. bwd_y = that block's raw conv output (same shape as y). Only honoured where
   * cvb_conv3x3_fprop_fuses_bwd_stats() says so; mutually exclusive with stat_partials / scale. */
  cvb_view bwd_y;
  const float* bwd_scale;
  const float* bwd_shift;
  float* bwd_partials;
} cvb_conv_epilogue;

/* Rows of the stat_partials buffer the conv writes (== its grid size; depends only on the device). */
CVB_API int cvb_conv_stat_rows(void);
/* y[n,h,w,co] = sum_{tap,ci} x[n,h+r-1,w+s-1,ci] * wpack[co][tap][ci]   (taps = 9, or 1 for a 1x1 GEMM).
 * x.c must equal cin_pad (multiple of 64), y.c == cout_pad (multiple of 64). The same entry point computes the
 * data gradient when given dy as `x` and the cvb_pack_weights_dgrad matrix. */
CVB_API int cvb_conv3x3_fprop(cvb_view x, const void* wpack, int taps, cvb_view y, const cvb_conv_epilogue* ep, void* stream);
/* 1 if cvb_conv3x3_fprop on these views runs a kernel that implements the bwd_* epilogue (today: the transposed
 * cout_pad = 64 kernel), 0 if the caller has to run cvb_bn_relu_bwd_reduce itself. Pure host query. */
CVB_API int cvb_conv3x3_fprop_fuses_bwd_stats(cvb_view x, cvb_view y, int taps);

/* Weight gradient (aten::convolution_backward weight grad): dw[co][ci][r][s] = sum_{n,h,w} dy[n,h,w,co] *
 * x[n,h+r-1,w+s-1,ci], written as OIHW fp32 [cout,cin,3,3] (taps==1: [cout, cin9] with x the im2col view).
 * `workspace` holds split-K partial tiles; size from cvb_conv3x3_wgrad_workspace_bytes. */
CVB_API int64_t cvb_conv3x3_wgrad_workspace_bytes(cvb_view x, cvb_view dy, int taps);
CVB_API int cvb_conv3x3_wgrad(cvb_view x, cvb_view dy, int taps, float* dw, int cout, int cin, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Optimizer: torch.optim.AdamW(net.parameters(), lr, weight_decay).step(), train.py:100,133, for every
 * parameter tensor in one launch (amsgrad = False, maximize = False; decoupled weight decay).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  float* param;        /* fp32, updated in place */
  const float* grad;   /* fp32 */
  float* exp_avg;      /* fp32 state, updated in place */
  float* exp_avg_sq;   /* fp32 state, updated in place */
  int64_t numel;
} cvb_adamw_entry;
/* Elements one CUDA block updates; the caller cuts every tensor into ceil(numel / this) chunks. */
CVB_API int cvb_adamw_chunk_elems(void);
/* table: DEVICE array of entries; chunks: DEVICE array of n_chunks (entry index, chunk index) int32 pairs.
 * step = the 1-based step count of this update (bias corrections 1 - beta^step). */
CVB_API int cvb_adamw_step(const cvb_adamw_entry* table, const int32_t* chunks, int n_chunks, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int64_t step, void* stream);
/* The same update for use inside a CUDA graph (kernel arguments are frozen at capture; lr under OneCycleLR,
 * train.py:102-104,134, and the bias corrections change every step): the per-step scalar factors are computed on the
 * host by cvb_adamw_factors into 8 floats, copied by the caller to `factors_dev` (stream-ordered, outside the graph)
 * and read by the kernel from device memory. */
CVB_API int cvb_adamw_factors(float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                      float* factors_host);
CVB_API int cvb_adamw_step_dev(const cvb_adamw_entry* table, const int32_t* chunks, int n_chunks, const float* factors_dev,
                       void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * BatchNorm2d (+ReLU) in training mode (nn.BatchNorm2d + nn.ReLU: models/unet.py:12-13, models/segnet.py:9-10).
 * ------------------------------------------------------------------------------------------------------------- */
/* Per-channel partial sums of a view (used when the producer did not emit them): partials fp32 [rows][2][c]. */
CVB_API int cvb_bn_stats(cvb_view y, float* partials, int rows, void* stream);
/* Reduces `rows` partial rows over `count` elements per channel and produces, for the first c channels:
 *   mean, invstd = 1/sqrt(var_biased + eps), scale = gamma*invstd, shift = beta - mean*scale  (fp32 [c] each);
 * if running_mean != NULL updates running stats like torch (momentum, unbiased variance), adding conv_bias to the
 * mean when given (the conv kernels skip the bias: it cancels under batch statistics).
 * Channels [c, c_pad) get scale = shift = 0 so padded lanes stay zero. */
CVB_API int cvb_bn_finalize(const float* partials, int rows, int c, int c_pad, int64_t count, const float* gamma,
                    const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                    float momentum, float eps, float* mean, float* invstd, float* scale, float* shift, void* stream);
/* a = relu(y*scale + shift).
 * `reverse` (ABI 4, here and in the two backward passes below): != 0 walks the tensor from its last pixel to its first --
 * same result, different visiting order, for callers who know which end of the tensor the previous kernel left in L2
 * (the drop-in modules pass 0: measured neutral at their sizes). */
CVB_API int cvb_bn_relu_apply(cvb_view y, const float* scale, const float* shift, cvb_view a, int reverse, void* stream);
/* Backward pass 1: g = da * [y*scale+shift > 0]; partials fp32 [rows][2][c] of (sum g, sum g*y). */
CVB_API int cvb_bn_relu_bwd_reduce(cvb_view da, cvb_view y, const float* scale, const float* shift, float* partials,
                           int rows, int reverse, void* stream);
/* Backward finalize: dgamma, dbeta (fp32 [c]) and the per-channel coefficients (coef fp32 [3][c_pad]) with
 * dy = g*coef0 + y*coef1 + coef2 used by cvb_bn_relu_bwd_apply. */
CVB_API int cvb_bn_bwd_finalize(const float* partials, int rows, int c, int c_pad, int64_t count, const float* gamma,
                        const float* mean, const float* invstd, float* dgamma, float* dbeta, float* coef,
                        void* stream);
/* The network's LAST block fused with the module boundary (ABI 4; cross-layer fusion). Its activation is the fp32 NCHW
 * logits tensor the module returns (models/unet.py:156, models/segnet.py:119): dst[n,c,h,w] = float(bf16(relu(y*scale +
 * shift))) for c < c_dst -- cvb_bn_relu_apply + cvb_nhwc_bf16_to_nchw_f32 in one pass, bit-identical to the pair; the bf16
 * activation is never written. y.c must be 8 or 16 (the narrow class tensor), else CVB_ERR_UNSUPPORTED. */
CVB_API int cvb_bn_relu_apply_nchw_f32(cvb_view y, const float* scale, const float* shift, float* dst, int c_dst,
                               void* stream);
/* ... and the entry of its backward pass (train.py:131, the gradient autograd hands to the module): src = dlogits fp32
 * NCHW [n,c_src,h,w] -> da = bf16 NHWC view (channels >= c_src zero) AND, in the same pass, partials fp32 [rows][2][y.c] of
 * (sum g, sum g*y), g = da * [y*scale+shift > 0] -- cvb_nchw_f32_to_nhwc_bf16 + cvb_bn_relu_bwd_reduce. Grid = rows
 * blocks; y.c must be 8 or 16. */
CVB_API int cvb_nchw_f32_to_nhwc_bf16_bn_reduce(const float* src, int c_src, cvb_view da, cvb_view y, const float* scale,
                                        const float* shift, float* partials, int rows, void* stream);
/* Backward pass 2: dy = (da*[a>0])*coef0 + y*coef1 + coef2  (bf16 view). */
CVB_API int cvb_bn_relu_bwd_apply(cvb_view da, cvb_view y, const float* scale, const float* shift, const float* coef,
                          cvb_view dy, int reverse, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MaxPool2d(2,2) with indices / MaxUnpool2d(2)  (models/unet.py:92, models/segnet.py:79-80).
 * Index code: uint8 per output element, 0..3 = (dh*2+dw) of the FIRST maximum in scan order (0,0),(0,1),(1,0),(1,1),
 * NaN wins over anything and the LAST NaN wins (torch's `val > max || isnan(val)` rule).
 * ------------------------------------------------------------------------------------------------------------- */
/* out[n,ho,wo,c] = max over the 2x2 window (floor mode: ho = h/2, wo = w/2). code may be NULL. */
CVB_API int cvb_maxpool2x2_fwd(cvb_view x, cvb_view out, uint8_t* code, void* stream);
/* Same, with the BatchNorm+ReLU of the producer fused in: a = relu(y*scale+shift) is written to `a` and pooled. */
CVB_API int cvb_bn_relu_maxpool2x2_fwd(cvb_view y, const float* scale, const float* shift, cvb_view a, cvb_view out,
                               uint8_t* code, void* stream);
/* dx[window] = dout at the coded position, 0 elsewhere; rows/cols outside any window (odd h/w) get 0.
 * accumulate != 0: dx += instead (UNet skip connection + pool path). code == NULL: recompute argmax from x. */
CVB_API int cvb_maxpool2x2_bwd(cvb_view dout, const uint8_t* code, cvb_view x_or_null, cvb_view dx, int accumulate,
                       void* stream);
/* cvb_maxpool2x2_bwd fused with cvb_bn_relu_bwd_reduce of the block whose activation was pooled (cross-layer fusion,
 * producer side): writes dx exactly as cvb_maxpool2x2_bwd does and, in the same pass, partials fp32 [rows][2][dx.c] of
 * (sum g, sum g*y), g = dx * [y*scale+shift > 0], taken from the bf16 values it stores (bit-identical to running the
 * two kernels). y = that block's raw conv output (same shape as dx); code = the window codes the forward wrote
 * (required). Grid = rows blocks. */
CVB_API int cvb_maxpool2x2_bwd_bn_reduce(cvb_view dout, const uint8_t* code, cvb_view y, const float* scale,
                                 const float* shift, cvb_view dx, int accumulate, float* partials, int rows,
                                 void* stream);
/* out (size given by the view, models/segnet.py:104 output_size=) = zeros, out[window pos by code] = x. */
CVB_API int cvb_maxunpool2x2_fwd(cvb_view x, const uint8_t* code, cvb_view out, void* stream);
/* dx[n,ho,wo,c] = dout at the coded position. */
CVB_API int cvb_maxunpool2x2_bwd(cvb_view dout, const uint8_t* code, cvb_view dx, void* stream);
/* Parity export: torch-style int64 flat indices (h*W + w per (n,c) plane, NCHW [n,c,ho,wo]) from the codes. */
CVB_API int cvb_pool_code_to_index(const uint8_t* code, int n, int ho, int wo, int c, int w_in, int64_t* idx_nchw,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)  (models/unet.py:25,29)
 * ------------------------------------------------------------------------------------------------------------- */
CVB_API int cvb_bilinear2x_fwd(cvb_view x, cvb_view out, void* stream);
/* The same with the BatchNorm+ReLU of the block that produced the source fused in (ABI 4, cross-layer fusion, the
 * `BasicConv2d -> UpSample2d` hand-over of models/unet.py:112-147): out = upsample(bf16(relu(y*scale + shift))), bit-identical
 * to cvb_bn_relu_apply followed by cvb_bilinear2x_fwd; that block's activation is never written. */
CVB_API int cvb_bn_relu_bilinear2x_fwd(cvb_view y, const float* scale, const float* shift, cvb_view out, void* stream);
CVB_API int cvb_bilinear2x_bwd(cvb_view dout, cvb_view dx, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Label tensors (targets / ground truth / predictions) are torch int64 at the reference API (`masks.cuda()`,
 * train.py:127; `preds.argmax(dim=1)`, train.py:191) or the uint8 masks the device input stage keeps resident
 * (cvb_input_stage_u8 below): every function that reads labels takes the element type.
 * ------------------------------------------------------------------------------------------------------------- */
typedef enum { CVB_LABEL_U8 = 1, CVB_LABEL_I64 = 8 } cvb_label_type; /* value = bytes per label */

/* ---------------------------------------------------------------------------------------------------------------
 * nn.CrossEntropyLoss() forward + backward fused (train.py:105,130-131; eval.py:42,58).
 * logits: NCHW fp32 [n,c,h,w] (the module boundary); target [n,h,w] of `label_type`.
 * A pixel is counted when target != ignore_index and 0 <= target < c.
 * mean != 0 (reduction='mean', the reference's): loss = sum / counted, gradients scaled by grad_scale / counted --
 *   the counted total is taken by a pre-pass over the target inside this call, so loss and gradient always agree
 *   (ignore_index inside or outside [0,c)). mean == 0 (reduction='sum'): loss = sum, gradients scaled by grad_scale.
 * scratch: double [4], ZEROED by the caller, private to this call (sum, counted, invalid labels, block ticket).
 * loss_out: device float, written by the last block to finish. A label that is neither ignore_index nor in [0,c)
 *   (torch raises a device-side assert) makes the loss NaN: it cannot pass unnoticed, and nothing is killed.
 * dlogits (may be NULL): same layout as logits, receives (softmax - onehot) * scale, 0 for pixels not counted.
 * ------------------------------------------------------------------------------------------------------------- */
CVB_API int cvb_softmax_ce_nchw_f32(const float* logits, const void* target, int label_type, int n, int c, int h, int w,
                            int64_t ignore_index, int mean, double* scratch, float* loss_out, float* dlogits,
                            float grad_scale, void* stream);
/* Same on the model's internal NHWC bf16 logits view (first c channels); dlogits (ptr may be NULL) is an NHWC bf16
 * view whose channels >= c are written as zero. */
CVB_API int cvb_softmax_ce_nhwc_bf16(cvb_view logits, int c, const void* target, int label_type, int64_t ignore_index,
                             int mean, double* scratch, float* loss_out, cvb_view dlogits, float grad_scale,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * preds.argmax(dim=1) + confusion matrix (train.py:191-194, utils.py:162-228, legacy/metrics.py:22-30).
 * cm: int64 [c][c], rows = ground truth, cols = prediction, ACCUMULATED into (caller zeroes).
 * cvb_confusion_matrix: pred and gt share `label_type`; pixels with gt == ignore_label are dropped (utils.py:178);
 * with clamp_oob == 0 a pair with either label outside [0,c) is dropped (sklearn `labels=range(C)`,
 * legacy/metrics.py:29); with clamp_oob != 0 such labels are counted in class c-1 (lets the caller keep an explicit
 * out-of-range bucket, used to reproduce np.histogram's per-array range handling in utils.py:183-187).
 * ------------------------------------------------------------------------------------------------------------- */
CVB_API int cvb_confusion_matrix(const void* pred, const void* gt, int label_type, int64_t count, int c,
                                 int64_t ignore_label, int clamp_oob, int64_t* cm, void* stream);
/* Fused: first-max argmax over c channels of NCHW fp32 logits; optionally also writes pred (int64 [n,h,w]).
 * gt is of `label_type`. */
CVB_API int cvb_argmax_confusion_nchw_f32(const float* logits, const void* gt, int label_type, int n, int c, int h, int w,
                                  int64_t* pred_or_null, int64_t* cm, void* stream);
CVB_API int cvb_argmax_confusion_nhwc_bf16(cvb_view logits, int c, const void* gt, int label_type, int64_t* pred_or_null,
                                   int64_t* cm, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Device input stage: transforms.ToTensor + transforms.Normalize (transforms.py:485-538) and the
 * mask's `.long()` (transforms.py:503), moved behind the host->device copy of train.py:126-127 so that the uint8
 * image and mask cross PCIe instead of fp32 / int64 tensors (5x fewer bytes).
 *   img_u8   DEVICE uint8 [n,h,w,c], HWC as cv2.imread / the dataset yield it (dataset/camvid.py:161-173), c <= 4
 *   out_nchw DEVICE fp32 [n,c,h,w] = ((float(u8) / 255) - mean[c]) / std[c], each operation rounded to fp32 exactly as
 *            torch does it (IEEE division, subtraction, IEEE division: bit-exact against the reference transforms)
 *   mean_host / std_host: HOST arrays of c floats (conf/settings.py:8-9 for CamVid BGR)
 *   mask_u8  DEVICE uint8 [n,h,w]; mask_i64 DEVICE int64 [n,h,w] or NULL (the loss and metric kernels read uint8
 *            labels directly: cvb_label_type). img_u8 / out_nchw may both be NULL to convert a mask only.
 * ------------------------------------------------------------------------------------------------------------- */
CVB_API int cvb_input_stage_u8(const uint8_t* img_u8, int n, int h, int w, int c, const float* mean_host,
                               const float* std_host, float* out_nchw, const uint8_t* mask_u8, int64_t* mask_i64,
                               void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink / NVSwitch peer memory (the path's one exchange step; reference
 * precedent: legacy/train_tpu.py:115 `xm.optimizer_step(optimizer)`, the XLA all-reduce of the gradients). One
 * process per GPU. Every rank allocates its flat fp32 gradient buffer and a flag pad of cvb_comm_flag_words() uint32
 * (zeroed once) as peer-accessible ("symmetric") memory and exchanges the device pointers; the kernels use no shared
 * memory and few registers, so they run on the same SMs as the backward pass's convolution kernels.
 * ------------------------------------------------------------------------------------------------------------- */
#define CVB_COMM_MAX_WORLD 8     /* ranks: the GPUs of one NVSwitch box */
#define CVB_COMM_MAX_BUCKETS 64  /* all-reduce calls per step */
typedef struct {
  void* const* peer_bufs_host;   /* HOST array [world] of DEVICE pointers: rank p's gradient buffer as mapped in THIS process */
  void* const* peer_flags_host;  /* HOST array [world] of DEVICE pointers: rank p's flag pad */
  void* multicast_buf;           /* DEVICE multicast (NVLS) address of the gradient buffer, or NULL: with it the reduction
                                    runs inside the NVSwitch (multimem.ld_reduce / multimem.st), one load + one store per
                                    16 bytes whatever the rank count */
  int32_t rank, world;
} cvb_comm;
CVB_API int cvb_comm_flag_words(void);
/* In place: buf[offset, offset+count) (fp32 elements) := mean over the ranks, on every rank, identical bits everywhere
 * (rank r reduces slice r in rank order and stores it to all ranks). `bucket` numbers the calls of one step
 * (0 .. CVB_COMM_MAX_BUCKETS-1, same sequence on every rank), `epoch` the steps (nonzero, increasing). Asynchronous on
 * `stream`; the results are complete on a stream once cvb_allreduce_wait has run on it. */
CVB_API int cvb_allreduce_mean_f32(const cvb_comm* comm, int64_t offset, int64_t count, int bucket, uint32_t epoch, int ctas,
                           void* stream);
/* Makes `stream` wait until buckets 0 .. n_buckets-1 of `epoch` have arrived from every rank. */
CVB_API int cvb_allreduce_wait(const cvb_comm* comm, int n_buckets, uint32_t epoch, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Utilities
 * ------------------------------------------------------------------------------------------------------------- */
/* Zero-fills a bf16 view (pad rows of concat buffers). */
CVB_API int cvb_zero_view(cvb_view v, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAMVID_B200_H_ */
