/* sfvos.h -- C ABI of libsfvos.so: the B200 (sm_100a) kernels behind the SlowFast-VOS hot path.
 *
 * The reference (ChantalMP/Applying-SlowFast-networks-to-video-object-segmentation) has no native code and no
 * FFI: every entry point below replaces a *dependency kernel* the reference reaches through torch.nn /
 * torchvision.ops.  The reference call site each one stands in for is cited per function
 * (code/... = /root/reference/code/...; TV/... = torchvision 0.26).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host
 *   - the caller owns every buffer; no entry point allocates device memory, frees, or synchronises
 *   - every launch goes to the cudaStream_t passed as the last argument (0 = legacy default stream)
 *   - return 0 on success; otherwise an SFVOS_ERR_* code, message via sfvos_last_error() (thread-local)
 *   - there is NO CPU fallback: without an sm_100 device sfvos_device_check() fails and launches error out
 *   - activations are channels-last: N(D)HWC, i.e. [B, T, H, W, C] with C contiguous; `cstride` arguments give
 *     the element distance between consecutive pixels so a kernel can read/write a channel slice of a wider
 *     (concatenated) buffer -- that is how torch.cat at code/helpers/model.py:115,162 disappears.
 */
#ifndef SFVOS_H
#define SFVOS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFVOS_OK 0
#define SFVOS_ERR_INVALID 1      /* bad argument / unsupported shape */
#define SFVOS_ERR_CUDA 2         /* CUDA runtime / driver error */
#define SFVOS_ERR_UNSUPPORTED 3  /* no sm_100 device */

#define SFVOS_F32 0
#define SFVOS_BF16 1
#define SFVOS_F64 2              /* only for the BatchNorm statistics of the fp32 validation mode */

typedef void* sfvos_stream;      /* cudaStream_t */

int sfvos_version(void);
const char* sfvos_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (B200).  No fallback exists. */
int sfvos_device_check(void);
/* Name of the kernel the last sfvos_conv_* / sfvos_wgrad_* call on this thread dispatched to ("conv_umma",
 * "conv_tstack", "wgrad_umma", "wgrad_halo", "wgrad_c32", "wgrad_stack", ...): used by the benchmark to attribute
 * measured launch times to kernels. */
const char* sfvos_last_kernel(void);

/* 1 iff the driver accepts tensor maps whose batch stride is smaller than a clip's extent, i.e. x_bstride = one frame:
 * clip b = frames [b, b+T) of ONE channels-last sequence buffer.  That is how the reference's clips relate (consecutive
 * windows of the cached per-frame features, code/helpers/model.py:318-337), and it lets the layout pass convert every frame
 * once instead of once per window.  dev_ptr: any 16-byte aligned device pointer (not dereferenced). */
int sfvos_tma_overlap_supported(const void* dev_ptr);

/* Measurement probe (tools/bench_l2.py), not part of the path: `iters` passes over buf touching every 16-byte vector once,
 * kind 0 = streaming loads (L2 -> SM read bandwidth when buf fits the L2), kind 1 = red.global.add.v4.f32 (the L2's
 * vector-reduction throughput, the denominator of the ROIAlign backward).  sink: one f32 on the device. */
int sfvos_probe_l2(int32_t kind, void* buf, int64_t nbytes, int32_t iters, float* sink, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution (fprop and dgrad share one kernel).
 * Replaces aten::conv3d / cudnn_convolution reached from nn.Conv3d at code/helpers/model.py:72-76,83-90
 * (invoked :112,120,124,132,136,144,147) and aten::conv2d / conv_transpose2d of MaskRCNNHeads /
 * MaskRCNNPredictor (TV/models/detection/mask_rcnn.py:284-296,342-344); the dgrad use replaces their
 * autograd backward-data kernels.
 *
 *   y[b,t,h,w,n] = act( scale[n] * sum_{a,i,j,c} x[b, t+a-pad_t, h+i-pad_h, w+j-pad_w, c] * Wp[n, (a,i,j), c]
 *                       + shift[n] )
 * Out-of-range input coordinates read as zero (TMA out-of-bounds fill = the conv padding).  Spatial size is
 * preserved (stride 1); To is the output temporal extent.
 * Packed weights Wp (see sfvos_pack_weights): umma = bf16 [N][taps*Cp] (K-major), simt = f32 [taps*Cp][N].
 * ------------------------------------------------------------------------------------------------------- */
typedef struct sfvos_conv_params {
    const void* x;            /* input activations, bf16 (umma) or f32 (simt), [B,T,H,W,*] */
    int64_t B, T, H, W;
    int64_t C;                /* input channels read per tap */
    int64_t x_cstride;        /* elements between consecutive pixels (w -> w+1) of x (>= C) */
    int64_t x_hstride, x_tstride, x_bstride;   /* element strides of the h / t / b axes; 0 = dense
                                                  (W*cstride, H*W*cstride, T*H*W*cstride).  Non-dense strides let
                                                  the ConvTranspose2d backward read every 2nd row/column of dy */
    const void* w;            /* packed weights */
    int64_t Cp;               /* per-tap K extent of the packed weights (C rounded up to 64 for umma) */
    int64_t N;                /* output channels: a multiple of 32 up to 256 (one N tile, activations fetched once),
                                 or a multiple of 256 (256-column N tiles; the fully-connected layers of the box
                                 head, TV/models/detection/faster_rcnn.py:286-307, run as 1x1x1 convolutions over
                                 [1,1,1,M rois] "pixels" with C = in_features) */
    int64_t kt, kh, kw;
    int64_t pad_t, pad_h, pad_w;
    int64_t To;
    void* y;                  /* output, already offset to its first channel */
    int32_t y_dtype;          /* SFVOS_F32 | SFVOS_BF16 */
    int32_t relu;
    int64_t y_cstride;
    const float* scale;       /* optional [N] (NULL = 1) */
    const float* shift;       /* optional [N] (NULL = 0) */
    float* sum;               /* optional [N]: += column sums of the raw fp32 accumulators (train-mode BN stats) */
    float* sumsq;             /* optional [N]: += column sums of squares */
    int32_t accumulate;       /* y += result (f32 y only): merges the two gradient paths into the fast pathway */
    int32_t reserved;
    /* output pixel scatter (identity: OH=H, OW=W, mul=1, off=0); ConvTranspose2d k2 s2 uses mul=2, off=i|j */
    int64_t OH, OW, oy_mul, oy_off, ox_mul, ox_off;
    /* optional fused ReLU backward (umma, dgrad use): relu_mask = the bf16 post-ReLU activation this gradient flows into,
     * same pixels / channels as y (cstride in elements).  y = (relu_mask > 0) ? result : 0, and `sum` (if given; sumsq must
     * then be NULL) receives the column sums of the STORED y = the bias gradient of the layer below.  Replaces the separate
     * relu_bwd pass between two 3x3 convolutions of MaskRCNNHeads (TV mask_rcnn.py:284-296). */
    const void* relu_mask;
    int64_t relu_mask_cstride;
    /* optional addend (umma): y = act(...) + addend[pixel, n], an f32 or bf16 tensor indexed like y (cstride in elements).
     * accumulate = 1 is the special case addend = y (f32).  Lets the second of the two gradient paths into the fast
     * pathway (a lateral dgrad on top of the fast convolution's dgrad, code/helpers/model.py:128-131,140-143) read the f32
     * partial sum and store the total ONCE, in bf16, for the BatchNorm backward that consumes it. */
    const void* addend;
    int32_t addend_dtype;     /* SFVOS_F32 | SFVOS_BF16 */
    int32_t reserved2;
    int64_t addend_cstride;
} sfvos_conv_params;

/* tcgen05/TMEM/TMA bf16 kernel (the product path). */
int sfvos_conv_umma(const sfvos_conv_params* p, sfvos_stream stream);
/* CUDA-core kernel: the "fp32 validation mode" of the north star (<=1e-4), same semantics.  f32 operands and results,
 * products accumulated in fp64 and rounded once (a pre-activation's sign is that of the exact sum, so ReLU masks -- and
 * with them the gradients -- do not depend on summation order); no atomics, bit-reproducible. */
int sfvos_conv_simt(const sfvos_conv_params* p, sfvos_stream stream);

/* Weight-gradient GEMM: dw[(a,i,j)][c][n] += sum_{b,t,h,w} x[b,t+a-pad_t,h+i-pad_h,w+j-pad_w,c] * dy[b,t,h,w,n].
 * Replaces cudnn_convolution_backward_weight for the same modules (and addmm's weight gradient for the box head's
 * nn.Linear layers).  dw is f32 [taps][C][N], accumulated.  N: a multiple of 32 up to 256, or a multiple of 256. */
typedef struct sfvos_wgrad_params {
    const void* x;  int64_t B, T, H, W, C, x_cstride;
    int64_t x_hstride, x_tstride, x_bstride;      /* 0 = dense */
    const void* dy; int64_t To, N, dy_cstride;
    int64_t dy_hstride, dy_tstride, dy_bstride;   /* 0 = dense, as for sfvos_conv_params.x_*stride */
    int64_t kt, kh, kw, pad_t, pad_h, pad_w;
    float* dw;
    /* sfvos_wgrad_simt only: scratch for the fp64 split-K partial tiles, reduced in split order (no atomics).  NULL or too
     * small = no split-K (still deterministic, fewer CTAs).  sfvos_wgrad_simt_workspace_bytes() gives the useful size. */
    void* workspace;
    int64_t workspace_bytes;
} sfvos_wgrad_params;
int sfvos_wgrad_umma(const sfvos_wgrad_params* p, sfvos_stream stream);   /* x, dy bf16; split-K merged with f32 atomics */
/* x, dy f32; fp64 accumulation; the CTA that owns a dw tile adds to it without atomics, so concurrent calls must not share
 * dw (the validation mode gives every pyramid level its own accumulator and sums them in level order). */
int sfvos_wgrad_simt(const sfvos_wgrad_params* p, sfvos_stream stream);
int64_t sfvos_wgrad_simt_workspace_bytes(const sfvos_wgrad_params* p);

/* fp32 master weight [Cout,Cin,kt,kh,kw] (the state_dict tensor) -> packed operands.
 *   mode 0 fprop : N=Cout, taps in (a,i,j) order, K=(tap,cin)
 *   mode 1 dgrad : N=Cin, taps flipped, K=(tap',cout)            (input coord = out + tap' - (k-1-pad))
 *   mode 2 convT : weight is [Cin,Cout,kh,kw] (ConvTranspose2d); packs the single tap (i,j)=(tap_i,tap_j):
 *                  N=Cout, K=cin  (TV/models/detection/mask_rcnn.py:342)
 *   mode 3 convT dgrad : N=Cin, K=cout for tap (tap_i,tap_j)
 * out_dtype SFVOS_BF16 -> [N][taps*Cp]; SFVOS_F32 -> [taps*Cp][N].  Padding entries are written as zero. */
int sfvos_pack_weights(const float* w, void* out, int32_t out_dtype, int32_t mode, int64_t Cout, int64_t Cin,
                       int64_t kt, int64_t kh, int64_t kw, int64_t Cp, int64_t tap_i, int64_t tap_j,
                       sfvos_stream stream);
/* dw [taps][C][N] f32 -> grad += in the state_dict layout.  mode as above (0: [Cout=N,Cin=C,kt,kh,kw];
 * 2: ConvTranspose2d [Cin=C,Cout=N,kh,kw] single tap (tap_i,tap_j) from a [1][C][N] dw). */
int sfvos_unpack_wgrad(const float* dw, float* grad, int32_t mode, int64_t Cout, int64_t Cin, int64_t kt,
                       int64_t kh, int64_t kw, int64_t tap_i, int64_t tap_j, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * BatchNorm3d pieces.  Replace aten::native_batch_norm (+backward) and aten::relu_ reached from
 * nn.BatchNorm3d / nn.ReLU at code/helpers/model.py:69,78,92 (invoked :113-114,121-122,...,148).
 * ------------------------------------------------------------------------------------------------------- */
/* Bytes of fp64 scratch the deterministic per-channel reductions (sfvos_channel_stats, sfvos_bn_bwd_reduce with a
 * workspace) need for npix pixels of C channels: one row of 2*C partial sums per CTA. */
int64_t sfvos_reduce_workspace_bytes(int64_t npix, int64_t C);
/* per-channel sum / sumsq over npix pixels of an f32 channels-last tensor (validation mode; the umma kernels fuse this
 * into their epilogue).  stats = fp64 [2*C] (sums, then sums of squares), WRITTEN: fp64 so that |mean| >> std does not
 * cancel in sumsq/n - mean^2, fixed reduction order (per-CTA rows in the workspace, merged in CTA order) so that two
 * runs agree bit for bit. */
int sfvos_channel_stats(const float* x, int64_t npix, int64_t C, int64_t cstride, double* stats, void* workspace,
                        int64_t workspace_bytes, sfvos_stream stream);
/* train: mean/var from (sum,sumsq,count) of the bias-free conv output (stats_dtype SFVOS_F32: float sums from the umma
 * epilogues; SFVOS_F64: sfvos_channel_stats); writes scale=gamma*rstd,
 * shift=beta-mean*scale, mean, rstd; updates running stats with PyTorch's rules (momentum, unbiased var,
 * conv bias added back to the mean) and num_batches_tracked += 1 (int64, may be NULL). */
int sfvos_bn_finalize(const void* sum, const void* sumsq, int32_t stats_dtype, double count, const float* conv_bias,
                      const float* gamma, const float* beta, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, double momentum, double eps, float* scale, float* shift,
                      float* mean, float* rstd, int64_t C, sfvos_stream stream);
/* Running-statistics update of ONE BatchNorm for n_calls consecutive train-mode forward calls (the pyramid levels of one
 * temporally_enhance_features call, code/helpers/model.py:155-162), applied in call order: the same sequence of EMA steps
 * as n_calls sfvos_bn_finalize calls with running buffers.  Lets the calls themselves run concurrently (finalize with NULL
 * running buffers) and keeps the buffers bit-identical to the sequential order.  num_batches_tracked += n_calls. */
#define SFVOS_BN_MAX_CALLS 8
typedef struct sfvos_bn_running_params {
    const void* sum[SFVOS_BN_MAX_CALLS];       /* per call: [C] sums of the bias-free conv output (f32 | f64) */
    const void* sumsq[SFVOS_BN_MAX_CALLS];
    double count[SFVOS_BN_MAX_CALLS];          /* per call: elements per channel */
    int32_t n_calls, stats_dtype;              /* SFVOS_F32 | SFVOS_F64 */
    const float* conv_bias;                    /* may be NULL */
    float* running_mean;
    float* running_var;
    int64_t* num_batches_tracked;              /* may be NULL */
    double momentum;
    int64_t C;
} sfvos_bn_running_params;
int sfvos_bn_running_update(const sfvos_bn_running_params* p, sfvos_stream stream);
/* eval: scale=gamma/sqrt(rv+eps), shift=beta+(conv_bias-rm)*scale; optionally (may be NULL) mean = rm - conv_bias and
 * rstd = 1/sqrt(rv+eps), the constants the eval-mode backward needs in terms of the bias-free conv output. */
int sfvos_bn_fold_eval(const float* conv_bias, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, double eps, float* scale, float* shift, float* mean, float* rstd,
                       int64_t C, sfvos_stream stream);
/* y = act(x*scale+shift) over a channels-last slice; x f32|bf16, y f32|bf16. */
int sfvos_affine_act(const void* x, int32_t x_dtype, int64_t x_cstride, void* y, int32_t y_dtype,
                     int64_t y_cstride, const float* scale, const float* shift, int32_t relu, int64_t npix,
                     int64_t C, sfvos_stream stream);
/* BN(+ReLU) backward, pass 1: sums[0][c] += sum dy_m, sums[1][c] += sum dy_m*xhat, with
 * dy_m = dy * (relu ? (x*scale+shift > 0) : 1), xhat = (x-mean)*rstd.  x is the saved raw conv output (f32).
 * workspace NULL: per-CTA sums merged with f32 atomics (product path).  workspace given (validation mode, f32 dy,
 * sfvos_reduce_workspace_bytes bytes): fp64 sums merged in CTA order, ``sums`` written instead of accumulated. */
int sfvos_bn_bwd_reduce(const void* dy, int32_t dy_dtype, int64_t dy_cstride, const float* x, int64_t x_cstride,
                        const float* scale, const float* shift, const float* mean, const float* rstd,
                        int32_t relu, int64_t npix, int64_t C, float* sums, void* workspace, int64_t workspace_bytes,
                        sfvos_stream stream);
/* pass 2: dx = gamma*rstd*(dy_m - sums0/n - xhat*sums1/n) written as bf16|f32; block 0 also does
 * dgamma += sums1, dbeta += sums0 (may be NULL).  fixed_stats != 0 = eval-mode BatchNorm (mean / rstd are the running
 * statistics, constants): dx = gamma*rstd*dy_m, and dbias (may be NULL) += gamma*rstd*sums0, the gradient of the conv bias
 * (which train-mode BN cancels exactly). */
int sfvos_bn_bwd_apply(const void* dy, int32_t dy_dtype, int64_t dy_cstride, const float* x, int64_t x_cstride,
                       const float* scale, const float* shift, const float* mean, const float* rstd,
                       const float* gamma, int32_t relu, int64_t npix, int64_t C, const float* sums, void* dx,
                       int32_t dx_dtype, int64_t dx_cstride, float* dgamma, float* dbeta, int32_t fixed_stats,
                       float* dbias, sfvos_stream stream);
/* bias+ReLU layers (mask head): dx = dy * (y > 0) as bf16|f32 and dbias[c] += sum dx. */
int sfvos_relu_bwd(const void* dy, int32_t dy_dtype, int64_t dy_cstride, const void* y, int32_t y_dtype,
                   int64_t y_cstride, void* dx, int32_t dx_dtype, int64_t dx_cstride, float* dbias, int64_t npix,
                   int64_t C, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Layout.  Replaces torch.stack(...).transpose(1,2) at code/helpers/model.py:157-158 and
 * torch.cat(...).squeeze at :162 (the latter by writing both pathways into one channels-last buffer).
 * ------------------------------------------------------------------------------------------------------- */
/* src (f32|bf16) [F, C, HW] (frame stride src_fstride elements) -> dst (f32|bf16) [F, HW, cstride].  bf16 sources are what
 * a feature cache kept in half the bytes hands over (host-resident windows cross PCIe at half the size). */
int sfvos_nchw_to_nhwc(const void* src, int32_t src_dtype, int64_t src_fstride, void* dst, int32_t dst_dtype,
                       int64_t dst_cstride, int64_t F, int64_t C, int64_t HW, sfvos_stream stream);
/* src (f32|bf16) [F, HW, cstride] -> dst f32 [F, C, HW]. */
int sfvos_nhwc_to_nchw(const void* src, int32_t src_dtype, int64_t src_cstride, float* dst, int64_t F, int64_t C,
                       int64_t HW, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * ROIAlign (legacy aligned=False), multi-level, one launch for all levels.
 * Replaces torch.ops.torchvision.roi_align / _roi_align_backward (TV/ops/roi_align.py:258) and the
 * LevelMapper + per-level gather/scatter glue of MultiScaleRoIAlign (TV/ops/poolers.py:73-95,147-227), reached
 * from code/helpers/model.py:346 via TV/models/detection/roi_heads.py:772,815.
 * ------------------------------------------------------------------------------------------------------- */
/* levels[k] = clamp(floor(4 + log2(sqrt(area)/224) + 1e-6), k_min, k_max) - k_min, fp32 like the reference. */
int sfvos_roi_levels(const float* rois /*[K,5]*/, int64_t K, int32_t k_min, int32_t k_max, int32_t* levels,
                     sfvos_stream stream);
typedef struct sfvos_roi_params {
    const void* feat[4];      /* per level, channels-last [N, H_l, W_l, C] (f32|bf16) */
    void* dfeat[4];           /* backward only: f32 gradient buffers, same shape, pre-zeroed by the caller */
    int64_t H[4], W[4];
    float scale[4];
    int32_t n_levels, feat_dtype;
    int64_t N, C, cstride;    /* C multiple of 8 */
    const float* rois;        /* [K,5] = (batch idx, x1, y1, x2, y2) in image pixels */
    const int32_t* levels;    /* [K] from sfvos_roi_levels */
    int64_t K;
    int32_t P, sampling_ratio;   /* output P x P; sampling_ratio > 0 */
    void* out;                /* fwd: output; bwd: grad of the output */
    int32_t out_dtype;        /* SFVOS_F32 | SFVOS_BF16 */
    int32_t out_nchw;         /* 0: [K,P,P,C] (feeds the mask head); 1: [K,C,P,P] (torch box head order) */
} sfvos_roi_params;
int sfvos_roi_align_fwd(const sfvos_roi_params* p, sfvos_stream stream);
int sfvos_roi_align_bwd(const sfvos_roi_params* p, sfvos_stream stream);
/* project_masks_on_boxes (TV/models/detection/roi_heads.py:85-97): masks u8 [n_obj,H,W], rois [K,5] with the
 * matched object index in column 0, output f32 [K,M,M]; spatial_scale 1, adaptive sampling (sampling_ratio=-1). */
int sfvos_mask_targets(const uint8_t* masks, int64_t n_obj, int64_t H, int64_t W, const float* rois, int64_t K,
                       int32_t M, float* out, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Mask predictor tail + loss.  Replaces mask_fcn_logits (1x1 conv 256->n_cls, TV/.../mask_rcnn.py:344),
 * the class-channel gather and F.binary_cross_entropy_with_logits of maskrcnn_loss (TV/.../roi_heads.py:100-129)
 * and the sigmoid/select of maskrcnn_inference (:56-82).
 * ------------------------------------------------------------------------------------------------------- */
/* x (bf16|f32) [K,S,S,C] -> logits f32 [K,n_cls,S,S] (NCHW like the reference).
 * pixel_order (this and the two backward entries): 0 = the rows of x are the S x S pixels in raster order; 1 = 2x2
 * space-to-depth order, row (h*(S/2)+w)*4 + 2i+j = pixel (2h+i, 2w+j) -- the layout in which ConvTranspose2d(k2,s2)
 * (conv5_mask) leaves its output when its four taps run as ONE 1x1 convolution with 4*C output channels.  logits and
 * glogits are always in the reference's raster order. */
int sfvos_mask_logits_fwd(const void* x, int32_t x_dtype, const float* w /*[n_cls,C]*/, const float* b,
                          float* logits, int64_t K, int64_t S, int64_t C, int32_t n_cls, int32_t pixel_order,
                          sfvos_stream stream);
/* loss[0] = mean over K*S*S of BCE-with-logits(logits[k,labels[k]], targets[k]). */
int sfvos_mask_bce_fwd(const float* logits, const int64_t* labels, const float* targets, float* loss, int64_t K,
                       int64_t S, int32_t n_cls, sfvos_stream stream);
/* backward of the 1x1 conv: dx[k,p,:] = sum_cls glogits[k,cls,p]*w[cls,:] (bf16|f32, same dtype as x);
 * dw [n_cls,C] += sum glogits*x; db [n_cls] += sum glogits. */
int sfvos_mask_logits_bwd(const void* x, int32_t x_dtype, const float* w, const float* glogits, void* dx,
                          int32_t dx_dtype, float* dw, float* db, int64_t K, int64_t S, int64_t C, int32_t n_cls,
                          int32_t pixel_order, sfvos_stream stream);
/* The same with the backward of the ReLU that produced x fused in (x = relu(conv5_mask(...)), TV mask_rcnn.py:342-343):
 * dx = (x > 0) ? sum_cls glogits*w : 0 and dbias_x[c] += sum dx (the ConvTranspose2d bias gradient). */
int sfvos_mask_logits_relu_bwd(const void* x, int32_t x_dtype, const float* w, const float* glogits, void* dx,
                               int32_t dx_dtype, float* dw, float* db, float* dbias_x, int64_t K, int64_t S, int64_t C,
                               int32_t n_cls, int32_t pixel_order, sfvos_stream stream);
/* backward of the loss: glogits[k,cls,p] = gloss*(sigmoid(z)-t)/(K*S*S) on the label channel, 0 elsewhere. */
int sfvos_mask_bce_bwd(const float* logits, const int64_t* labels, const float* targets, const float* gloss,
                       float* glogits, int64_t K, int64_t S, int32_t n_cls, sfvos_stream stream);
/* maskrcnn_inference: prob[k,0,s,s] = sigmoid(logits[k,labels[k]]). */
int sfvos_mask_probs(const float* logits, const int64_t* labels, float* prob, int64_t K, int64_t S, int32_t n_cls,
                     sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Box-branch losses.  Replaces fastrcnn_loss (TV/models/detection/roi_heads.py:12-53: F.cross_entropy over the
 * class logits + F.smooth_l1_loss(beta, reduction="sum") / M over the matched class's 4 box deltas of the
 * positive ROIs), reached from code/helpers/model.py:346 via RoIHeads.forward (TV roi_heads.py:783).
 * cls_logits [M,n_cls] and box_reg [M,4*n_cls] are f32 with a row stride (elements) each -- they may be column
 * slices of one fused predictor output; labels int64 [M]; reg_targets f32 [M,4] dense.
 * ------------------------------------------------------------------------------------------------------- */
/* losses[0] = loss_classifier, losses[1] = loss_box_reg (written, not accumulated; deterministic). */
int sfvos_fastrcnn_loss_fwd(const float* cls_logits, int64_t cls_stride, const float* box_reg, int64_t box_stride,
                            const int64_t* labels, const float* reg_targets, int64_t M, int32_t n_cls, float beta,
                            float* losses, sfvos_stream stream);
/* gloss [2] = upstream gradients of the two losses.  Writes every element of dcls [M,n_cls] and dbox [M,4*n_cls]
 * (row strides in elements; zeros outside the matched class / for background ROIs). */
int sfvos_fastrcnn_loss_bwd(const float* cls_logits, int64_t cls_stride, const float* box_reg, int64_t box_stride,
                            const int64_t* labels, const float* reg_targets, const float* gloss, int64_t M,
                            int32_t n_cls, float beta, float* dcls, int64_t dcls_stride, float* dbox,
                            int64_t dbox_stride, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Mask paste-back.  Replaces paste_masks_in_image (TV/models/detection/roi_heads.py:415-501: expand_masks, expand_boxes,
 * per-mask F.interpolate(bilinear, align_corners=False) + slice assignment in a Python loop), reached from
 * code/helpers/model.py:347 via GeneralizedRCNNTransform.postprocess (TV/models/detection/transform.py:257-279).
 * masks f32 [K,1,M,M] (probabilities), boxes f32 [K,4] in the OUTPUT image's pixels, out f32 [K,1,im_h,im_w] (every
 * element written; zeros outside the grown box).  One launch for all K masks (K <= 65535).
 * ------------------------------------------------------------------------------------------------------- */
int sfvos_paste_masks(const float* masks, const float* boxes, int64_t K, int32_t M, int32_t padding, int64_t im_h,
                      int64_t im_w, float* out, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * FPN top-down merge.  Replaces F.interpolate(last_inner, size=..., mode="nearest") + the addition at
 * TV/ops/feature_pyramid_network.py (FeaturePyramidNetwork.forward), reached from code/helpers/model.py:204 through
 * maskrcnn_model.backbone.  inner f32 [N,H,W,C] (the lateral 1x1 convolution's output) += top f32 [N,Ht,Wt,C] at the
 * nearest coarser pixel (ATen's floor(dst * in / out) rule), in place; out_bf16 (may be NULL) receives the bf16 copy that
 * the 3x3 output convolution reads.  top == NULL: only the copy (coarsest level).
 * ------------------------------------------------------------------------------------------------------- */
int sfvos_upsample_add(const float* top, int64_t Ht, int64_t Wt, float* inner, void* out_bf16, int64_t N, int64_t H,
                       int64_t W, int64_t C, sfvos_stream stream);
/* ResNet-50 body helpers (the frozen backbone's convolution stack, TV/models/resnet.py reached through
 * torchvision...backbone_utils.BackboneWithFPN at code/helpers/model.py:204; the convolutions themselves run on
 * sfvos_conv_umma with the frozen BatchNorm folded into scale / shift):
 *   sfvos_im2col       f32 NCHW image [N,Cin,H,W] -> patch rows [N*Ho*Wo, Kp] (bf16|f32), column (i*kw + j)*Cin + c, zeros
 *                      outside the image and in the padding columns: the 7x7 stride-2 stem on 3 channels as ONE GEMM
 *   sfvos_maxpool3x3s2 F.max_pool2d(kernel 3, stride 2, padding 1) on a channels-last (bf16|f32) tensor
 *   sfvos_add_relu     out = relu(a + b) over f32 (the residual stream), plus an optional bf16 copy for the next block's convs */
int sfvos_im2col(const float* x, void* rows, int32_t rows_dtype, int64_t N, int64_t Cin, int64_t H, int64_t W, int64_t kh,
                 int64_t kw, int64_t stride, int64_t pad, int64_t Kp, sfvos_stream stream);
int sfvos_maxpool3x3s2(const void* x, void* y, int32_t dtype, int64_t N, int64_t H, int64_t W, int64_t C, sfvos_stream stream);
int sfvos_add_relu(const float* a, const float* b, float* out, void* out_bf16, int64_t n, sfvos_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimiser-side helpers for the data-parallel step (the one collective is NCCL all-reduce, called from Python).
 * ------------------------------------------------------------------------------------------------------- */
/* y[i] = a*x[i] + b*y[i]  (flat f32), used to fold 1/world_size into the reduced gradient bucket. */
int sfvos_axpby(const float* x, float* y, float a, float b, int64_t n, sfvos_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SFVOS_H */
