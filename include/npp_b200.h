/*
 * npp_b200.h — C ABI of libnpp_b200.so: hand-written sm_100a kernels for the NPPNet
 * (GuHuangAI/NPP) conv training / inference hot path.
 *
 * The reference has NO native boundary: every FLOP goes through torch.nn
 * modules that dispatch to cuDNN/ATen.  Each entry point below therefore replaces a
 * *library call class* issued by a reference module; the reference file:line whose
 * arithmetic it reproduces is cited per function.  Host code (npp_b200/*.py) keeps the
 * reference's module API and binds these symbols with ctypes.
 *
 * Conventions
 *  - all tensors are device pointers owned by the caller (PyTorch caching allocator);
 *    the library never allocates or frees device memory, never synchronises the device and
 *    only ever launches on the stream it is given;
 *  - activations are NHWC views (`npp_view4`): channel stride 1, explicit pixel/row/image
 *    strides in ELEMENTS so channel slices of a wider concat buffer are first-class;
 *  - dtype: NPP_BF16 storage (fp32 math / accumulation) is the product path, NPP_F32 is the
 *    fp32 validation mode of BASELINE.json (north_star: "1e-4 in an fp32 validation mode");
 *  - every function returns 0 on success, a negative NPP_E_* code otherwise.  No exceptions
 *    cross the ABI, there is no CPU fallback and no cuDNN fallback.
 */
#ifndef NPP_B200_H_
#define NPP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* npp_stream_t; /* cudaStream_t */

enum { NPP_F32 = 0, NPP_BF16 = 1 };

enum {
  NPP_OK = 0,
  NPP_E_INVALID = -1,     /* bad argument (null pointer, negative size, misaligned view) */
  NPP_E_UNSUPPORTED = -2, /* shape / dtype outside what the kernels implement */
  NPP_E_CUDA = -3,        /* a CUDA runtime / driver call failed (see npp_last_error) */
  NPP_E_NODRIVER = -4     /* cuTensorMapEncodeTiled could not be resolved */
};

/* NHWC activation view.  c = channels visible through the view; sw/sh/sn = element strides
 * between neighbouring pixels / rows / images (sw >= c; sw > c for a channel slice). */
typedef struct {
  void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
} npp_view4;

const char* npp_version(void);
const char* npp_last_error(void); /* thread-local text of the last NPP_E_CUDA */
int npp_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense convolution as implicit GEMM on tcgen05 / TMEM, operands staged by TMA (bf16 only).
 * Replaces nn.Conv2d(groups=1) in ReLUConvBN (models/operations.py:69-82), the pointwise
 * half of DilConvS (:214), FactorizedReduce (:149-150), Pooled_Conv (:239), the stems, layer
 * convs and heads of models/model_augment.py:244-398 and extra_conv (:592-596).
 *   w     bf16 [cout, kh, kw, cin]   (OHWI == torch channels_last weight)
 *   bias  fp32 [cout] or NULL
 *   y[n,ho,wo,co] = bias[co] + sum_{r,s,ci} x[n, ho*stride - pad + r*dil, wo*stride - pad + s*dil, ci] * w[co,r,s,ci]
 * stats (optional, fp32 [2*cout], must be zeroed by the caller): per-channel sum and sum of
 * squares of the bf16-rounded outputs, accumulated with atomics (BatchNorm batch statistics,
 * operations.py:79 nn.BatchNorm2d in training mode).
 * in_h_off/in_w_off: extra input offset (FactorizedReduce's x[:, :, 1:, 1:] branch :155).
 * ---------------------------------------------------------------------------------------- */
int npp_conv2d_fwd(const npp_view4* x, const void* w, const float* bias, const npp_view4* y,
                   int kh, int kw, int stride, int pad, int dil, int in_h_off, int in_w_off,
                   float* stats, npp_stream_t stream);

/* dgrad: dx = conv_transpose(dy, w).  wt bf16 [cin, kh, kw, cout] (npp_pack_weight_t).
 * dx is fully overwritten (positions no output touches are written as zero). */
int npp_conv2d_dgrad(const npp_view4* dy, const void* wt, const npp_view4* dx, int kh, int kw,
                     int stride, int pad, int dil, int in_h_off, int in_w_off,
                     npp_stream_t stream);

/* wgrad: dw fp32 [dw_cout, kh, kw, dw_cin] += sum_pixels dy (x) x ; caller zeroes dw first
 * (split-K partial sums are combined with red.global.add.f32).  dw_cout <= dy->c and
 * dw_cin <= x->c are the un-padded channel counts of the fp32 master weight. */
int npp_conv2d_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin,
                     int kh, int kw, int stride, int pad, int dil, int in_h_off, int in_w_off,
                     npp_stream_t stream);

/* fp32 master weight [cout,taps,cin] -> bf16 [cout_pad,taps,cin_pad] (w) and/or transposed
 * [cin_pad,taps,cout_pad] (wt), zero padded; either output may be NULL.  Activation buffers
 * keep channel counts that are multiples of 8 (16-byte TMA rows), hence the padding. */
int npp_pack_weight(const float* w32, void* w, void* wt, int cout, int taps, int cin,
                    int cout_pad, int cin_pad, npp_stream_t stream);

/* Validation-mode / cross-check convolution on CUDA cores (fp32 accumulate, dtype-templated).
 * Same arithmetic as the three functions above, any dtype, groups==1. w/dw are fp32 OHWI
 * when dtype==NPP_F32 and bf16 (w) / fp32 (dw) when dtype==NPP_BF16. */
int npp_conv2d_direct_fwd(const npp_view4* x, const void* w, const float* bias,
                          const npp_view4* y, int kh, int kw, int stride, int pad, int dil,
                          int in_h_off, int in_w_off, int dtype, npp_stream_t stream);
int npp_conv2d_direct_dgrad(const npp_view4* dy, const void* w, const npp_view4* dx, int kh,
                            int kw, int stride, int pad, int dil, int in_h_off, int in_w_off,
                            int dtype, npp_stream_t stream);
int npp_conv2d_direct_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout,
                            int dw_cin, int kh, int kw, int stride, int pad, int dil,
                            int in_h_off, int in_w_off, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise (groups == C) dilated k x k convolution, HBM-bound.
 * Replaces DilConvS.net[1] (models/operations.py:213).  w fp32 [kh*kw, c] (tap-major).
 * relu_in: apply the leading nn.ReLU (:212) while loading.
 * bwd: dx (masked by x>0 when relu_in) and dw fp32 [kh*kw, c] (+=, caller zeroes).
 * ---------------------------------------------------------------------------------------- */
int npp_dwconv_fwd(const npp_view4* x, const float* w, const npp_view4* y, int k, int stride,
                   int pad, int dil, int relu_in, int dtype, npp_stream_t stream);
int npp_dwconv_bwd(const npp_view4* x, const float* w, const npp_view4* dy, const npp_view4* dx,
                   float* dw, int k, int stride, int pad, int dil, int relu_in, int dtype,
                   npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d, training and eval (operations.py:61,79,97,117,151 ...; momentum 0.1, eps 1e-5).
 *  stats:    sums[0:c] += sum x, sums[c:2c] += sum x^2 over all pixels (fp32, caller zeroes).
 *            In SyncBN mode the caller all-reduces `sums` over ranks between stats and finalize.
 *  finalize: mean/var from sums and `count`; writes scale = gamma*invstd, shift = beta-mean*scale,
 *            save_mean, save_invstd; updates running_mean/var (unbiased var, momentum) if non-NULL.
 *  eval_coef: scale/shift from running statistics.
 *  apply:    y = x*scale + shift (+ res) (relu)   — y may be a channel slice of a concat buffer.
 *  bwd_reduce: sums[0:c] += sum dy, sums[c:2c] += sum dy * (x-mean)*invstd
 *            (if relu_mask_y != NULL, dy is first masked by y>0: BN followed by nn.ReLU).
 *  bwd_apply: dx = gamma*invstd * (dy - sum_dy/count - xhat * sum_dy_xhat/count)
 * ---------------------------------------------------------------------------------------- */
int npp_bn_stats(const npp_view4* x, float* sums, int dtype, npp_stream_t stream);
int npp_bn_finalize(const float* sums, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps,
                    float* scale, float* shift, float* save_mean, float* save_invstd, int c,
                    npp_stream_t stream);
int npp_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean,
                     const float* running_var, float eps, float* scale, float* shift, int c,
                     npp_stream_t stream);
int npp_bn_apply(const npp_view4* x, const float* scale, const float* shift,
                 const npp_view4* res, int relu, const npp_view4* y, int dtype,
                 npp_stream_t stream);
int npp_bn_bwd_reduce(const npp_view4* dy, const npp_view4* x, const npp_view4* relu_mask_y,
                      const float* save_mean, const float* save_invstd, float* sums, int dtype,
                      npp_stream_t stream);
int npp_bn_bwd_apply(const npp_view4* dy, const npp_view4* x, const npp_view4* relu_mask_y,
                     const float* gamma, const float* save_mean, const float* save_invstd,
                     const float* sums, double count, const npp_view4* dx, int dtype,
                     npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Elementwise / data movement (all NHWC views, c % 8 == 0 for bf16, % 4 for fp32).
 *  relu_fwd: y = max(x,0)                     nn.ReLU (operations.py:76,95,212,239)
 *  relu_bwd: dx = x>0 ? dy : 0
 *  add:      y = a + b (b may be NULL: copy)  `s = h1 + h2` (model_augment.py:60), torch.cat slices (:62)
 *  axpby:    y = alpha*a + beta*b             MixedOp weighted sum (model_search_interact.py:59)
 *  cast:     dtype conversion of flat arrays (n elements)
 *  fill:     zero / constant fp32
 * ---------------------------------------------------------------------------------------- */
int npp_relu_fwd(const npp_view4* x, const npp_view4* y, int dtype, npp_stream_t stream);
int npp_relu_bwd(const npp_view4* x, const npp_view4* dy, const npp_view4* dx, int dtype,
                 npp_stream_t stream);
int npp_add(const npp_view4* a, const npp_view4* b, const npp_view4* y, int dtype,
            npp_stream_t stream);
int npp_axpby(const npp_view4* a, const float* alpha, const npp_view4* b, const float* beta,
              const npp_view4* y, int dtype, npp_stream_t stream);
int npp_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n,
             npp_stream_t stream);
/* NCHW fp32 <-> NHWC (dtype) layout changes at the module edge (images in, logits out). */
int npp_nchw_to_nhwc(const float* src, int src_c, const npp_view4* dst, int dtype,
                     npp_stream_t stream); /* channels >= src_c of dst are zero filled */
int npp_nhwc_to_nchw(const npp_view4* src, float* dst, int dst_c, int dtype, npp_stream_t stream);
int npp_fill_zero(const npp_view4* y, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Pooling.  maxpool3x3: nn.MaxPool2d(3, stride, 1) (operations.py:55), -inf padding.
 * bwd routes dy to the FIRST maximum in window scan order (ATen semantics).
 * avgpool3x3: nn.AvgPool2d(3, stride, 1, count_include_pad=False) (:57).
 * avgpool2x2: nn.AvgPool2d(2) (:115, :237).   gap: nn.AdaptiveAvgPool2d(1) (:111).
 * ---------------------------------------------------------------------------------------- */
int npp_maxpool3x3_fwd(const npp_view4* x, const npp_view4* y, int stride, int dtype,
                       npp_stream_t stream);
int npp_maxpool3x3_bwd(const npp_view4* x, const npp_view4* dy, const npp_view4* dx, int stride,
                       int dtype, npp_stream_t stream);
int npp_avgpool3x3_fwd(const npp_view4* x, const npp_view4* y, int stride, int dtype,
                       npp_stream_t stream);
int npp_avgpool3x3_bwd(const npp_view4* dy, const npp_view4* dx, int stride, int dtype,
                       npp_stream_t stream);
int npp_avgpool2x2_fwd(const npp_view4* x, const npp_view4* y, int dtype, npp_stream_t stream);
int npp_avgpool2x2_bwd(const npp_view4* dy, const npp_view4* dx, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SE_Block (operations.py:105-129): w = sigmoid(W2 relu(W1 gap(x) + b1) + b2); out = x*w.
 *  gap_fwd:  g[n,c] = mean_hw x  (fp32 [n,c])
 *  se_fc_fwd: h = relu(W1 g + b1) [n, c/2];  s = sigmoid(W2 h + b2) [n, c]   (fp32 weights [out,in])
 *  se_scale_fwd: y = x * s[n,c]
 *  se_scale_bwd: dx_partial = dy * s ; ds[n,c] = sum_hw dy*x   (fp32)
 *  se_fc_bwd: from ds: dW2, db2, dW1, db1 (+=) and dg[n,c]
 *  gap_bwd_add: dx += dg[n,c]/(h*w)
 * ---------------------------------------------------------------------------------------- */
int npp_gap_fwd(const npp_view4* x, float* g, int dtype, npp_stream_t stream);
int npp_se_fc_fwd(const float* g, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* hbuf, float* s, int n, int c, npp_stream_t stream);
int npp_se_scale_fwd(const npp_view4* x, const float* s, const npp_view4* y, int dtype,
                     npp_stream_t stream);
int npp_se_scale_bwd(const npp_view4* x, const float* s, const npp_view4* dy,
                     const npp_view4* dx, float* ds, int dtype, npp_stream_t stream);
int npp_se_fc_bwd(const float* g, const float* hbuf, const float* s, const float* ds,
                  const float* w1, const float* w2, float* dw1, float* db1, float* dw2,
                  float* db2, float* dg, int n, int c, npp_stream_t stream);
int npp_gap_bwd_add(const float* dg, const npp_view4* dx, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Resampling.  bilinear: F.interpolate(mode='bilinear', align_corners=True|False)
 * (model_augment.py:116,539-543; criterion.py:181; function.py:927 uses align_corners=False).
 * nearest: F.interpolate default mode (model_search_interact.py:63-64).
 * bwd accumulates into dx, which the caller zeroes.
 * ---------------------------------------------------------------------------------------- */
int npp_bilinear_fwd(const npp_view4* x, const npp_view4* y, int align_corners, int dtype,
                     npp_stream_t stream);
int npp_bilinear_bwd(const npp_view4* dy, const npp_view4* dx, int align_corners, int dtype,
                     npp_stream_t stream);
int npp_nearest_fwd(const npp_view4* x, const npp_view4* y, int dtype, npp_stream_t stream);
int npp_nearest_bwd(const npp_view4* dy, const npp_view4* dx, int dtype, npp_stream_t stream);

/* per-channel column sum: out[c] += sum_pixels x  (conv bias gradient) */
int npp_colsum(const npp_view4* x, float* out, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Losses (core/criterion.py).  Logits are read at head resolution (NHWC, channel-padded) and
 * bilinearly upsampled (align_corners=True, :181,194) on the fly to the label resolution;
 * the upsampled tensors never touch HBM.
 *
 * par_loss_pixels: per label pixel p: softmax over classes, prob[p] = p(target) (2.0f for
 *   ignored pixels so they sort last / never pass `< thr`), loss[p] = -w[t]*log p(target).
 *   Also counts valid pixels (count[0], int64).                       (criterion.py:54-64)
 * ohem_select: thr = max(kth smallest prob (k = min(min_kept, n_valid-1)), thres) via a
 *   radix select on the fp32 bit patterns (exactly the element a full sort would pick, :65-67);
 *   out[0] = sum of kept losses, out[1] = number kept (:69-72 mean = out0/out1), out[2] = thr.
 * par_loss_bwd: dlogits (head resolution, fp32 accumulate via atomics then cast by caller)
 *   of  gscale * mean_kept(loss).
 * edge_loss: weighted 2-class CE, weights from pos/neg counts of this batch (:161-166,196).
 * mse_loss: sum over all elements of (pred-gt)^2 (criterion.py:98-128 gives per-joint means;
 *   with equal-sized joints their sum/num_joints == total_sum/(B*J*H*W) * ... see criterion.py).
 * ---------------------------------------------------------------------------------------- */
int npp_par_loss_pixels(const npp_view4* logits, const int64_t* target, int th, int tw,
                        int num_classes, const float* class_w, int ignore_index, float* prob,
                        float* loss, int64_t* count, int dtype, npp_stream_t stream);
int npp_ohem_select(const float* prob, const float* loss, int64_t npix, const int64_t* count,
                    int min_kept, float thres, float* out3, void* workspace,
                    int64_t workspace_bytes, npp_stream_t stream);
int npp_par_loss_bwd(const npp_view4* logits, const int64_t* target, int th, int tw,
                     int num_classes, const float* class_w, int ignore_index, const float* prob,
                     const float* out3, const float* gscale, float* dlogits_f32, int dtype,
                     npp_stream_t stream);
int npp_edge_count(const int64_t* target, int64_t npix, int64_t* posneg, npp_stream_t stream);
int npp_edge_loss_fwd(const npp_view4* logits, const int64_t* target, int th, int tw,
                      int ignore_index, const int64_t* posneg, float* out2, int dtype,
                      npp_stream_t stream);
int npp_edge_loss_bwd(const npp_view4* logits, const int64_t* target, int th, int tw,
                      int ignore_index, const int64_t* posneg, const float* out2,
                      const float* gscale, float* dlogits_f32, int dtype, npp_stream_t stream);
int npp_mse_fwd(const npp_view4* pred, const float* target_nchw, int num_joints, float* out,
                int dtype, npp_stream_t stream);
int npp_mse_bwd(const npp_view4* pred, const float* target_nchw, int num_joints,
                const float* gscale, const npp_view4* dpred, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Integer evaluation kernels — bit-exact.
 * confusion_hist: utils/utils.py:192-218 get_confusion_matrix: argmax over classes (first
 *   maximum on ties), skip label == ignore, hist[gt*C + pred] += 1 (int64 [C*C], caller zeroes).
 *   logits NCHW fp32 (the reference's layout at that call site, function.py:955).
 * heatmap_argmax: core/evaluate.py:13-41 get_max_preds: per (n,j) first arg-max and max value.
 * pck_counts: core/evaluate.py:43-99: hit[j], valid[j] int64 from pred/gt argmax coordinates.
 * ---------------------------------------------------------------------------------------- */
int npp_confusion_hist(const float* logits_nchw, const int64_t* label, int n, int c, int h,
                       int w, int label_h, int label_w, int ignore, int64_t* hist,
                       npp_stream_t stream);
int npp_heatmap_argmax(const float* hm_nchw, int n, int j, int h, int w, int32_t* idx,
                       float* maxval, npp_stream_t stream);
int npp_pck_counts(const int32_t* pred_idx, const float* pred_max, const int32_t* gt_idx,
                   const float* gt_max, int n, int j, int h, int w, float thr, int64_t* hit,
                   int64_t* valid, npp_stream_t stream);

/* MixedOp channel interleave (model_search_interact.py:22-36,70-71 cat + channel_shuffle(2)):
 *   out[..., 2c] = a[..., c], out[..., 2c+1] = b[..., c];  bwd splits. */
int npp_interleave2_fwd(const npp_view4* a, const npp_view4* b, const npp_view4* y, int dtype,
                        npp_stream_t stream);
int npp_interleave2_bwd(const npp_view4* dy, const npp_view4* da, const npp_view4* db,
                        int dtype, npp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NPP_B200_H_ */
