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
long long npp_launch_count(void); /* kernels launched by this library so far (process-wide) */

/* ------------------------------------------------------------------------------------------
 * Dense convolution as implicit GEMM on tcgen05 / TMEM, operands staged by TMA (bf16 only).
 * Replaces nn.Conv2d(groups=1) in ReLUConvBN (models/operations.py:69-82), the pointwise
 * half of DilConvS (:214), FactorizedReduce (:149-150), Pooled_Conv (:239), the stems, layer
 * convs and heads of models/model_augment.py:244-398 and extra_conv (:592-596).
 *   w     bf16 [cout, kh, kw, cin]   (packed OHWI, see npp_pack_weight)
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

/* wgrad: dw fp32 in torch OIHW order [dw_cout, dw_cin, kh, kw] += sum_pixels dy (x) x ; caller zeroes dw first
 * (split-K partial sums are combined with red.global.add.f32).  dw_cout <= dy->c and
 * dw_cin <= x->c are the un-padded channel counts of the fp32 master weight. */
int npp_conv2d_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin,
                     int kh, int kw, int stride, int pad, int dil, int in_h_off, int in_w_off,
                     npp_stream_t stream);

/* Same, with a caller-provided workspace (npp_conv2d_wgrad_workspace_bytes() bytes, 16-byte aligned, reusable by
 * every call on one stream): the split-K partial tiles are written there with plain stores and folded into dw (+=)
 * by a second kernel instead of millions of scattered fp32 atomics.  Too small / NULL = the atomic path. */
int64_t npp_conv2d_wgrad_workspace_bytes(void);
int npp_conv2d_wgrad_ws(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin,
                        int kh, int kw, int stride, int pad, int dil, int in_h_off, int in_w_off,
                        void* workspace, int64_t workspace_bytes, npp_stream_t stream);

/* fp32 master weight in torch OIHW order [cout,cin,taps] -> packed OHWI [cout_pad,taps,cin_pad]
 * (w) and/or its transpose [cin_pad,taps,cout_pad] (wt), zero padded, stored as out_dtype
 * (NPP_BF16 for the tcgen05 path, NPP_F32 for validation mode); either output may be NULL.
 * Activation buffers keep channel counts that are multiples of 8 (16-byte TMA rows), hence
 * the padding. */
int npp_pack_weight(const float* w32, void* w, void* wt, int cout, int taps, int cin,
                    int cout_pad, int cin_pad, int out_dtype, npp_stream_t stream);

/* Every dense-conv weight of a model in one launch (bf16 outputs).  table: device array of ntensors x
 * {const float* w32; bf16* w; bf16* wt; int32 cout, taps, cin, cout_pad, cin_pad, pad} (48 bytes each, either
 * output may be NULL); chunk_tensor / chunk_index: for every block, which tensor and which chunk of chunk_elems
 * packed elements it converts. */
int npp_pack_weights_multi(const void* table, int ntensors, const int32_t* chunk_tensor,
                           const int32_t* chunk_index, int nchunks, int chunk_elems, npp_stream_t stream);

/* Same table, tile form: block b packs the 32 (Cout) x 32 (Cin) x taps tile tile_index[b] (row-major over
 * ceil(cout_pad/32) x ceil(cin_pad/32)) of tensor tile_tensor[b]; taps <= 9; rows with pad != 0 are not supported. */
int npp_pack_weights_tiles(const void* table, int ntensors, const int32_t* tile_tensor,
                           const int32_t* tile_index, int ntiles, npp_stream_t stream);

/* Pixel-pair layout for 3x3 / stride-1 / pad-1 convolutions with 32 input and 32 output channels (the C = 32 cells
 * of the first encoder stage, models/model_augment.py:274-295): [N,H,W,32] read as [N,H,W/2,64] turns the layer into a
 * 64 -> 64 convolution on half as many (full 128-byte) rows.  pack_weight_pair: fp32 OIHW [32,32,3,3] -> bf16
 * [64,9,64] (+ transposed for dgrad); in npp_pack_weights_multi a table row with pad == 1 selects the same mapping.
 * fold_pair_wgrad: dw[32,32,3,3] += the [64,64,3,3] gradient of the paired convolution folded back. */
int npp_pack_weight_pair(const float* w32, void* w, void* wt, npp_stream_t stream);
int npp_fold_pair_wgrad(const float* dw_pair, float* dw, npp_stream_t stream);

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
 *            save_mean, save_invstd; updates running_mean/var[0:c_running] (unbiased var, momentum)
 *            if non-NULL (c_running < c when the activation carries zero padding channels).
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
                    int c_running, npp_stream_t stream);
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
/* 3x3 im2col of a 3-channel bf16 image (x: c == 8 padded) into y: c == 32 with y[.., ci*9 + r*3 + s] =
 * x[.., ho*stride - pad + r, wo*stride - pad + s, ci]: the stem convolutions Conv2d(3, C, 3, 2, 1)
 * (models/model_augment.py:244-272) then run as 1x1 convolutions whose [Cout, 27] weight matrix is the OIHW weight. */
int npp_im2col3x3_c3(const npp_view4* x, const npp_view4* y, int stride, int pad, npp_stream_t stream);
int npp_fill_zero(const npp_view4* y, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Pooling.  maxpool3x3: nn.MaxPool2d(3, stride, 1) (operations.py:55), -inf padding.
 * argmax (optional, dense uint8 [n,ho,wo,c]): winning window position 0..8 = FIRST maximum in
 * scan order (ATen semantics); bwd gathers dy through it.
 * avgpool3x3: nn.AvgPool2d(3, stride, 1, count_include_pad=False) (:57).
 * avgpool2x2: nn.AvgPool2d(2) (:115, :237).   gap: nn.AdaptiveAvgPool2d(1) (:111).
 * ---------------------------------------------------------------------------------------- */
int npp_maxpool3x3_fwd(const npp_view4* x, const npp_view4* y, uint8_t* argmax, int stride,
                       int dtype, npp_stream_t stream);
int npp_maxpool3x3_bwd(const uint8_t* argmax, const npp_view4* dy, const npp_view4* dx,
                       int stride, int dtype, npp_stream_t stream);
int npp_avgpool3x3_fwd(const npp_view4* x, const npp_view4* y, int stride, int dtype,
                       npp_stream_t stream);
int npp_avgpool3x3_bwd(const npp_view4* dy, const npp_view4* dx, int stride, int dtype,
                       npp_stream_t stream);
int npp_avgpool2x2_fwd(const npp_view4* x, const npp_view4* y, int dtype, npp_stream_t stream);
int npp_avgpool2x2_bwd(const npp_view4* dy, const npp_view4* dx, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SE_Block (operations.py:105-129): w = sigmoid(W2 relu(W1 gap(x) + b1) + b2); out = x*w.
 *  gap_fwd:       g[n,c] += mean_hw x           (fp32 [n,c], caller zeroes)
 *  se_fc_fwd:     h = relu(W1 g + b1) [n,c/2];  s = sigmoid(W2 h + b2) [n,c]  (fp32, W [out,in])
 *  se_scale_fwd:  y = x * s[n,c]
 *  se_bwd_reduce: ds[n,c] += sum_hw dy*x        (fp32, caller zeroes)
 *  se_fc_bwd:     from ds: dW2, db2, dW1, db1 (+=) and dg[n,c]
 *  se_bwd_apply:  dx = dy * s[n,c] + dg[n,c]/(h*w)
 * ---------------------------------------------------------------------------------------- */
int npp_gap_fwd(const npp_view4* x, float* g, int dtype, npp_stream_t stream);
int npp_se_fc_fwd(const float* g, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* hbuf, float* s, int n, int c, npp_stream_t stream);
/* se_fc_bwd in two kernels without weight-gradient atomics (same results up to fp32 summation order): per-image
 * vectors first, then dW2 += dz2^T h and dW1 += dz1^T g with one thread per element.  scratch: n * (c + c/2) floats. */
int npp_se_fc_bwd2(const float* g, const float* hbuf, const float* sg, const float* ds, const float* w1,
                   const float* w2, float* dw1, float* db1, float* dw2, float* db2, float* dg,
                   float* scratch, int n, int c, npp_stream_t stream);
int npp_se_scale_fwd(const npp_view4* x, const float* s, const npp_view4* y, int dtype,
                     npp_stream_t stream);
int npp_se_bwd_reduce(const npp_view4* x, const npp_view4* dy, float* ds, int dtype,
                      npp_stream_t stream);
int npp_se_fc_bwd(const float* g, const float* hbuf, const float* s, const float* ds,
                  const float* w1, const float* w2, float* dw1, float* db1, float* dw2,
                  float* db2, float* dg, int n, int c, npp_stream_t stream);
int npp_se_bwd_apply(const npp_view4* dy, const float* s, const float* dg, const npp_view4* dx,
                     int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Resampling.  bilinear: F.interpolate(mode='bilinear', align_corners=True|False)
 * (model_augment.py:116,539-543; operations.py:241; function.py:927 uses align_corners=False).
 * nearest: F.interpolate default mode (model_search_interact.py:63-64).
 * scale_h/scale_w: the scale_factor passed to F.interpolate (0 when the call gave size=); only
 * used where ATen uses it (align_corners=False and nearest).  Backward is a deterministic gather.
 * ---------------------------------------------------------------------------------------- */
int npp_bilinear_fwd(const npp_view4* x, const npp_view4* y, int align_corners, double scale_h,
                     double scale_w, int dtype, npp_stream_t stream);
int npp_bilinear_bwd(const npp_view4* dy, const npp_view4* dx, int align_corners, double scale_h,
                     double scale_w, int dtype, npp_stream_t stream);
/* Separable bilinear backward (same result as npp_bilinear_bwd up to fp32 summation order): column pass into the
 * caller's fp32 scratch tmp[n * dy->h * dx->w * c] (16-byte aligned), then row pass.  Reads dY once. */
int npp_bilinear_bwd_sep(const npp_view4* dy, const npp_view4* dx, float* tmp, int align_corners,
                         double scale_h, double scale_w, int dtype, npp_stream_t stream);
int npp_nearest_fwd(const npp_view4* x, const npp_view4* y, double scale_h, double scale_w,
                    int dtype, npp_stream_t stream);
int npp_nearest_bwd(const npp_view4* dy, const npp_view4* dx, double scale_h, double scale_w,
                    int dtype, npp_stream_t stream);

/* per-channel column sum: out[c] += sum_pixels x  (conv bias gradient) */
int npp_colsum(const npp_view4* x, float* out, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Losses (core/criterion.py) on the reference's boundary tensors: NCHW fp32 logits at head
 * resolution [n,c,h,w], int64 labels [n,lh,lw].  The bilinear upsample to label resolution
 * (align_corners=True, criterion.py:181,194) is evaluated on the fly from a shared-memory
 * tile of head logits; the upsampled [n,c,lh,lw] tensors never exist in HBM.
 *
 * par_loss_pixels (OhemCrossEntropy.forward :54-64): per label pixel softmax; prob[p] =
 *   p(target) (2.0f for ignored pixels: never below any threshold), loss[p] = -w[t]*log p(t)
 *   (0 for ignored); n_valid[0] += number of non-ignored pixels (int64, caller zeroes).
 * ohem_select (:65-72): thr = max(k-th smallest valid prob, thres), k = min(min_kept, n_valid-1)
 *   by an 8-bit x 4-pass radix select on the fp32 bit patterns — exactly the element the
 *   reference's full sort() picks; out3 = {sum of kept losses, number kept, thr}; the loss is
 *   out3[0]/out3[1] (plain mean over pixels with prob < thr).  workspace: >= 4200 bytes, zeroed.
 * par_loss_bwd: dlogits[n,c,h,w] += gscale[0]/n_kept * w[t]*(softmax - onehot) pulled back
 *   through the bilinear upsample (caller zeroes dlogits).
 * edge_*: weighted 2-class CE (:161-166,196): weights [pos/(pos+neg), neg/(pos+neg)] from the
 *   batch label counts posneg = {#(t==1), #(t==0)}; out2 = {sum w[t]*nll, sum w[t]};
 *   loss = out2[0]/out2[1].
 * mse_*: out[0] += sum (w*(pred-target))^2 over n elements, w = row_w[i / row_len] or 1 when
 *   row_w is NULL (criterion.py:100-104 target_weight); dpred = gscale[0]*2*w^2*(pred-target).
 * ---------------------------------------------------------------------------------------- */
int npp_par_loss_pixels(const float* logits, int n, int c, int h, int w, const int64_t* target,
                        int lh, int lw, const float* class_w, int ignore_index,
                        int align_corners, float* prob, float* loss, int64_t* n_valid,
                        npp_stream_t stream);
int npp_ohem_select(const float* prob, const float* loss, int64_t npix, const int64_t* n_valid,
                    int min_kept, float thres, float* out3, void* workspace,
                    npp_stream_t stream);
int npp_par_loss_bwd(const float* logits, int n, int c, int h, int w, const int64_t* target,
                     int lh, int lw, const float* class_w, int ignore_index, int align_corners,
                     const float* prob, const float* out3, const float* gscale, float* dlogits,
                     npp_stream_t stream);
/* Per-pixel form of the two cross-entropy backward passes: G[n, y, x, 0:cq] (fp32 NHWC at LABEL resolution, cq =
 * classes rounded up to a multiple of 4, 16-byte aligned) = d loss / d up-sampled logit, zero for ignored /
 * unselected pixels and padding channels.  d loss / d head-logit = bilinear backward of G (npp_bilinear_bwd_sep with
 * the same align_corners), converted to NCHW — same result as npp_par_loss_bwd / npp_edge_loss_bwd up to fp32
 * summation order, without the shared-memory atomics. */
int npp_par_loss_grad_pixels(const float* logits, int n, int c, int h, int w, const int64_t* target, int lh,
                             int lw, const float* class_w, int ignore_index, int align_corners,
                             const float* prob, const float* out3, const float* gscale, float* G, int cq,
                             npp_stream_t stream);
int npp_edge_loss_grad_pixels(const float* logits, int n, int h, int w, const int64_t* target, int lh, int lw,
                              int ignore_index, int align_corners, const int64_t* posneg, const float* out2,
                              const float* gscale, float* G, int cq, npp_stream_t stream);
int npp_edge_count(const int64_t* target, int64_t npix, int64_t* posneg, npp_stream_t stream);
int npp_edge_loss_fwd(const float* logits, int n, int h, int w, const int64_t* target, int lh,
                      int lw, int ignore_index, int align_corners, const int64_t* posneg,
                      float* out2, npp_stream_t stream);
int npp_edge_loss_bwd(const float* logits, int n, int h, int w, const int64_t* target, int lh,
                      int lw, int ignore_index, int align_corners, const int64_t* posneg,
                      const float* out2, const float* gscale, float* dlogits,
                      npp_stream_t stream);
int npp_mse_fwd(const float* pred, const float* target, int64_t n, const float* row_w,
                int64_t row_len, float* out, npp_stream_t stream);
int npp_mse_bwd(const float* pred, const float* target, int64_t n, const float* row_w,
                int64_t row_len, const float* gscale, float* dpred, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Integer evaluation kernels — bit-exact.
 * confusion_hist: utils/utils.py:192-218 get_confusion_matrix: argmax over classes (first
 *   maximum on ties, numpy.argmax), skip label == ignore, hist[gt*C + pred] += 1
 *   (int64 [C*C], caller zeroes).  logits NCHW fp32 [n,c,h,w]; label int64 [n,label_h,label_w]
 *   cropped to [:h,:w] like the reference (:201-202).
 * tta_merge: core/function.py:927-939 flip-test merge: bilinear (align_corners=False) resize of
 *   both predictions to (oh,ow), the reference's aliasing left/right channel "swap" (channels
 *   14,16,18 of the flipped prediction read 15,17,19; 15,17,19 keep their own), horizontal flip,
 *   0.5*(a+b).  swap_lr=0 gives the pascal variant (function_ppp.py:923-928).
 * heatmap_argmax: core/evaluate.py:13-41 get_max_preds: per (n,j) first arg-max and max value.
 * pck_counts: core/evaluate.py:43-99: per joint hit/valid counters from pred/gt arg-max
 *   (valid: gt x>=1 or y>=1; hit: ||(p-t)/(h/10, w/10)|| < thr), int64, caller zeroes.
 * ---------------------------------------------------------------------------------------- */
int npp_confusion_hist(const float* logits_nchw, const int64_t* label, int n, int c, int h,
                       int w, int label_h, int label_w, int ignore, int64_t* hist,
                       npp_stream_t stream);
int npp_tta_merge(const float* pred, const float* flip_pred, int n, int c, int h, int w, int oh,
                  int ow, int swap_lr, float* out, npp_stream_t stream);
int npp_heatmap_argmax(const float* hm_nchw, int nj, int h, int w, int32_t* idx, float* maxval,
                       npp_stream_t stream);
/* LIP pose post-process of validate_sync (core/function.py:962-986), NCHW fp32:
 *   pose_merge: out[n,j] = 0.5*(cv2.resize(pred[n,j], (ow,oh), INTER_LINEAR) +
 *               cv2.flip(cv2.resize(flip_pred[n, flip_idx[j]], (ow,oh), INTER_LINEAR), 1))      (:973-979);
 *               flip_idx: HOST array of nj (<= 32) joint indices (flipped_poseidx, :908).
 *   gaussian_filter: scipy.ndimage.gaussian_filter(x, sigma) per [h,w] plane (:980): separable, rows then columns,
 *               radius int(truncate*sigma+0.5), mode 'reflect', fp64 accumulation, each pass rounded to fp32; tmp is a
 *               caller-provided scratch of the same size; dst may alias src.
 *   The peak (:981-986) is npp_heatmap_argmax on the filtered maps. */
int npp_pose_merge(const float* pred, const float* flip_pred, int n, int nj, int h, int w,
                   const int* flip_idx, int oh, int ow, float* out, npp_stream_t stream);
int npp_gaussian_filter(const float* src, float* tmp, float* dst, int planes, int h, int w, double sigma,
                        double truncate, npp_stream_t stream);
/* pascal validate_sync (core/function_ppp.py:905,957-958): out[n,j] = 0.5*(pred[n,j] + flip_pred[n, flip_idx[j]]) in
 * heat-map space (the mirrored image's maps are joint-permuted, not mirrored back); out may alias pred. */
int npp_heatmap_flip_avg(const float* pred, const float* flip_pred, int n, int nj, int h, int w,
                         const int* flip_idx, float* out, npp_stream_t stream);
int npp_pck_counts(const int32_t* pred_idx, const float* pred_max, const int32_t* gt_idx,
                   const float* gt_max, int n, int j, int h, int w, float thr, int64_t* hit,
                   int64_t* valid, npp_stream_t stream);
/* pckh_counts: utils/calc_pckh.py:35-97 on fp64 image-space coordinates pred/gt [n,p,2]
 *   (gt < 0 marks a missing joint): per joint valid (head size != 0 and joint present) and hit
 *   (||gt-pred||/head <= thr) counters; PCKh = 100*hit/valid. */
int npp_pckh_counts(const double* pred, const double* gt, int n, int p, double thr, int64_t* hit,
                    int64_t* valid, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor Adam (torch.optim.Adam arithmetic, no amsgrad; augment_lip_sync.py:210-212):
 * tensor_table: device array of ntensors x {float* p; const float* g; float* m; float* v; int64 n;
 * float lr; float wd; int64* step} (56 bytes each; step = that parameter's own device counter, as
 * torch keeps one per parameter — incremented by the call); chunk_tensor/chunk_index: for every
 * block, which tensor and which chunk of chunk_elems elements it updates.
 * ---------------------------------------------------------------------------------------- */
int npp_adam_step(const void* tensor_table, int ntensors, const int32_t* chunk_tensor,
                  const int32_t* chunk_index, int nchunks, int chunk_elems, float beta1, float beta2,
                  float eps, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused cell node (csrc/node.cu).  A cell node is s = op_a(h_a) + op_b(h_b)
 * (models/model_augment.py:48-62, 90-106, 153-174); primitives end in BatchNorm2d
 * (models/operations.py:61,79,97,215,240) and consumers start with nn.ReLU (:76,95,212,239).
 *   node_fwd: y = f_a(a) [+ f_b(b)], f = x*scale[c]+shift[c] (scale == NULL: identity), written
 *     as y_raw and/or relu(y) = y_relu (either may be NULL, not both; views may be channel slices
 *     of a concat buffer — replaces torch.cat, model_augment.py:62).
 *   node_bwd_reduce: g = g_raw + [relu_out > 0]*g_relu (either gradient may be NULL; g_out is
 *     written when given) and, per BatchNorm input (a / b non-NULL), per-block partial sums of
 *     (sum g, sum g*xhat) into partials[blocks][nq][C] with nq = 2 per BatchNorm input and
 *     blocks = npp_node_bwd_blocks(n,h,w,c,dtype); npp_reduce_partials(partials, blocks, nq*C, out)
 *     folds them (deterministic, no atomics) into out = [sum g, sum g*xhat_a, sum g, sum g*xhat_b].
 *   node_bwd_apply: d_in = gamma*invstd*(g - s1/count - xhat*s2/count) per BatchNorm input, both
 *     from one read of g; sums_x = [2][C] as produced above (all-reduced across ranks for SyncBN,
 *     with count the global pixel count).
 * ---------------------------------------------------------------------------------------- */
int npp_node_fwd(const npp_view4* a, const float* scale_a, const float* shift_a, const npp_view4* b,
                 const float* scale_b, const float* shift_b, const npp_view4* y_raw,
                 const npp_view4* y_relu, int dtype, npp_stream_t stream);
/* node_fwd with the BatchNorm finalize of either side folded into the kernel (replaces one npp_bn_finalize launch
 * per side): fin_x != NULL => side x is normalised with the batch statistics in fin_x->stats ([2][C] sums over
 * `count` elements per channel), the kernel writes fin_x->coef = [scale | shift | mean | invstd] (4C floats, the
 * backward kernels read it) and updates running_mean / running_var[0:c_run] with `momentum` (unbiased variance), as
 * nn.BatchNorm2d does in training mode (models/operations.py:27,61,79).  fin_x == NULL => scale_x / shift_x as in
 * npp_node_fwd. */
typedef struct {
  const float* stats;
  const float* gamma;        /* NULL = 1 */
  const float* beta;         /* NULL = 0 */
  float* running_mean;       /* NULL = not tracked */
  float* running_var;
  float* coef;
  float momentum, eps;
  int32_t c_run;
} npp_bn_fin;
int npp_node_fwd_bn(const npp_view4* a, const npp_bn_fin* fin_a, const float* scale_a, const float* shift_a,
                    const npp_view4* b, const npp_bn_fin* fin_b, const float* scale_b, const float* shift_b,
                    const npp_view4* y_raw, const npp_view4* y_relu, double count, int dtype,
                    npp_stream_t stream);
int npp_node_bwd_blocks(int n, int h, int w, int c, int dtype);
int npp_node_bwd_reduce(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                        const npp_view4* a, const float* mean_a, const float* invstd_a,
                        const npp_view4* b, const float* mean_b, const float* invstd_b,
                        const npp_view4* g_out, float* partials, int dtype, npp_stream_t stream);
int npp_reduce_partials(const float* partials, int rows, int len, float* out, npp_stream_t stream);
/* node_bwd_reduce without the partials buffer / second kernel: every block adds its per-channel sums with fp32
 * atomics into sums = [sum g, sum g*xhat_a, sum g, sum g*xhat_b] ([nq][C], zeroed by the caller) and, when
 * acc[i] != NULL, also into acc[i][0:acc_valid[i]] (the BatchNorm parameters' gradient slots: d beta, d gamma per
 * side).  acc / acc_valid: HOST arrays of nq entries, or NULL.  Summation order is not deterministic. */
int npp_node_bwd_reduce_atomic(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                               const npp_view4* a, const float* mean_a, const float* invstd_a,
                               const npp_view4* b, const float* mean_b, const float* invstd_b,
                               const npp_view4* g_out, float* sums, float* const* acc,
                               const int* acc_valid, int dtype, npp_stream_t stream);
/* Two-kernel BatchNorm backward of a node without the partials buffer and the fold kernel: reduce blocks add their
 * per-channel sums with fp32 atomics into copy (block % stripes) of sums = [stripes][nq][C] (zeroed by the caller; the
 * striping keeps same-address atomic chains short), apply folds the copies while it builds its coefficients and its
 * first pixel block adds d beta / d gamma into acc[i][0:acc_valid[i]] (HOST arrays of nq entries in sums-row order,
 * entries may be NULL; acc may be NULL).  Not usable when the sums must be all-reduced in between (SyncBN). */
int npp_node_bwd_reduce_striped(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                                const npp_view4* a, const float* mean_a, const float* invstd_a,
                                const npp_view4* b, const float* mean_b, const float* invstd_b,
                                const npp_view4* g_out, float* sums, int stripes, int dtype, npp_stream_t stream);
/* General form of the first backward kernel: g = g_raw + g_raw2 + [relu_out > 0] * (g_relu + g_relu2).  g_raw2 /
 * g_relu2 are second gradients of the same outputs (the channel slice of the cell-output gradient that the consumers
 * of the concat buffer hand down, model_augment.py:62) — summed here instead of by a separate strided add kernel;
 * g_x2 requires g_x.  Exactly one of partials (per-block rows, fold with npp_reduce_partials) / sums (striped atomic
 * totals, see above) when a BatchNorm side exists. */
int npp_node_bwd_reduce2(const npp_view4* g_raw, const npp_view4* g_raw2, const npp_view4* g_relu,
                         const npp_view4* g_relu2, const npp_view4* relu_out, const npp_view4* a,
                         const float* mean_a, const float* invstd_a, const npp_view4* b, const float* mean_b,
                         const float* invstd_b, const npp_view4* g_out, float* partials, float* sums,
                         int stripes, int dtype, npp_stream_t stream);
/* Same reduction with two type-flexible extra gradient inputs: an output of the node that is read by several consumers
 * (more primitives of the cell, the concat route) hands every consumer its own handle, so up to three gradients per
 * node arrive separately and are summed HERE instead of by autograd's accumulation kernels (one add launch and three
 * tensor passes per extra consumer in the reference).  extraK_is_relu: the slot is a gradient of the relu output
 * (masked like g_relu) or of the raw output; the primary gradient of that kind must be present; slots fill in order. */
int npp_node_bwd_reduce3(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                         const npp_view4* extra0, int extra0_is_relu, const npp_view4* extra1, int extra1_is_relu,
                         const npp_view4* a, const float* mean_a, const float* invstd_a, const npp_view4* b,
                         const float* mean_b, const float* invstd_b, const npp_view4* g_out, float* partials, int dtype,
                         npp_stream_t stream);
int npp_node_bwd_apply_striped(const npp_view4* g, const npp_view4* a, const float* gamma_a, const float* mean_a,
                               const float* invstd_a, const npp_view4* da, const npp_view4* b,
                               const float* gamma_b, const float* mean_b, const float* invstd_b,
                               const npp_view4* db, const float* sums, int stripes, float* const* acc,
                               const int* acc_valid, double count, int dtype, npp_stream_t stream);
/* reduce_partials that also accumulates segment s = columns [s*seg_len,(s+1)*seg_len) of the folded row into
 * acc[s][0:acc_valid[s]] (+=; NULL = skip): BatchNorm d beta / d gamma go straight into the optimizer's flat
 * gradient buffer.  acc / acc_valid: HOST arrays of len/seg_len (<= NPP_ACC_MAX) entries. */
#define NPP_ACC_MAX 16
int npp_reduce_partials_acc(const float* partials, int rows, int len, float* out, int seg_len,
                            float* const* acc, const int* acc_valid, npp_stream_t stream);
int npp_node_bwd_apply(const npp_view4* g, const npp_view4* a, const float* gamma_a,
                       const float* mean_a, const float* invstd_a, const float* sums_a,
                       const npp_view4* da, const npp_view4* b, const float* gamma_b,
                       const float* mean_b, const float* invstd_b, const float* sums_b,
                       const npp_view4* db, double count, int dtype, npp_stream_t stream);

/* MixedOp channel interleave (model_search_interact.py:22-36,70-71 cat + channel_shuffle(2)):
 *   out[..., 2c] = a[..., c], out[..., 2c+1] = b[..., c];  bwd splits. */
int npp_interleave2_fwd(const npp_view4* a, const npp_view4* b, const npp_view4* y, int dtype,
                        npp_stream_t stream);
int npp_interleave2_bwd(const npp_view4* dy, const npp_view4* da, const npp_view4* db,
                        int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Search supernet: MixedOp weighted sum and beta-weighted node sums (csrc/mix.cu).
 * models/model_search_interact.py:56-74 MixedOp.forward: temp1 = sum_k w_k*op_k(x[:, :C/2]);
 * ans = channel_shuffle(cat(temp1, x[:, C/2:]), 2), i.e. ans[2c] = temp1[c], ans[2c+1] = x[C/2+c];
 * :352-356, :648-654 node sums  s = base + sum_j beta_j * MixedOp_j(h_j, alpha_j).
 * A branch is either a plain tensor or the not-yet-normalised input of a BatchNorm2d (every
 * candidate primitive ends in one, operations.py:61,79,215,240; model_search_interact.py:48-49)
 * described by scale/shift (forward) and mean/invstd/gamma (backward).
 *   mix_fwd:        out[.., c] = sum_k w[k]*(y_k*scale_k + shift_k)[c]; with desc.interleave the
 *                   output has 2C channels: out[.., 2c] = that sum, out[.., 2c+1] = pass[.., c].
 *                   w: device fp32 [k] (NULL = all ones: plain n-ary sum).
 *   mix_bwd_reduce: per-block partials [blocks][k+1][C] of S_0 = sum g and S_{1+k} = sum g*xhat_k
 *                   (BatchNorm branch) / sum g*y_k (plain); g = dout[.., 2c] when interleaved;
 *                   blocks = npp_node_bwd_blocks(n,h,w,C,dtype); fold with npp_reduce_partials.
 *   mix_dw:         dw[k] = sum_c gamma_k[c]*S_{1+k}[c] + beta_k[c]*S_0[c]  (= <g, branch_k>, the
 *                   architecture-weight gradient); beta: host array of k device pointers or NULL.
 *   mix_bwd_apply:  desc.dy[k] = w[k]*gamma_k*invstd_k*(g - S_0/count - xhat_k*S_{1+k}/count)
 *                   (BatchNorm) or w[k]*g (plain) for every k with dy[k].ptr != NULL, and
 *                   dpass = dout[.., 2c+1]; sums = [k+1][C] (all-reduced over ranks for SyncBN).
 * ---------------------------------------------------------------------------------------- */
#define NPP_MIX_MAX 8
typedef struct {
  int32_t k;          /* number of branches, 1..NPP_MIX_MAX */
  int32_t interleave; /* 1: the mixed half is interleaved with a pass-through half */
  npp_view4 y[NPP_MIX_MAX];
  const float* scale[NPP_MIX_MAX]; /* forward:  NULL = plain branch */
  const float* shift[NPP_MIX_MAX];
  const float* mean[NPP_MIX_MAX];  /* backward: NULL = plain branch */
  const float* invstd[NPP_MIX_MAX];
  const float* gamma[NPP_MIX_MAX]; /* backward: NULL = 1 */
  npp_view4 dy[NPP_MIX_MAX];       /* backward outputs; ptr NULL = not wanted */
} npp_mix_desc;

int npp_mix_fwd(const npp_mix_desc* d, const float* w, const npp_view4* pass, const npp_view4* out,
                int dtype, npp_stream_t stream);
int npp_mix_bwd_reduce(const npp_mix_desc* d, const npp_view4* g, float* partials, int dtype,
                       npp_stream_t stream);
int npp_mix_dw(const npp_mix_desc* d, const float* sums, const float* const* beta, float* dw,
               npp_stream_t stream);
int npp_mix_bwd_apply(const npp_mix_desc* d, const npp_view4* g, const float* w, const float* sums,
                      double count, const npp_view4* dpass, int dtype, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * On-GPU label synthesis (csrc/labels.cu) — what the reference's data loader builds per image on the
 * host (dataset/data_loader.py:239-285), for a whole batch per launch.
 *   pose_target:  dataset/target_generation.py:94-121,146-168 gen_pose_target + gen_single_gaussian_map.
 *                 joints fp64 [b, j, 2] (x, y in crop pixels), visible int32 [b, j]; out fp32 [b, j+1, grid_y, grid_x]
 *                 (channel j = background = 1 - max); call twice (sigma, 2*sigma) for the aux maps (:109-121).
 *   edge_label:   target_generation.py:210-239 generate_edge (+ data_loader.py:284 edge[label==255] = 255):
 *                 label / out int64 [b, h, w]; edge_width odd (3).
 *   flip_parsing: target_generation.py:44-56 (mirror + left/right relabel); out must not alias label.
 * ---------------------------------------------------------------------------------------- */
int npp_pose_target(const double* joints, const int* visible, int b, int j, double stride, int grid_x, int grid_y,
                    double sigma, float* out, npp_stream_t stream);
int npp_edge_label(const int64_t* label, int b, int h, int w, int edge_width, int64_t* out, npp_stream_t stream);
int npp_flip_parsing(const int64_t* label, int b, int h, int w, int64_t* out, npp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SyncBN statistics exchange over NVLink peer memory (csrc/peer.cu).
 * Replaces the per-BatchNorm collectives of torch.nn.SyncBatchNorm that the reference installs with
 * nn.SyncBatchNorm.convert_sync_batchnorm (augment_lip_sync.py:191, search_lip_sync.py:268): forward all_gather of
 * (mean, invstd, count), backward all_reduce of (sum_dy, sum_dy_xmu) — here one all-reduce (SUM) of the raw
 * 2C..4C-float vectors per cell node and direction, done by ONE single-block kernel that reads the peers' staging
 * buffers directly (one-shot all-reduce; rank-ordered sums, bit-identical on all ranks).
 *
 * One process per GPU.  Setup (host, once): every rank allocates a communication buffer (npp_peer_alloc: cudaMalloc of
 * npp_peer_buffer_bytes(), zeroed), exports it (npp_peer_export: 64-byte cudaIpcMemHandle_t), the handles travel
 * through the caller's own control plane (torch.distributed all_gather_object), every rank maps its peers' buffers
 * (npp_peer_open) and fills an npp_peer_comm.  All ranks must issue the same sequence of npp_peer_allreduce calls.
 *   src0/dst0/n0 (+ optional src1/dst1/n1): fp32 vectors, n0 + n1 <=
 *   NPP_PEER_MAX_FLOATS; dst may alias src.  timeout_ms (0 = 30 s): a peer that does not show up sets the error word
 *   (npp_peer_status; the kernel then continues with whatever it can read instead of hanging the GPU).
 * ---------------------------------------------------------------------------------------- */
#define NPP_PEER_MAX_RANKS 8
#define NPP_PEER_RING 4
#define NPP_PEER_MAX_FLOATS 8192
#define NPP_PEER_HANDLE_BYTES 64
typedef struct {
  void* bufs[NPP_PEER_MAX_RANKS]; /* bufs[p]: rank p's communication buffer as mapped in THIS process */
  int32_t rank, world;
  int32_t timeout_ms;
  int32_t reserved;
} npp_peer_comm;

int64_t npp_peer_buffer_bytes(void);
int npp_peer_alloc(void** ptr);
int npp_peer_free(void* ptr);
int npp_peer_export(void* ptr, void* handle64);
int npp_peer_open(const void* handle64, void** ptr);
int npp_peer_close(void* ptr);
int npp_peer_allreduce(const npp_peer_comm* comm, const float* src0, float* dst0, int n0, const float* src1,
                       float* dst1, int n1, npp_stream_t stream);
int npp_peer_status(const npp_peer_comm* comm, unsigned int* seq, unsigned int* error);

#ifdef __cplusplus
}
#endif
#endif /* NPP_B200_H_ */
