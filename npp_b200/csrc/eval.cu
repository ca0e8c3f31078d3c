// Evaluation kernels.  The integer ones are bit-exact restatements of the reference's numpy code:
//   confusion_hist  : utils/utils.py:192-218 get_confusion_matrix
//   heatmap_argmax  : core/evaluate.py:13-41 get_max_preds
//   pck_counts      : core/evaluate.py:43-99 calc_dists / dist_acc / accuracy   (counts only)
//   pckh_counts     : utils/calc_pckh.py:35-97 get_head_size / get_norm_dist / compute_pck (counts only)
//   tta_merge       : core/function.py:927-939 flip-test merge of the parsing logits
// Counts are int64; every floating-point comparison that decides a count is done in fp64 with the
// same operation order as numpy (no FMA contraction) so the counters match the reference exactly.
#include "common.cuh"
#include <math.h>
#include "resample.cuh"

namespace npp {

// ---- confusion matrix ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ label, int C, int H, int W, int LH,
                 int LW, int ignore, unsigned long long* __restrict__ hist) {
  pdl_wait();
  extern __shared__ unsigned int sh[];  // C*C
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int n = blockIdx.y;
  const int64_t hw = (int64_t)H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < hw; p += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p % W);
    if (y >= LH || x >= LW) continue;
    const int64_t gt = label[((int64_t)n * LH + y) * LW + x];
    if (gt == ignore) continue;
    const float* lp = logits + (int64_t)n * C * hw + p;
    float best = lp[0];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = lp[(int64_t)c * hw];
      if (v > best) { best = v; arg = c; }  // first maximum wins ties (numpy.argmax)
    }
    if (gt < 0 || gt >= C) continue;  // index beyond C*C: never read back by the reference (:211-217)
    atomicAdd(&sh[(int)gt * C + arg], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// ---- flip-test merge ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tta_merge_kernel(const float* __restrict__ pred, const float* __restrict__ flip, int N, int C, int H, int W, int OH,
                 int OW, Axis ah, Axis aw, int swap_lr, float* __restrict__ out) {
  pdl_wait();
  const int64_t total = (int64_t)N * C * OH * OW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const int y = (int)((i / OW) % OH);
    const int c = (int)((i / ((int64_t)OW * OH)) % C);
    const int n = (int)(i / ((int64_t)OW * OH * C));
    // function.py:931-937: `tmp` aliases flip_pred_par, so 14<-15 then 15<-(new)14: both end up as old 15, etc.
    int cf = c;
    if (swap_lr && c >= 14 && c <= 19) cf = c | 1;
    int h0, h1, w0, w1; float lh0, lh1, lw0, lw1;
    bilinear_taps(ah, y, h0, h1, lh0, lh1);
    bilinear_taps(aw, x, w0, w1, lw0, lw1);
    const float* a = pred + ((int64_t)n * C + c) * H * W;
    const float va = lh0 * (lw0 * a[h0 * W + w0] + lw1 * a[h0 * W + w1]) + lh1 * (lw0 * a[h1 * W + w0] + lw1 * a[h1 * W + w1]);
    const int xf = OW - 1 - x;
    bilinear_taps(aw, xf, w0, w1, lw0, lw1);
    const float* b = flip + ((int64_t)n * C + cf) * H * W;
    const float vb = lh0 * (lw0 * b[h0 * W + w0] + lw1 * b[h0 * W + w1]) + lh1 * (lw0 * b[h1 * W + w0] + lw1 * b[h1 * W + w1]);
    out[i] = 0.5f * (va + vb);
  }
}

// ---- LIP pose post-process of validate_sync (core/function.py:962-986) --------------------------
//   merged[n, j] = 0.5 * (cv2.resize(pred[n, j]) + cv2.flip(cv2.resize(flip[n, flipped_poseidx[j]]), 1))
// cv2.resize(INTER_LINEAR) on float32: half-pixel centres, source index clamped into the image with weight 0 at the
// borders, horizontal pass then vertical pass, every product and sum rounded to float32 (no FMA contraction — the
// arg-max that follows decides integer pixel coordinates).  Restated and pinned in oracle/pose_post_ref.py.
__device__ __forceinline__ void cv2_linear_tap(int d, double scale, int src, int& i0, int& i1, float& w1) {
  double f = ((double)d + 0.5) * scale - 0.5;
  int s = (int)floor(f);
  f -= (double)s;
  if (s < 0) { s = 0; f = 0.0; }
  if (s >= src - 1) { s = src - 1; f = 0.0; }
  i0 = s;
  i1 = s + 1 < src ? s + 1 : src - 1;
  w1 = (float)f;
}
__device__ __forceinline__ float cv2_bilinear(const float* __restrict__ img, int W, int y0, int y1, float ay, int x0,
                                              int x1, float ax) {
  const float bx = __fsub_rn(1.0f, ax), by = __fsub_rn(1.0f, ay);
  const float r0 = __fadd_rn(__fmul_rn(img[y0 * W + x0], bx), __fmul_rn(img[y0 * W + x1], ax));
  const float r1 = __fadd_rn(__fmul_rn(img[y1 * W + x0], bx), __fmul_rn(img[y1 * W + x1], ax));
  return __fadd_rn(__fmul_rn(r0, by), __fmul_rn(r1, ay));
}

struct FlipIdx { int idx[32]; };

__global__ void __launch_bounds__(256)
pose_merge_kernel(const float* __restrict__ pred, const float* __restrict__ flip, int N, int J, int H, int W, int OH,
                  int OW, FlipIdx fi, float* __restrict__ out) {
  pdl_wait();
  const double sy = (double)H / (double)OH, sx = (double)W / (double)OW;
  const int64_t total = (int64_t)N * J * OH * OW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const int y = (int)((i / OW) % OH);
    const int j = (int)((i / ((int64_t)OW * OH)) % J);
    const int n = (int)(i / ((int64_t)OW * OH * J));
    int y0, y1, x0, x1;
    float ay, ax;
    cv2_linear_tap(y, sy, H, y0, y1, ay);
    cv2_linear_tap(x, sx, W, x0, x1, ax);
    const float va = cv2_bilinear(pred + ((int64_t)n * J + j) * H * W, W, y0, y1, ay, x0, x1, ax);
    cv2_linear_tap(OW - 1 - x, sx, W, x0, x1, ax);   // cv2.flip(.., 1) of the resized flipped map
    const float vb = cv2_bilinear(flip + ((int64_t)n * J + fi.idx[j]) * H * W, W, y0, y1, ay, x0, x1, ax);
    out[i] = __fmul_rn(__fadd_rn(va, vb), 0.5f);
  }
}

// scipy.ndimage.gaussian_filter: one separable pass along `axis` (0 = rows, 1 = columns) of [planes, H, W] float32
// maps, mode 'reflect' (d c b a | a b c d | d c b a), float64 accumulation in tap order, result rounded to float32.
struct GaussTaps { double k[64]; int radius; };

__device__ __forceinline__ int reflect_index(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - i - 1;
  return i;
}

__global__ void __launch_bounds__(256)
gauss1d_reflect_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes, int H, int W, int axis,
                       GaussTaps g) {
  pdl_wait();
  const int64_t total = (int64_t)planes * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const float* p = src + (i / ((int64_t)W * H)) * (int64_t)W * H;
    double acc = 0.0;
    for (int t = 0; t <= 2 * g.radius; ++t) {
      const int o = t - g.radius;
      const float v = axis == 0 ? p[reflect_index(y + o, H) * W + x] : p[y * W + reflect_index(x + o, W)];
      acc = __dadd_rn(acc, __dmul_rn(g.k[t], (double)v));
    }
    dst[i] = (float)acc;
  }
}

// ---- heat-map arg-max ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
heatmap_argmax_kernel(const float* __restrict__ hm, int hw, int32_t* __restrict__ idx, float* __restrict__ maxval) {
  pdl_wait();
  __shared__ float sv[256];
  __shared__ int si[256];
  const float* p = hm + (int64_t)blockIdx.x * hw;
  float best = -3.4e38f;
  int arg = 0x7fffffff;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float v = p[i];
    if (v > best) { best = v; arg = i; }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = arg;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float v = sv[threadIdx.x + s];
      const int a = si[threadIdx.x + s];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && a < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = a; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { idx[blockIdx.x] = si[0]; maxval[blockIdx.x] = sv[0]; }
}

// ---- PCK on arg-max coordinates (heat-map space) ----------------------------------------------------------------
__global__ void pck_counts_kernel(const int32_t* __restrict__ pidx, const float* __restrict__ pmax,
                                  const int32_t* __restrict__ gidx, const float* __restrict__ gmax, int N, int J, int H,
                                  int W, double thr, unsigned long long* __restrict__ hit,
                                  unsigned long long* __restrict__ valid) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * J) return;
  const int j = i % J;
  // get_max_preds: x = idx % W, y = floor(idx / W), zeroed when maxval <= 0 (evaluate.py:33-40)
  const float px = pmax[i] > 0.f ? (float)(pidx[i] % W) : 0.f, py = pmax[i] > 0.f ? (float)(pidx[i] / W) : 0.f;
  const float tx = gmax[i] > 0.f ? (float)(gidx[i] % W) : 0.f, ty = gmax[i] > 0.f ? (float)(gidx[i] / W) : 0.f;
  if (tx < 1.f && ty < 1.f) return;  // calc_dists: dists = -1 (evaluate.py:49-50)
  // norm = ones((N,2)) * [h, w] / 10 applied to (x, y) — x is divided by h/10, as in the reference (:82-84)
  const double nx = (double)H / 10.0, ny = (double)W / 10.0;
  const double dx = __dsub_rn(__ddiv_rn((double)px, nx), __ddiv_rn((double)tx, nx));
  const double dy = __dsub_rn(__ddiv_rn((double)py, ny), __ddiv_rn((double)ty, ny));
  const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  atomicAdd(valid + j, 1ull);
  if (d < thr) atomicAdd(hit + j, 1ull);
}

// ---- PCKh on image-space coordinates (LIP csv evaluation) ----------------------------------------------------------
__global__ void pckh_counts_kernel(const double* __restrict__ pred, const double* __restrict__ gt, int N, int P,
                                   double thr, unsigned long long* __restrict__ hit,
                                   unsigned long long* __restrict__ valid) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * P) return;
  const int n = i / P, p = i % P;
  const double* g = gt + (int64_t)n * P * 2;
  // get_head_size (calc_pckh.py:35-41): ||gt[9]-gt[8]||, 0 when either x coordinate is negative
  const double hx = __dsub_rn(g[9 * 2], g[8 * 2]), hy = __dsub_rn(g[9 * 2 + 1], g[8 * 2 + 1]);
  double head = __dsqrt_rn(__dadd_rn(__dmul_rn(hx, hx), __dmul_rn(hy, hy)));
  if (g[8 * 2] < 0 || g[9 * 2] < 0) head = 0;
  if (head == 0) return;                          // dist = -1 for the whole image (:49-50)
  if (g[p * 2] < 0 || g[p * 2 + 1] < 0) return;  // dist = -1 for a missing joint (:54-55)
  const double dx = __dsub_rn(g[p * 2], pred[(int64_t)i * 2]), dy = __dsub_rn(g[p * 2 + 1], pred[(int64_t)i * 2 + 1]);
  const double d = __ddiv_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), head);
  if (!(d >= 0)) return;
  atomicAdd(valid + p, 1ull);
  if (d <= thr) atomicAdd(hit + p, 1ull);
}

// pascal validate_sync (core/function_ppp.py:957-958): pred[:, j] = 0.5 * (pred[:, j] + flip_pred[:, flipped_poseidx[j]])
// in heat-map space — the mirrored image's maps are joint-permuted but NOT mirrored back (reference behaviour).
__global__ void heatmap_flip_avg_kernel(const float* __restrict__ pred, const float* __restrict__ flip, int nj, int hw,
                                        FlipIdx fi, int64_t total, float* __restrict__ out) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int px = (int)(i % hw);
    const int64_t t = i / hw;
    const int j = (int)(t % nj);
    const int64_t n = t / nj;
    out[i] = 0.5f * (pred[i] + flip[(n * nj + fi.idx[j]) * hw + px]);
  }
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_confusion_hist(const float* logits, const int64_t* label, int n, int c, int h, int w, int label_h, int label_w,
                       int ignore, int64_t* hist, npp_stream_t s) {
  if (!logits || !label || !hist || n <= 0 || c <= 0 || c > 64 || h <= 0 || w <= 0) return NPP_E_INVALID;
  const int64_t hw = (int64_t)h * w;
  int gx = (int)((hw + 255) / 256);
  const int cap = sm_count() * 8 / n + 1;
  if (gx > cap) gx = cap;
  dim3 grid(gx, n);
  NPP_LAUNCH((confusion_kernel), grid, 256, (size_t)c * c * sizeof(unsigned int), as_stream(s), 
      logits, label, c, h, w, label_h, label_w, ignore, reinterpret_cast<unsigned long long*>(hist));
  NPP_CHECK_LAUNCH("confusion_kernel");
  return NPP_OK;
}

int npp_tta_merge(const float* pred, const float* flip_pred, int n, int c, int h, int w, int oh, int ow, int swap_lr,
                  float* out, npp_stream_t s) {
  if (!pred || !flip_pred || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0 || oh <= 0 || ow <= 0) return NPP_E_INVALID;
  if (swap_lr && c < 20) return NPP_E_INVALID;
  const Axis ah = make_axis(h, oh, 0, 0.0), aw = make_axis(w, ow, 0, 0.0);
  const int64_t total = (int64_t)n * c * oh * ow;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  NPP_LAUNCH((tta_merge_kernel), (int)grid, 256, 0, as_stream(s), pred, flip_pred, n, c, h, w, oh, ow, ah, aw, swap_lr, out);
  NPP_CHECK_LAUNCH("tta_merge_kernel");
  return NPP_OK;
}

int npp_pose_merge(const float* pred, const float* flip_pred, int n, int nj, int h, int w, const int* flip_idx, int oh,
                   int ow, float* out, npp_stream_t s) {
  if (!pred || !flip_pred || !out || !flip_idx || n <= 0 || nj <= 0 || nj > 32 || h <= 0 || w <= 0 || oh <= 0 || ow <= 0)
    return NPP_E_INVALID;
  FlipIdx fi;
  for (int j = 0; j < 32; ++j) fi.idx[j] = 0;
  for (int j = 0; j < nj; ++j) {
    if (flip_idx[j] < 0 || flip_idx[j] >= nj) return NPP_E_INVALID;
    fi.idx[j] = flip_idx[j];
  }
  const int64_t total = (int64_t)n * nj * oh * ow;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  NPP_LAUNCH((pose_merge_kernel), (int)grid, 256, 0, as_stream(s), pred, flip_pred, n, nj, h, w, oh, ow, fi, out);
  NPP_CHECK_LAUNCH("pose_merge_kernel");
  return NPP_OK;
}

int npp_heatmap_flip_avg(const float* pred, const float* flip_pred, int n, int nj, int h, int w, const int* flip_idx,
                         float* out, npp_stream_t s) {
  if (!pred || !flip_pred || !out || !flip_idx || n <= 0 || nj <= 0 || nj > 32 || h <= 0 || w <= 0) return NPP_E_INVALID;
  FlipIdx fi;
  for (int j = 0; j < 32; ++j) fi.idx[j] = 0;
  for (int j = 0; j < nj; ++j) {
    if (flip_idx[j] < 0 || flip_idx[j] >= nj) return NPP_E_INVALID;
    fi.idx[j] = flip_idx[j];
  }
  const int64_t total = (int64_t)n * nj * h * w;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  NPP_LAUNCH((heatmap_flip_avg_kernel), (int)grid, 256, 0, as_stream(s), pred, flip_pred, nj, h * w, fi, total, out);
  NPP_CHECK_LAUNCH("heatmap_flip_avg_kernel");
  return NPP_OK;
}

int npp_gaussian_filter(const float* src, float* tmp, float* dst, int planes, int h, int w, double sigma,
                        double truncate, npp_stream_t s) {
  if (!src || !tmp || !dst || planes <= 0 || h <= 0 || w <= 0 || !(sigma > 0.0) || !(truncate > 0.0)) return NPP_E_INVALID;
  GaussTaps g;
  g.radius = (int)(truncate * sigma + 0.5);   // scipy: int(truncate * sd + 0.5)
  if (g.radius < 0 || 2 * g.radius + 1 > 64) return NPP_E_UNSUPPORTED;
  double sum = 0.0;
  for (int t = 0; t <= 2 * g.radius; ++t) {
    const double x = (double)(t - g.radius);
    g.k[t] = exp(-0.5 / (sigma * sigma) * x * x);
    sum += g.k[t];
  }
  for (int t = 0; t <= 2 * g.radius; ++t) g.k[t] /= sum;
  for (int t = 2 * g.radius + 1; t < 64; ++t) g.k[t] = 0.0;
  const int64_t total = (int64_t)planes * h * w;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  NPP_LAUNCH((gauss1d_reflect_kernel), (int)grid, 256, 0, as_stream(s), src, tmp, planes, h, w, 0, g);
  NPP_CHECK_LAUNCH("gauss1d_reflect_kernel(axis 0)");
  NPP_LAUNCH((gauss1d_reflect_kernel), (int)grid, 256, 0, as_stream(s), tmp, dst, planes, h, w, 1, g);
  NPP_CHECK_LAUNCH("gauss1d_reflect_kernel(axis 1)");
  return NPP_OK;
}

int npp_heatmap_argmax(const float* hm, int nj, int h, int w, int32_t* idx, float* maxval, npp_stream_t s) {
  if (!hm || !idx || !maxval || nj <= 0 || h <= 0 || w <= 0) return NPP_E_INVALID;
  NPP_LAUNCH((heatmap_argmax_kernel), nj, 256, 0, as_stream(s), hm, h * w, idx, maxval);
  NPP_CHECK_LAUNCH("heatmap_argmax_kernel");
  return NPP_OK;
}

int npp_pck_counts(const int32_t* pred_idx, const float* pred_max, const int32_t* gt_idx, const float* gt_max, int n,
                   int j, int h, int w, float thr, int64_t* hit, int64_t* valid, npp_stream_t s) {
  if (!pred_idx || !pred_max || !gt_idx || !gt_max || !hit || !valid || n <= 0 || j <= 0) return NPP_E_INVALID;
  NPP_LAUNCH((pck_counts_kernel), (n * j + 127) / 128, 128, 0, as_stream(s), 
      pred_idx, pred_max, gt_idx, gt_max, n, j, h, w, (double)thr, reinterpret_cast<unsigned long long*>(hit),
      reinterpret_cast<unsigned long long*>(valid));
  NPP_CHECK_LAUNCH("pck_counts_kernel");
  return NPP_OK;
}

int npp_pckh_counts(const double* pred, const double* gt, int n, int p, double thr, int64_t* hit, int64_t* valid,
                    npp_stream_t s) {
  if (!pred || !gt || !hit || !valid || n <= 0 || p < 10) return NPP_E_INVALID;
  NPP_LAUNCH((pckh_counts_kernel), (n * p + 127) / 128, 128, 0, as_stream(s), pred, gt, n, p, thr,
                                                                     reinterpret_cast<unsigned long long*>(hit),
                                                                     reinterpret_cast<unsigned long long*>(valid));
  NPP_CHECK_LAUNCH("pckh_counts_kernel");
  return NPP_OK;
}

}  // extern "C"
