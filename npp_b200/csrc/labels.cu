// On-GPU label synthesis: the targets the reference's data loader builds per image on the host
// (dataset/target_generation.py, called from dataset/data_loader.py:239-285) for a whole batch in one launch each.
//
//   pose_target   gen_pose_target / gen_single_gaussian_map (:94-168): per joint a clamped Gaussian window on the
//                 stride grid, exponent cut at 4.6052 (exp(-4.6052) = 1 %), background channel = 1 - max over joints;
//                 the aux maps use 2*sigma.  numpy computes in float64 and the loop casts to float32 afterwards
//                 (core/function.py:78-79): the same double arithmetic in the same order here, rounded to fp32 once.
//   edge_label    generate_edge (:210-239): label differences towards four neighbours (ignoring 255), 3x3 dilation
//                 (cv2.dilate, rectangular kernel, constant border), then edge[label == 255] = 255
//                 (data_loader.py:281-285).  Integer, exact.
//   flip_parsing  gen_parsing_target's flip branch (:44-56): horizontal mirror + left/right relabel 14<->15, 16<->17,
//                 18<->19.  Integer, exact.
#include "common.cuh"

namespace npp {

__global__ void __launch_bounds__(256)
pose_target_kernel(const double* __restrict__ joints, const int* __restrict__ vis, int B, int J, double stride, int gx,
                   int gy, double sigma, float* __restrict__ out) {
  pdl_wait();
  const int64_t total = (int64_t)B * gy * gx;
  const double start = stride / 2.0 - 0.5;
  const double max_dist = ceil(sqrt(4.6052 * sigma * sigma * 2.0));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g_x = (int)(i % gx);
    const int g_y = (int)((i / gx) % gy);
    const int b = (int)(i / ((int64_t)gx * gy));
    const double x = start + g_x * stride;
    const double y = start + g_y * stride;
    double mx = 0.0;
    for (int j = 0; j < J; ++j) {
      double v = 0.0;
      if (vis[b * J + j]) {
        const double cx = joints[((int64_t)b * J + j) * 2], cy = joints[((int64_t)b * J + j) * 2 + 1];
        const int sx = (int)fmax(0.0, floor((cx - max_dist - start) / stride));
        const int ex = (int)fmin((double)gx, ceil((cx + max_dist - start) / stride));
        const int sy = (int)fmax(0.0, floor((cy - max_dist - start) / stride));
        const int ey = (int)fmin((double)gy, ceil((cy + max_dist - start) / stride));
        if (g_x >= sx && g_x < ex && g_y >= sy && g_y < ey) {
          const double d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
          const double e = d2 / 2.0 / sigma / sigma;
          if (!(e > 4.6052)) {
            v = exp(-e);
            if (v > 1.0) v = 1.0;
          }
        }
      }
      out[(((int64_t)b * (J + 1) + j) * gy + g_y) * gx + g_x] = (float)v;
      mx = fmax(mx, v);
    }
    out[(((int64_t)b * (J + 1) + J) * gy + g_y) * gx + g_x] = (float)(1.0 - mx);
  }
}

__device__ __forceinline__ int edge_seed(const int64_t* __restrict__ lab, int h, int w, int y, int x) {
  // the four rules of generate_edge before dilation, for pixel (y, x)
  const int64_t c = lab[(int64_t)y * w + x];
  if (c == 255) return 0;
  if (y >= 1) {                                   // "right":       label[1:h] vs label[:h-1]
    const int64_t o = lab[(int64_t)(y - 1) * w + x];
    if (o != 255 && o != c) return 1;
  }
  if (x + 1 < w) {                                // "up":          label[:, :w-1] vs label[:, 1:w]
    const int64_t o = lab[(int64_t)y * w + x + 1];
    if (o != 255 && o != c) return 1;
  }
  if (y + 1 < h && x + 1 < w) {                   // "upright":     label[:h-1, :w-1] vs label[1:h, 1:w]
    const int64_t o = lab[(int64_t)(y + 1) * w + x + 1];
    if (o != 255 && o != c) return 1;
  }
  if (y + 1 < h && x >= 1) {                      // "bottomright": label[:h-1, 1:w] vs label[1:h, :w-1]
    const int64_t o = lab[(int64_t)(y + 1) * w + x - 1];
    if (o != 255 && o != c) return 1;
  }
  return 0;
}

__global__ void __launch_bounds__(256)
edge_label_kernel(const int64_t* __restrict__ label, int B, int h, int w, int radius, int64_t* __restrict__ out) {
  pdl_wait();
  const int64_t total = (int64_t)B * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int64_t* lab = label + (i / ((int64_t)h * w)) * h * w;
    int64_t v = 0;
    if (lab[(int64_t)y * w + x] == 255) {
      v = 255;
    } else {
      for (int dy = -radius; dy <= radius && !v; ++dy)
        for (int dx = -radius; dx <= radius && !v; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < h && xx >= 0 && xx < w && edge_seed(lab, h, w, yy, xx)) v = 1;
        }
    }
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
flip_parsing_kernel(const int64_t* __restrict__ label, int64_t total, int w, int64_t* __restrict__ out) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    int64_t v = label[i - x + (w - 1 - x)];
    if (v >= 14 && v <= 19) v ^= 1;               // 14<->15, 16<->17, 18<->19
    out[i] = v;
  }
}

static inline int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_pose_target(const double* joints, const int* visible, int b, int j, double stride, int grid_x, int grid_y,
                    double sigma, float* out, npp_stream_t s) {
  if (!joints || !visible || !out || b <= 0 || j <= 0 || grid_x <= 0 || grid_y <= 0 || !(stride > 0) || !(sigma > 0))
    return NPP_E_INVALID;
  NPP_LAUNCH((pose_target_kernel), grid_for((int64_t)b * grid_x * grid_y), 256, 0, as_stream(s), joints, visible, b, j, stride, grid_x,
                                                                                    grid_y, sigma, out);
  NPP_CHECK_LAUNCH("pose_target_kernel");
  return NPP_OK;
}

int npp_edge_label(const int64_t* label, int b, int h, int w, int edge_width, int64_t* out, npp_stream_t s) {
  if (!label || !out || b <= 0 || h <= 0 || w <= 0 || edge_width < 1 || edge_width % 2 == 0 || edge_width > 15)
    return NPP_E_INVALID;
  NPP_LAUNCH((edge_label_kernel), grid_for((int64_t)b * h * w), 256, 0, as_stream(s), label, b, h, w, edge_width / 2, out);
  NPP_CHECK_LAUNCH("edge_label_kernel");
  return NPP_OK;
}

int npp_flip_parsing(const int64_t* label, int b, int h, int w, int64_t* out, npp_stream_t s) {
  if (!label || !out || label == out || b <= 0 || h <= 0 || w <= 0) return NPP_E_INVALID;
  const int64_t total = (int64_t)b * h * w;
  NPP_LAUNCH((flip_parsing_kernel), grid_for(total), 256, 0, as_stream(s), label, total, w, out);
  NPP_CHECK_LAUNCH("flip_parsing_kernel");
  return NPP_OK;
}

}  // extern "C"
