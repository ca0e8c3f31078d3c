// Device-side NHWC view + generic vectorised launchers for the bandwidth-bound kernels.
#pragma once
#include "common.cuh"

namespace npp {

template <typename T>
struct DView {
  T* p;
  int n, h, w, c;
  int64_t sn, sh, sw;
  __device__ __forceinline__ T* at(int in, int ih, int iw, int ic) const {
    return p + in * sn + ih * sh + iw * sw + ic;
  }
};

template <typename T>
static inline DView<T> dview(const npp_view4* v) {
  DView<T> d;
  d.p = static_cast<T*>(v->ptr);
  d.n = v->n; d.h = v->h; d.w = v->w; d.c = v->c;
  d.sn = v->sn; d.sh = v->sh; d.sw = v->sw;
  return d;
}

// One thread per (pixel, channel-vector) of an [n,h,w,c] index space, grid-stride.
// f(n, h, w, c0) is called with c0 a multiple of the vector width.
template <int VEC, typename F>
__global__ void __launch_bounds__(256) foreach_vec_kernel(int N, int H, int W, int C, F f) {
  const int cv = C / VEC;
  const int64_t total = (int64_t)N * H * W * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * VEC;
    int64_t p = i / cv;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    f(n, h, w, c0);
  }
}

template <int VEC, typename F>
static inline int foreach_vec(int N, int H, int W, int C, cudaStream_t st, const char* name, F f) {
  const int64_t total = (int64_t)N * H * W * (C / VEC);
  if (total <= 0) return NPP_OK;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;  // 16 x 256 threads resident per SM
  if (grid > cap) grid = cap;
  foreach_vec_kernel<VEC, F><<<(int)grid, 256, 0, st>>>(N, H, W, C, f);
  NPP_CHECK_LAUNCH(name);
  return NPP_OK;
}

// Per-channel reduction over pixels of an [n,h,w,c] space: f(n,h,w,c0, acc[K][VEC]) accumulates K
// quantities per channel; results are atomically added to out[k*out_kstride + (zoff + c)*out_cstride].
// grid = (pixel chunks, channel-vector groups, Z) where Z = N if per_image else 1.
template <int VEC, int K, typename F>
__global__ void __launch_bounds__(256) reduce_ch_kernel(int N, int H, int W, int C, int per_image, float* out,
                                                        int64_t out_kstride, int64_t out_cstride, F f) {
  __shared__ float red[256 * VEC];
  const int cv = C / VEC;
  const int cv_pb = cv < 256 ? cv : 256;        // channel vectors per block
  const int rows_pb = 256 / cv_pb;              // pixel rows per block iteration
  const int tcv = threadIdx.x % cv_pb;
  const int trow = threadIdx.x / cv_pb;
  const int mycv = blockIdx.y * cv_pb + tcv;
  const bool active = trow < rows_pb && mycv < cv;
  float acc[K][VEC];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
  const int nimg = per_image ? 1 : N;
  const int n_base = per_image ? blockIdx.z : 0;
  const int64_t npix = (int64_t)nimg * H * W;
  if (active) {
    for (int64_t p = (int64_t)blockIdx.x * rows_pb + trow; p < npix; p += (int64_t)gridDim.x * rows_pb) {
      const int w = (int)(p % W);
      const int h = (int)((p / W) % H);
      const int n = n_base + (int)(p / ((int64_t)W * H));
      f(n, h, w, mycv * VEC, acc);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[threadIdx.x * VEC + v] = acc[k][v];
    __syncthreads();
    if (active && trow == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float s = 0.f;
        for (int r = 0; r < rows_pb; ++r) s += red[(r * cv_pb + tcv) * VEC + v];
        atomicAdd(out + k * out_kstride + ((per_image ? (int64_t)blockIdx.z * C : 0) + mycv * VEC + v) * out_cstride, s);
      }
    }
  }
}

template <int VEC, int K, typename F>
static inline int reduce_ch(int N, int H, int W, int C, bool per_image, float* out, int64_t out_kstride, cudaStream_t st,
                            const char* name, F f, int64_t out_cstride = 1) {
  const int cv = C / VEC;
  const int cv_pb = cv < 256 ? cv : 256;
  const int rows_pb = 256 / cv_pb;
  const int gy = (cv + cv_pb - 1) / cv_pb;
  const int64_t npix = (int64_t)(per_image ? 1 : N) * H * W;
  int64_t gx = (npix + rows_pb - 1) / rows_pb;
  // enough blocks to fill the machine a few times, but keep >= ~8 pixels per thread to amortise the atomics
  const int64_t cap = (int64_t)sm_count() * 8 / (gy * (per_image ? N : 1)) + 1;
  if (gx > cap) gx = cap;
  const int64_t min_rows = 8;
  if (gx > (npix + rows_pb * min_rows - 1) / (rows_pb * min_rows)) gx = (npix + rows_pb * min_rows - 1) / (rows_pb * min_rows);
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)(per_image ? N : 1));
  reduce_ch_kernel<VEC, K, F><<<grid, 256, 0, st>>>(N, H, W, C, per_image ? 1 : 0, out, out_kstride, out_cstride, f);
  NPP_CHECK_LAUNCH(name);
  return NPP_OK;
}

#define NPP_DISPATCH_DTYPE(dtype, ...)                       \
  do {                                                       \
    if ((dtype) == NPP_F32) {                                \
      using T = float;                                       \
      __VA_ARGS__                                            \
    } else if ((dtype) == NPP_BF16) {                        \
      using T = __nv_bfloat16;                               \
      __VA_ARGS__                                            \
    } else {                                                 \
      return NPP_E_UNSUPPORTED;                              \
    }                                                        \
  } while (0)

}  // namespace npp
