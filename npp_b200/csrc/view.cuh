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

// ---- launch geometry shared by the two launchers ------------------------------------------------------
// A 256-thread block is laid out as (cvb channel-vectors) x (rows pixels): consecutive threads touch
// consecutive 16-byte channel vectors of one pixel (coalesced), the block then strides over pixels.
// Pixel index -> (n, h, w) is one 32-bit div/mod pair per pixel (64-bit divisions were the bottleneck of
// the first version of these kernels).
struct VecGeom {
  int cv;      // channel vectors per pixel
  int cvb;     // channel vectors handled by one block (<= 256)
  int rows;    // pixels handled per block iteration
  int gy;      // blocks along the channel axis
};
static inline VecGeom vec_geom(int C, int VEC) {
  VecGeom g;
  g.cv = C / VEC;
  g.cvb = g.cv < 256 ? g.cv : 256;
  g.rows = 256 / g.cvb;
  g.gy = (g.cv + g.cvb - 1) / g.cvb;
  return g;
}

// f(n, h, w, c0) for every pixel and every channel vector (c0 multiple of VEC).
template <int VEC, typename F>
__global__ void __launch_bounds__(256) foreach_vec_kernel(int npix, int H, int W, VecGeom g, F f) {
  pdl_wait();
  const int tcv = threadIdx.x % g.cvb;
  const int trow = threadIdx.x / g.cvb;
  const int mycv = blockIdx.y * g.cvb + tcv;
  if (trow >= g.rows || mycv >= g.cv) return;
  const int c0 = mycv * VEC;
  const int step = gridDim.x * g.rows;
  for (int p = blockIdx.x * g.rows + trow; p < npix; p += step) {
    const int w = p % W;
    const int t = p / W;
    const int h = t % H;
    const int n = t / H;
    f(n, h, w, c0);
  }
}

template <int VEC, typename F>
static inline int foreach_vec(int N, int H, int W, int C, cudaStream_t st, const char* name, F f) {
  const int64_t npix64 = (int64_t)N * H * W;
  if (npix64 <= 0 || C <= 0) return NPP_OK;
  if (npix64 > 0x7fffffff) return NPP_E_UNSUPPORTED;
  const VecGeom g = vec_geom(C, VEC);
  int64_t gx = (npix64 + g.rows - 1) / g.rows;
  const int64_t cap = ((int64_t)sm_count() * 16 + g.gy - 1) / g.gy;  // ~16 resident 256-thread blocks per SM
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)g.gy);
  NPP_LAUNCH((foreach_vec_kernel<VEC, F>), grid, 256, 0, st, (int)npix64, H, W, g, f);
  NPP_CHECK_LAUNCH(name);
  return NPP_OK;
}

// Column sums of a [rows][ncol] fp32 array in shared memory (ncol = channel vectors of the block x VEC), called by all
// 256 threads after the array was written and a __syncthreads(); on return red[0..ncol) holds the sums (and a
// __syncthreads() has been passed).  The first version let the `trow == 0` threads walk all rows serially: with few
// channels per pixel (C = 32: 4 threads x 8 channels x 64 rows) that tail cost more than the streaming loop.
__device__ __forceinline__ void block_colsum(float* red, int rows, int ncol) {
  if (ncol <= 128) {
    const int groups = 256 / ncol;  // >= 2 row groups, each thread folds rows rg, rg + groups, ...
    const int col = threadIdx.x % ncol, rg = threadIdx.x / ncol;
    float s = 0.f;
    if (rg < groups)
      for (int r = rg; r < rows; r += groups) s += red[r * ncol + col];
    __syncthreads();
    if (rg < groups) red[rg * ncol + col] = s;
    __syncthreads();
    if (threadIdx.x < ncol) {
      float t = 0.f;
      for (int g = 0; g < groups; ++g) t += red[g * ncol + threadIdx.x];
      red[threadIdx.x] = t;
    }
  } else {
    float t[8];  // ncol <= 256 * VEC <= 2048 columns
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = threadIdx.x + j * 256;
      t[j] = 0.f;
      if (col < ncol)
        for (int r = 0; r < rows; ++r) t[j] += red[r * ncol + col];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = threadIdx.x + j * 256;
      if (col < ncol) red[col] = t[j];
    }
  }
  __syncthreads();
}

// Per-channel reduction over pixels of an [n,h,w,c] space: f(n,h,w,c0, acc[K][VEC]) accumulates K
// quantities per channel; results are atomically added to out[k*out_kstride + (zoff + c)*out_cstride].
// grid = (pixel chunks, channel-vector groups, Z) where Z = N if per_image else 1.
template <int VEC, int K, typename F>
__global__ void __launch_bounds__(256) reduce_ch_kernel(int npix, int H, int W, int C, VecGeom g, int per_image,
                                                        float* out, int64_t out_kstride, int64_t out_cstride, F f) {
  pdl_wait();
  __shared__ float red[256 * VEC];
  const int tcv = threadIdx.x % g.cvb;
  const int trow = threadIdx.x / g.cvb;
  const int mycv = blockIdx.y * g.cvb + tcv;
  const bool active = trow < g.rows && mycv < g.cv;
  float acc[K][VEC];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
  const int n_base = per_image ? blockIdx.z : 0;
  if (active) {
    const int step = gridDim.x * g.rows;
#pragma unroll 4
    for (int p = blockIdx.x * g.rows + trow; p < npix; p += step) {
      const int w = p % W;
      const int t = p / W;
      const int h = t % H;
      const int n = n_base + t / H;
      f(n, h, w, mycv * VEC, acc);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[threadIdx.x * VEC + v] = acc[k][v];
    __syncthreads();
    block_colsum(red, g.rows, g.cvb * VEC);
    if (active && trow == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        atomicAdd(out + k * out_kstride + ((per_image ? (int64_t)blockIdx.z * C : 0) + mycv * VEC + v) * out_cstride,
                  red[tcv * VEC + v]);
    }
  }
}

template <int VEC, int K, typename F>
static inline int reduce_ch(int N, int H, int W, int C, bool per_image, float* out, int64_t out_kstride, cudaStream_t st,
                            const char* name, F f, int64_t out_cstride = 1) {
  const VecGeom g = vec_geom(C, VEC);
  const int64_t npix = (int64_t)(per_image ? 1 : N) * H * W;
  if (npix > 0x7fffffff) return NPP_E_UNSUPPORTED;
  const int z = per_image ? N : 1;
  int64_t gx = (npix + g.rows - 1) / g.rows;
  // Every block ends with one atomic per (channel, quantity) on the SAME out[] addresses (per image when per_image):
  // same-address atomics serialise in L2 (~30 clk each), so with 8 blocks per SM the tail of a whole-batch reduction
  // (1184 blocks -> 1184 serialised adds per address = 18 us) cost more than streaming the tensor.  Two blocks per SM
  // keep the tail under 5 us; the pixel loop is unrolled x4 so 512 threads still hold 32 KB of loads in flight per SM.
  const int bps = per_image ? 8 : 2;
  const int64_t cap = ((int64_t)sm_count() * bps + (int64_t)g.gy * z - 1) / ((int64_t)g.gy * z);
  if (gx > cap) gx = cap;
  const int64_t max_by_work = (npix + (int64_t)g.rows * 16 - 1) / ((int64_t)g.rows * 16);
  if (gx > max_by_work) gx = max_by_work;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)g.gy, (unsigned)z);
  NPP_LAUNCH((reduce_ch_kernel<VEC, K, F>), grid, 256, 0, st, (int)npix, H, W, C, g, per_image ? 1 : 0, out, out_kstride,
                                                    out_cstride, f);
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
