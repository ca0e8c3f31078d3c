// Elementwise / data-movement kernels over NHWC views: 16-byte vector access, coalesced along the
// channel axis, grid sized to the SM count (view.cuh).  All HBM-bound.
#include "view.cuh"

namespace npp {

template <typename T>
static int relu_fwd_t(const npp_view4* x, const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  return foreach_vec<V>(x->n, x->h, x->w, x->c, st, "relu_fwd", [=] __device__(int n, int h, int w, int c) {
    float v[V];
    Pack<T>::load(X.at(n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], 0.f);
    Pack<T>::store(Y.at(n, h, w, c), v);
  });
}

template <typename T>
static int relu_bwd_t(const npp_view4* x, const npp_view4* dy, const npp_view4* dx, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  return foreach_vec<V>(x->n, x->h, x->w, x->c, st, "relu_bwd", [=] __device__(int n, int h, int w, int c) {
    float a[V], g[V];
    Pack<T>::load(X.at(n, h, w, c), a);
    Pack<T>::load(DY.at(n, h, w, c), g);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = a[i] > 0.f ? g[i] : 0.f;
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

template <typename T>
static int add_t(const npp_view4* a, const npp_view4* b, const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto A = dview<const T>(a);
  const auto Y = dview<T>(y);
  if (b == nullptr) {
    return foreach_vec<V>(a->n, a->h, a->w, a->c, st, "copy", [=] __device__(int n, int h, int w, int c) {
      *reinterpret_cast<uint4*>(Y.at(n, h, w, c)) = *reinterpret_cast<const uint4*>(A.at(n, h, w, c));
    });
  }
  const auto B = dview<const T>(b);
  return foreach_vec<V>(a->n, a->h, a->w, a->c, st, "add", [=] __device__(int n, int h, int w, int c) {
    float u[V], v[V];
    Pack<T>::load(A.at(n, h, w, c), u);
    Pack<T>::load(B.at(n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) u[i] += v[i];
    Pack<T>::store(Y.at(n, h, w, c), u);
  });
}

template <typename T>
static int axpby_t(const npp_view4* a, const float* alpha, const npp_view4* b, const float* beta, const npp_view4* y,
                   cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto A = dview<const T>(a);
  const auto Y = dview<T>(y);
  if (b == nullptr) {
    return foreach_vec<V>(a->n, a->h, a->w, a->c, st, "scale", [=] __device__(int n, int h, int w, int c) {
      const float al = *alpha;
      float u[V];
      Pack<T>::load(A.at(n, h, w, c), u);
#pragma unroll
      for (int i = 0; i < V; ++i) u[i] *= al;
      Pack<T>::store(Y.at(n, h, w, c), u);
    });
  }
  const auto B = dview<const T>(b);
  return foreach_vec<V>(a->n, a->h, a->w, a->c, st, "axpby", [=] __device__(int n, int h, int w, int c) {
    const float al = *alpha, be = *beta;
    float u[V], v[V];
    Pack<T>::load(A.at(n, h, w, c), u);
    Pack<T>::load(B.at(n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) u[i] = al * u[i] + be * v[i];
    Pack<T>::store(Y.at(n, h, w, c), u);
  });
}

template <typename T>
static int fill_zero_t(const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto Y = dview<T>(y);
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "fill_zero", [=] __device__(int n, int h, int w, int c) {
    *reinterpret_cast<uint4*>(Y.at(n, h, w, c)) = make_uint4(0, 0, 0, 0);
  });
}

namespace tc {
int fill_zero_view_bf16(const npp_view4* v, cudaStream_t st) { return fill_zero_t<__nv_bfloat16>(v, st); }
}  // namespace tc

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = from_f<D>(to_f<S>(s[i]));
}

// NCHW fp32 -> NHWC T through a 32x32 shared-memory transpose of (C, HW) per image; channels >= C
// of the destination view are written as zero (channel padding).
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, DView<T> dst, int C) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int hw = dst.h * dst.w;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < hw) ? src[((int64_t)n * C + c) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < hw && c < dst.c) *dst.at(n, p / dst.w, p % dst.w, c) = from_f<T>(tile[threadIdx.x][j]);
  }
}

// Images (C <= 8 -> one 16-byte bf16 vector per pixel): a thread per pixel reads its C channel planes (coalesced
// across the warp) and writes one uint4.  The generic 32 x 32 transpose above wastes 29 of 32 tile rows on a
// 3-channel image and writes 2-byte elements (ncu: 154 GB/s on the 32 x 3 x 384 x 384 input batch).
template <int DUMMY>
__global__ void __launch_bounds__(256) nchw_to_nhwc8_kernel(const float* __restrict__ src, DView<__nv_bfloat16> dst,
                                                            int C) {
  pdl_wait();
  const int hw = dst.h * dst.w;
  const int n = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < C ? src[((int64_t)n * C + c) * hw + p] : 0.f;
    Pack<__nv_bfloat16>::store(dst.at(n, p / dst.w, p % dst.w, 0), v);
  }
}

// 3x3 im2col of a 3-channel image (internal NHWC bf16, channels padded to 8) for the network stems
// (models/model_augment.py:244-272: Conv2d(3, C, 3, stride 2, padding 1)): y[n, ho, wo, ci*9 + r*3 + s] =
// x[n, ho*stride - pad + r, wo*stride - pad + s, ci] (zero outside), channels 27..31 zero.  With the taps folded into
// the channel axis the stem is a 1x1 convolution with K = 32 whose weight matrix [Cout, 27] IS the OIHW master
// weight — the implicit-GEMM kernel otherwise spends nine 64-wide K blocks on 3 real channels each (7.6 TFLOP/s).
__global__ void __launch_bounds__(256) im2col3x3_c3_kernel(DView<const __nv_bfloat16> X, DView<__nv_bfloat16> Y,
                                                           int stride, int pad) {
  pdl_wait();
  const int npix = Y.n * Y.h * Y.w;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int wo = p % Y.w;
    const int t = p / Y.w;
    const int ho = t % Y.h;
    const int n = t / Y.h;
    uint2 in[9];  // channels 0..3 of the nine taps (channel 3 is padding)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int h = ho * stride - pad + r, w = wo * stride - pad + q;
        in[r * 3 + q] = (h >= 0 && h < X.h && w >= 0 && w < X.w) ? *reinterpret_cast<const uint2*>(X.at(n, h, w, 0))
                                                                   : make_uint2(0u, 0u);
      }
    // 16-bit element k = ci*9 + tap of the output row
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t pair = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * j + e;
        if (k < 27) {
          const int ci = k / 9, tap = k % 9;
          const uint32_t word = ci < 2 ? in[tap].x : in[tap].y;
          const uint32_t el = (ci & 1) ? (word >> 16) : (word & 0xffffu);
          pair |= el << (16 * e);
        }
      }
      o[j] = pair;
    }
    uint4* dst = reinterpret_cast<uint4*>(Y.at(n, ho, wo, 0));
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(DView<const T> src, float* __restrict__ dst, int C) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int hw = src.h * src.w;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < hw && c < C) ? to_f<T>(*src.at(n, p / src.w, p % src.w, c)) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < hw) dst[((int64_t)n * C + c) * hw + p] = tile[threadIdx.x][j];
  }
}

template <typename T>
static int interleave2_fwd_t(const npp_view4* a, const npp_view4* b, const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto A = dview<const T>(a);
  const auto B = dview<const T>(b);
  const auto Y = dview<T>(y);
  return foreach_vec<V>(a->n, a->h, a->w, a->c, st, "interleave2_fwd", [=] __device__(int n, int h, int w, int c) {
    float u[V], v[V], o0[V], o1[V];
    Pack<T>::load(A.at(n, h, w, c), u);
    Pack<T>::load(B.at(n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V / 2; ++i) {
      o0[2 * i] = u[i]; o0[2 * i + 1] = v[i];
      o1[2 * i] = u[V / 2 + i]; o1[2 * i + 1] = v[V / 2 + i];
    }
    Pack<T>::store(Y.at(n, h, w, 2 * c), o0);
    Pack<T>::store(Y.at(n, h, w, 2 * c + V), o1);
  });
}

template <typename T>
static int interleave2_bwd_t(const npp_view4* dy, const npp_view4* da, const npp_view4* db, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DYv = dview<const T>(dy);
  const auto DA = dview<T>(da);
  const auto DB = dview<T>(db);
  return foreach_vec<V>(da->n, da->h, da->w, da->c, st, "interleave2_bwd", [=] __device__(int n, int h, int w, int c) {
    float u[V], v[V], o0[V], o1[V];
    Pack<T>::load(DYv.at(n, h, w, 2 * c), o0);
    Pack<T>::load(DYv.at(n, h, w, 2 * c + V), o1);
#pragma unroll
    for (int i = 0; i < V / 2; ++i) {
      u[i] = o0[2 * i]; v[i] = o0[2 * i + 1];
      u[V / 2 + i] = o1[2 * i]; v[V / 2 + i] = o1[2 * i + 1];
    }
    Pack<T>::store(DA.at(n, h, w, c), u);
    Pack<T>::store(DB.at(n, h, w, c), v);
  });
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_relu_fwd(const npp_view4* x, const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !same_shape(x, y)) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return relu_fwd_t<T>(x, y, as_stream(s)););
}
int npp_relu_bwd(const npp_view4* x, const npp_view4* dy, const npp_view4* dx, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(dy, dtype) || !view_ok(dx, dtype) || !same_shape(x, dy) || !same_shape(x, dx))
    return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return relu_bwd_t<T>(x, dy, dx, as_stream(s)););
}
int npp_add(const npp_view4* a, const npp_view4* b, const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(a, dtype) || !view_ok(y, dtype) || !same_shape(a, y)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(a, b))) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return add_t<T>(a, b, y, as_stream(s)););
}
int npp_axpby(const npp_view4* a, const float* alpha, const npp_view4* b, const float* beta, const npp_view4* y,
              int dtype, npp_stream_t s) {
  if (!view_ok(a, dtype) || !view_ok(y, dtype) || !same_shape(a, y) || !alpha) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(a, b) || !beta)) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return axpby_t<T>(a, alpha, b, beta, y, as_stream(s)););
}
int npp_fill_zero(const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(y, dtype)) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return fill_zero_t<T>(y, as_stream(s)););
}
int npp_cast(const void* src, int sd, void* dst, int dd, int64_t n, npp_stream_t s) {
  if (!src || !dst || n < 0) return NPP_E_INVALID;
  if (n == 0) return NPP_OK;
  int64_t grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  cudaStream_t st = as_stream(s);
  if (sd == NPP_F32 && dd == NPP_BF16)
    NPP_LAUNCH((cast_kernel<float, __nv_bfloat16>), (int)grid, 256, 0, st, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == NPP_BF16 && dd == NPP_F32)
    NPP_LAUNCH((cast_kernel<__nv_bfloat16, float>), (int)grid, 256, 0, st, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == NPP_F32 && dd == NPP_F32)
    NPP_LAUNCH((cast_kernel<float, float>), (int)grid, 256, 0, st, (const float*)src, (float*)dst, n);
  else if (sd == NPP_BF16 && dd == NPP_BF16)
    NPP_LAUNCH((cast_kernel<__nv_bfloat16, __nv_bfloat16>), (int)grid, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else
    return NPP_E_UNSUPPORTED;
  NPP_CHECK_LAUNCH("cast_kernel");
  return NPP_OK;
}
int npp_nchw_to_nhwc(const float* src, int src_c, const npp_view4* dst, int dtype, npp_stream_t s) {
  if (!src || !view_ok(dst, dtype) || src_c <= 0 || src_c > dst->c) return NPP_E_INVALID;
  if (dtype == NPP_BF16 && dst->c == 8 && dst->n <= 65535) {
    const int hw = dst->h * dst->w;
    int gx = (hw + 255) / 256;
    if (gx > 2048) gx = 2048;
    NPP_LAUNCH((nchw_to_nhwc8_kernel<0>), dim3(gx, dst->n), 256, 0, as_stream(s), src, dview<__nv_bfloat16>(dst), src_c);
    NPP_CHECK_LAUNCH("nchw_to_nhwc8_kernel");
    return NPP_OK;
  }
  dim3 grid((dst->h * dst->w + 31) / 32, (dst->c + 31) / 32, dst->n), block(32, 8);
  NPP_DISPATCH_DTYPE(dtype, NPP_LAUNCH((nchw_to_nhwc_kernel<T>), grid, block, 0, as_stream(s), src, dview<T>(dst), src_c););
  NPP_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return NPP_OK;
}
int npp_im2col3x3_c3(const npp_view4* x, const npp_view4* y, int stride, int pad, npp_stream_t s) {
  if (!view_ok(x, NPP_BF16) || !view_ok(y, NPP_BF16) || x->c != 8 || y->c != 32 || x->n != y->n || stride < 1 || pad < 0)
    return NPP_E_INVALID;
  if (y->h != (x->h + 2 * pad - 3) / stride + 1 || y->w != (x->w + 2 * pad - 3) / stride + 1) return NPP_E_INVALID;
  const int64_t npix = (int64_t)y->n * y->h * y->w;
  if (npix > 0x7fffffff) return NPP_E_UNSUPPORTED;
  int64_t grid = (npix + 255) / 256;
  if (grid > (int64_t)sm_count() * 32) grid = (int64_t)sm_count() * 32;
  NPP_LAUNCH((im2col3x3_c3_kernel), (unsigned)grid, 256, 0, as_stream(s), dview<const __nv_bfloat16>(x), dview<__nv_bfloat16>(y),
                                                               stride, pad);
  NPP_CHECK_LAUNCH("im2col3x3_c3_kernel");
  return NPP_OK;
}
int npp_nhwc_to_nchw(const npp_view4* src, float* dst, int dst_c, int dtype, npp_stream_t s) {
  if (!dst || !view_ok(src, dtype) || dst_c <= 0 || dst_c > src->c) return NPP_E_INVALID;
  dim3 grid((src->h * src->w + 31) / 32, (dst_c + 31) / 32, src->n), block(32, 8);
  NPP_DISPATCH_DTYPE(dtype, NPP_LAUNCH((nhwc_to_nchw_kernel<T>), grid, block, 0, as_stream(s), dview<const T>(src), dst, dst_c););
  NPP_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return NPP_OK;
}
int npp_interleave2_fwd(const npp_view4* a, const npp_view4* b, const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(a, dtype) || !view_ok(b, dtype) || !view_ok(y, dtype) || !same_shape(a, b) || y->c != 2 * a->c ||
      y->n != a->n || y->h != a->h || y->w != a->w)
    return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return interleave2_fwd_t<T>(a, b, y, as_stream(s)););
}
int npp_interleave2_bwd(const npp_view4* dy, const npp_view4* da, const npp_view4* db, int dtype, npp_stream_t s) {
  if (!view_ok(da, dtype) || !view_ok(db, dtype) || !view_ok(dy, dtype) || !same_shape(da, db) || dy->c != 2 * da->c ||
      dy->n != da->n || dy->h != da->h || dy->w != da->w)
    return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return interleave2_bwd_t<T>(dy, da, db, as_stream(s)););
}

}  // extern "C"
