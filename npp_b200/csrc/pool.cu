// Pooling over NHWC views (HBM-bound, vectorised along channels).
//  maxpool3x3 : nn.MaxPool2d(3, stride, 1)                        models/operations.py:55
//  avgpool3x3 : nn.AvgPool2d(3, stride, 1, count_include_pad=False)  :57
//  avgpool2x2 : nn.AvgPool2d(2)                                     :115 (SE_Block.pool2), :237 (Pooled_Conv)
//  gap        : nn.AdaptiveAvgPool2d(1)                             :111
// Backward passes are written in gather form (each dx element pulls from the outputs whose
// window contains it) so they need no atomics; max-pool keeps a 1-byte winner index per output.
#include "view.cuh"
#include <math_constants.h>

namespace npp {

// Forward optionally records, per output element, which of the 9 window positions won (first maximum in
// row-major scan order with strict '>': the element ATen's max_pool2d backward routes the gradient to).
template <typename T>
static int maxpool_fwd_t(const npp_view4* x, const npp_view4* y, uint8_t* idx, int stride, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  const int H = x->h, W = x->w, Ho = y->h, Wo = y->w, C = y->c;
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "maxpool3x3_fwd", [=] __device__(int n, int ho, int wo, int c) {
    float m[V];
    uint8_t arg[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { m[i] = -CUDART_INF_F; arg[i] = 0; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hi = ho * stride - 1 + r;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int wi = wo * stride - 1 + s;
        if (wi < 0 || wi >= W) continue;
        float v[V];
        Pack<T>::load(X.at(n, hi, wi, c), v);
#pragma unroll
        for (int i = 0; i < V; ++i)
          if (v[i] > m[i]) { m[i] = v[i]; arg[i] = (uint8_t)(r * 3 + s); }
      }
    }
    Pack<T>::store(Y.at(n, ho, wo, c), m);
    if (idx != nullptr) {
      uint8_t* ip = idx + (((int64_t)n * Ho + ho) * Wo + wo) * C + c;
      if (V == 8) {
        uint2 pk;
        pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | ((uint32_t)arg[3] << 24);
        pk.y = arg[4 % V] | (arg[5 % V] << 8) | (arg[6 % V] << 16) | ((uint32_t)arg[7 % V] << 24);
        *reinterpret_cast<uint2*>(ip) = pk;
      } else {
        *reinterpret_cast<uint32_t*>(ip) = arg[0] | (arg[1] << 8) | (arg[2] << 16) | ((uint32_t)arg[3] << 24);
      }
    }
  });
}

// dx[h,w] = sum over output windows (ho,wo) containing (h,w) whose recorded winner is (h,w) of dy[ho,wo]
// (gather form: no atomics; 9 (stride 1) or 4 (stride 2) candidate windows per element).
template <typename T>
static int maxpool_bwd_t(const uint8_t* idx, const npp_view4* dy, const npp_view4* dx, int stride, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  const int Ho = dy->h, Wo = dy->w, C = dy->c;
  return foreach_vec<V>(dx->n, dx->h, dx->w, dx->c, st, "maxpool3x3_bwd", [=] __device__(int n, int h, int w, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    for (int ho = (h - 1 + stride - 1) / stride; ho * stride - 1 <= h; ++ho) {
      if (ho < 0) continue;
      if (ho >= Ho) break;
      const int r = h - (ho * stride - 1);
      for (int wo = (w - 1 + stride - 1) / stride; wo * stride - 1 <= w; ++wo) {
        if (wo < 0) continue;
        if (wo >= Wo) break;
        const uint32_t me = (uint32_t)(r * 3 + (w - (wo * stride - 1)));
        const uint8_t* ip = idx + (((int64_t)n * Ho + ho) * Wo + wo) * C + c;
        uint32_t a[2];
        if (V == 8) {
          const uint2 pk = *reinterpret_cast<const uint2*>(ip);
          a[0] = pk.x; a[1] = pk.y;
        } else {
          a[0] = *reinterpret_cast<const uint32_t*>(ip); a[1] = 0;
        }
        float d[V];
        Pack<T>::load(DY.at(n, ho, wo, c), d);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += (((a[i >> 2] >> ((i & 3) * 8)) & 0xffu) == me) ? d[i] : 0.f;
      }
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

__device__ __forceinline__ int win_count3(int o, int stride, int L) {
  int cnt = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int i = o * stride - 1 + r;
    cnt += (i >= 0 && i < L) ? 1 : 0;
  }
  return cnt;
}

template <typename T>
static int avgpool3_fwd_t(const npp_view4* x, const npp_view4* y, int stride, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  const int H = x->h, W = x->w;
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "avgpool3x3_fwd", [=] __device__(int n, int ho, int wo, int c) {
    float a[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hi = ho * stride - 1 + r;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int wi = wo * stride - 1 + s;
        if (wi < 0 || wi >= W) continue;
        float v[V];
        Pack<T>::load(X.at(n, hi, wi, c), v);
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] += v[i];
      }
    }
    const float inv = 1.f / (float)(win_count3(ho, stride, H) * win_count3(wo, stride, W));
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] *= inv;
    Pack<T>::store(Y.at(n, ho, wo, c), a);
  });
}

template <typename T>
static int avgpool3_bwd_t(const npp_view4* dy, const npp_view4* dx, int stride, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  const int H = dx->h, W = dx->w, Ho = dy->h, Wo = dy->w;
  return foreach_vec<V>(dx->n, dx->h, dx->w, dx->c, st, "avgpool3x3_bwd", [=] __device__(int n, int h, int w, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    for (int ho = (h - 1 + stride - 1) / stride; ho * stride - 1 <= h; ++ho) {
      if (ho < 0) continue;
      if (ho >= Ho) break;
      for (int wo = (w - 1 + stride - 1) / stride; wo * stride - 1 <= w; ++wo) {
        if (wo < 0) continue;
        if (wo >= Wo) break;
        const float inv = 1.f / (float)(win_count3(ho, stride, H) * win_count3(wo, stride, W));
        float d[V];
        Pack<T>::load(DY.at(n, ho, wo, c), d);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += d[i] * inv;
      }
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

template <typename T>
static int avgpool2_fwd_t(const npp_view4* x, const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "avgpool2x2_fwd", [=] __device__(int n, int ho, int wo, int c) {
    float a[V], v[V];
    Pack<T>::load(X.at(n, 2 * ho, 2 * wo, c), a);
    Pack<T>::load(X.at(n, 2 * ho, 2 * wo + 1, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] += v[i];
    Pack<T>::load(X.at(n, 2 * ho + 1, 2 * wo, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] += v[i];
    Pack<T>::load(X.at(n, 2 * ho + 1, 2 * wo + 1, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] = (a[i] + v[i]) * 0.25f;
    Pack<T>::store(Y.at(n, ho, wo, c), a);
  });
}

template <typename T>
static int avgpool2_bwd_t(const npp_view4* dy, const npp_view4* dx, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  const int Ho = dy->h, Wo = dy->w;
  return foreach_vec<V>(dx->n, dx->h, dx->w, dx->c, st, "avgpool2x2_bwd", [=] __device__(int n, int h, int w, int c) {
    float g[V];
    const int ho = h >> 1, wo = w >> 1;
    if (ho < Ho && wo < Wo) {
      Pack<T>::load(DY.at(n, ho, wo, c), g);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] *= 0.25f;
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = 0.f;
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

}  // namespace npp

using namespace npp;

static bool pool3_shapes_ok(const npp_view4* x, const npp_view4* y, int stride) {
  if (stride != 1 && stride != 2) return false;
  return x->n == y->n && x->c == y->c && y->h == (x->h + 2 - 3) / stride + 1 && y->w == (x->w + 2 - 3) / stride + 1;
}

extern "C" {

int npp_maxpool3x3_fwd(const npp_view4* x, const npp_view4* y, uint8_t* argmax, int stride, int dtype,
                       npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !pool3_shapes_ok(x, y, stride)) return NPP_E_INVALID;
  if (argmax && reinterpret_cast<uintptr_t>(argmax) % 8) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return maxpool_fwd_t<T>(x, y, argmax, stride, as_stream(s)););
}
int npp_maxpool3x3_bwd(const uint8_t* argmax, const npp_view4* dy, const npp_view4* dx, int stride, int dtype,
                       npp_stream_t s) {
  if (!argmax || !view_ok(dy, dtype) || !view_ok(dx, dtype) || !pool3_shapes_ok(dx, dy, stride)) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return maxpool_bwd_t<T>(argmax, dy, dx, stride, as_stream(s)););
}
int npp_avgpool3x3_fwd(const npp_view4* x, const npp_view4* y, int stride, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !pool3_shapes_ok(x, y, stride)) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return avgpool3_fwd_t<T>(x, y, stride, as_stream(s)););
}
int npp_avgpool3x3_bwd(const npp_view4* dy, const npp_view4* dx, int stride, int dtype, npp_stream_t s) {
  if (!view_ok(dy, dtype) || !view_ok(dx, dtype) || !pool3_shapes_ok(dx, dy, stride)) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return avgpool3_bwd_t<T>(dy, dx, stride, as_stream(s)););
}
int npp_avgpool2x2_fwd(const npp_view4* x, const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || x->n != y->n || x->c != y->c || y->h != x->h / 2 || y->w != x->w / 2)
    return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return avgpool2_fwd_t<T>(x, y, as_stream(s)););
}
int npp_avgpool2x2_bwd(const npp_view4* dy, const npp_view4* dx, int dtype, npp_stream_t s) {
  if (!view_ok(dy, dtype) || !view_ok(dx, dtype) || dx->n != dy->n || dx->c != dy->c || dy->h != dx->h / 2 ||
      dy->w != dx->w / 2)
    return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return avgpool2_bwd_t<T>(dy, dx, as_stream(s)););
}

}  // extern "C"
