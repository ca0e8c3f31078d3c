// Depthwise (groups == C) dilated k x k convolution over NHWC views — DilConvS.net[1]
// (models/operations.py:213, nn.Conv2d(C, C, k, stride, padding, dilation, groups=C, bias=False))
// with the leading nn.ReLU (:212) fused into the loads.  ~2 flop/byte: pure HBM; each thread owns a
// 16-byte channel vector of one output pixel, neighbouring pixels' taps are served by L1/L2.
// Weights are read in the torch layout [C, k*k] (fp32 master weight viewed flat).
#include "view.cuh"

namespace npp {

template <typename T, int K>
static int dw_fwd_t(const npp_view4* x, const float* w, const npp_view4* y, int stride, int pad, int dil, int relu_in,
                    cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  const int H = x->h, W = x->w;
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "dwconv_fwd", [=] __device__(int n, int ho, int wo, int c) {
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int hi = ho * stride - pad + r * dil;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int wi = wo * stride - pad + s * dil;
        if (wi < 0 || wi >= W) continue;
        float v[V];
        Pack<T>::load(X.at(n, hi, wi, c), v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xv = relu_in ? fmaxf(v[i], 0.f) : v[i];
          acc[i] = fmaf(xv, __ldg(w + (c + i) * (K * K) + r * K + s), acc[i]);
        }
      }
    }
    Pack<T>::store(Y.at(n, ho, wo, c), acc);
  });
}

template <typename T, int K>
static int dw_bwd_data_t(const npp_view4* x, const float* w, const npp_view4* dy, const npp_view4* dx, int stride,
                         int pad, int dil, int relu_in, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  const int Ho = dy->h, Wo = dy->w;
  return foreach_vec<V>(x->n, x->h, x->w, x->c, st, "dwconv_bwd_data", [=] __device__(int n, int h, int wi, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      int ho = h + pad - r * dil;
      if (ho < 0 || (ho % stride)) continue;
      ho /= stride;
      if (ho >= Ho) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        int wo = wi + pad - s * dil;
        if (wo < 0 || (wo % stride)) continue;
        wo /= stride;
        if (wo >= Wo) continue;
        float d[V];
        Pack<T>::load(DY.at(n, ho, wo, c), d);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = fmaf(d[i], __ldg(w + (c + i) * (K * K) + r * K + s), g[i]);
      }
    }
    if (relu_in) {
      float v[V];
      Pack<T>::load(X.at(n, h, wi, c), v);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = v[i] > 0.f ? g[i] : 0.f;
    }
    Pack<T>::store(DX.at(n, h, wi, c), g);
  });
}

// dw[c, r*K+s] += sum_pixels dy[ho,wo,c] * relu(x)[ho*stride-pad+r*dil, wo*stride-pad+s*dil, c].
// ROWS kernel rows per launch: all 9 taps of a 3x3 in one pass (72 accumulators), one row at a time for 5x5.
template <typename T, int K, int ROWS>
static int dw_bwd_weight_rows(const npp_view4* x, const npp_view4* dy, float* dw, int r0, int stride, int pad, int dil,
                              int relu_in, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto DY = dview<const T>(dy);
  const int H = x->h, W = x->w;
  return reduce_ch<V, ROWS * K>(
      dy->n, dy->h, dy->w, dy->c, false, dw + r0 * K, 1, st, "dwconv_bwd_weight",
      [=] __device__(int n, int ho, int wo, int c, float (&acc)[ROWS * K][V]) {
        float d[V];
        Pack<T>::load(DY.at(n, ho, wo, c), d);
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) {
          const int hi = ho * stride - pad + (r0 + rr) * dil;
          if (hi < 0 || hi >= H) continue;
#pragma unroll
          for (int s = 0; s < K; ++s) {
            const int wi = wo * stride - pad + s * dil;
            if (wi < 0 || wi >= W) continue;
            float v[V];
            Pack<T>::load(X.at(n, hi, wi, c), v);
#pragma unroll
            for (int i = 0; i < V; ++i)
              acc[rr * K + s][i] = fmaf(d[i], relu_in ? fmaxf(v[i], 0.f) : v[i], acc[rr * K + s][i]);
          }
        }
      },
      K * K);
}

template <typename T, int K>
static int dw_bwd_weight_t(const npp_view4* x, const npp_view4* dy, float* dw, int stride, int pad, int dil,
                           int relu_in, cudaStream_t st) {
  if (K == 3) return dw_bwd_weight_rows<T, K, (K == 3 ? 3 : 1)>(x, dy, dw, 0, stride, pad, dil, relu_in, st);
  for (int r = 0; r < K; ++r) {
    int rc = dw_bwd_weight_rows<T, K, 1>(x, dy, dw, r, stride, pad, dil, relu_in, st);
    if (rc) return rc;
  }
  return NPP_OK;
}

// dwconv_tile.cu: shared-memory halo-staged 3x3 kernels; NPP_E_UNSUPPORTED = use the gather kernels above
template <typename T>
int dw_tile_fwd(const npp_view4* x, const float* w, const npp_view4* y, int stride, int pad, int dil, int relu_in,
                cudaStream_t st);
template <typename T>
int dw_tile_dgrad(const npp_view4* x, const float* w, const npp_view4* dy, const npp_view4* dx, int stride, int pad,
                  int dil, int relu_in, cudaStream_t st);
template <typename T>
int dw_tile_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int stride, int pad, int dil, int relu_in,
                  cudaStream_t st);

}  // namespace npp

using namespace npp;

static bool dw_shapes_ok(const npp_view4* x, const npp_view4* y, int k, int stride, int pad, int dil) {
  if ((k != 3 && k != 5) || (stride != 1 && stride != 2) || pad < 0 || dil < 1) return false;
  const int ho = (x->h + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  const int wo = (x->w + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  return x->n == y->n && x->c == y->c && y->h == ho && y->w == wo;
}

extern "C" {

int npp_dwconv_fwd(const npp_view4* x, const float* w, const npp_view4* y, int k, int stride, int pad, int dil,
                   int relu_in, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !w) return NPP_E_INVALID;
  if (!dw_shapes_ok(x, y, k, stride, pad, dil)) return NPP_E_UNSUPPORTED;
  NPP_DISPATCH_DTYPE(
      dtype,
      if (k == 3) {
        const int rc = dw_tile_fwd<T>(x, w, y, stride, pad, dil, relu_in, as_stream(s));
        if (rc != NPP_E_UNSUPPORTED) return rc;
        return dw_fwd_t<T, 3>(x, w, y, stride, pad, dil, relu_in, as_stream(s));
      } return dw_fwd_t<T, 5>(x, w, y, stride, pad, dil, relu_in, as_stream(s)););
}

int npp_dwconv_bwd(const npp_view4* x, const float* w, const npp_view4* dy, const npp_view4* dx, float* dw, int k,
                   int stride, int pad, int dil, int relu_in, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(dy, dtype) || !w) return NPP_E_INVALID;
  if (!dw_shapes_ok(x, dy, k, stride, pad, dil)) return NPP_E_UNSUPPORTED;
  if (dx && (!view_ok(dx, dtype) || !same_shape(x, dx))) return NPP_E_INVALID;
  cudaStream_t st = as_stream(s);
  int rc = NPP_OK;
  NPP_DISPATCH_DTYPE(
      dtype,
      if (dx) {
        rc = (k == 3) ? dw_tile_dgrad<T>(x, w, dy, dx, stride, pad, dil, relu_in, st) : NPP_E_UNSUPPORTED;
        if (rc == NPP_E_UNSUPPORTED)
          rc = (k == 3) ? dw_bwd_data_t<T, 3>(x, w, dy, dx, stride, pad, dil, relu_in, st)
                        : dw_bwd_data_t<T, 5>(x, w, dy, dx, stride, pad, dil, relu_in, st);
        if (rc) return rc;
      } if (dw) {
        rc = (k == 3) ? dw_tile_wgrad<T>(x, dy, dw, stride, pad, dil, relu_in, st) : NPP_E_UNSUPPORTED;
        if (rc == NPP_E_UNSUPPORTED)
          rc = (k == 3) ? dw_bwd_weight_t<T, 3>(x, dy, dw, stride, pad, dil, relu_in, st)
                        : dw_bwd_weight_t<T, 5>(x, dy, dw, stride, pad, dil, relu_in, st);
      } return rc;);
}

}  // extern "C"
