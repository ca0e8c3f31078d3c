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

// ---- bf16 / stride 1 fast paths -----------------------------------------------------------------------------
// The generic kernels above issue one conditional load per window position and run the max / select arithmetic on
// unpacked floats: ~300 instructions and 9 dependent L1 round trips per 16-byte output (ncu: 122 us forward, 206 us
// backward for 128 channels @ 96x96 x 32 against 29 / 35 us of HBM time).  Here a thread owns a vertical strip of
// four outputs of one column and one 8-channel vector: the 6 x 3 input vectors it needs are loaded up front
// (clamped addresses, validity kept as bit masks — 18 independent loads in flight, each reused by up to three
// outputs) and the running maximum is kept as packed bf16x2 selected by a packed `>` mask (exact: no rounding is
// involved; `>` with -inf initial value and row-major scan order = ATen's first-maximum rule, identical to the
// generic kernel).
__device__ __forceinline__ uint32_t bf2_gt_mask(uint32_t a, uint32_t b) {  // 0xffff per 16-bit lane where a > b
  uint32_t m;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
  return m;
}

constexpr int kPoolStrip = 4;

static int maxpool_fwd_strip_bf16(const npp_view4* x, const npp_view4* y, uint8_t* idx, cudaStream_t st) {
  const auto X = dview<const __nv_bfloat16>(x);
  const auto Y = dview<__nv_bfloat16>(y);
  const int H = x->h, W = x->w, C = y->c;
  const int HB = (H + kPoolStrip - 1) / kPoolStrip;
  return foreach_vec<8>(y->n, HB, W, C, st, "maxpool3x3_fwd", [=] __device__(int n, int hb, int wo, int c) {
    const int h0 = hb * kPoolStrip;
    uint4 v[kPoolStrip + 2][3];
    uint32_t vh = 0, vw = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) vw |= (wo - 1 + s >= 0 && wo - 1 + s < W) ? (1u << s) : 0u;
#pragma unroll
    for (int r = 0; r < kPoolStrip + 2; ++r) {
      const int hi = h0 - 1 + r;
      vh |= (hi >= 0 && hi < H) ? (1u << r) : 0u;
      const int hc = min(max(hi, 0), H - 1);
#pragma unroll
      for (int s = 0; s < 3; ++s) v[r][s] = ldraw(X.at(n, hc, min(max(wo - 1 + s, 0), W - 1), c));
    }
#pragma unroll
    for (int j = 0; j < kPoolStrip; ++j) {
      const int ho = h0 + j;
      if (ho >= H) break;
      uint32_t m[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u};  // -inf
      uint32_t a[4] = {0u, 0u, 0u, 0u};                                      // winner index per 16-bit lane
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (!((vh >> (j + r)) & 1u)) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          if (!((vw >> s) & 1u)) continue;
          const uint32_t pos = (uint32_t)(r * 3 + s) * 0x00010001u;
          const uint32_t u[4] = {v[j + r][s].x, v[j + r][s].y, v[j + r][s].z, v[j + r][s].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t gt = bf2_gt_mask(u[k], m[k]);
            a[k] = (a[k] & ~gt) | (pos & gt);
            m[k] = (m[k] & ~gt) | (u[k] & gt);
          }
        }
      }
      *reinterpret_cast<uint4*>(Y.at(n, ho, wo, c)) = make_uint4(m[0], m[1], m[2], m[3]);
      if (idx != nullptr) {
        uint2 pk;
        pk.x = __byte_perm(a[0], a[1], 0x6420);
        pk.y = __byte_perm(a[2], a[3], 0x6420);
        *reinterpret_cast<uint2*>(idx + (((int64_t)n * H + ho) * W + wo) * C + c) = pk;
      }
    }
  });
}

static int maxpool_bwd_strip_bf16(const uint8_t* idx, const npp_view4* dy, const npp_view4* dx, cudaStream_t st) {
  const auto DY = dview<const __nv_bfloat16>(dy);
  const auto DX = dview<__nv_bfloat16>(dx);
  const int H = dx->h, W = dx->w, C = dy->c;
  const int HB = (H + kPoolStrip - 1) / kPoolStrip;
  return foreach_vec<8>(dx->n, HB, W, C, st, "maxpool3x3_bwd", [=] __device__(int n, int hb, int w, int c) {
    const int h0 = hb * kPoolStrip;
    // candidate windows: ho in h-1..h+1, wo in w-1..w+1; window (ho, wo) holds (h, w) at position (h-ho+1, w-wo+1)
    uint4 d[kPoolStrip + 2][3];
    uint2 ix[kPoolStrip + 2][3];
    uint32_t vh = 0, vw = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) vw |= (w - 1 + s >= 0 && w - 1 + s < W) ? (1u << s) : 0u;
#pragma unroll
    for (int r = 0; r < kPoolStrip + 2; ++r) {
      const int ho = h0 - 1 + r;
      vh |= (ho >= 0 && ho < H) ? (1u << r) : 0u;
      const int hc = min(max(ho, 0), H - 1);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int wc = min(max(w - 1 + s, 0), W - 1);
        d[r][s] = ldraw(DY.at(n, hc, wc, c));
        ix[r][s] = *reinterpret_cast<const uint2*>(idx + (((int64_t)n * H + hc) * W + wc) * C + c);
      }
    }
#pragma unroll
    for (int j = 0; j < kPoolStrip; ++j) {
      const int h = h0 + j;
      if (h >= H) break;
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {       // window row ho = h - 1 + r  -> strip row j + r, position row 2 - r
        if (!((vh >> (j + r)) & 1u)) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {     // window column wo = w - 1 + s -> position column 2 - s
          if (!((vw >> s) & 1u)) continue;
          const uint32_t me = (uint32_t)((2 - r) * 3 + (2 - s)) * 0x01010101u;
          const uint32_t e0 = __vcmpeq4(ix[j + r][s].x, me), e1 = __vcmpeq4(ix[j + r][s].y, me);  // 0xff per winner byte
          const uint32_t lm[4] = {__byte_perm(e0, 0u, 0x1100), __byte_perm(e0, 0u, 0x3322),
                                  __byte_perm(e1, 0u, 0x1100), __byte_perm(e1, 0u, 0x3322)};
          const uint32_t u[4] = {d[j + r][s].x & lm[0], d[j + r][s].y & lm[1], d[j + r][s].z & lm[2],
                                 d[j + r][s].w & lm[3]};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            g[2 * k] += __uint_as_float(u[k] << 16);
            g[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
          }
        }
      }
      Pack<__nv_bfloat16>::store(DX.at(n, h, w, c), g);
    }
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
  if (dtype == NPP_BF16 && stride == 1) return maxpool_fwd_strip_bf16(x, y, argmax, as_stream(s));
  NPP_DISPATCH_DTYPE(dtype, return maxpool_fwd_t<T>(x, y, argmax, stride, as_stream(s)););
}
int npp_maxpool3x3_bwd(const uint8_t* argmax, const npp_view4* dy, const npp_view4* dx, int stride, int dtype,
                       npp_stream_t s) {
  if (!argmax || !view_ok(dy, dtype) || !view_ok(dx, dtype) || !pool3_shapes_ok(dx, dy, stride)) return NPP_E_INVALID;
  if (dtype == NPP_BF16 && stride == 1) return maxpool_bwd_strip_bf16(argmax, dy, dx, as_stream(s));
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
