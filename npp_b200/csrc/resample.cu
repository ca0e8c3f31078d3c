// Bilinear / nearest resampling over NHWC views.
//  bilinear: F.interpolate(mode='bilinear', align_corners=True)  models/model_augment.py:116,539-543,
//            nn.UpsamplingBilinear2d (operations.py:241), criterion.py:181; align_corners=False for the
//            flip-TTA merge (core/function.py:927).  Index arithmetic follows ATen's
//            area_pixel_compute_source_index in fp32 so the sampled taps are identical.
//  nearest : F.interpolate default mode (model_search_interact.py:63-64, model_augment.py:165-166).
// Backward is in gather form: each dx element scans the (small) range of outputs that can touch it and
// re-derives the forward taps, so results are deterministic and need no atomics / fp32 scratch.
#include "view.cuh"
#include <stdlib.h>
#include "resample.cuh"

namespace npp {

template <typename T>
static int bilinear_fwd_t(const npp_view4* x, const npp_view4* y, Axis ah, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "bilinear_fwd", [=] __device__(int n, int ho, int wo, int c) {
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    bilinear_taps(ah, ho, h0, h1, lh0, lh1);
    bilinear_taps(aw, wo, w0, w1, lw0, lw1);
    float a[V], b[V], cc[V], d[V], o[V];
    Pack<T>::load(X.at(n, h0, w0, c), a);
    Pack<T>::load(X.at(n, h0, w1, c), b);
    Pack<T>::load(X.at(n, h1, w0, c), cc);
    Pack<T>::load(X.at(n, h1, w1, c), d);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = lh0 * (lw0 * a[i] + lw1 * b[i]) + lh1 * (lw0 * cc[i] + lw1 * d[i]);
    Pack<T>::store(Y.at(n, ho, wo, c), o);
  });
}

// Up-sampling (>= x2 along w): a thread produces four neighbouring output columns of one row from ONE fetch of the
// (at most four) input columns they touch in the two source rows — 8 loads per 4 outputs instead of 16.  The gather
// form above reads every input element ~4 * scale^2 times through L2 (x4: 302 MB written, 1.2 GB read), which is
// what bounds it (0.33 of HBM speed).  Same arithmetic order as bilinear_fwd_t: bit-identical results.
template <typename T>
static int bilinear_fwd_strip_t(const npp_view4* x, const npp_view4* y, Axis ah, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  constexpr int KW = 4;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  const int Wo = y->w, Wi = x->w;
  return foreach_vec<V>(y->n, y->h, (Wo + KW - 1) / KW, y->c, st, "bilinear_fwd_strip",
                        [=] __device__(int n, int ho, int ws, int c) {
    int h0, h1;
    float lh0, lh1;
    bilinear_taps(ah, ho, h0, h1, lh0, lh1);
    const int wo0 = ws * KW;
    int w0[KW], w1[KW];
    float lw0[KW], lw1[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) bilinear_taps(aw, min(wo0 + k, Wo - 1), w0[k], w1[k], lw0[k], lw1[k]);
    const int base = w0[0];
    uint4 top[KW], bot[KW];  // input columns base .. base + 3 (clamped) of rows h0 and h1
#pragma unroll
    for (int j = 0; j < KW; ++j) {
      const int wc = min(base + j, Wi - 1);
      top[j] = ldraw(X.at(n, h0, wc, c));
      bot[j] = ldraw(X.at(n, h1, wc, c));
    }
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      if (wo0 + k >= Wo) break;
      const int r0 = w0[k] - base, r1 = w1[k] - base;  // 0..3 for scale <= 0.5
      uint4 qa = top[0], qb = top[0], qc = bot[0], qd = bot[0];
#pragma unroll
      for (int j = 1; j < KW; ++j) {
        if (r0 == j) { qa = top[j]; qc = bot[j]; }
        if (r1 == j) { qb = top[j]; qd = bot[j]; }
      }
      float a[V], b[V], cc[V], d[V], o[V];
      Pack<T>::unpack(qa, a);
      Pack<T>::unpack(qb, b);
      Pack<T>::unpack(qc, cc);
      Pack<T>::unpack(qd, d);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = lh0 * (lw0[k] * a[i] + lw1[k] * b[i]) + lh1 * (lw0[k] * cc[i] + lw1[k] * d[i]);
      Pack<T>::store(Y.at(n, ho, wo0 + k, c), o);
    }
  });
}

template <typename T>
static int bilinear_bwd_t(const npp_view4* dy, const npp_view4* dx, Axis ah, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  // (Tried: deriving the column weights once per element instead of per (row, column) candidate — 113 registers
  // instead of 40 cut the resident warps 3x and the kernel, which lives on memory-level parallelism, went from 3.1
  // to 5.7 ms per step even with the branch disabled at run time.  The fix for this kernel is a tiled two-pass
  // separable reduction, not fewer tap evaluations.)
  return foreach_vec<V>(dx->n, dx->h, dx->w, dx->c, st, "bilinear_bwd", [=] __device__(int n, int h, int w, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    int hlo, hhi, wlo, whi;
    bilinear_range(ah, h, hlo, hhi);
    bilinear_range(aw, w, wlo, whi);
    for (int ho = hlo; ho <= hhi; ++ho) {
      int h0, h1;
      float lh0, lh1;
      bilinear_taps(ah, ho, h0, h1, lh0, lh1);
      const float wh = (h0 == h ? lh0 : 0.f) + (h1 == h ? lh1 : 0.f);
      if (wh == 0.f) continue;
      for (int wo = wlo; wo <= whi; ++wo) {
        int w0, w1;
        float lw0, lw1;
        bilinear_taps(aw, wo, w0, w1, lw0, lw1);
        const float ww = (w0 == w ? lw0 : 0.f) + (w1 == w ? lw1 : 0.f);
        if (ww == 0.f) continue;
        float d[V];
        Pack<T>::load(DY.at(n, ho, wo, c), d);
        const float f = wh * ww;
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = fmaf(f, d[i], g[i]);
      }
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

// Separable form of the bilinear backward (the default since the end of round 1; functional._state["bilinear_sep"],
// NPP_BILINEAR_SEP=0 goes back to the gather kernel above.  Parity-green on a B200, see
// tests/test_gpu_zz_bilinear_sep.py; its timing is the first thing to check in round 2).
// The gather kernel above re-reads every dY element through L1/L2 once per input pixel it touches in BOTH axes
// (~(2s+1)^2 candidates at scale s, 4 real contributions) and evaluates the taps of every candidate pair.  Bilinear
// weights factor as wh(ho,h) * ww(wo,w), so
//     T[n, ho, w, c]  = sum_wo ww(wo, w) * dy[n, ho, wo, c]        (pass 1, fp32 scratch of N*Ho*Wi*C floats)
//     dx[n, h, w, c]  = sum_ho wh(ho, h) * T[n, ho, w, c]          (pass 2)
// reads dY once, touches 2s+1 candidates per axis instead of (2s+1)^2, and keeps the same tap arithmetic
// (bilinear_taps / bilinear_range); only the summation order differs from the gather form (fp32 rounding).
// First pass of the separable backward with all candidate loads in flight: the loop form below tests every candidate
// and loads inside the branch, so the (up to 2 / scale) contributing dY vectors of an input column arrive one L2 round
// trip after the other.  Here the candidate window is a fixed MAXC columns from a tight lower bound, every load is
// issued unconditionally (clamped address), the tap test only decides the weight (0 outside).  Same tap arithmetic.
template <typename T, int MAXC>
static int bilinear_bwd_sep_w_batched(const npp_view4* dy, float* tmp, int Wi, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const int Ho = dy->h, Wo = dy->w, C = dy->c;
  const float inv = 1.f / aw.scale, off = aw.align ? 0.f : 0.5f;
  return foreach_vec<V>(dy->n, Ho, Wi, C, st, "bilinear_bwd_sep(w, batched)", [=] __device__(int n, int ho, int w, int c) {
    // taps of wo reach w only if src(wo) is in (w - 1, w + 1): wo > ((w - 1) + off) / scale - off  (one column of slack
    // for fp rounding; the weights below are exact whatever the window)
    int lo = (int)floorf(((float)w - 1.f + off) * inv - off) - 1;
    if (lo < 0) lo = 0;
    if (lo > Wo - MAXC) lo = max(Wo - MAXC, 0);
    uint4 q[MAXC];
    float ww[MAXC];
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int wo = min(lo + j, Wo - 1);
      q[j] = ldraw(DY.at(n, ho, wo, c));
      int w0, w1;
      float lw0, lw1;
      bilinear_taps(aw, wo, w0, w1, lw0, lw1);
      ww[j] = (lo + j < Wo) ? ((w0 == w ? lw0 : 0.f) + (w1 == w ? lw1 : 0.f)) : 0.f;
    }
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      float d[V];
      Pack<T>::unpack(q[j], d);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = fmaf(ww[j], d[i], g[i]);
    }
    float* tp = tmp + (((int64_t)n * Ho + ho) * Wi + w) * C + c;
#pragma unroll
    for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(tp + i) = make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]);
  });
}

template <typename T>
static int bilinear_bwd_sep_t(const npp_view4* dy, const npp_view4* dx, float* tmp, Axis ah, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  const int Ho = dy->h, Wi = dx->w, C = dx->c;
  // measured: 83.42 ms per train step with the batched first pass against 82.92 ms without (profiles/r02_bilinear_ab.txt):
  // the extra registers cost more occupancy than the overlapped loads win -> off by default
  static const int batched = []() { const char* e = getenv("NPP_BILINEAR_BWD_BATCHED"); return (e && *e) ? atoi(e) : 0; }();
  int rc;
  // window = 2 / scale + 4 columns: x2 up-sampling -> 8 (scale >= 0.45 incl. align_corners 47/95), x4 -> 13
  if (batched && aw.scale >= 0.45f && aw.scale <= 1.f && dy->w >= 8)
    rc = bilinear_bwd_sep_w_batched<T, 8>(dy, tmp, Wi, aw, st);
  else if (batched && aw.scale >= 0.2f && aw.scale < 0.45f && dy->w >= 13)
    rc = bilinear_bwd_sep_w_batched<T, 13>(dy, tmp, Wi, aw, st);
  else
  rc = foreach_vec<V>(dy->n, Ho, Wi, C, st, "bilinear_bwd_sep(w)", [=] __device__(int n, int ho, int w, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    int wlo, whi;
    bilinear_range(aw, w, wlo, whi);
    for (int wo = wlo; wo <= whi; ++wo) {
      int w0, w1;
      float lw0, lw1;
      bilinear_taps(aw, wo, w0, w1, lw0, lw1);
      const float ww = (w0 == w ? lw0 : 0.f) + (w1 == w ? lw1 : 0.f);
      if (ww == 0.f) continue;
      float d[V];
      Pack<T>::load(DY.at(n, ho, wo, c), d);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = fmaf(ww, d[i], g[i]);
    }
    float* tp = tmp + (((int64_t)n * Ho + ho) * Wi + w) * C + c;
#pragma unroll
    for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(tp + i) = make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]);
  });
  if (rc) return rc;
  return foreach_vec<V>(dx->n, dx->h, Wi, C, st, "bilinear_bwd_sep(h)", [=] __device__(int n, int h, int w, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    int hlo, hhi;
    bilinear_range(ah, h, hlo, hhi);
    for (int ho = hlo; ho <= hhi; ++ho) {
      int h0, h1;
      float lh0, lh1;
      bilinear_taps(ah, ho, h0, h1, lh0, lh1);
      const float wh = (h0 == h ? lh0 : 0.f) + (h1 == h ? lh1 : 0.f);
      if (wh == 0.f) continue;
      const float* tp = tmp + (((int64_t)n * Ho + ho) * Wi + w) * C + c;
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(tp + i);
        g[i] = fmaf(wh, t.x, g[i]); g[i + 1] = fmaf(wh, t.y, g[i + 1]);
        g[i + 2] = fmaf(wh, t.z, g[i + 2]); g[i + 3] = fmaf(wh, t.w, g[i + 3]);
      }
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

// nearest: src = min(floor(dst * scale), in-1), scale = 1/scale_factor (ATen nearest_neighbor_compute_source_index)
__device__ __forceinline__ int nearest_src(const Axis& a, int o) {
  int s = (int)floorf((float)o * a.scale);
  return s < a.in - 1 ? s : a.in - 1;
}

template <typename T>
static int nearest_fwd_t(const npp_view4* x, const npp_view4* y, Axis ah, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  return foreach_vec<V>(y->n, y->h, y->w, y->c, st, "nearest_fwd", [=] __device__(int n, int ho, int wo, int c) {
    *reinterpret_cast<uint4*>(Y.at(n, ho, wo, c)) =
        *reinterpret_cast<const uint4*>(X.at(n, nearest_src(ah, ho), nearest_src(aw, wo), c));
  });
}

template <typename T>
static int nearest_bwd_t(const npp_view4* dy, const npp_view4* dx, Axis ah, Axis aw, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  return foreach_vec<V>(dx->n, dx->h, dx->w, dx->c, st, "nearest_bwd", [=] __device__(int n, int h, int w, int c) {
    float g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    const float ih = ah.scale > 0.f ? 1.f / ah.scale : 0.f, iw = aw.scale > 0.f ? 1.f / aw.scale : 0.f;
    int hlo = (int)floorf((float)h * ih) - 2, hhi = (int)ceilf((float)(h + 1) * ih) + 2;
    int wlo = (int)floorf((float)w * iw) - 2, whi = (int)ceilf((float)(w + 1) * iw) + 2;
    if (hlo < 0) hlo = 0;
    if (wlo < 0) wlo = 0;
    if (hhi > ah.out - 1) hhi = ah.out - 1;
    if (whi > aw.out - 1) whi = aw.out - 1;
    for (int ho = hlo; ho <= hhi; ++ho) {
      if (nearest_src(ah, ho) != h) continue;
      for (int wo = wlo; wo <= whi; ++wo) {
        if (nearest_src(aw, wo) != w) continue;
        float d[V];
        Pack<T>::load(DY.at(n, ho, wo, c), d);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += d[i];
      }
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

}  // namespace npp

using namespace npp;

extern "C" {

/* scale_h/scale_w: the scale_factor given to F.interpolate (0 if the call gave `size=`); only used
 * when align_corners == 0, mirroring ATen (align_corners=True always uses (in-1)/(out-1)). */
int npp_bilinear_fwd(const npp_view4* x, const npp_view4* y, int align_corners, double scale_h, double scale_w,
                     int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || x->n != y->n || x->c != y->c) return NPP_E_INVALID;
  const Axis ah = make_axis(x->h, y->h, align_corners, scale_h), aw = make_axis(x->w, y->w, align_corners, scale_w);
  static const int strip = []() { const char* e = getenv("NPP_BILINEAR_STRIP"); return (e && *e) ? atoi(e) : 1; }();
  if (strip && aw.scale > 0.f && aw.scale <= 0.5f && y->w >= 8) {
    NPP_DISPATCH_DTYPE(dtype, return bilinear_fwd_strip_t<T>(x, y, ah, aw, as_stream(s)););
  }
  NPP_DISPATCH_DTYPE(dtype, return bilinear_fwd_t<T>(x, y, ah, aw, as_stream(s)););
}
int npp_bilinear_bwd(const npp_view4* dy, const npp_view4* dx, int align_corners, double scale_h, double scale_w,
                     int dtype, npp_stream_t s) {
  if (!view_ok(dx, dtype) || !view_ok(dy, dtype) || dx->n != dy->n || dx->c != dy->c) return NPP_E_INVALID;
  const Axis ah = make_axis(dx->h, dy->h, align_corners, scale_h), aw = make_axis(dx->w, dy->w, align_corners, scale_w);
  NPP_DISPATCH_DTYPE(dtype, return bilinear_bwd_t<T>(dy, dx, ah, aw, as_stream(s)););
}
int npp_bilinear_bwd_sep(const npp_view4* dy, const npp_view4* dx, float* tmp, int align_corners, double scale_h,
                         double scale_w, int dtype, npp_stream_t s) {
  if (!view_ok(dx, dtype) || !view_ok(dy, dtype) || dx->n != dy->n || dx->c != dy->c || !tmp ||
      (reinterpret_cast<uintptr_t>(tmp) & 15))
    return NPP_E_INVALID;
  const Axis ah = make_axis(dx->h, dy->h, align_corners, scale_h), aw = make_axis(dx->w, dy->w, align_corners, scale_w);
  NPP_DISPATCH_DTYPE(dtype, return bilinear_bwd_sep_t<T>(dy, dx, tmp, ah, aw, as_stream(s)););
}
int npp_nearest_fwd(const npp_view4* x, const npp_view4* y, double scale_h, double scale_w, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || x->n != y->n || x->c != y->c) return NPP_E_INVALID;
  const Axis ah = make_axis(x->h, y->h, 0, scale_h), aw = make_axis(x->w, y->w, 0, scale_w);
  NPP_DISPATCH_DTYPE(dtype, return nearest_fwd_t<T>(x, y, ah, aw, as_stream(s)););
}
int npp_nearest_bwd(const npp_view4* dy, const npp_view4* dx, double scale_h, double scale_w, int dtype,
                    npp_stream_t s) {
  if (!view_ok(dx, dtype) || !view_ok(dy, dtype) || dx->n != dy->n || dx->c != dy->c) return NPP_E_INVALID;
  const Axis ah = make_axis(dx->h, dy->h, 0, scale_h), aw = make_axis(dx->w, dy->w, 0, scale_w);
  NPP_DISPATCH_DTYPE(dtype, return nearest_bwd_t<T>(dy, dx, ah, aw, as_stream(s)););
}

}  // extern "C"
