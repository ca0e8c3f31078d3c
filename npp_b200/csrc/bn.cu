// BatchNorm2d (training + eval) as HBM-bound passes over NHWC views.
// Reference: nn.BatchNorm2d(momentum=0.1, eps=1e-5) in models/operations.py:61,79,97,117,151,163,184,215,240
// and models/model_augment.py:246-398.  The batch statistics are kept as raw per-channel sums so
// the SyncBN all-reduce (augment_lip_sync.py:191) is one 2C-float message between stats and finalize.
#include "view.cuh"

namespace npp {

template <typename T>
static int bn_stats_t(const npp_view4* x, float* sums, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  return reduce_ch<V, 2>(x->n, x->h, x->w, x->c, false, sums, x->c, st, "bn_stats",
                         [=] __device__(int n, int h, int w, int c, float (&acc)[2][V]) {
                           float v[V];
                           Pack<T>::load(X.at(n, h, w, c), v);
#pragma unroll
                           for (int i = 0; i < V; ++i) {
                             acc[0][i] += v[i];
                             acc[1][i] += v[i] * v[i];
                           }
                         });
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd, int C, int C_run) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = (double)sums[c] / count;
  double var = (double)sums[C + c] / count - mean * mean;  // biased variance, as ATen normalises with
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - (float)mean * sc;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean && c < C_run) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var && c < C_run) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                    float* __restrict__ scale, float* __restrict__ shift, int C) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(rv[c] + eps);
  const float sc = (gamma ? gamma[c] : 1.f) * invstd;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

template <typename T>
static int bn_apply_t(const npp_view4* x, const float* scale, const float* shift, const npp_view4* res, int relu,
                      const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  const bool has_res = res != nullptr;
  DView<const T> R = has_res ? dview<const T>(res) : X;
  return foreach_vec<V>(x->n, x->h, x->w, x->c, st, "bn_apply", [=] __device__(int n, int h, int w, int c) {
    float v[V], sc[V], sh[V];
    Pack<T>::load(X.at(n, h, w, c), v);
    Pack<float>::load(scale + c, reinterpret_cast<float(&)[4]>(sc[0]));
    Pack<float>::load(shift + c, reinterpret_cast<float(&)[4]>(sh[0]));
    if (V == 8) {
      Pack<float>::load(scale + c + 4, reinterpret_cast<float(&)[4]>(sc[V - 4]));
      Pack<float>::load(shift + c + 4, reinterpret_cast<float(&)[4]>(sh[V - 4]));
    }
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
    if (has_res) {
      float r[V];
      Pack<T>::load(R.at(n, h, w, c), r);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] += r[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    Pack<T>::store(Y.at(n, h, w, c), v);
  });
}

template <typename T>
static int bn_bwd_reduce_t(const npp_view4* dy, const npp_view4* x, const npp_view4* my, const float* mean,
                           const float* invstd, float* sums, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto X = dview<const T>(x);
  const bool masked = my != nullptr;
  DView<const T> M = masked ? dview<const T>(my) : X;
  return reduce_ch<V, 2>(x->n, x->h, x->w, x->c, false, sums, x->c, st, "bn_bwd_reduce",
                         [=] __device__(int n, int h, int w, int c, float (&acc)[2][V]) {
                           float g[V], v[V];
                           Pack<T>::load(DY.at(n, h, w, c), g);
                           Pack<T>::load(X.at(n, h, w, c), v);
                           if (masked) {
                             float m[V];
                             Pack<T>::load(M.at(n, h, w, c), m);
#pragma unroll
                             for (int i = 0; i < V; ++i) g[i] = m[i] > 0.f ? g[i] : 0.f;
                           }
#pragma unroll
                           for (int i = 0; i < V; ++i) {
                             acc[0][i] += g[i];
                             acc[1][i] += g[i] * (v[i] - mean[c + i]) * invstd[c + i];
                           }
                         });
}

template <typename T>
static int bn_bwd_apply_t(const npp_view4* dy, const npp_view4* x, const npp_view4* my, const float* gamma,
                          const float* mean, const float* invstd, const float* sums, double count, const npp_view4* dx,
                          cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto X = dview<const T>(x);
  const auto DX = dview<T>(dx);
  const bool masked = my != nullptr;
  DView<const T> M = masked ? dview<const T>(my) : X;
  const int C = x->c;
  const float inv_count = (float)(1.0 / count);
  return foreach_vec<V>(x->n, x->h, x->w, x->c, st, "bn_bwd_apply", [=] __device__(int n, int h, int w, int c) {
    float g[V], v[V];
    Pack<T>::load(DY.at(n, h, w, c), g);
    Pack<T>::load(X.at(n, h, w, c), v);
    if (masked) {
      float m[V];
      Pack<T>::load(M.at(n, h, w, c), m);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = m[i] > 0.f ? g[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float is = invstd[c + i];
      const float xh = (v[i] - mean[c + i]) * is;
      const float ga = gamma ? gamma[c + i] : 1.f;
      g[i] = ga * is * (g[i] - sums[c + i] * inv_count - xh * sums[C + c + i] * inv_count);
    }
    Pack<T>::store(DX.at(n, h, w, c), g);
  });
}

template <typename T>
static int colsum_t(const npp_view4* x, float* out, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  return reduce_ch<V, 1>(x->n, x->h, x->w, x->c, false, out, x->c, st, "colsum",
                         [=] __device__(int n, int h, int w, int c, float (&acc)[1][V]) {
                           float v[V];
                           Pack<T>::load(X.at(n, h, w, c), v);
#pragma unroll
                           for (int i = 0; i < V; ++i) acc[0][i] += v[i];
                         });
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_bn_stats(const npp_view4* x, float* sums, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !sums) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return bn_stats_t<T>(x, sums, as_stream(s)););
}
int npp_bn_finalize(const float* sums, double count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                    float* save_invstd, int c, int c_running, npp_stream_t s) {
  if (!sums || !scale || !shift || c <= 0 || count <= 0 || c_running > c) return NPP_E_INVALID;
  NPP_LAUNCH((bn_finalize_kernel), (c + 127) / 128, 128, 0, as_stream(s), sums, count, gamma, beta, running_mean, running_var,
                                                                momentum, eps, scale, shift, save_mean, save_invstd, c, c_running);
  NPP_CHECK_LAUNCH("bn_finalize_kernel");
  return NPP_OK;
}
int npp_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     float eps, float* scale, float* shift, int c, npp_stream_t s) {
  if (!running_mean || !running_var || !scale || !shift || c <= 0) return NPP_E_INVALID;
  NPP_LAUNCH((bn_eval_coef_kernel), (c + 127) / 128, 128, 0, as_stream(s), gamma, beta, running_mean, running_var, eps, scale,
                                                                 shift, c);
  NPP_CHECK_LAUNCH("bn_eval_coef_kernel");
  return NPP_OK;
}
int npp_bn_apply(const npp_view4* x, const float* scale, const float* shift, const npp_view4* res, int relu,
                 const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !same_shape(x, y) || !scale || !shift) return NPP_E_INVALID;
  if (res && (!view_ok(res, dtype) || !same_shape(x, res))) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return bn_apply_t<T>(x, scale, shift, res, relu, y, as_stream(s)););
}
int npp_bn_bwd_reduce(const npp_view4* dy, const npp_view4* x, const npp_view4* relu_mask_y, const float* save_mean,
                      const float* save_invstd, float* sums, int dtype, npp_stream_t s) {
  if (!view_ok(dy, dtype) || !view_ok(x, dtype) || !same_shape(x, dy) || !save_mean || !save_invstd || !sums)
    return NPP_E_INVALID;
  if (relu_mask_y && (!view_ok(relu_mask_y, dtype) || !same_shape(x, relu_mask_y))) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return bn_bwd_reduce_t<T>(dy, x, relu_mask_y, save_mean, save_invstd, sums, as_stream(s)););
}
int npp_bn_bwd_apply(const npp_view4* dy, const npp_view4* x, const npp_view4* relu_mask_y, const float* gamma,
                     const float* save_mean, const float* save_invstd, const float* sums, double count,
                     const npp_view4* dx, int dtype, npp_stream_t s) {
  if (!view_ok(dy, dtype) || !view_ok(x, dtype) || !view_ok(dx, dtype) || !same_shape(x, dy) || !same_shape(x, dx) ||
      !save_mean || !save_invstd || !sums || count <= 0)
    return NPP_E_INVALID;
  if (relu_mask_y && (!view_ok(relu_mask_y, dtype) || !same_shape(x, relu_mask_y))) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return bn_bwd_apply_t<T>(dy, x, relu_mask_y, gamma, save_mean, save_invstd, sums, count,
                                                     dx, as_stream(s)););
}
int npp_colsum(const npp_view4* x, float* out, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !out) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return colsum_t<T>(x, out, as_stream(s)););
}

}  // extern "C"
