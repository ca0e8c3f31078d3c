// SE_Block (models/operations.py:105-129):
//   w = sigmoid(conv2(relu(conv1(AdaptiveAvgPool2d(1)(x)))));  out = x * w
// Three HBM passes forward (squeeze read, scale read+write); the two 1x1 convolutions on a
// [N, C, 1, 1] tensor are per-image GEMVs done by one CTA per image in shared memory.
#include "view.cuh"

namespace npp {

template <typename T>
static int gap_fwd_t(const npp_view4* x, float* g, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const float inv = 1.f / (float)((int64_t)x->h * x->w);
  return reduce_ch<V, 1>(x->n, x->h, x->w, x->c, true, g, 0, st, "gap_fwd",
                         [=] __device__(int n, int h, int w, int c, float (&acc)[1][V]) {
                           float v[V];
                           Pack<T>::load(X.at(n, h, w, c), v);
#pragma unroll
                           for (int i = 0; i < V; ++i) acc[0][i] += v[i] * inv;
                         });
}

// one block (1024 threads: the two GEMVs are latency chains, 32 warps hide them 4x better than 8) per image;
// g [N,C], w1 [C/2, C], w2 [C, C/2]
__global__ void __launch_bounds__(1024) se_fc_fwd_kernel(const float* __restrict__ g, const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ hbuf,
                                 float* __restrict__ s, int C) {
  pdl_wait();
  extern __shared__ float sm[];
  float* sg = sm;       // C
  float* sh = sm + C;   // C/2
  const int n = blockIdx.x, Ch = C / 2;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sg[i] = g[(int64_t)n * C + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Ch; j += nw) {
    float a = 0.f;
    for (int i = lane; i < C; i += 32) a = fmaf(w1[(int64_t)j * C + i], sg[i], a);
    a = warp_sum(a);
    if (lane == 0) {
      a = fmaxf(a + (b1 ? b1[j] : 0.f), 0.f);
      sh[j] = a;
      hbuf[(int64_t)n * Ch + j] = a;
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    float a = 0.f;
    for (int j = lane; j < Ch; j += 32) a = fmaf(w2[(int64_t)c * Ch + j], sh[j], a);
    a = warp_sum(a);
    if (lane == 0) s[(int64_t)n * C + c] = 1.f / (1.f + __expf(-(a + (b2 ? b2[c] : 0.f))));
  }
}

// ds [N,C] = d loss / d s.  One block per image; parameter gradients accumulated with atomics.
__global__ void __launch_bounds__(1024) se_fc_bwd_kernel(const float* __restrict__ g, const float* __restrict__ hbuf, const float* __restrict__ s,
                                 const float* __restrict__ ds, const float* __restrict__ w1, const float* __restrict__ w2,
                                 float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                 float* __restrict__ db2, float* __restrict__ dg, int C) {
  pdl_wait();
  extern __shared__ float sm[];
  float* dz2 = sm;            // C
  float* sg = sm + C;         // C
  float* sh = sm + 2 * C;     // C/2
  float* dz1 = sh + C / 2;    // C/2
  const int n = blockIdx.x, Ch = C / 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sv = s[(int64_t)n * C + c];
    const float d = ds[(int64_t)n * C + c] * sv * (1.f - sv);
    dz2[c] = d;
    sg[c] = g[(int64_t)n * C + c];
    if (db2) atomicAdd(db2 + c, d);
  }
  for (int j = threadIdx.x; j < Ch; j += blockDim.x) sh[j] = hbuf[(int64_t)n * Ch + j];
  __syncthreads();
  // dW2[c,j] += dz2[c]*h[j]
  for (int i = threadIdx.x; i < C * Ch; i += blockDim.x) atomicAdd(dw2 + i, dz2[i / Ch] * sh[i % Ch]);
  // dh[j] = sum_c W2[c,j] dz2[c]; dz1 = dh * (h>0)
  for (int j = threadIdx.x; j < Ch; j += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(w2[(int64_t)c * Ch + j], dz2[c], a);
    a = sh[j] > 0.f ? a : 0.f;
    dz1[j] = a;
    if (db1) atomicAdd(db1 + j, a);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Ch * C; i += blockDim.x) atomicAdd(dw1 + i, dz1[i / C] * sg[i % C]);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < Ch; ++j) a = fmaf(w1[(int64_t)j * C + c], dz1[j], a);
    dg[(int64_t)n * C + c] = a;
  }
}

// ---- two-kernel backward of the SE bottleneck (round-2 candidate, functional._state["se_bwd2"]) ------------------
// se_fc_bwd_kernel above adds every image's outer products dz2 (x) h and dz1 (x) g into dW2 / dW1 with global
// atomics: 2 * C * C/2 atomics per image on the same addresses from all N blocks (2 M for C = 256, N = 32), which is
// what its 17 us per launch are spent on.  Here the per-image vectors are written once (kernel A, no weight
// atomics) and the weight gradients are formed as [C x N] x [N x C/2] products by one thread per element (kernel B):
// every dW element is written exactly once.
__global__ void __launch_bounds__(1024) se_fc_bwd_vec_kernel(const float* __restrict__ hbuf, const float* __restrict__ s,
                                                             const float* __restrict__ ds, const float* __restrict__ w1,
                                                             const float* __restrict__ w2, float* __restrict__ db1,
                                                             float* __restrict__ db2, float* __restrict__ dz2g,
                                                             float* __restrict__ dz1g, float* __restrict__ dg, int C) {
  pdl_wait();
  extern __shared__ float sm[];
  float* dz2 = sm;          // C
  float* sh = sm + C;       // C/2
  float* dz1 = sh + C / 2;  // C/2
  const int n = blockIdx.x, Ch = C / 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sv = s[(int64_t)n * C + c];
    const float d = ds[(int64_t)n * C + c] * sv * (1.f - sv);
    dz2[c] = d;
    dz2g[(int64_t)n * C + c] = d;
    if (db2) atomicAdd(db2 + c, d);
  }
  for (int j = threadIdx.x; j < Ch; j += blockDim.x) sh[j] = hbuf[(int64_t)n * Ch + j];
  __syncthreads();
  for (int j = threadIdx.x; j < Ch; j += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(w2[(int64_t)c * Ch + j], dz2[c], a);
    a = sh[j] > 0.f ? a : 0.f;
    dz1[j] = a;
    dz1g[(int64_t)n * Ch + j] = a;
    if (db1) atomicAdd(db1 + j, a);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < Ch; ++j) a = fmaf(w1[(int64_t)j * C + c], dz1[j], a);
    dg[(int64_t)n * C + c] = a;
  }
}

// dW2[c, j] += sum_n dz2[n, c] * h[n, j]   (elements [0, C*Ch));   dW1[j, i] += sum_n dz1[n, j] * g[n, i]   (the rest)
__global__ void __launch_bounds__(256) se_fc_bwd_outer_kernel(const float* __restrict__ g, const float* __restrict__ hbuf,
                                                              const float* __restrict__ dz2g,
                                                              const float* __restrict__ dz1g, float* __restrict__ dw1,
                                                              float* __restrict__ dw2, int N, int C) {
  pdl_wait();
  const int Ch = C / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * C * Ch) return;
  float a = 0.f;
  if (e < C * Ch) {
    const int c = e / Ch, j = e % Ch;
    for (int n = 0; n < N; ++n) a = fmaf(dz2g[(int64_t)n * C + c], hbuf[(int64_t)n * Ch + j], a);
    dw2[e] += a;
  } else {
    const int k = e - C * Ch, j = k / C, i = k % C;
    for (int n = 0; n < N; ++n) a = fmaf(dz1g[(int64_t)n * Ch + j], g[(int64_t)n * C + i], a);
    dw1[k] += a;
  }
}

template <typename T>
static int se_scale_fwd_t(const npp_view4* x, const float* s, const npp_view4* y, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto Y = dview<T>(y);
  const int C = x->c;
  return foreach_vec<V>(x->n, x->h, x->w, x->c, st, "se_scale_fwd", [=] __device__(int n, int h, int w, int c) {
    float v[V];
    Pack<T>::load(X.at(n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] *= s[(int64_t)n * C + c + i];
    Pack<T>::store(Y.at(n, h, w, c), v);
  });
}

// ds[n,c] += sum_hw dy*x
template <typename T>
static int se_bwd_reduce_t(const npp_view4* x, const npp_view4* dy, float* ds, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto X = dview<const T>(x);
  const auto DY = dview<const T>(dy);
  return reduce_ch<V, 1>(x->n, x->h, x->w, x->c, true, ds, 0, st, "se_bwd_reduce",
                         [=] __device__(int n, int h, int w, int c, float (&acc)[1][V]) {
                           float v[V], d[V];
                           Pack<T>::load(X.at(n, h, w, c), v);
                           Pack<T>::load(DY.at(n, h, w, c), d);
#pragma unroll
                           for (int i = 0; i < V; ++i) acc[0][i] = fmaf(v[i], d[i], acc[0][i]);
                         });
}

// dx = dy * s[n,c] + dg[n,c] / (H*W)
template <typename T>
static int se_bwd_apply_t(const npp_view4* dy, const float* s, const float* dg, const npp_view4* dx, cudaStream_t st) {
  constexpr int V = Pack<T>::N;
  const auto DY = dview<const T>(dy);
  const auto DX = dview<T>(dx);
  const int C = dy->c;
  const float inv = 1.f / (float)((int64_t)dy->h * dy->w);
  return foreach_vec<V>(dy->n, dy->h, dy->w, dy->c, st, "se_bwd_apply", [=] __device__(int n, int h, int w, int c) {
    float d[V];
    Pack<T>::load(DY.at(n, h, w, c), d);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = fmaf(d[i], s[(int64_t)n * C + c + i], dg[(int64_t)n * C + c + i] * inv);
    Pack<T>::store(DX.at(n, h, w, c), d);
  });
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_gap_fwd(const npp_view4* x, float* g, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !g) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return gap_fwd_t<T>(x, g, as_stream(s)););
}
int npp_se_fc_fwd(const float* g, const float* w1, const float* b1, const float* w2, const float* b2, float* hbuf,
                  float* sg, int n, int c, npp_stream_t s) {
  if (!g || !w1 || !w2 || !hbuf || !sg || n <= 0 || c <= 1 || (c & 1)) return NPP_E_INVALID;
  const size_t smem = (size_t)(c + c / 2) * sizeof(float);
  if (smem > 48 * 1024) return NPP_E_UNSUPPORTED;
  NPP_LAUNCH((se_fc_fwd_kernel), n, 1024, smem, as_stream(s), g, w1, b1, w2, b2, hbuf, sg, c);
  NPP_CHECK_LAUNCH("se_fc_fwd_kernel");
  return NPP_OK;
}
int npp_se_fc_bwd(const float* g, const float* hbuf, const float* sg, const float* ds, const float* w1, const float* w2,
                  float* dw1, float* db1, float* dw2, float* db2, float* dg, int n, int c, npp_stream_t s) {
  if (!g || !hbuf || !sg || !ds || !w1 || !w2 || !dw1 || !dw2 || !dg || n <= 0 || c <= 1 || (c & 1))
    return NPP_E_INVALID;
  const size_t smem = (size_t)(3 * c) * sizeof(float);
  if (smem > 48 * 1024) return NPP_E_UNSUPPORTED;
  NPP_LAUNCH((se_fc_bwd_kernel), n, 1024, smem, as_stream(s), g, hbuf, sg, ds, w1, w2, dw1, db1, dw2, db2, dg, c);
  NPP_CHECK_LAUNCH("se_fc_bwd_kernel");
  return NPP_OK;
}
int npp_se_fc_bwd2(const float* g, const float* hbuf, const float* sg, const float* ds, const float* w1, const float* w2,
                   float* dw1, float* db1, float* dw2, float* db2, float* dg, float* scratch, int n, int c,
                   npp_stream_t s) {
  if (!g || !hbuf || !sg || !ds || !w1 || !w2 || !dw1 || !dw2 || !dg || !scratch || n <= 0 || c <= 1 || (c & 1))
    return NPP_E_INVALID;
  const size_t smem = (size_t)(2 * c) * sizeof(float);
  if (smem > 48 * 1024) return NPP_E_UNSUPPORTED;
  float* dz2g = scratch;                      // [n][c]
  float* dz1g = scratch + (size_t)n * c;      // [n][c/2]
  NPP_LAUNCH((se_fc_bwd_vec_kernel), n, 1024, smem, as_stream(s), hbuf, sg, ds, w1, w2, db1, db2, dz2g, dz1g, dg, c);
  NPP_CHECK_LAUNCH("se_fc_bwd_vec_kernel");
  const int total = c * c;                    // 2 * c * c/2 weight-gradient elements
  NPP_LAUNCH((se_fc_bwd_outer_kernel), (total + 255) / 256, 256, 0, as_stream(s), g, hbuf, dz2g, dz1g, dw1, dw2, n, c);
  NPP_CHECK_LAUNCH("se_fc_bwd_outer_kernel");
  return NPP_OK;
}
int npp_se_scale_fwd(const npp_view4* x, const float* sg, const npp_view4* y, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !same_shape(x, y) || !sg) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return se_scale_fwd_t<T>(x, sg, y, as_stream(s)););
}
int npp_se_bwd_reduce(const npp_view4* x, const npp_view4* dy, float* ds, int dtype, npp_stream_t s) {
  if (!view_ok(x, dtype) || !view_ok(dy, dtype) || !same_shape(x, dy) || !ds) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return se_bwd_reduce_t<T>(x, dy, ds, as_stream(s)););
}
int npp_se_bwd_apply(const npp_view4* dy, const float* sg, const float* dg, const npp_view4* dx, int dtype,
                     npp_stream_t s) {
  if (!view_ok(dy, dtype) || !view_ok(dx, dtype) || !same_shape(dx, dy) || !sg || !dg) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return se_bwd_apply_t<T>(dy, sg, dg, dx, as_stream(s)););
}

}  // extern "C"
