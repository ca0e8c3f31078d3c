// Direct (CUDA-core) dense convolution: the fp32 validation mode of the north star and the
// on-device cross-check for the tcgen05 path.  Deliberately simple: one thread per output
// element, fp32 accumulation, no tiling.  Not a product path for bf16 (Network refuses to use
// it outside validation mode) — see conv_tcgen05.cu for the real kernels.
#include "common.cuh"

namespace npp {

template <typename T, typename WT>
__global__ void direct_fwd_kernel(const T* __restrict__ x, const WT* __restrict__ w, const float* __restrict__ bias,
                                  T* __restrict__ y, int N, int H, int W, int Cin, int64_t xsn, int64_t xsh,
                                  int64_t xsw, int Ho, int Wo, int Cout, int64_t ysn, int64_t ysh, int64_t ysw, int kh,
                                  int kw, int stride, int pad, int dil, int hoff, int woff) {
  pdl_wait();
  const int64_t total = (int64_t)N * Ho * Wo * Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    int64_t p = i / Cout;
    const int wo = (int)(p % Wo);
    p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float acc = bias ? bias[co] : 0.f;
    for (int r = 0; r < kh; ++r) {
      const int hi = ho * stride - pad + r * dil + hoff;
      if (hi < 0 || hi >= H) continue;
      for (int s = 0; s < kw; ++s) {
        const int wi = wo * stride - pad + s * dil + woff;
        if (wi < 0 || wi >= W) continue;
        const T* xp = x + n * xsn + hi * xsh + wi * xsw;
        const WT* wp = w + ((int64_t)co * kh * kw + r * kw + s) * Cin;
        for (int ci = 0; ci < Cin; ++ci) acc += to_f<T>(xp[ci]) * to_f<WT>(wp[ci]);
      }
    }
    y[n * ysn + ho * ysh + wo * ysw + co] = from_f<T>(acc);
  }
}

template <typename T, typename WT>
__global__ void direct_dgrad_kernel(const T* __restrict__ dy, const WT* __restrict__ w, T* __restrict__ dx, int N,
                                    int H, int W, int Cin, int64_t xsn, int64_t xsh, int64_t xsw, int Ho, int Wo,
                                    int Cout, int64_t ysn, int64_t ysh, int64_t ysw, int kh, int kw, int stride,
                                    int pad, int dil, int hoff, int woff) {
  pdl_wait();
  const int64_t total = (int64_t)N * H * W * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t p = i / Cin;
    const int wi = (int)(p % W);
    p /= W;
    const int hi = (int)(p % H);
    const int n = (int)(p / H);
    float acc = 0.f;
    for (int r = 0; r < kh; ++r) {
      int ho = hi + pad - r * dil - hoff;
      if (ho < 0 || ho % stride) continue;
      ho /= stride;
      if (ho >= Ho) continue;
      for (int s = 0; s < kw; ++s) {
        int wo = wi + pad - s * dil - woff;
        if (wo < 0 || wo % stride) continue;
        wo /= stride;
        if (wo >= Wo) continue;
        const T* dp = dy + n * ysn + ho * ysh + wo * ysw;
        const WT* wp = w + (int64_t)(r * kw + s) * Cin + ci;
        for (int co = 0; co < Cout; ++co) acc += to_f<T>(dp[co]) * to_f<WT>(wp[(int64_t)co * kh * kw * Cin]);
      }
    }
    dx[n * xsn + hi * xsh + wi * xsw + ci] = from_f<T>(acc);
  }
}

// grid.x over dW elements [dw_cout, taps, dw_cin], grid.y over pixel chunks; dw += partial.
template <typename T>
__global__ void direct_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int N,
                                    int H, int W, int64_t xsn, int64_t xsh, int64_t xsw, int Ho, int Wo,
                                    int64_t ysn, int64_t ysh, int64_t ysw, int dw_cout, int dw_cin, int kh, int kw,
                                    int stride, int pad, int dil, int hoff, int woff, int pix_per_chunk) {
  pdl_wait();
  const int64_t total = (int64_t)dw_cout * kh * kw * dw_cin;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // OIHW index
  if (i >= total) return;
  const int t = (int)(i % (kh * kw));
  const int ci = (int)((i / (kh * kw)) % dw_cin);
  const int co = (int)(i / ((int64_t)dw_cin * kh * kw));
  const int r = t / kw, s = t % kw;
  const int64_t npix = (int64_t)N * Ho * Wo;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_chunk;
  int64_t p1 = p0 + pix_per_chunk;
  if (p1 > npix) p1 = npix;
  float acc = 0.f;
  for (int64_t p = p0; p < p1; ++p) {
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int n = (int)(p / ((int64_t)Wo * Ho));
    const int hi = ho * stride - pad + r * dil + hoff, wi = wo * stride - pad + s * dil + woff;
    if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
    acc += to_f<T>(dy[n * ysn + ho * ysh + wo * ysw + co]) * to_f<T>(x[n * xsn + hi * xsh + wi * xsw + ci]);
  }
  atomicAdd(dw + i, acc);
}

static int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T, typename WT>
static int direct_fwd_t(const npp_view4* x, const void* w, const float* bias, const npp_view4* y, int kh, int kw,
                        int stride, int pad, int dil, int hoff, int woff, cudaStream_t st) {
  const int64_t total = (int64_t)y->n * y->h * y->w * y->c;
  NPP_LAUNCH((direct_fwd_kernel<T, WT>), grid_for(total), 256, 0, st, 
      static_cast<const T*>(x->ptr), static_cast<const WT*>(w), bias, static_cast<T*>(y->ptr), x->n, x->h, x->w, x->c,
      x->sn, x->sh, x->sw, y->h, y->w, y->c, y->sn, y->sh, y->sw, kh, kw, stride, pad, dil, hoff, woff);
  NPP_CHECK_LAUNCH("direct_fwd_kernel");
  return NPP_OK;
}
template <typename T, typename WT>
static int direct_dgrad_t(const npp_view4* dy, const void* w, const npp_view4* dx, int kh, int kw, int stride, int pad,
                          int dil, int hoff, int woff, cudaStream_t st) {
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * dx->c;
  NPP_LAUNCH((direct_dgrad_kernel<T, WT>), grid_for(total), 256, 0, st, 
      static_cast<const T*>(dy->ptr), static_cast<const WT*>(w), static_cast<T*>(dx->ptr), dx->n, dx->h, dx->w, dx->c,
      dx->sn, dx->sh, dx->sw, dy->h, dy->w, dy->c, dy->sn, dy->sh, dy->sw, kh, kw, stride, pad, dil, hoff, woff);
  NPP_CHECK_LAUNCH("direct_dgrad_kernel");
  return NPP_OK;
}
template <typename T>
static int direct_wgrad_t(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, int kh, int kw,
                          int stride, int pad, int dil, int hoff, int woff, cudaStream_t st) {
  const int64_t total = (int64_t)dw_cout * kh * kw * dw_cin;
  const int64_t npix = (int64_t)dy->n * dy->h * dy->w;
  const int chunk = 2048;
  dim3 grid((unsigned)((total + 127) / 128), (unsigned)((npix + chunk - 1) / chunk));
  NPP_LAUNCH((direct_wgrad_kernel<T>), grid, 128, 0, st, static_cast<const T*>(x->ptr), static_cast<const T*>(dy->ptr), dw, x->n,
                                              x->h, x->w, x->sn, x->sh, x->sw, dy->h, dy->w, dy->sn, dy->sh, dy->sw,
                                              dw_cout, dw_cin, kh, kw, stride, pad, dil, hoff, woff, chunk);
  NPP_CHECK_LAUNCH("direct_wgrad_kernel");
  return NPP_OK;
}

namespace tc {
int conv_fwd(const npp_view4* x, const void* w, const float* bias, const npp_view4* y, int kh, int kw, int stride,
             int pad, int dil, int hoff, int woff, float* stats, cudaStream_t st);
int conv_dgrad(const npp_view4* dy, const void* wt, const npp_view4* dx, int kh, int kw, int stride, int pad, int dil,
               int hoff, int woff, cudaStream_t st);
int conv_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, int kh, int kw, int stride,
               int pad, int dil, int hoff, int woff, float* ws, size_t ws_bytes, cudaStream_t st);
int pack_weight(const float* w32, void* w, void* wt, int cout, int taps, int cin, int cout_pad, int cin_pad,
                int out_dtype, cudaStream_t st);
int pack_weights_multi(const void* table, int ntensors, const int* chunk_tensor, const int* chunk_index, int nchunks,
                       int chunk_elems, cudaStream_t st);
int pack_weights_tiles(const void* table, int ntensors, const int* tile_tensor, const int* tile_index, int ntiles,
                       cudaStream_t st);
int pack_weight_pair(const float* w32, void* w, void* wt, cudaStream_t st);
int fold_pair_wgrad(const float* dwp, float* dw, cudaStream_t st);
}  // namespace tc

}  // namespace npp

using namespace npp;

extern "C" {

int npp_conv2d_fwd(const npp_view4* x, const void* w, const float* bias, const npp_view4* y, int kh, int kw, int stride,
                   int pad, int dil, int in_h_off, int in_w_off, float* stats, npp_stream_t stream) {
  if (!x || !y) return NPP_E_INVALID;
  return tc::conv_fwd(x, w, bias, y, kh, kw, stride, pad, dil, in_h_off, in_w_off, stats, as_stream(stream));
}
int npp_conv2d_dgrad(const npp_view4* dy, const void* wt, const npp_view4* dx, int kh, int kw, int stride, int pad,
                     int dil, int in_h_off, int in_w_off, npp_stream_t stream) {
  if (!dy || !dx) return NPP_E_INVALID;
  return tc::conv_dgrad(dy, wt, dx, kh, kw, stride, pad, dil, in_h_off, in_w_off, as_stream(stream));
}
int npp_conv2d_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, int kh, int kw,
                     int stride, int pad, int dil, int in_h_off, int in_w_off, npp_stream_t stream) {
  if (!x || !dy) return NPP_E_INVALID;
  return tc::conv_wgrad(x, dy, dw, dw_cout, dw_cin, kh, kw, stride, pad, dil, in_h_off, in_w_off, nullptr, 0,
                        as_stream(stream));
}
int64_t npp_conv2d_wgrad_workspace_bytes(void) {
  // one fp32 partial tile set per CTA: <= sm_count CTAs x max(3 x 128 x 128, 128 x 256) floats
  return (int64_t)sm_count() * 3 * 128 * 128 * (int64_t)sizeof(float);
}
int npp_conv2d_wgrad_ws(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, int kh, int kw,
                        int stride, int pad, int dil, int in_h_off, int in_w_off, void* workspace,
                        int64_t workspace_bytes, npp_stream_t stream) {
  if (!x || !dy || workspace_bytes < 0) return NPP_E_INVALID;
  return tc::conv_wgrad(x, dy, dw, dw_cout, dw_cin, kh, kw, stride, pad, dil, in_h_off, in_w_off,
                        static_cast<float*>(workspace), (size_t)workspace_bytes, as_stream(stream));
}
int npp_pack_weights_multi(const void* table, int ntensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                           int nchunks, int chunk_elems, npp_stream_t stream) {
  return tc::pack_weights_multi(table, ntensors, chunk_tensor, chunk_index, nchunks, chunk_elems, as_stream(stream));
}
int npp_pack_weights_tiles(const void* table, int ntensors, const int32_t* tile_tensor, const int32_t* tile_index,
                           int ntiles, npp_stream_t stream) {
  return tc::pack_weights_tiles(table, ntensors, tile_tensor, tile_index, ntiles, as_stream(stream));
}
int npp_pack_weight_pair(const float* w32, void* w, void* wt, npp_stream_t stream) {
  return tc::pack_weight_pair(w32, w, wt, as_stream(stream));
}
int npp_fold_pair_wgrad(const float* dwp, float* dw, npp_stream_t stream) {
  return tc::fold_pair_wgrad(dwp, dw, as_stream(stream));
}
int npp_pack_weight(const float* w32, void* w, void* wt, int cout, int taps, int cin, int cout_pad, int cin_pad,
                    int out_dtype, npp_stream_t stream) {
  return tc::pack_weight(w32, w, wt, cout, taps, cin, cout_pad, cin_pad, out_dtype, as_stream(stream));
}

int npp_conv2d_direct_fwd(const npp_view4* x, const void* w, const float* bias, const npp_view4* y, int kh, int kw,
                          int stride, int pad, int dil, int in_h_off, int in_w_off, int dtype, npp_stream_t stream) {
  if (!view_ok(x, dtype) || !view_ok(y, dtype) || !w || x->n != y->n) return NPP_E_INVALID;
  if (dtype == NPP_F32)
    return direct_fwd_t<float, float>(x, w, bias, y, kh, kw, stride, pad, dil, in_h_off, in_w_off, as_stream(stream));
  return direct_fwd_t<__nv_bfloat16, __nv_bfloat16>(x, w, bias, y, kh, kw, stride, pad, dil, in_h_off, in_w_off,
                                                    as_stream(stream));
}
int npp_conv2d_direct_dgrad(const npp_view4* dy, const void* w, const npp_view4* dx, int kh, int kw, int stride,
                            int pad, int dil, int in_h_off, int in_w_off, int dtype, npp_stream_t stream) {
  if (!view_ok(dy, dtype) || !view_ok(dx, dtype) || !w || dx->n != dy->n) return NPP_E_INVALID;
  if (dtype == NPP_F32)
    return direct_dgrad_t<float, float>(dy, w, dx, kh, kw, stride, pad, dil, in_h_off, in_w_off, as_stream(stream));
  return direct_dgrad_t<__nv_bfloat16, __nv_bfloat16>(dy, w, dx, kh, kw, stride, pad, dil, in_h_off, in_w_off,
                                                      as_stream(stream));
}
int npp_conv2d_direct_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, int kh,
                            int kw, int stride, int pad, int dil, int in_h_off, int in_w_off, int dtype,
                            npp_stream_t stream) {
  if (!view_ok(x, dtype) || !view_ok(dy, dtype) || !dw || x->n != dy->n) return NPP_E_INVALID;
  if (dw_cout > dy->c || dw_cin > x->c) return NPP_E_INVALID;
  if (dtype == NPP_F32)
    return direct_wgrad_t<float>(x, dy, dw, dw_cout, dw_cin, kh, kw, stride, pad, dil, in_h_off, in_w_off,
                                 as_stream(stream));
  return direct_wgrad_t<__nv_bfloat16>(x, dy, dw, dw_cout, dw_cin, kh, kw, stride, pad, dil, in_h_off, in_w_off,
                                       as_stream(stream));
}

}  // extern "C"
