// Depthwise dilated 3x3 convolution with shared-memory halo staging — DilConvS.net[1]
// (models/operations.py:213) forward, data gradient and weight gradient.
//
// The gather kernels in dwconv.cu read every input element nine times through L1/L2 (dilation 2/4 spreads the taps
// over rows that never share a cache line), which caps them near 25 % of HBM bandwidth.  Here a CTA stages the
// input tile of its TH x TW output pixels (+ halo of `dil` pixels per side) for 8 channel vectors (64 bf16 / 32
// fp32 channels = one 128-byte row per pixel) in shared memory with fully coalesced 16-byte loads — the leading
// nn.ReLU (:212) is applied once while staging — and every tap is then a conflict-free LDS.128: each input
// element crosses the L2 -> SM link (TH+2d)(TW+2d)/(TH*TW) times instead of nine.
//
//   fwd   : y[p]  = sum_t relu(x)[p*stride - pad + t*dil] * w[t]
//   dgrad : dx[p] = [x[p] > 0] * sum_t dy[p + pad - t*dil] * w[t]            (stride 1; same kernel, flipped taps)
//   wgrad : dw[c,t] += sum_p dy[p] * relu(x)[p*stride - pad + t*dil]         persistent CTAs keep the 9 x 8
//           accumulators of their channel vector in registers over all their tiles, one atomic per (CTA, c, t).
//
// A block is CVL channel-vector lanes x 256 / CVL pixel lanes.  CVL = 8 (64 bf16 channels per block) for wide layers;
// the search supernet runs its depthwise primitives on C / 4 = 16 or 32 channels (MixedOp, model_search_interact.py:
// 39-74), where 8 lanes per pixel would leave 50-75 % of the threads without a channel vector: CVL = 4 / 2 there.
#include "view.cuh"
#include <stdlib.h>

namespace npp {

struct DwTileGeom {
  int Hi, Wi, Ho, Wo, C, N;   // staged-tensor extents, produced-tensor extents
  int stride, dil, off_h, off_w;  // staged row of tap r for produced row h: h*stride + off + r*dil
  int TW, TH, RPP, PPT;       // produced tile, rows per pass (pixel lanes / TW), pixels per thread
  int IH, IW;                 // staged tile extents
  int tiles_w, tiles_h, cblocks, ntiles, slots;
};

template <typename T>
__device__ __forceinline__ uint4 relu_packed(uint4 v);
template <>
__device__ __forceinline__ uint4 relu_packed<float>(uint4 v) {
  v.x = (v.x & 0x80000000u) ? 0u : v.x; v.y = (v.y & 0x80000000u) ? 0u : v.y;
  v.z = (v.z & 0x80000000u) ? 0u : v.z; v.w = (v.w & 0x80000000u) ? 0u : v.w;
  return v;
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t u) {
  if (u & 0x00008000u) u &= 0xffff0000u;
  if (u & 0x80000000u) u &= 0x0000ffffu;
  return u;
}
template <>
__device__ __forceinline__ uint4 relu_packed<__nv_bfloat16>(uint4 v) {
  v.x = relu_bf16x2(v.x); v.y = relu_bf16x2(v.y); v.z = relu_bf16x2(v.z); v.w = relu_bf16x2(v.w);
  return v;
}

// stage the input tile of (n, th, tw, channel block) into shared memory (zero outside the tensor = zero padding)
template <typename T, int CVL>
__device__ __forceinline__ void dw_stage(uint4* tile, const DView<const T>& X, const DwTileGeom& g, int n, int ih0,
                                         int iw0, int c0, int cvn, bool relu) {
  constexpr int V = Pack<T>::N;
  const int total = g.IH * g.IW * CVL;
  // four independent 16-byte loads in flight per thread before the first shared-memory store (the one-load-per-
  // iteration loop of the first version left the staging phase latency-bound)
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * 256) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < total) {
        const int cv = i % CVL;
        const int p = i / CVL;
        const int pw = p % g.IW, ph = p / g.IW;
        const int h = ih0 + ph, w = iw0 + pw;
        if (cv < cvn && h >= 0 && h < g.Hi && w >= 0 && w < g.Wi) v[u] = ldraw(X.at(n, h, w, c0 + cv * V));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i < total) tile[i] = relu ? relu_packed<T>(v[u]) : v[u];
    }
  }
}

template <typename T, bool BWD, int CVL>
__global__ void __launch_bounds__(256) dw_tile_kernel(const DView<const T> X, const float* __restrict__ wgt,
                                                      const DView<T> Y, const DView<const T> M, const DwTileGeom g,
                                                      int relu_in) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  extern __shared__ uint4 dw_tile_smem[];
  uint4* tile = dw_tile_smem;
  int b = blockIdx.x;
  const int cb = b % g.cblocks; b /= g.cblocks;
  const int tw = b % g.tiles_w; b /= g.tiles_w;
  const int th = b % g.tiles_h;
  const int n = b / g.tiles_h;
  const int c0 = cb * CVL * V;
  int cvn = (g.C - c0) / V;
  if (cvn > CVL) cvn = CVL;
  const int oh0 = th * g.TH, ow0 = tw * g.TW;
  dw_stage<T, CVL>(tile, X, g, n, oh0 * g.stride + g.off_h, ow0 * g.stride + g.off_w, c0, cvn, !BWD && relu_in);
  __syncthreads();
  const int cv = threadIdx.x % CVL, lane = threadIdx.x / CVL;
  if (cv >= cvn) return;
  const int c = c0 + cv * V;
  float wr[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) wr[t][i] = __ldg(wgt + (c + i) * 9 + (BWD ? 8 - t : t));
  const int pw = lane % g.TW, pr = lane / g.TW;
  const int ow = ow0 + pw;
  if (ow >= g.Wo) return;
  // dgrad: the ReLU mask (x at the produced pixel) of the NEXT row is fetched while the current row is computed —
  // ncu showed the row loop stalled on this global load (long scoreboard) once per row
  uint4 mnext = make_uint4(0u, 0u, 0u, 0u);
  if (BWD && relu_in && pr < g.TH && oh0 + pr < g.Ho) mnext = ldraw(M.at(n, oh0 + pr, ow, c));
  for (int k = 0; k < g.PPT; ++k) {
    const int ph = pr + k * g.RPP;
    const int oh = oh0 + ph;
    if (ph >= g.TH || oh >= g.Ho) break;
    const uint4 mcur = mnext;
    if (BWD && relu_in && k + 1 < g.PPT && ph + g.RPP < g.TH && oh + g.RPP < g.Ho)
      mnext = ldraw(M.at(n, oh + g.RPP, ow, c));
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    const uint4* base = tile + ((ph * g.stride) * g.IW + pw * g.stride) * CVL + cv;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float v[V];
        Pack<T>::unpack(base[(r * g.dil * g.IW + q * g.dil) * CVL], v);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wr[r * 3 + q][i], acc[i]);
      }
    if (BWD && relu_in) {
      float xv[V];
      Pack<T>::unpack(mcur, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = xv[i] > 0.f ? acc[i] : 0.f;
    }
    Pack<T>::store(Y.at(n, oh, ow, c), acc);
  }
}

template <typename T, int CVL>
__global__ void __launch_bounds__(256) dw_tile_wgrad_kernel(const DView<const T> X, const DView<const T> DY,
                                                            float* __restrict__ dw, const DwTileGeom g, int relu_in) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  extern __shared__ uint4 dw_tile_smem[];
  uint4* tile = dw_tile_smem;
  const int cb = blockIdx.x % g.cblocks;
  const int slot = blockIdx.x / g.cblocks;
  const int c0 = cb * CVL * V;
  int cvn = (g.C - c0) / V;
  if (cvn > CVL) cvn = CVL;
  const int cv = threadIdx.x % CVL, lane = threadIdx.x / CVL;
  const int c = c0 + cv * V;
  const int pw = lane % g.TW, pr = lane / g.TW;
  float acc[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[t][i] = 0.f;
  for (int tix = slot; tix < g.ntiles; tix += g.slots) {
    int b = tix;
    const int tw = b % g.tiles_w; b /= g.tiles_w;
    const int th = b % g.tiles_h;
    const int n = b / g.tiles_h;
    const int oh0 = th * g.TH, ow0 = tw * g.TW;
    __syncthreads();  // the previous tile's readers are done
    dw_stage<T, CVL>(tile, X, g, n, oh0 * g.stride + g.off_h, ow0 * g.stride + g.off_w, c0, cvn, relu_in != 0);
    __syncthreads();
    const int ow = ow0 + pw;
    if (cv < cvn && ow < g.Wo) {
      // dY of the next row is in flight while the nine taps of the current row are accumulated
      uint4 dnext = make_uint4(0u, 0u, 0u, 0u);
      if (pr < g.TH && oh0 + pr < g.Ho) dnext = ldraw(DY.at(n, oh0 + pr, ow, c));
      for (int k = 0; k < g.PPT; ++k) {
        const int ph = pr + k * g.RPP;
        const int oh = oh0 + ph;
        if (ph >= g.TH || oh >= g.Ho) break;
        const uint4 dcur = dnext;
        if (k + 1 < g.PPT && ph + g.RPP < g.TH && oh + g.RPP < g.Ho) dnext = ldraw(DY.at(n, oh + g.RPP, ow, c));
        float d[V];
        Pack<T>::unpack(dcur, d);
        const uint4* base = tile + ((ph * g.stride) * g.IW + pw * g.stride) * CVL + cv;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float v[V];
            Pack<T>::unpack(base[(r * g.dil * g.IW + q * g.dil) * CVL], v);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[r * 3 + q][i] = fmaf(d[i], v[i], acc[r * 3 + q][i]);
          }
      }
    }
  }
  // fold the 32 pixel lanes of every channel vector, one atomic per (CTA, channel, tap)
  float* red = reinterpret_cast<float*>(tile);  // 256 * V floats <= the staged tile
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = acc[t][i];
    __syncthreads();
    if (threadIdx.x < CVL * V) {
      const int cvj = threadIdx.x / V, ij = threadIdx.x % V;
      if (cvj < cvn) {
        float s = 0.f;
#pragma unroll 8
        for (int l = 0; l < 256 / CVL; ++l) s += red[((l * CVL) + cvj) * V + ij];
        atomicAdd(dw + (int64_t)(c0 + threadIdx.x) * 9 + t, s);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pipelined variants: persistent CTAs, TWO staged tiles in shared memory, cp.async staging.
// ncu on the kernels above (128 channels @ 96x96 x 32): 23 % of the warp slots active, top stalls = the shared-memory
// store waiting for the staging loads and the CTA barrier behind it — every CTA runs "load tile, wait, compute" back
// to back and two or three resident CTAs do not hide a DRAM round trip.  Here the copy of tile i+1 (cp.async, zero
// fill outside the tensor, no registers) is in flight while tile i is computed; the leading ReLU is applied in place
// by the thread that issued the copy.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int CVL>
__device__ __forceinline__ void dw_stage_async(uint4* tile, const DView<const T>& X, const DwTileGeom& g, int n, int ih0,
                                               int iw0, int c0, int cvn) {
  constexpr int V = Pack<T>::N;
  const int total = g.IH * g.IW * CVL;
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  for (int i = threadIdx.x; i < total; i += 256) {
    const int cv = i % CVL;
    const int p = i / CVL;
    const int pw = p % g.IW, ph = p / g.IW;
    const int h = ih0 + ph, w = iw0 + pw;
    const bool ok = cv < cvn && h >= 0 && h < g.Hi && w >= 0 && w < g.Wi;
    cp_async16_zfill(base + i * 16, ok ? static_cast<const void*>(X.at(n, h, w, c0 + cv * V)) : static_cast<const void*>(X.p), ok);
  }
}
// in-place ReLU of the elements this thread copied (visible to it after its own cp.async.wait_group)
template <typename T, int CVL>
__device__ __forceinline__ void dw_relu_own(uint4* tile, const DwTileGeom& g) {
  const int total = g.IH * g.IW * CVL;
  for (int i = threadIdx.x; i < total; i += 256) tile[i] = relu_packed<T>(tile[i]);
}

__device__ __forceinline__ void dw_tile_coords(const DwTileGeom& g, int tix, int& n, int& oh0, int& ow0) {
  const int tw = tix % g.tiles_w;
  tix /= g.tiles_w;
  const int th = tix % g.tiles_h;
  n = tix / g.tiles_h;
  oh0 = th * g.TH;
  ow0 = tw * g.TW;
}

template <typename T, bool BWD, int CVL>
__global__ void __launch_bounds__(256) dw_pipe_kernel(const DView<const T> X, const float* __restrict__ wgt,
                                                      const DView<T> Y, const DView<const T> M, const DwTileGeom g,
                                                      int relu_in, int tile_elems) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  extern __shared__ uint4 dw_tile_smem[];
  const int cb = blockIdx.x % g.cblocks;
  const int slot = blockIdx.x / g.cblocks;
  const int c0 = cb * CVL * V;
  int cvn = (g.C - c0) / V;
  if (cvn > CVL) cvn = CVL;
  const int cv = threadIdx.x % CVL, lane = threadIdx.x / CVL;
  const int c = c0 + cv * V;
  const int pw = lane % g.TW, pr = lane / g.TW;
  const bool cv_ok = cv < cvn;
  float wr[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) wr[t][i] = cv_ok ? __ldg(wgt + (c + i) * 9 + (BWD ? 8 - t : t)) : 0.f;
  const bool relu_stage = !BWD && relu_in;
  int it = 0;
  {
    int n, oh0, ow0;
    if (slot < g.ntiles) {
      dw_tile_coords(g, slot, n, oh0, ow0);
      dw_stage_async<T, CVL>(dw_tile_smem, X, g, n, oh0 * g.stride + g.off_h, ow0 * g.stride + g.off_w, c0, cvn);
    }
    cp_async_commit();
  }
  for (int tix = slot; tix < g.ntiles; tix += g.slots, ++it) {
    uint4* cur = dw_tile_smem + (it & 1) * tile_elems;
    uint4* nxt = dw_tile_smem + ((it + 1) & 1) * tile_elems;
    if (tix + g.slots < g.ntiles) {
      int n2, oh2, ow2;
      dw_tile_coords(g, tix + g.slots, n2, oh2, ow2);
      dw_stage_async<T, CVL>(nxt, X, g, n2, oh2 * g.stride + g.off_h, ow2 * g.stride + g.off_w, c0, cvn);
    }
    cp_async_commit();       // (possibly empty) group: keeps the group count uniform
    cp_async_wait<1>();      // everything but the newest group has landed: tile `tix` is in `cur`
    if (relu_stage) dw_relu_own<T, CVL>(cur, g);
    __syncthreads();
    int n, oh0, ow0;
    dw_tile_coords(g, tix, n, oh0, ow0);
    const int ow = ow0 + pw;
    if (cv_ok && ow < g.Wo) {
      uint4 mnext = make_uint4(0u, 0u, 0u, 0u);
      if (BWD && relu_in && pr < g.TH && oh0 + pr < g.Ho) mnext = ldraw(M.at(n, oh0 + pr, ow, c));
      for (int k = 0; k < g.PPT; ++k) {
        const int ph = pr + k * g.RPP;
        const int oh = oh0 + ph;
        if (ph >= g.TH || oh >= g.Ho) break;
        const uint4 mcur = mnext;
        if (BWD && relu_in && k + 1 < g.PPT && ph + g.RPP < g.TH && oh + g.RPP < g.Ho)
          mnext = ldraw(M.at(n, oh + g.RPP, ow, c));
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        const uint4* base = cur + ((ph * g.stride) * g.IW + pw * g.stride) * CVL + cv;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float v[V];
            Pack<T>::unpack(base[(r * g.dil * g.IW + q * g.dil) * CVL], v);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], wr[r * 3 + q][i], acc[i]);
          }
        if (BWD && relu_in) {
          float xv[V];
          Pack<T>::unpack(mcur, xv);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] = xv[i] > 0.f ? acc[i] : 0.f;
        }
        Pack<T>::store(Y.at(n, oh, ow, c), acc);
      }
    }
    __syncthreads();  // `cur` is the copy target of the next iteration
  }
  cp_async_wait<0>();
}

template <typename T, int CVL>
__global__ void __launch_bounds__(256) dw_pipe_wgrad_kernel(const DView<const T> X, const DView<const T> DY,
                                                            float* __restrict__ dw, const DwTileGeom g, int relu_in,
                                                            int tile_elems) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  extern __shared__ uint4 dw_tile_smem[];
  const int cb = blockIdx.x % g.cblocks;
  const int slot = blockIdx.x / g.cblocks;
  const int c0 = cb * CVL * V;
  int cvn = (g.C - c0) / V;
  if (cvn > CVL) cvn = CVL;
  const int cv = threadIdx.x % CVL, lane = threadIdx.x / CVL;
  const int c = c0 + cv * V;
  const int pw = lane % g.TW, pr = lane / g.TW;
  float acc[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[t][i] = 0.f;
  int it = 0;
  {
    int n, oh0, ow0;
    if (slot < g.ntiles) {
      dw_tile_coords(g, slot, n, oh0, ow0);
      dw_stage_async<T, CVL>(dw_tile_smem, X, g, n, oh0 * g.stride + g.off_h, ow0 * g.stride + g.off_w, c0, cvn);
    }
    cp_async_commit();
  }
  for (int tix = slot; tix < g.ntiles; tix += g.slots, ++it) {
    uint4* cur = dw_tile_smem + (it & 1) * tile_elems;
    uint4* nxt = dw_tile_smem + ((it + 1) & 1) * tile_elems;
    if (tix + g.slots < g.ntiles) {
      int n2, oh2, ow2;
      dw_tile_coords(g, tix + g.slots, n2, oh2, ow2);
      dw_stage_async<T, CVL>(nxt, X, g, n2, oh2 * g.stride + g.off_h, ow2 * g.stride + g.off_w, c0, cvn);
    }
    cp_async_commit();
    cp_async_wait<1>();
    if (relu_in) dw_relu_own<T, CVL>(cur, g);
    __syncthreads();
    int n, oh0, ow0;
    dw_tile_coords(g, tix, n, oh0, ow0);
    const int ow = ow0 + pw;
    if (cv < cvn && ow < g.Wo) {
      uint4 dnext = make_uint4(0u, 0u, 0u, 0u);
      if (pr < g.TH && oh0 + pr < g.Ho) dnext = ldraw(DY.at(n, oh0 + pr, ow, c));
      for (int k = 0; k < g.PPT; ++k) {
        const int ph = pr + k * g.RPP;
        const int oh = oh0 + ph;
        if (ph >= g.TH || oh >= g.Ho) break;
        const uint4 dcur = dnext;
        if (k + 1 < g.PPT && ph + g.RPP < g.TH && oh + g.RPP < g.Ho) dnext = ldraw(DY.at(n, oh + g.RPP, ow, c));
        float d[V];
        Pack<T>::unpack(dcur, d);
        const uint4* base = cur + ((ph * g.stride) * g.IW + pw * g.stride) * CVL + cv;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float v[V];
            Pack<T>::unpack(base[(r * g.dil * g.IW + q * g.dil) * CVL], v);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[r * 3 + q][i] = fmaf(d[i], v[i], acc[r * 3 + q][i]);
          }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  // fold the 32 pixel lanes of every channel vector, one atomic per (CTA, channel, tap)
  float* red = reinterpret_cast<float*>(dw_tile_smem);  // 256 * V floats <= one staged tile
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = acc[t][i];
    __syncthreads();
    if (threadIdx.x < CVL * V) {
      const int cvj = threadIdx.x / V, ij = threadIdx.x % V;
      if (cvj < cvn) {
        float s = 0.f;
#pragma unroll 8
        for (int l = 0; l < 256 / CVL; ++l) s += red[((l * CVL) + cvj) * V + ij];
        atomicAdd(dw + (int64_t)(c0 + threadIdx.x) * 9 + t, s);
      }
    }
  }
}

static int dw_pipe_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NPP_DW_PIPE");
    v = (e && *e) ? atoi(e) : 1;
  }
  return v;
}

// Channel-vector lanes per pixel for a layer of C channels (V channels per 16-byte vector)
static int dw_cvl(int C, int V) {
  const int cv = (C + V - 1) / V;
  return cv > 4 ? 8 : (cv > 2 ? 4 : 2);
}

// Tile geometry; returns false when the shape should stay on the gather kernels (stride-2 tiles that would not fit).
static bool dw_tile_geom(DwTileGeom& g, int N, int Hi, int Wi, int Ho, int Wo, int C, int V, int cvl, int stride, int dil,
                         int off_h, int off_w, size_t* smem) {
  g.N = N; g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo; g.C = C;
  g.stride = stride; g.dil = dil; g.off_h = off_h; g.off_w = off_w;
  const int lanes = 256 / cvl;  // pixel lanes of a block
  int best = 32;
  int64_t best_pad = -1;
  for (int tw = 32; tw >= 8; tw >>= 1) {
    const int64_t padded = cdiv64(Wo, tw) * tw;
    if (best_pad < 0 || padded < best_pad) { best_pad = padded; best = tw; }
  }
  g.TW = best;
  g.RPP = lanes / g.TW;
  int th = 8 * g.RPP;             // eight passes per tile: 256 / 512 / 1024 pixels for 8 / 4 / 2 vector lanes
  const int hround = (int)(cdiv64(Ho, g.RPP) * g.RPP);
  if (th > hround) th = hround;
  g.TH = th;
  g.PPT = g.TH / g.RPP;
  g.IH = (g.TH - 1) * stride + 2 * dil + 1;
  g.IW = (g.TW - 1) * stride + 2 * dil + 1;
  g.tiles_w = (int)cdiv64(Wo, g.TW);
  g.tiles_h = (int)cdiv64(Ho, g.TH);
  g.cblocks = (int)cdiv64(C, cvl * V);
  const int64_t ntiles = (int64_t)N * g.tiles_h * g.tiles_w;
  if (ntiles * g.cblocks > 0x7fffffff) return false;
  g.ntiles = (int)ntiles;
  g.slots = 1;
  *smem = (size_t)g.IH * g.IW * cvl * 16;
  if (*smem < (size_t)256 * V * 4) *smem = (size_t)256 * V * 4;
  return *smem <= 100 * 1024;
}

// opt-in to > 48 KB dynamic shared memory, once per kernel instantiation (K is a distinct type per instantiation
// only through TAG)
template <int TAG, typename K>
static int dw_set_smem(K kernel, const char* name) {
  static bool done = false;
  if (done) return NPP_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) { set_error(name, e); return NPP_E_CUDA; }
  done = true;
  return NPP_OK;
}

template <int TAG, typename K>
static int dw_set_smem2(K kernel, const char* name) {
  static bool done = false;
  if (done) return NPP_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  if (e != cudaSuccess) { set_error(name, e); return NPP_E_CUDA; }
  done = true;
  return NPP_OK;
}
// persistent grid of the pipelined kernels: two CTAs (2 x <= 112 KB) per SM, every CTA keeps one channel block
static void dw_pipe_slots(DwTileGeom& g) {
  int slots = (2 * sm_count()) / g.cblocks;
  if (slots < 1) slots = 1;
  if (slots > g.ntiles) slots = g.ntiles;
  g.slots = slots;
}

template <typename T, int CVL>
static int dw_fwd_launch(const npp_view4* x, const float* w, const npp_view4* y, DwTileGeom& g, size_t smem, int relu_in,
                         cudaStream_t st) {
  constexpr int TAG = sizeof(T) * 10 + CVL * 100;
  const auto X = dview<const T>(x);
  if (dw_pipe_enabled() && 2 * smem <= 112 * 1024) {
    int rc = dw_set_smem2<TAG + 3>(dw_pipe_kernel<T, false, CVL>, "cudaFuncSetAttribute(dw_pipe_fwd)");
    if (rc) return rc;
    dw_pipe_slots(g);
    NPP_LAUNCH((dw_pipe_kernel<T, false, CVL>), g.slots * g.cblocks, 256, 2 * smem, st, X, w, dview<T>(y), X, g, relu_in,
                                                                              (int)(smem / 16));
    NPP_CHECK_LAUNCH("dw_pipe_fwd");
    return NPP_OK;
  }
  int rc = dw_set_smem<TAG + 0>(dw_tile_kernel<T, false, CVL>, "cudaFuncSetAttribute(dw_tile_fwd)");
  if (rc) return rc;
  NPP_LAUNCH((dw_tile_kernel<T, false, CVL>), g.ntiles * g.cblocks, 256, smem, st, X, w, dview<T>(y), X, g, relu_in);
  NPP_CHECK_LAUNCH("dw_tile_fwd");
  return NPP_OK;
}

template <typename T, int CVL>
static int dw_dgrad_launch(const npp_view4* x, const float* w, const npp_view4* dy, const npp_view4* dx, DwTileGeom& g,
                           size_t smem, int relu_in, cudaStream_t st) {
  constexpr int TAG = sizeof(T) * 10 + CVL * 100;
  if (dw_pipe_enabled() && 2 * smem <= 112 * 1024) {
    int rc = dw_set_smem2<TAG + 4>(dw_pipe_kernel<T, true, CVL>, "cudaFuncSetAttribute(dw_pipe_dgrad)");
    if (rc) return rc;
    dw_pipe_slots(g);
    NPP_LAUNCH((dw_pipe_kernel<T, true, CVL>), g.slots * g.cblocks, 256, 2 * smem, st, dview<const T>(dy), w, dview<T>(dx),
                                                                             dview<const T>(x), g, relu_in, (int)(smem / 16));
    NPP_CHECK_LAUNCH("dw_pipe_dgrad");
    return NPP_OK;
  }
  int rc = dw_set_smem<TAG + 1>(dw_tile_kernel<T, true, CVL>, "cudaFuncSetAttribute(dw_tile_dgrad)");
  if (rc) return rc;
  NPP_LAUNCH((dw_tile_kernel<T, true, CVL>), g.ntiles * g.cblocks, 256, smem, st, dview<const T>(dy), w, dview<T>(dx),
                                                                         dview<const T>(x), g, relu_in);
  NPP_CHECK_LAUNCH("dw_tile_dgrad");
  return NPP_OK;
}

template <typename T, int CVL>
static int dw_wgrad_launch(const npp_view4* x, const npp_view4* dy, float* dw, DwTileGeom& g, size_t smem, int relu_in,
                           cudaStream_t st) {
  constexpr int TAG = sizeof(T) * 10 + CVL * 100;
  if (dw_pipe_enabled() && 2 * smem <= 112 * 1024) {
    int rc = dw_set_smem2<TAG + 5>(dw_pipe_wgrad_kernel<T, CVL>, "cudaFuncSetAttribute(dw_pipe_wgrad)");
    if (rc) return rc;
    dw_pipe_slots(g);
    NPP_LAUNCH((dw_pipe_wgrad_kernel<T, CVL>), g.slots * g.cblocks, 256, 2 * smem, st, dview<const T>(x), dview<const T>(dy), dw,
                                                                             g, relu_in, (int)(smem / 16));
    NPP_CHECK_LAUNCH("dw_pipe_wgrad");
    return NPP_OK;
  }
  int rc = dw_set_smem<TAG + 2>(dw_tile_wgrad_kernel<T, CVL>, "cudaFuncSetAttribute(dw_tile_wgrad)");
  if (rc) return rc;
  int slots = (2 * sm_count()) / g.cblocks;  // (4 per SM measured slower: twice the same-address atomics at the end)
  if (slots < 1) slots = 1;
  if (slots > g.ntiles) slots = g.ntiles;
  g.slots = slots;
  NPP_LAUNCH((dw_tile_wgrad_kernel<T, CVL>), slots * g.cblocks, 256, smem, st, dview<const T>(x), dview<const T>(dy), dw, g,
                                                                             relu_in);
  NPP_CHECK_LAUNCH("dw_tile_wgrad");
  return NPP_OK;
}

#define NPP_DW_CVL(cvl, expr8, expr4, expr2) ((cvl) == 8 ? (expr8) : ((cvl) == 4 ? (expr4) : (expr2)))

// returns NPP_E_UNSUPPORTED when the caller should use the gather kernels instead
template <typename T>
int dw_tile_fwd(const npp_view4* x, const float* w, const npp_view4* y, int stride, int pad, int dil, int relu_in,
                cudaStream_t st) {
  DwTileGeom g;
  size_t smem;
  const int cvl = dw_cvl(y->c, Pack<T>::N);
  if (!dw_tile_geom(g, y->n, x->h, x->w, y->h, y->w, y->c, Pack<T>::N, cvl, stride, dil, -pad, -pad, &smem))
    return NPP_E_UNSUPPORTED;
  return NPP_DW_CVL(cvl, (dw_fwd_launch<T, 8>(x, w, y, g, smem, relu_in, st)), (dw_fwd_launch<T, 4>(x, w, y, g, smem, relu_in, st)),
                    (dw_fwd_launch<T, 2>(x, w, y, g, smem, relu_in, st)));
}

template <typename T>
int dw_tile_dgrad(const npp_view4* x, const float* w, const npp_view4* dy, const npp_view4* dx, int stride, int pad,
                  int dil, int relu_in, cudaStream_t st) {
  if (stride != 1) return NPP_E_UNSUPPORTED;
  DwTileGeom g;
  size_t smem;
  const int cvl = dw_cvl(dx->c, Pack<T>::N);
  // dx[h] = sum_t dy[h + pad - t*dil] w[t] = sum_t' dy[h + pad - 2*dil + t'*dil] w[2 - t']
  if (!dw_tile_geom(g, dx->n, dy->h, dy->w, dx->h, dx->w, dx->c, Pack<T>::N, cvl, 1, dil, pad - 2 * dil, pad - 2 * dil, &smem))
    return NPP_E_UNSUPPORTED;
  return NPP_DW_CVL(cvl, (dw_dgrad_launch<T, 8>(x, w, dy, dx, g, smem, relu_in, st)),
                    (dw_dgrad_launch<T, 4>(x, w, dy, dx, g, smem, relu_in, st)),
                    (dw_dgrad_launch<T, 2>(x, w, dy, dx, g, smem, relu_in, st)));
}

template <typename T>
int dw_tile_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int stride, int pad, int dil, int relu_in,
                  cudaStream_t st) {
  DwTileGeom g;
  size_t smem;
  const int cvl = dw_cvl(dy->c, Pack<T>::N);
  if (!dw_tile_geom(g, dy->n, x->h, x->w, dy->h, dy->w, dy->c, Pack<T>::N, cvl, stride, dil, -pad, -pad, &smem))
    return NPP_E_UNSUPPORTED;
  return NPP_DW_CVL(cvl, (dw_wgrad_launch<T, 8>(x, dy, dw, g, smem, relu_in, st)),
                    (dw_wgrad_launch<T, 4>(x, dy, dw, g, smem, relu_in, st)),
                    (dw_wgrad_launch<T, 2>(x, dy, dw, g, smem, relu_in, st)));
}

template int dw_tile_fwd<float>(const npp_view4*, const float*, const npp_view4*, int, int, int, int, cudaStream_t);
template int dw_tile_fwd<__nv_bfloat16>(const npp_view4*, const float*, const npp_view4*, int, int, int, int,
                                        cudaStream_t);
template int dw_tile_dgrad<float>(const npp_view4*, const float*, const npp_view4*, const npp_view4*, int, int, int,
                                  int, cudaStream_t);
template int dw_tile_dgrad<__nv_bfloat16>(const npp_view4*, const float*, const npp_view4*, const npp_view4*, int, int,
                                          int, int, cudaStream_t);
template int dw_tile_wgrad<float>(const npp_view4*, const npp_view4*, float*, int, int, int, int, cudaStream_t);
template int dw_tile_wgrad<__nv_bfloat16>(const npp_view4*, const npp_view4*, float*, int, int, int, int, cudaStream_t);

}  // namespace npp
