// Fused losses of core/criterion.py on NCHW fp32 logits at head resolution.
//
//  Criterion_par (criterion.py:148-217): per stage, OhemCrossEntropy on the bilinearly upsampled
//  parsing logits (:54-72, upsample :181 align_corners=True) + weighted 2-class CE on the upsampled
//  edge logits (:161-166,194-197).  The reference materialises the [B,20,384,384] upsampled tensor,
//  its softmax, the per-pixel CE, a gather and a full sort() of every valid pixel.  Here each CTA
//  owns a 32x32 tile of label pixels, stages the (<= 12x12) patch of head logits it interpolates from
//  in shared memory and evaluates upsample+softmax+CE in registers; the OHEM threshold is a 4-pass
//  radix select on the fp32 bit patterns (the exact element sort() would put at index k); backward
//  re-derives the softmax and pulls the gradient through the bilinear taps into a shared-memory
//  accumulator tile that is flushed with one atomic per head element.
//  Criterion_pose (:74-145): sum of per-joint MSEs == one sum of squared differences.
#include "common.cuh"
#include "resample.cuh"

namespace npp {

constexpr int TILE = 32;       // label pixels per CTA side
constexpr int CMAX = 32;       // max classes held in registers
constexpr int kLossThreads = 256;

struct LossGeom {
  int n, c, h, w, lh, lw;
  Axis ah, aw;
  int tiles_x, tiles_y;
  int HR, WR;  // head patch extent staged per tile
};

__device__ __forceinline__ int axis_first(const Axis& a, int o) {
  int i0, i1; float l0, l1;
  bilinear_taps(a, o, i0, i1, l0, l1);
  return i0;
}

// mode 0: parsing (writes prob / loss per pixel, counts valid);  mode 1: edge (accumulates num/den)
template <int MODE>
__global__ void __launch_bounds__(kLossThreads)
ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, LossGeom g,
              const float* __restrict__ class_w, const int64_t* __restrict__ posneg, int ignore,
              float* __restrict__ prob, float* __restrict__ loss, unsigned long long* __restrict__ n_valid,
              float* __restrict__ out2) {
  pdl_wait();
  extern __shared__ float patch[];  // [c][HR][WR]
  __shared__ float red[2][kLossThreads / 32];
  __shared__ unsigned int cnt_sh;
  const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE, n = blockIdx.z;
  const int hy0 = axis_first(g.ah, ty0), hx0 = axis_first(g.aw, tx0);
  const int plane = g.HR * g.WR;
  for (int i = threadIdx.x; i < g.c * plane; i += blockDim.x) {
    const int c = i / plane, r = (i % plane) / g.WR, q = i % g.WR;
    const int hy = min(hy0 + r, g.h - 1), hx = min(hx0 + q, g.w - 1);
    patch[i] = logits[(((int64_t)n * g.c + c) * g.h + hy) * g.w + hx];
  }
  if (threadIdx.x == 0) cnt_sh = 0;
  __syncthreads();
  float w_edge[2] = {0.f, 0.f};
  if (MODE == 1) {
    const float pos = (float)posneg[0], neg = (float)posneg[1];
    w_edge[0] = pos / (pos + neg);  // weight of class 0 ("weight_neg", criterion.py:165)
    w_edge[1] = neg / (pos + neg);  // weight of class 1 ("weight_pos", :164)
  }
  float num = 0.f, den = 0.f;
  unsigned int valid = 0;
  const int lx = threadIdx.x % TILE;
  for (int ly = threadIdx.x / TILE; ly < TILE; ly += kLossThreads / TILE) {
    const int y = ty0 + ly, x = tx0 + lx;
    if (y >= g.lh || x >= g.lw) continue;
    const int64_t pidx = ((int64_t)n * g.lh + y) * g.lw + x;
    const int64_t t = target[pidx];
    if (t == ignore || t < 0 || t >= g.c) {
      if (MODE == 0) { prob[pidx] = 2.0f; loss[pidx] = 0.f; }
      continue;
    }
    int h0, h1, w0, w1; float lh0, lh1, lw0, lw1;
    bilinear_taps(g.ah, y, h0, h1, lh0, lh1);
    bilinear_taps(g.aw, x, w0, w1, lw0, lw1);
    const int o00 = (h0 - hy0) * g.WR + (w0 - hx0), o01 = (h0 - hy0) * g.WR + (w1 - hx0);
    const int o10 = (h1 - hy0) * g.WR + (w0 - hx0), o11 = (h1 - hy0) * g.WR + (w1 - hx0);
    float v[CMAX];
    float mx = -3.4e38f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) {
        const float* p = patch + c * plane;
        v[c] = lh0 * (lw0 * p[o00] + lw1 * p[o01]) + lh1 * (lw0 * p[o10] + lw1 * p[o11]);
        mx = fmaxf(mx, v[c]);
      }
    }
    float sum = 0.f, vt = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) {
        sum += expf(v[c] - mx);
        if (c == (int)t) vt = v[c];
      }
    }
    const float logp = vt - mx - logf(sum);
    if (MODE == 0) {
      prob[pidx] = expf(vt - mx) / sum;
      loss[pidx] = -class_w[t] * logp;
      ++valid;
    } else {
      const float wt = w_edge[t];
      num += -wt * logp;
      den += wt;
    }
  }
  if (MODE == 0) {
    valid = __reduce_add_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&cnt_sh, valid);
    __syncthreads();
    if (threadIdx.x == 0 && cnt_sh) atomicAdd(n_valid, (unsigned long long)cnt_sh);
  } else {
    num = warp_sum(num);
    den = warp_sum(den);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int i = 0; i < kLossThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; }
      atomicAdd(out2, a);
      atomicAdd(out2 + 1, b);
    }
  }
}

// d logits[n,c,h,w] += coef(p) * (softmax_c - [c==t]) pulled back through the bilinear taps.
//  MODE 0: coef = gscale/n_kept * w[t] for pixels with prob < thr;  MODE 1: coef = gscale * w[t] / den.
template <int MODE>
__global__ void __launch_bounds__(kLossThreads)
ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, LossGeom g,
              const float* __restrict__ class_w, const int64_t* __restrict__ posneg, int ignore,
              const float* __restrict__ prob, const float* __restrict__ sel, const float* __restrict__ gscale,
              float* __restrict__ dlogits) {
  pdl_wait();
  extern __shared__ float sm[];
  const int plane = g.HR * g.WR;
  float* patch = sm;                 // [c][HR][WR]
  float* acc = sm + g.c * plane;     // same shape
  const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE, n = blockIdx.z;
  const int hy0 = axis_first(g.ah, ty0), hx0 = axis_first(g.aw, tx0);
  for (int i = threadIdx.x; i < g.c * plane; i += blockDim.x) {
    const int c = i / plane, r = (i % plane) / g.WR, q = i % g.WR;
    const int hy = min(hy0 + r, g.h - 1), hx = min(hx0 + q, g.w - 1);
    patch[i] = logits[(((int64_t)n * g.c + c) * g.h + hy) * g.w + hx];
    acc[i] = 0.f;
  }
  __syncthreads();
  float w_edge[2] = {0.f, 0.f};
  float base;
  float thr = 0.f;
  if (MODE == 1) {
    const float pos = (float)posneg[0], neg = (float)posneg[1];
    w_edge[0] = pos / (pos + neg);
    w_edge[1] = neg / (pos + neg);
    base = gscale[0] / sel[1];           // sel = out2 {num, den}
  } else {
    base = sel[1] > 0.f ? gscale[0] / sel[1] : 0.f;  // sel = out3 {sum, n_kept, thr}
    thr = sel[2];
  }
  const int lx = threadIdx.x % TILE;
  for (int ly = threadIdx.x / TILE; ly < TILE; ly += kLossThreads / TILE) {
    const int y = ty0 + ly, x = tx0 + lx;
    if (y >= g.lh || x >= g.lw) continue;
    const int64_t pidx = ((int64_t)n * g.lh + y) * g.lw + x;
    const int64_t t = target[pidx];
    if (t == ignore || t < 0 || t >= g.c) continue;
    if (MODE == 0 && !(prob[pidx] < thr)) continue;
    int h0, h1, w0, w1; float lh0, lh1, lw0, lw1;
    bilinear_taps(g.ah, y, h0, h1, lh0, lh1);
    bilinear_taps(g.aw, x, w0, w1, lw0, lw1);
    const int o00 = (h0 - hy0) * g.WR + (w0 - hx0), o01 = (h0 - hy0) * g.WR + (w1 - hx0);
    const int o10 = (h1 - hy0) * g.WR + (w0 - hx0), o11 = (h1 - hy0) * g.WR + (w1 - hx0);
    float v[CMAX];
    float mx = -3.4e38f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) {
        const float* p = patch + c * plane;
        v[c] = lh0 * (lw0 * p[o00] + lw1 * p[o01]) + lh1 * (lw0 * p[o10] + lw1 * p[o11]);
        mx = fmaxf(mx, v[c]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) { v[c] = expf(v[c] - mx); sum += v[c]; }
    }
    const float coef = base * (MODE == 0 ? class_w[t] : w_edge[t]);
    const float inv = 1.f / sum;
    const float f00 = lh0 * lw0, f01 = lh0 * lw1, f10 = lh1 * lw0, f11 = lh1 * lw1;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) {
        const float gl = coef * (v[c] * inv - (c == (int)t ? 1.f : 0.f));
        float* a = acc + c * plane;
        atomicAdd(a + o00, gl * f00);
        atomicAdd(a + o01, gl * f01);
        atomicAdd(a + o10, gl * f10);
        atomicAdd(a + o11, gl * f11);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.c * plane; i += blockDim.x) {
    const float a = acc[i];
    if (a == 0.f) continue;
    const int c = i / plane, r = (i % plane) / g.WR, q = i % g.WR;
    const int hy = hy0 + r, hx = hx0 + q;
    if (hy < g.h && hx < g.w) atomicAdd(dlogits + (((int64_t)n * g.c + c) * g.h + hy) * g.w + hx, a);
  }
}

// Gather form of the same backward (default; NPP_CE_BWD_ATOMIC=1 selects the kernel above).  ncu on ce_bwd_kernel:
// 1.47 ms per parsing launch, 80 shared-memory atomics per label pixel, 4-8-way conflicted because neighbouring
// pixels share their source taps.  Pulling a gradient back through a bilinear up-sampling is separable, so the tile
// is processed in chunks of 8 label rows (one per warp, as above) without a single atomic:
//   1. every thread computes its pixel's gradient w.r.t. the up-sampled logits, G[row][class][x]  (plain stores,
//      consecutive lanes -> consecutive words);
//   2. R[row][class][w] = sum_x weight_x(x -> w) * G[row][class][x]   over the <= 9 label columns that touch head
//      column w (their range per w is tabulated once per tile);
//   3. acc[class][h][w] += sum_row weight_y(row -> h) * R[row][class][w]   for the <= 4 head rows the chunk touches —
//      every (class, h, w) has one owner thread per chunk, so plain read-modify-write.
// The tile's accumulator is flushed with one global atomic per head element as before.  Summation order inside a tile
// is fixed (deterministic up to the cross-tile flush).
template <int MODE>
__global__ void __launch_bounds__(kLossThreads)
ce_bwd_gather_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, LossGeom g,
                     const float* __restrict__ class_w, const int64_t* __restrict__ posneg, int ignore,
                     const float* __restrict__ prob, const float* __restrict__ sel, const float* __restrict__ gscale,
                     float* __restrict__ dlogits) {
  pdl_wait();
  constexpr int ROWS = kLossThreads / TILE;   // 8 label rows per chunk
  constexpr int GP = TILE + 1;                // padded row of G
  extern __shared__ float sm[];
  const int plane = g.HR * g.WR;
  float* patch = sm;                          // [c][HR][WR]
  float* acc = patch + g.c * plane;           // [c][HR][WR]
  float* G = acc + g.c * plane;               // [ROWS][c][GP]
  float* R = G + ROWS * g.c * GP;             // [ROWS][c][WR]
  __shared__ int xw0[TILE], xw1[TILE], yh0[TILE], yh1[TILE], xs[TILE], xe[TILE];
  __shared__ float xl0[TILE], xl1[TILE], yl0[TILE], yl1[TILE];
  const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE, n = blockIdx.z;
  const int hy0 = axis_first(g.ah, ty0), hx0 = axis_first(g.aw, tx0);
  for (int i = threadIdx.x; i < g.c * plane; i += blockDim.x) {
    const int c = i / plane, r = (i % plane) / g.WR, q = i % g.WR;
    const int hy = min(hy0 + r, g.h - 1), hx = min(hx0 + q, g.w - 1);
    patch[i] = logits[(((int64_t)n * g.c + c) * g.h + hy) * g.w + hx];
    acc[i] = 0.f;
  }
  if (threadIdx.x < TILE) {
    int i0, i1; float l0, l1;
    bilinear_taps(g.aw, min(tx0 + (int)threadIdx.x, g.lw - 1), i0, i1, l0, l1);
    xw0[threadIdx.x] = i0 - hx0; xw1[threadIdx.x] = i1 - hx0; xl0[threadIdx.x] = l0; xl1[threadIdx.x] = l1;
  } else if (threadIdx.x < 2 * TILE) {
    const int t = threadIdx.x - TILE;
    int i0, i1; float l0, l1;
    bilinear_taps(g.ah, min(ty0 + t, g.lh - 1), i0, i1, l0, l1);
    yh0[t] = i0 - hy0; yh1[t] = i1 - hy0; yl0[t] = l0; yl1[t] = l1;
  }
  __syncthreads();
  if (threadIdx.x < g.WR) {   // label columns whose taps touch head column w: contiguous, taps are monotone in x
    const int w = threadIdx.x;
    int a = TILE, b = 0;
    for (int x = 0; x < TILE; ++x)
      if (xw0[x] == w || xw1[x] == w) { a = min(a, x); b = max(b, x + 1); }
    xs[w] = a; xe[w] = b;
  }
  float w_edge[2] = {0.f, 0.f};
  float base;
  float thr = 0.f;
  if (MODE == 1) {
    const float pos = (float)posneg[0], neg = (float)posneg[1];
    w_edge[0] = pos / (pos + neg);
    w_edge[1] = neg / (pos + neg);
    base = gscale[0] / sel[1];           // sel = out2 {num, den}
  } else {
    base = sel[1] > 0.f ? gscale[0] / sel[1] : 0.f;  // sel = out3 {sum, n_kept, thr}
    thr = sel[2];
  }
  const int lx = threadIdx.x % TILE, wr = threadIdx.x / TILE;
  for (int chunk = 0; chunk < TILE / ROWS; ++chunk) {
    const int ly = chunk * ROWS + wr;
    // ---- 1. per-pixel gradient w.r.t. the up-sampled logits
    {
      const int y = ty0 + ly, x = tx0 + lx;
      float* grow = G + (wr * g.c) * GP + lx;
      bool live = y < g.lh && x < g.lw;
      int64_t t = 0;
      if (live) {
        const int64_t pidx = ((int64_t)n * g.lh + y) * g.lw + x;
        t = target[pidx];
        live = !(t == ignore || t < 0 || t >= g.c);
        if (live && MODE == 0 && !(prob[pidx] < thr)) live = false;
      }
      if (!live) {
        for (int c = 0; c < g.c; ++c) grow[c * GP] = 0.f;
      } else {
        const float lh0 = yl0[ly], lh1 = yl1[ly], lw0 = xl0[lx], lw1 = xl1[lx];
        const int o00 = yh0[ly] * g.WR + xw0[lx], o01 = yh0[ly] * g.WR + xw1[lx];
        const int o10 = yh1[ly] * g.WR + xw0[lx], o11 = yh1[ly] * g.WR + xw1[lx];
        float v[CMAX];
        float mx = -3.4e38f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
          if (c < g.c) {
            const float* p = patch + c * plane;
            v[c] = lh0 * (lw0 * p[o00] + lw1 * p[o01]) + lh1 * (lw0 * p[o10] + lw1 * p[o11]);
            mx = fmaxf(mx, v[c]);
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
          if (c < g.c) { v[c] = expf(v[c] - mx); sum += v[c]; }
        }
        const float coef = base * (MODE == 0 ? class_w[t] : w_edge[t]);
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
          if (c < g.c) grow[c * GP] = coef * (v[c] * inv - (c == (int)t ? 1.f : 0.f));
        }
      }
    }
    __syncthreads();
    // ---- 2. along x
    for (int o = threadIdx.x; o < ROWS * g.c * g.WR; o += kLossThreads) {
      const int w = o % g.WR, rc = o / g.WR;    // rc = row * c + class
      const float* gr = G + rc * GP;
      float a = 0.f;
      for (int x = xs[w]; x < xe[w]; ++x) {
        const float wt = (xw0[x] == w ? xl0[x] : 0.f) + (xw1[x] == w ? xl1[x] : 0.f);
        a = fmaf(wt, gr[x], a);
      }
      R[o] = a;
    }
    __syncthreads();
    // ---- 3. along y, into the tile accumulator
    const int hlo = yh0[chunk * ROWS], hhi = yh1[chunk * ROWS + ROWS - 1];
    const int nh = hhi - hlo + 1;
    for (int o = threadIdx.x; o < g.c * nh * g.WR; o += kLossThreads) {
      const int w = o % g.WR, hh = hlo + (o / g.WR) % nh, c = o / (g.WR * nh);
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int ry = chunk * ROWS + r;
        const float wt = (yh0[ry] == hh ? yl0[ry] : 0.f) + (yh1[ry] == hh ? yl1[ry] : 0.f);
        a = fmaf(wt, R[(r * g.c + c) * g.WR + w], a);
      }
      acc[c * plane + hh * g.WR + w] += a;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < g.c * plane; i += blockDim.x) {
    const float a = acc[i];
    if (a == 0.f) continue;
    const int c = i / plane, r = (i % plane) / g.WR, q = i % g.WR;
    const int hy = hy0 + r, hx = hx0 + q;
    if (hy < g.h && hx < g.w) atomicAdd(dlogits + (((int64_t)n * g.c + c) * g.h + hy) * g.w + hx, a);
  }
}

static size_t ce_bwd_gather_smem(const LossGeom& g) {
  return ((size_t)2 * g.c * g.HR * g.WR + (size_t)(kLossThreads / TILE) * g.c * (TILE + 1 + g.WR)) * sizeof(float);
}

static bool ce_bwd_atomic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NPP_CE_BWD_ATOMIC");
    v = (e && *e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

// Per-pixel form of the cross-entropy backward (round-2 candidate, criterion._state "ce_bwd_sep"): the kernel above
// pulls every label pixel's gradient through the four bilinear taps of every class into a shared-memory accumulator
// with atomics — 80 shared-memory atomics per pixel, 4-8-way conflicted because neighbouring pixels share their
// source taps (ncu: 1.47 ms per parsing launch, 30 % SM throughput, 24 % warp slots).  Pulling a gradient back
// through a bilinear up-sampling IS the bilinear backward, for which a separable kernel exists
// (npp_bilinear_bwd_sep), so this kernel only writes G[n, y, x, c] = d loss / d up-sampled-logit (fp32 NHWC, CQ
// channels = classes rounded up to a multiple of 4, zero for ignored / unselected pixels); the caller runs the
// separable bilinear backward on G and converts the NHWC result to the NCHW logit gradient.
template <int MODE>
__global__ void __launch_bounds__(kLossThreads)
ce_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, LossGeom g,
               const float* __restrict__ class_w, const int64_t* __restrict__ posneg, int ignore,
               const float* __restrict__ prob, const float* __restrict__ sel, const float* __restrict__ gscale,
               float* __restrict__ G, int CQ) {
  pdl_wait();
  extern __shared__ float patch[];  // [c][HR][WR]
  const int plane = g.HR * g.WR;
  const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE, n = blockIdx.z;
  const int hy0 = axis_first(g.ah, ty0), hx0 = axis_first(g.aw, tx0);
  for (int i = threadIdx.x; i < g.c * plane; i += blockDim.x) {
    const int c = i / plane, r = (i % plane) / g.WR, q = i % g.WR;
    const int hy = min(hy0 + r, g.h - 1), hx = min(hx0 + q, g.w - 1);
    patch[i] = logits[(((int64_t)n * g.c + c) * g.h + hy) * g.w + hx];
  }
  __syncthreads();
  float w_edge[2] = {0.f, 0.f};
  float base;
  float thr = 0.f;
  if (MODE == 1) {
    const float pos = (float)posneg[0], neg = (float)posneg[1];
    w_edge[0] = pos / (pos + neg);
    w_edge[1] = neg / (pos + neg);
    base = gscale[0] / sel[1];
  } else {
    base = sel[1] > 0.f ? gscale[0] / sel[1] : 0.f;
    thr = sel[2];
  }
  const int lx = threadIdx.x % TILE;
  for (int ly = threadIdx.x / TILE; ly < TILE; ly += kLossThreads / TILE) {
    const int y = ty0 + ly, x = tx0 + lx;
    if (y >= g.lh || x >= g.lw) continue;
    const int64_t pidx = ((int64_t)n * g.lh + y) * g.lw + x;
    float4* gp = reinterpret_cast<float4*>(G + pidx * CQ);
    const int64_t t = target[pidx];
    bool live = !(t == ignore || t < 0 || t >= g.c);
    if (live && MODE == 0 && !(prob[pidx] < thr)) live = false;
    if (!live) {
      for (int k = 0; k < CQ / 4; ++k) gp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    int h0, h1, w0, w1; float lh0, lh1, lw0, lw1;
    bilinear_taps(g.ah, y, h0, h1, lh0, lh1);
    bilinear_taps(g.aw, x, w0, w1, lw0, lw1);
    const int o00 = (h0 - hy0) * g.WR + (w0 - hx0), o01 = (h0 - hy0) * g.WR + (w1 - hx0);
    const int o10 = (h1 - hy0) * g.WR + (w0 - hx0), o11 = (h1 - hy0) * g.WR + (w1 - hx0);
    float v[CMAX];
    float mx = -3.4e38f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) {
        const float* p = patch + c * plane;
        v[c] = lh0 * (lw0 * p[o00] + lw1 * p[o01]) + lh1 * (lw0 * p[o10] + lw1 * p[o11]);
        mx = fmaxf(mx, v[c]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < g.c) { v[c] = expf(v[c] - mx); sum += v[c]; }
    }
    const float coef = base * (MODE == 0 ? class_w[t] : w_edge[t]);
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) v[c] = c < g.c ? coef * (v[c] * inv - (c == (int)t ? 1.f : 0.f)) : 0.f;
#pragma unroll
    for (int k = 0; k < CMAX / 4; ++k)
      if (4 * k < CQ) gp[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
}

template <typename K>
static int ensure_smem(K kernel, size_t bytes) {
  if (bytes > 200 * 1024) return NPP_E_UNSUPPORTED;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(ce kernel)", e); return NPP_E_CUDA; }
  }
  return NPP_OK;
}

static int make_geom(LossGeom* g, int n, int c, int h, int w, int lh, int lw, int align) {
  if (n <= 0 || c <= 0 || c > CMAX || h <= 0 || w <= 0 || lh <= 0 || lw <= 0) return NPP_E_INVALID;
  if (lh < h || lw < w) return NPP_E_UNSUPPORTED;  // the criterion only ever upsamples
  g->n = n; g->c = c; g->h = h; g->w = w; g->lh = lh; g->lw = lw;
  g->ah = make_axis(h, lh, align, 0.0);
  g->aw = make_axis(w, lw, align, 0.0);
  g->tiles_x = (lw + TILE - 1) / TILE;
  g->tiles_y = (lh + TILE - 1) / TILE;
  g->HR = (int)((TILE - 1) * g->ah.scale) + 3;
  g->WR = (int)((TILE - 1) * g->aw.scale) + 3;
  return NPP_OK;
}

// ------------------------------------------------------------------------------------------------
// OHEM threshold: k-th smallest of the fp32 probabilities via 4 x 8-bit radix select
// workspace layout (uint32): [0..255] histogram, [256] prefix, [257] k remaining, [258] pass index
// ------------------------------------------------------------------------------------------------
__global__ void ohem_init_kernel(unsigned int* ws, const unsigned long long* n_valid, int min_kept) {
  pdl_wait();
  if (threadIdx.x < 256) ws[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    const long long nv = (long long)*n_valid;
    long long k = nv - 1;
    if (k > min_kept) k = min_kept;
    ws[256] = 0;
    ws[257] = k < 0 ? 0xffffffffu : (unsigned int)k;
    ws[258] = 0;
  }
}

__global__ void __launch_bounds__(256) ohem_hist_kernel(const float* __restrict__ prob, int64_t npix, unsigned int* ws) {
  pdl_wait();
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int pass = ws[258];
  const unsigned int prefix = ws[256];
  const int shift = 24 - 8 * (int)pass;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned int key = __float_as_uint(prob[i]);
    const bool match = pass == 0 ? true : ((key >> (shift + 8)) == prefix);
    if (match) atomicAdd(&h[(key >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&ws[threadIdx.x], h[threadIdx.x]);
}

__global__ void ohem_scan_kernel(unsigned int* ws) {
  pdl_wait();  // 1 block of 256 threads
  __shared__ unsigned int h[256];
  h[threadIdx.x] = ws[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int k = ws[257];
    if (k != 0xffffffffu) {
      unsigned int cum = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (k < cum + h[b]) break;
        cum += h[b];
      }
      ws[256] = (ws[256] << 8) | (unsigned int)b;
      ws[257] = k - cum;
    }
    ws[258] += 1;
  }
  __syncthreads();
  ws[threadIdx.x] = 0;
}

// out3 = {sum of kept losses, number of kept pixels, threshold}.  The kept count is accumulated as a 64-bit INTEGER in
// the workspace (ws[260..261]; exact beyond 2^24 pixels and independent of the atomic order) and converted to fp32 once
// by the last block to finish (ticket in ws[262]); the workspace arrives zeroed.
__global__ void __launch_bounds__(256)
ohem_sum_kernel(const float* __restrict__ prob, const float* __restrict__ loss, int64_t npix, unsigned int* ws,
                float thres, float* __restrict__ out3) {
  pdl_wait();
  __shared__ float rs[8];
  __shared__ unsigned int rc[8];
  const float kth = ws[257] == 0xffffffffu ? 0.f : __uint_as_float(ws[256]);
  const float thr = fmaxf(kth, thres);
  float s = 0.f;
  unsigned int c = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    if (prob[i] < thr) { s += loss[i]; ++c; }
  }
  s = warp_sum(s);
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    unsigned long long b = 0;
    for (int i = 0; i < 8; ++i) { a += rs[i]; b += rc[i]; }
    unsigned long long* kept = reinterpret_cast<unsigned long long*>(ws + 260);
    atomicAdd(out3, a);
    if (b) atomicAdd(kept, b);
    if (blockIdx.x == 0) out3[2] = thr;
    __threadfence();
    if (atomicAdd(ws + 262, 1u) == gridDim.x - 1) {   // last block: every count has landed
      __threadfence();
      out3[1] = (float)*reinterpret_cast<volatile unsigned long long*>(kept);
    }
  }
}

__global__ void __launch_bounds__(256)
edge_count_kernel(const int64_t* __restrict__ t, int64_t npix, unsigned long long* __restrict__ posneg) {
  pdl_wait();
  unsigned int p = 0, q = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t[i];
    p += v == 1;
    q += v == 0;
  }
  p = __reduce_add_sync(0xffffffffu, p);
  q = __reduce_add_sync(0xffffffffu, q);
  if ((threadIdx.x & 31) == 0) {
    if (p) atomicAdd(posneg, (unsigned long long)p);
    if (q) atomicAdd(posneg + 1, (unsigned long long)q);
  }
}

__global__ void __launch_bounds__(256)
mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, const float* __restrict__ row_w,
               int64_t row_len, float* __restrict__ out) {
  pdl_wait();
  __shared__ float rs[8];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    if (row_w) d *= row_w[i / row_len];  // pred*w - gt*w (criterion.py:100-104 use_target_weight)
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += rs[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, const float* __restrict__ row_w,
               int64_t row_len, const float* __restrict__ gscale, float* __restrict__ da) {
  pdl_wait();
  const float g2 = 2.f * gscale[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float w = row_w ? row_w[i / row_len] : 1.f;
    da[i] = g2 * w * w * (a[i] - b[i]);
  }
}

static int flat_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_par_loss_pixels(const float* logits, int n, int c, int h, int w, const int64_t* target, int lh, int lw,
                        const float* class_w, int ignore_index, int align_corners, float* prob, float* loss,
                        int64_t* n_valid, npp_stream_t s) {
  if (!logits || !target || !class_w || !prob || !loss || !n_valid) return NPP_E_INVALID;
  LossGeom g;
  int rc = make_geom(&g, n, c, h, w, lh, lw, align_corners);
  if (rc) return rc;
  const size_t smem = (size_t)c * g.HR * g.WR * sizeof(float);
  if ((rc = ensure_smem(ce_fwd_kernel<0>, smem))) return rc;
  dim3 grid(g.tiles_x, g.tiles_y, n);
  NPP_LAUNCH((ce_fwd_kernel<0>), grid, kLossThreads, smem, as_stream(s), logits, target, g, class_w, nullptr, ignore_index, prob,
                                                               loss, reinterpret_cast<unsigned long long*>(n_valid),
                                                               nullptr);
  NPP_CHECK_LAUNCH("ce_fwd_kernel<par>");
  return NPP_OK;
}

int npp_ohem_select(const float* prob, const float* loss, int64_t npix, const int64_t* n_valid, int min_kept,
                    float thres, float* out3, void* workspace, npp_stream_t s) {
  if (!prob || !loss || !n_valid || !out3 || !workspace || npix <= 0) return NPP_E_INVALID;
  cudaStream_t st = as_stream(s);
  unsigned int* ws = static_cast<unsigned int*>(workspace);
  if (min_kept < 1) min_kept = 1;  // criterion.py:48 max(1, min_kept)
  NPP_LAUNCH((ohem_init_kernel), 1, 256, 0, st, ws, reinterpret_cast<const unsigned long long*>(n_valid), min_kept);
  const int grid = flat_grid(npix);
  for (int pass = 0; pass < 4; ++pass) {
    NPP_LAUNCH((ohem_hist_kernel), grid, 256, 0, st, prob, npix, ws);
    NPP_LAUNCH((ohem_scan_kernel), 1, 256, 0, st, ws);
  }
  NPP_LAUNCH((ohem_sum_kernel), grid, 256, 0, st, prob, loss, npix, ws, thres, out3);
  NPP_CHECK_LAUNCH("ohem_select");
  count_launch(9);  // init + 4 x (hist, scan); the sum kernel is counted by the check above
  return NPP_OK;
}

int npp_par_loss_bwd(const float* logits, int n, int c, int h, int w, const int64_t* target, int lh, int lw,
                     const float* class_w, int ignore_index, int align_corners, const float* prob, const float* out3,
                     const float* gscale, float* dlogits, npp_stream_t s) {
  if (!logits || !target || !class_w || !prob || !out3 || !gscale || !dlogits) return NPP_E_INVALID;
  LossGeom g;
  int rc = make_geom(&g, n, c, h, w, lh, lw, align_corners);
  if (rc) return rc;
  dim3 grid(g.tiles_x, g.tiles_y, n);
  if (!ce_bwd_atomic() && g.WR <= TILE) {
    const size_t smem2 = ce_bwd_gather_smem(g);
    if ((rc = ensure_smem(ce_bwd_gather_kernel<0>, smem2))) return rc;
    NPP_LAUNCH((ce_bwd_gather_kernel<0>), grid, kLossThreads, smem2, as_stream(s), logits, target, g, class_w, nullptr, ignore_index,
                                                                        prob, out3, gscale, dlogits);
    NPP_CHECK_LAUNCH("ce_bwd_gather_kernel<par>");
    return NPP_OK;
  }
  const size_t smem = 2 * (size_t)c * g.HR * g.WR * sizeof(float);
  if ((rc = ensure_smem(ce_bwd_kernel<0>, smem))) return rc;
  NPP_LAUNCH((ce_bwd_kernel<0>), grid, kLossThreads, smem, as_stream(s), logits, target, g, class_w, nullptr, ignore_index, prob,
                                                               out3, gscale, dlogits);
  NPP_CHECK_LAUNCH("ce_bwd_kernel<par>");
  return NPP_OK;
}

int npp_par_loss_grad_pixels(const float* logits, int n, int c, int h, int w, const int64_t* target, int lh, int lw,
                             const float* class_w, int ignore_index, int align_corners, const float* prob,
                             const float* out3, const float* gscale, float* G, int cq, npp_stream_t s) {
  if (!logits || !target || !class_w || !prob || !out3 || !gscale || !G || cq < c || (cq & 3) || cq > CMAX ||
      (reinterpret_cast<uintptr_t>(G) & 15))
    return NPP_E_INVALID;
  LossGeom g;
  int rc = make_geom(&g, n, c, h, w, lh, lw, align_corners);
  if (rc) return rc;
  const size_t smem = (size_t)c * g.HR * g.WR * sizeof(float);
  if ((rc = ensure_smem(ce_grad_kernel<0>, smem))) return rc;
  dim3 grid(g.tiles_x, g.tiles_y, n);
  NPP_LAUNCH((ce_grad_kernel<0>), grid, kLossThreads, smem, as_stream(s), logits, target, g, class_w, nullptr, ignore_index, prob,
                                                                out3, gscale, G, cq);
  NPP_CHECK_LAUNCH("ce_grad_kernel<par>");
  return NPP_OK;
}
int npp_edge_loss_grad_pixels(const float* logits, int n, int h, int w, const int64_t* target, int lh, int lw,
                              int ignore_index, int align_corners, const int64_t* posneg, const float* out2,
                              const float* gscale, float* G, int cq, npp_stream_t s) {
  if (!logits || !target || !posneg || !out2 || !gscale || !G || cq < 2 || (cq & 3) || cq > CMAX ||
      (reinterpret_cast<uintptr_t>(G) & 15))
    return NPP_E_INVALID;
  LossGeom g;
  int rc = make_geom(&g, n, 2, h, w, lh, lw, align_corners);
  if (rc) return rc;
  const size_t smem = (size_t)2 * g.HR * g.WR * sizeof(float);
  if ((rc = ensure_smem(ce_grad_kernel<1>, smem))) return rc;
  dim3 grid(g.tiles_x, g.tiles_y, n);
  NPP_LAUNCH((ce_grad_kernel<1>), grid, kLossThreads, smem, as_stream(s), logits, target, g, nullptr, posneg, ignore_index,
                                                                nullptr, out2, gscale, G, cq);
  NPP_CHECK_LAUNCH("ce_grad_kernel<edge>");
  return NPP_OK;
}

int npp_edge_count(const int64_t* target, int64_t npix, int64_t* posneg, npp_stream_t s) {
  if (!target || !posneg || npix <= 0) return NPP_E_INVALID;
  NPP_LAUNCH((edge_count_kernel), flat_grid(npix), 256, 0, as_stream(s), target, npix,
                                                              reinterpret_cast<unsigned long long*>(posneg));
  NPP_CHECK_LAUNCH("edge_count_kernel");
  return NPP_OK;
}

int npp_edge_loss_fwd(const float* logits, int n, int h, int w, const int64_t* target, int lh, int lw,
                      int ignore_index, int align_corners, const int64_t* posneg, float* out2, npp_stream_t s) {
  if (!logits || !target || !posneg || !out2) return NPP_E_INVALID;
  LossGeom g;
  int rc = make_geom(&g, n, 2, h, w, lh, lw, align_corners);
  if (rc) return rc;
  const size_t smem = (size_t)2 * g.HR * g.WR * sizeof(float);
  if ((rc = ensure_smem(ce_fwd_kernel<1>, smem))) return rc;
  dim3 grid(g.tiles_x, g.tiles_y, n);
  NPP_LAUNCH((ce_fwd_kernel<1>), grid, kLossThreads, smem, as_stream(s), logits, target, g, nullptr, posneg, ignore_index,
                                                               nullptr, nullptr, nullptr, out2);
  NPP_CHECK_LAUNCH("ce_fwd_kernel<edge>");
  return NPP_OK;
}

int npp_edge_loss_bwd(const float* logits, int n, int h, int w, const int64_t* target, int lh, int lw,
                      int ignore_index, int align_corners, const int64_t* posneg, const float* out2,
                      const float* gscale, float* dlogits, npp_stream_t s) {
  if (!logits || !target || !posneg || !out2 || !gscale || !dlogits) return NPP_E_INVALID;
  LossGeom g;
  int rc = make_geom(&g, n, 2, h, w, lh, lw, align_corners);
  if (rc) return rc;
  dim3 grid(g.tiles_x, g.tiles_y, n);
  if (!ce_bwd_atomic() && g.WR <= TILE) {
    const size_t smem2 = ce_bwd_gather_smem(g);
    if ((rc = ensure_smem(ce_bwd_gather_kernel<1>, smem2))) return rc;
    NPP_LAUNCH((ce_bwd_gather_kernel<1>), grid, kLossThreads, smem2, as_stream(s), logits, target, g, nullptr, posneg, ignore_index,
                                                                        nullptr, out2, gscale, dlogits);
    NPP_CHECK_LAUNCH("ce_bwd_gather_kernel<edge>");
    return NPP_OK;
  }
  const size_t smem = 2 * (size_t)2 * g.HR * g.WR * sizeof(float);
  if ((rc = ensure_smem(ce_bwd_kernel<1>, smem))) return rc;
  NPP_LAUNCH((ce_bwd_kernel<1>), grid, kLossThreads, smem, as_stream(s), logits, target, g, nullptr, posneg, ignore_index,
                                                               nullptr, out2, gscale, dlogits);
  NPP_CHECK_LAUNCH("ce_bwd_kernel<edge>");
  return NPP_OK;
}

int npp_mse_fwd(const float* pred, const float* target, int64_t n, const float* row_w, int64_t row_len, float* out,
                npp_stream_t s) {
  if (!pred || !target || !out || n <= 0 || (row_w && row_len <= 0)) return NPP_E_INVALID;
  NPP_LAUNCH((mse_fwd_kernel), flat_grid(n), 256, 0, as_stream(s), pred, target, n, row_w, row_len, out);
  NPP_CHECK_LAUNCH("mse_fwd_kernel");
  return NPP_OK;
}

int npp_mse_bwd(const float* pred, const float* target, int64_t n, const float* row_w, int64_t row_len,
                const float* gscale, float* dpred, npp_stream_t s) {
  if (!pred || !target || !gscale || !dpred || n <= 0 || (row_w && row_len <= 0)) return NPP_E_INVALID;
  NPP_LAUNCH((mse_bwd_kernel), flat_grid(n), 256, 0, as_stream(s), pred, target, n, row_w, row_len, gscale, dpred);
  NPP_CHECK_LAUNCH("mse_bwd_kernel");
  return NPP_OK;
}

}  // extern "C"
