// Fused cell-node kernels: every node of an NPPNet cell is  s = op_a(h_a) + op_b(h_b)  (models/model_augment.py:48-62,
// 90-106, 153-174) where most primitives end in a training-mode BatchNorm2d (models/operations.py:61,79,97,215,240)
// and most consumers start with nn.ReLU (operations.py:76,95,212,239).  The reference runs BN-apply, BN-apply, add
// and one ReLU per consumer as separate full-tensor passes (and their four backward passes); here one HBM pass
// produces the node from the two raw conv outputs:
//
//   fwd     : y = f_a(a) [+ f_b(b)],  f(x) = x*scale[c] + shift[c] (BatchNorm with batch statistics folded into
//             scale/shift) or identity;  written as y (raw) and/or relu(y) — optionally into channel slices of a
//             concat buffer (the views carry the strides), so torch.cat (model_augment.py:62) costs nothing.
//   bwd 1/2 : g = g_raw + [relu_out > 0] * g_relu  (written once), and per BatchNorm input the per-channel partial
//             sums  (sum g, sum g*xhat)  — per-block partials, no atomics (deterministic), reduced by
//             npp_reduce_partials; SyncBN all-reduces the 2C-float result (augment_lip_sync.py:191).
//   bwd 2/2 : d_in = gamma*invstd*(g - s1/N - xhat*s2/N) for each BatchNorm input, both from one read of g.
//
// All HBM-bound: 16-byte vectors, consecutive threads on consecutive channel vectors of a pixel, per-thread channel
// coefficients loaded once (a thread keeps its channel vector for its whole pixel loop), 4 (fwd) / 2 (bwd) pixels of
// independent loads in flight per thread.
#include "view.cuh"
#include <string.h>

namespace npp {

// Pixel addressing: dense views (sh == w*sw, sn == h*sh) index by flat pixel; others by (n,h,w).
template <typename T>
struct PView {
  T* p;
  int64_t sn, sh, sw;
  int H, W;
  bool dense;
  __device__ __forceinline__ T* at(int pix, int c0) const {
    if (dense) return p + (int64_t)pix * sw + c0;
    const int w = pix % W;
    const int t = pix / W;
    const int h = t % H;
    const int n = t / H;
    return p + n * sn + h * sh + w * sw + c0;
  }
};
template <typename T>
static inline PView<T> pview(const npp_view4* v) {
  PView<T> d;
  d.p = static_cast<T*>(v->ptr);
  d.sn = v->sn; d.sh = v->sh; d.sw = v->sw;
  d.H = v->h; d.W = v->w;
  d.dense = (v->sh == (int64_t)v->w * v->sw) && (v->sn == (int64_t)v->h * v->sh);
  return d;
}
template <typename T>
static inline PView<T> pview_null() {
  PView<T> d;
  d.p = nullptr; d.sn = d.sh = d.sw = 0; d.H = d.W = 1; d.dense = true;
  return d;
}

template <int V>
__device__ __forceinline__ void load_coef(const float* p, int c0, float (&v)[V]) {
#pragma unroll
  for (int i = 0; i < V; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c0 + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// BatchNorm "finalize" of one side done inside the node kernel: batch sums -> (scale, shift, mean, invstd), running
// statistics update (nn.BatchNorm2d training semantics, same double-precision arithmetic as bn_finalize_kernel in
// bn.cu).  Every block derives the coefficients of its own channel group in shared memory; the first pixel block
// also stores them (coef = [scale | shift | mean | invstd], 4C floats, read by the backward kernels) and updates
// the running statistics.  Saves one tiny, latency-bound launch per BatchNorm (432 per training step).
struct BnFin {
  const float* stats;  // [2][C] sums (nullptr = this side is not finalized here)
  const float* gamma;  // nullable
  const float* beta;   // nullable
  float* rmean;        // nullable
  float* rvar;         // nullable
  float* coef;         // [4][C]
  float momentum, eps;
  int c_run;           // channels of the running-statistics vectors (<= C: C may be padded)
};

template <typename T>
struct NodeFwdArgs {
  PView<const T> a, b;
  const float *sa, *ta, *sb, *tb;  // scale/shift (nullptr = identity)
  PView<T> yraw, yrelu;            // p == nullptr = not written
  BnFin fa, fb;
  double count;
  int npix, C;
  VecGeom g;
};

__device__ __forceinline__ void bn_fin_channel(const BnFin& f, double count, int C, int ch, bool writer, float& sc,
                                               float& sh) {
  const double mean = (double)f.stats[ch] / count;
  double var = (double)f.stats[C + ch] / count - mean * mean;  // biased variance, as ATen normalises with
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
  const float g = f.gamma ? f.gamma[ch] : 1.f;
  const float b = f.beta ? f.beta[ch] : 0.f;
  sc = g * invstd;
  sh = b - (float)mean * sc;
  if (writer) {
    f.coef[ch] = sc;
    f.coef[C + ch] = sh;
    f.coef[2 * C + ch] = (float)mean;
    f.coef[3 * C + ch] = invstd;
    if (f.rmean && ch < f.c_run) f.rmean[ch] = (1.f - f.momentum) * f.rmean[ch] + f.momentum * (float)mean;
    if (f.rvar && ch < f.c_run) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      f.rvar[ch] = (1.f - f.momentum) * f.rvar[ch] + f.momentum * (float)unbiased;
    }
  }
}

template <typename T, int U>
__global__ void __launch_bounds__(256, 2) node_fwd_kernel(const NodeFwdArgs<T> A) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  extern __shared__ float node_coef_smem[];  // [4][cvb * V] when a side is finalized here
  const int tcv = threadIdx.x % A.g.cvb;
  const int trow = threadIdx.x / A.g.cvb;
  const int mycv = blockIdx.y * A.g.cvb + tcv;
  const bool fin_a = A.fa.stats != nullptr, fin_b = A.fb.stats != nullptr;
  const int nch = A.g.cvb * V;
  if (fin_a || fin_b) {
    const int cbase = blockIdx.y * nch;
    for (int i = threadIdx.x; i < nch; i += 256) {
      const int ch = cbase + i;
      if (ch < A.C) {
        float sc, sh;
        if (fin_a) {
          bn_fin_channel(A.fa, A.count, A.C, ch, blockIdx.x == 0, sc, sh);
          node_coef_smem[i] = sc;
          node_coef_smem[nch + i] = sh;
        }
        if (fin_b) {
          bn_fin_channel(A.fb, A.count, A.C, ch, blockIdx.x == 0, sc, sh);
          node_coef_smem[2 * nch + i] = sc;
          node_coef_smem[3 * nch + i] = sh;
        }
      }
    }
    __syncthreads();
  }
  if (trow >= A.g.rows || mycv >= A.g.cv) return;
  const int c0 = mycv * V;
  float sa[V], ta[V], sb[V], tb[V];
  const bool affa = A.sa != nullptr || fin_a, affb = A.sb != nullptr || fin_b, hasb = A.b.p != nullptr;
  if (fin_a) {
    load_coef<V>(node_coef_smem, tcv * V, sa); load_coef<V>(node_coef_smem + nch, tcv * V, ta);
  } else if (affa) {
    load_coef<V>(A.sa, c0, sa); load_coef<V>(A.ta, c0, ta);
  }
  if (fin_b) {
    load_coef<V>(node_coef_smem + 2 * nch, tcv * V, sb); load_coef<V>(node_coef_smem + 3 * nch, tcv * V, tb);
  } else if (affb) {
    load_coef<V>(A.sb, c0, sb); load_coef<V>(A.tb, c0, tb);
  }
  const int step = gridDim.x * A.g.rows;
  for (int p0 = blockIdx.x * A.g.rows + trow; p0 < A.npix; p0 += U * step) {
    uint4 rx[U], ry[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * step;
      if (p < A.npix) {
        rx[u] = ldraw(A.a.at(p, c0));
        if (hasb) ry[u] = ldraw(A.b.at(p, c0));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * step;
      if (p >= A.npix) continue;
      float x[1][V], y[1][V];
      Pack<T>::unpack(rx[u], x[0]);
      if (hasb) Pack<T>::unpack(ry[u], y[0]);
      if (affa) {
#pragma unroll
        for (int i = 0; i < V; ++i) x[0][i] = fmaf(x[0][i], sa[i], ta[i]);
      }
      if (hasb) {
        if (affb) {
#pragma unroll
          for (int i = 0; i < V; ++i) y[0][i] = fmaf(y[0][i], sb[i], tb[i]);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) x[0][i] += y[0][i];
      }
      if (A.yraw.p) Pack<T>::store(A.yraw.at(p, c0), x[0]);
      if (A.yrelu.p) {
#pragma unroll
        for (int i = 0; i < V; ++i) x[0][i] = fmaxf(x[0][i], 0.f);
        Pack<T>::store(A.yrelu.at(p, c0), x[0]);
      }
    }
  }
}

static inline int stream_grid(int64_t npix, const VecGeom& g, int per_thread, int blocks_per_sm) {
  int64_t gx = (npix + (int64_t)g.rows * per_thread - 1) / ((int64_t)g.rows * per_thread);
  const int64_t cap = ((int64_t)sm_count() * blocks_per_sm + g.gy - 1) / g.gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return (int)gx;
}

template <typename T>
static int node_fwd_t(const npp_view4* a, const float* sa, const float* ta, const npp_view4* b, const float* sb,
                      const float* tb, const npp_view4* yraw, const npp_view4* yrelu, cudaStream_t st,
                      const BnFin* fa = nullptr, const BnFin* fb = nullptr, double count = 1.0) {
  constexpr int V = Pack<T>::N;
  NodeFwdArgs<T> A;
  memset(&A.fa, 0, sizeof A.fa);
  memset(&A.fb, 0, sizeof A.fb);
  if (fa) A.fa = *fa;
  if (fb) A.fb = *fb;
  A.count = count;
  A.C = a->c;
  A.a = pview<const T>(a);
  A.b = b ? pview<const T>(b) : pview_null<const T>();
  A.sa = sa; A.ta = ta; A.sb = sb; A.tb = tb;
  A.yraw = yraw ? pview<T>(yraw) : pview_null<T>();
  A.yrelu = yrelu ? pview<T>(yrelu) : pview_null<T>();
  const int64_t npix = (int64_t)a->n * a->h * a->w;
  if (npix > 0x3fffffff) return NPP_E_UNSUPPORTED;
  A.npix = (int)npix;
  A.g = vec_geom(a->c, V);
  dim3 grid((unsigned)stream_grid(npix, A.g, 4, 8), (unsigned)A.g.gy);
  const size_t smem = (A.fa.stats || A.fb.stats) ? (size_t)4 * A.g.cvb * V * sizeof(float) : 0;
  NPP_LAUNCH((node_fwd_kernel<T, 4>), grid, 256, smem, st, A);
  NPP_CHECK_LAUNCH("node_fwd_kernel");
  return NPP_OK;
}

// ------------------------------------------------------------------------------------------------ backward 1/2
struct AccSeg { float* ptr; int valid; };
struct AccSegs { AccSeg s[NPP_ACC_MAX]; };

template <typename T>
struct NodeBwdReduceArgs {
  float* sums;                    // atomic mode: [nq][C] zero-initialised accumulators (partials == nullptr)
  AccSegs acc;                    // atomic mode: per-row parameter-gradient slots (d beta / d gamma), ptr may be null
  PView<const T> graw, grelu, r;  // gradients of the raw / relu outputs, relu output (mask)
  int x0_relu, x1_relu;           // kind of the two extra gradient slots below: 0 = added to graw, 1 = to grelu
  PView<const T> graw2, grelu2;   // second gradient of either output (the concat-buffer slice handed down by the
                                  // consumers of the cell output), summed here instead of by a separate add kernel
  PView<const T> a, b;            // BatchNorm inputs (p == nullptr = that input has no BatchNorm)
  const float *mean_a, *invstd_a, *mean_b, *invstd_b;
  PView<T> gout;                  // combined gradient (p == nullptr = not written)
  float* partials;                // [gridDim.x][4][C]: (sum g, sum g*xhat_a, sum g, sum g*xhat_b) — 2 rows if one input
  int stripes;                    // atomic mode: block b adds into copy b % stripes of sums ([stripes][nq][C])
  int npix, C;
  VecGeom g;
};

// HAS2 = the second-gradient inputs exist.  A separate instantiation (with one pixel in flight per thread instead of
// two): compiled into the common kernel the two extra 16-byte loads per pixel pushed it past the 128-register cap of
// two resident blocks per SM and it spilled (ptxas: 168 bytes), slowing every node backward of the step.
template <typename T, int U, bool HAS2>
__global__ void __launch_bounds__(256, 2) node_bwd_reduce_kernel(const NodeBwdReduceArgs<T> A) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  __shared__ float red[256 * V];
  const int tcv = threadIdx.x % A.g.cvb;
  const int trow = threadIdx.x / A.g.cvb;
  const int mycv = tcv;  // gy == 1 (C <= 256 * V), checked on the host
  const bool active = trow < A.g.rows && mycv < A.g.cv;
  const int c0 = mycv * V;
  const bool has_raw = A.graw.p != nullptr, has_relu = A.grelu.p != nullptr;
  const bool has_raw2 = HAS2 && A.graw2.p != nullptr, has_relu2 = HAS2 && A.grelu2.p != nullptr;
  const bool has_a = A.a.p != nullptr, has_b = A.b.p != nullptr;
  float ma[V], mb[V];
  float s0[V], s1[V], s2[V];  // sum g, sum g*(x_a - mean_a), sum g*(x_b - mean_b); invstd is applied at the end
#pragma unroll
  for (int i = 0; i < V; ++i) { s0[i] = 0.f; s1[i] = 0.f; s2[i] = 0.f; ma[i] = mb[i] = 0.f; }
  if (active) {
    if (has_a) load_coef<V>(A.mean_a, c0, ma);
    if (has_b) load_coef<V>(A.mean_b, c0, mb);
    const int step = gridDim.x * A.g.rows;
    for (int p0 = blockIdx.x * A.g.rows + trow; p0 < A.npix; p0 += U * step) {
      uint4 qg[U], qgr[U], qr[U], qa[U], qb[U], qg2[HAS2 ? U : 1], qgr2[HAS2 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u * step;
        if (p < A.npix) {
          if (has_raw) qg[u] = ldraw(A.graw.at(p, c0));
          if (has_relu) {
            qgr[u] = ldraw(A.grelu.at(p, c0));
            qr[u] = ldraw(A.r.at(p, c0));
          }
          if (HAS2 && has_raw2) qg2[HAS2 ? u : 0] = ldraw(A.graw2.at(p, c0));
          if (HAS2 && has_relu2) qgr2[HAS2 ? u : 0] = ldraw(A.grelu2.at(p, c0));
          if (has_a) qa[u] = ldraw(A.a.at(p, c0));
          if (has_b) qb[u] = ldraw(A.b.at(p, c0));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u * step;
        if (p >= A.npix) continue;
        float g[1][V], gr[1][V], rr[1][V], xa[1][V], xb[1][V];
        if (has_raw) Pack<T>::unpack(qg[u], g[0]);
        if (has_relu) { Pack<T>::unpack(qgr[u], gr[0]); Pack<T>::unpack(qr[u], rr[0]); }
        if (has_a) Pack<T>::unpack(qa[u], xa[0]);
        if (has_b) Pack<T>::unpack(qb[u], xb[0]);
        // two extra gradient slots (more consumers of the same output: the concat route, further primitives of the
        // cell).  Each slot is of raw or relu kind; the host guarantees that the primary gradient of that kind exists.
        if (HAS2 && has_raw2) {
          float t2[V];
          Pack<T>::unpack(qg2[HAS2 ? u : 0], t2);
          if (A.x0_relu) {
#pragma unroll
            for (int i = 0; i < V; ++i) gr[0][i] += t2[i];
          } else {
#pragma unroll
            for (int i = 0; i < V; ++i) g[0][i] += t2[i];
          }
        }
        if (HAS2 && has_relu2) {
          float t2[V];
          Pack<T>::unpack(qgr2[HAS2 ? u : 0], t2);
          if (A.x1_relu) {
#pragma unroll
            for (int i = 0; i < V; ++i) gr[0][i] += t2[i];
          } else {
#pragma unroll
            for (int i = 0; i < V; ++i) g[0][i] += t2[i];
          }
        }
        if (has_relu) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float m = rr[0][i] > 0.f ? gr[0][i] : 0.f;
            g[0][i] = has_raw ? g[0][i] + m : m;
          }
        }
        if (has_relu || has_raw2 || has_relu2) {
          if (A.gout.p) {
            Pack<T>::store(A.gout.at(p, c0), g[0]);
            if (sizeof(T) == 2) {  // the apply pass reads the rounded value: accumulate what it will see
              float t[V];
#pragma unroll
              for (int i = 0; i < V; ++i) t[i] = to_f<T>(from_f<T>(g[0][i]));
#pragma unroll
              for (int i = 0; i < V; ++i) g[0][i] = t[i];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          s0[i] += g[0][i];
          if (has_a) s1[i] = fmaf(g[0][i], xa[0][i] - ma[i], s1[i]);
          if (has_b) s2[i] = fmaf(g[0][i], xb[0][i] - mb[i], s2[i]);
        }
      }
    }
  }
  if (A.partials == nullptr && A.sums == nullptr) return;
  // block reduction over the pixel rows of the block, then one coalesced row of partials per quantity — or, in
  // atomic mode, one fp32 atomic per (block, channel, quantity) straight into the totals and the parameters'
  // gradient slots (no partials buffer, no second kernel)
  const int nq = (has_a ? 2 : 0) + (has_b ? 2 : 0);
  const bool atomic = A.partials == nullptr;
  float* out = atomic ? A.sums + (int64_t)(blockIdx.x % A.stripes) * nq * A.C
                      : A.partials + (int64_t)blockIdx.x * nq * A.C;
  int q = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k == 1 && !has_a) continue;
    if (k == 2 && !has_b) continue;
    const float* src = k == 0 ? s0 : (k == 1 ? s1 : s2);
    __syncthreads();
#pragma unroll
    for (int v = 0; v < V; ++v) red[threadIdx.x * V + v] = src[v];
    __syncthreads();
    block_colsum(red, A.g.rows, A.g.cvb * V);
    if (active && trow == 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float s = red[tcv * V + v];
        const int ch = c0 + v;
        if (!atomic) {
          if (k == 0) {
            if (has_a) out[ch] = s;
            if (has_b) out[(has_a ? 2 : 0) * A.C + ch] = s;
          } else if (k == 1) {
            out[A.C + ch] = s * A.invstd_a[ch];
          } else {
            out[((has_a ? 2 : 0) + 1) * A.C + ch] = s * A.invstd_b[ch];
          }
        } else {
          // rows: [sum g (a), sum g*xhat_a, sum g (b), sum g*xhat_b] (only the BatchNorm sides that exist)
          const int rb = has_a ? 2 : 0;
          if (k == 0) {
            if (has_a) {
              atomicAdd(out + ch, s);
              if (A.acc.s[0].ptr && ch < A.acc.s[0].valid) atomicAdd(A.acc.s[0].ptr + ch, s);
            }
            if (has_b) {
              atomicAdd(out + rb * A.C + ch, s);
              if (A.acc.s[rb].ptr && ch < A.acc.s[rb].valid) atomicAdd(A.acc.s[rb].ptr + ch, s);
            }
          } else if (k == 1) {
            const float t = s * A.invstd_a[ch];
            atomicAdd(out + A.C + ch, t);
            if (A.acc.s[1].ptr && ch < A.acc.s[1].valid) atomicAdd(A.acc.s[1].ptr + ch, t);
          } else {
            const float t = s * A.invstd_b[ch];
            atomicAdd(out + (rb + 1) * A.C + ch, t);
            if (A.acc.s[rb + 1].ptr && ch < A.acc.s[rb + 1].valid) atomicAdd(A.acc.s[rb + 1].ptr + ch, t);
          }
        }
      }
    }
  }
  (void)q;
}

// out[j] = sum_i partials[i][j]   (rows x len), 32 columns x 8 row groups per block
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int rows, int len,
                                                              float* __restrict__ out) {
  pdl_wait();
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;
  float s = 0.f;
  if (col < len) {
    int r = rg;
    for (; r + 24 < rows; r += 32) {
      const float a = partials[(int64_t)r * len + col], b = partials[(int64_t)(r + 8) * len + col];
      const float c = partials[(int64_t)(r + 16) * len + col], d = partials[(int64_t)(r + 24) * len + col];
      s += (a + b) + (c + d);
    }
    for (; r < rows; r += 8) s += partials[(int64_t)r * len + col];
  }
  red[rg][threadIdx.x & 31] = s;
  __syncthreads();
  if (rg == 0 && col < len) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[col] = t;
  }
}

// reduce_partials + parameter-gradient accumulation: segment s = columns [s*seg_len, (s+1)*seg_len) of the folded
// row is additionally added into acc[s].ptr[0:valid] (BatchNorm d beta / d gamma written straight into the
// optimizer's flat gradient buffer instead of going through one autograd accumulation kernel per parameter).
__global__ void __launch_bounds__(256) reduce_partials_acc_kernel(const float* __restrict__ partials, int rows, int len,
                                                                  float* __restrict__ out, int seg_len, AccSegs acc) {
  pdl_wait();
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;
  float s = 0.f;
  if (col < len) {
    for (int r = rg; r < rows; r += 8) s += partials[(int64_t)r * len + col];
  }
  red[rg][threadIdx.x & 31] = s;
  __syncthreads();
  if (rg == 0 && col < len) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[col] = t;
    const int seg = col / seg_len, off = col - seg * seg_len;
    if (seg < NPP_ACC_MAX && acc.s[seg].ptr && off < acc.s[seg].valid) acc.s[seg].ptr[off] += t;
  }
}

static int node_bwd_blocks(int64_t npix, int C, int V) {
  const VecGeom g = vec_geom(C, V);
  // ~105 registers x 256 threads: two blocks are resident per SM, so 2 x SMs blocks are one full wave — and half as
  // many partial rows for reduce_partials to fold as the former 4 x SMs
  return stream_grid(npix, g, 8, 2);
}

template <typename T>
static int node_bwd_reduce_t(const npp_view4* graw, const npp_view4* grelu, const npp_view4* r, const npp_view4* a,
                             const float* mean_a, const float* invstd_a, const npp_view4* b, const float* mean_b,
                             const float* invstd_b, const npp_view4* gout, float* partials, float* sums,
                             const AccSegs* acc, cudaStream_t st, int stripes = 1, const npp_view4* graw2 = nullptr,
                             const npp_view4* grelu2 = nullptr, int x0_relu = 0, int x1_relu = 1) {
  constexpr int V = Pack<T>::N;
  const npp_view4* ref = graw ? graw : grelu;
  NodeBwdReduceArgs<T> A;
  A.sums = sums;
  if (acc) A.acc = *acc; else memset(&A.acc, 0, sizeof A.acc);
  A.graw = graw ? pview<const T>(graw) : pview_null<const T>();
  A.grelu = grelu ? pview<const T>(grelu) : pview_null<const T>();
  A.r = r ? pview<const T>(r) : pview_null<const T>();
  A.x0_relu = x0_relu; A.x1_relu = x1_relu;
  A.graw2 = graw2 ? pview<const T>(graw2) : pview_null<const T>();
  A.grelu2 = grelu2 ? pview<const T>(grelu2) : pview_null<const T>();
  A.a = a ? pview<const T>(a) : pview_null<const T>();
  A.b = b ? pview<const T>(b) : pview_null<const T>();
  A.mean_a = mean_a; A.invstd_a = invstd_a; A.mean_b = mean_b; A.invstd_b = invstd_b;
  A.gout = gout ? pview<T>(gout) : pview_null<T>();
  A.partials = partials;
  A.stripes = stripes < 1 ? 1 : stripes;
  const int64_t npix = (int64_t)ref->n * ref->h * ref->w;
  if (npix > 0x3fffffff) return NPP_E_UNSUPPORTED;
  A.npix = (int)npix;
  A.C = ref->c;
  A.g = vec_geom(ref->c, V);
  if (A.g.gy != 1) return NPP_E_UNSUPPORTED;
  dim3 grid((unsigned)node_bwd_blocks(npix, ref->c, V), 1);
  if (graw2 || grelu2)
    NPP_LAUNCH((node_bwd_reduce_kernel<T, 1, true>), grid, 256, 0, st, A);
  else
    NPP_LAUNCH((node_bwd_reduce_kernel<T, 2, false>), grid, 256, 0, st, A);
  NPP_CHECK_LAUNCH("node_bwd_reduce_kernel");
  return NPP_OK;
}

// ------------------------------------------------------------------------------------------------ backward 2/2
template <typename T>
struct NodeBwdApplyArgs {
  PView<const T> g, a, b;
  const float *gamma_a, *mean_a, *invstd_a, *sums_a;  // sums: [2][C] = (sum g, sum g*xhat)
  const float *gamma_b, *mean_b, *invstd_b, *sums_b;
  PView<T> da, db;
  float inv_count;
  int npix, C;
  int stripes;        // sums_a / sums_b point at copy 0 of [stripes][nq][C] striped totals (1 = plain [2][C] rows)
  int stripe_stride;  // floats between copies (nq * C)
  AccSegs acc;        // striped mode: parameter-gradient slots (d beta_a, d gamma_a, d beta_b, d gamma_b), += by block 0
  VecGeom geom;
};

// d_in = A*g + B*x + K with A = gamma*invstd, B = -gamma*invstd^2*s2/N, K = -A*s1/N - B*mean
template <int V>
__device__ __forceinline__ void bn_bwd_coef(const float* gamma, const float* mean, const float* invstd,
                                            const float* sums, int C, int c0, float inv_count, float (&ca)[V],
                                            float (&cb)[V], float (&ck)[V], int stripes, int stripe_stride,
                                            float* slot_beta, int valid_beta, float* slot_gamma, int valid_gamma) {
  float ga[V], mu[V], is[V], s1[V], s2[V];
  if (gamma) load_coef<V>(gamma, c0, ga);
  load_coef<V>(mean, c0, mu);
  load_coef<V>(invstd, c0, is);
  load_coef<V>(sums, c0, s1);
  load_coef<V>(sums + C, c0, s2);
  for (int k = 1; k < stripes; ++k) {  // fold the striped copies the reduce blocks added into
    float t1[V], t2[V];
    load_coef<V>(sums + (int64_t)k * stripe_stride, c0, t1);
    load_coef<V>(sums + (int64_t)k * stripe_stride + C, c0, t2);
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] += t1[i]; s2[i] += t2[i]; }
  }
  // BatchNorm parameter gradients (d beta = sum g, d gamma = sum g*xhat) straight into the optimizer's flat gradient
  // buffer: one thread per channel vector of the first pixel block, plain += (this stream is the only writer)
  if (slot_beta != nullptr || slot_gamma != nullptr) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if (slot_beta != nullptr && c0 + i < valid_beta) slot_beta[c0 + i] += s1[i];
      if (slot_gamma != nullptr && c0 + i < valid_gamma) slot_gamma[c0 + i] += s2[i];
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float gm = gamma ? ga[i] : 1.f;
    ca[i] = gm * is[i];
    cb[i] = -ca[i] * is[i] * s2[i] * inv_count;
    ck[i] = -ca[i] * s1[i] * inv_count - cb[i] * mu[i];
  }
}

// STRIPED = striped totals / gradient-slot writes (npp_node_bwd_apply_striped); kept out of the default instantiation
// for the same register-budget reason as HAS2 above.
template <typename T, int U, bool STRIPED>
__global__ void __launch_bounds__(256, 2) node_bwd_apply_kernel(const NodeBwdApplyArgs<T> A) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  const int tcv = threadIdx.x % A.geom.cvb;
  const int trow = threadIdx.x / A.geom.cvb;
  const int mycv = blockIdx.y * A.geom.cvb + tcv;
  if (trow >= A.geom.rows || mycv >= A.geom.cv) return;
  const int c0 = mycv * V;
  const bool has_a = A.a.p != nullptr, has_b = A.b.p != nullptr;
  float aa[V], ab[V], ak[V], ba[V], bb[V], bk[V];
  const bool owner = STRIPED && blockIdx.x == 0 && trow == 0;  // one thread per channel vector writes the slots
  const int stripes = STRIPED ? A.stripes : 1;
  if (has_a)
    bn_bwd_coef<V>(A.gamma_a, A.mean_a, A.invstd_a, A.sums_a, A.C, c0, A.inv_count, aa, ab, ak, stripes,
                   A.stripe_stride, owner ? A.acc.s[0].ptr : nullptr, A.acc.s[0].valid, owner ? A.acc.s[1].ptr : nullptr,
                   A.acc.s[1].valid);
  if (has_b)
    bn_bwd_coef<V>(A.gamma_b, A.mean_b, A.invstd_b, A.sums_b, A.C, c0, A.inv_count, ba, bb, bk, stripes,
                   A.stripe_stride, owner ? A.acc.s[2].ptr : nullptr, A.acc.s[2].valid, owner ? A.acc.s[3].ptr : nullptr,
                   A.acc.s[3].valid);
  const int step = gridDim.x * A.geom.rows;
  for (int p0 = blockIdx.x * A.geom.rows + trow; p0 < A.npix; p0 += U * step) {
    uint4 qg[U], qa[U], qb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * step;
      if (p < A.npix) {
        qg[u] = ldraw(A.g.at(p, c0));
        if (has_a) qa[u] = ldraw(A.a.at(p, c0));
        if (has_b) qb[u] = ldraw(A.b.at(p, c0));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * step;
      if (p >= A.npix) continue;
      float g[1][V], xa[1][V], xb[1][V];
      Pack<T>::unpack(qg[u], g[0]);
      if (has_a) Pack<T>::unpack(qa[u], xa[0]);
      if (has_b) Pack<T>::unpack(qb[u], xb[0]);
      if (has_a) {
#pragma unroll
        for (int i = 0; i < V; ++i) xa[0][i] = fmaf(aa[i], g[0][i], fmaf(ab[i], xa[0][i], ak[i]));
        Pack<T>::store(A.da.at(p, c0), xa[0]);
      }
      if (has_b) {
#pragma unroll
        for (int i = 0; i < V; ++i) xb[0][i] = fmaf(ba[i], g[0][i], fmaf(bb[i], xb[0][i], bk[i]));
        Pack<T>::store(A.db.at(p, c0), xb[0]);
      }
    }
  }
}

template <typename T>
static int node_bwd_apply_t(const npp_view4* g, const npp_view4* a, const float* gamma_a, const float* mean_a,
                            const float* invstd_a, const float* sums_a, const npp_view4* da, const npp_view4* b,
                            const float* gamma_b, const float* mean_b, const float* invstd_b, const float* sums_b,
                            const npp_view4* db, double count, cudaStream_t st, int stripes = 1, int stripe_stride = 0,
                            const AccSegs* acc = nullptr) {
  constexpr int V = Pack<T>::N;
  NodeBwdApplyArgs<T> A;
  A.stripes = stripes < 1 ? 1 : stripes;
  A.stripe_stride = stripe_stride;
  if (acc) A.acc = *acc; else memset(&A.acc, 0, sizeof A.acc);
  A.g = pview<const T>(g);
  A.a = a ? pview<const T>(a) : pview_null<const T>();
  A.b = b ? pview<const T>(b) : pview_null<const T>();
  A.gamma_a = gamma_a; A.mean_a = mean_a; A.invstd_a = invstd_a; A.sums_a = sums_a;
  A.gamma_b = gamma_b; A.mean_b = mean_b; A.invstd_b = invstd_b; A.sums_b = sums_b;
  A.da = da ? pview<T>(da) : pview_null<T>();
  A.db = db ? pview<T>(db) : pview_null<T>();
  A.inv_count = (float)(1.0 / count);
  const int64_t npix = (int64_t)g->n * g->h * g->w;
  if (npix > 0x3fffffff) return NPP_E_UNSUPPORTED;
  A.npix = (int)npix;
  A.C = g->c;
  A.geom = vec_geom(g->c, V);
  dim3 grid((unsigned)stream_grid(npix, A.geom, 4, 8), (unsigned)A.geom.gy);
  if (A.stripes > 1 || acc)
    NPP_LAUNCH((node_bwd_apply_kernel<T, 4, true>), grid, 256, 0, st, A);
  else
    NPP_LAUNCH((node_bwd_apply_kernel<T, 4, false>), grid, 256, 0, st, A);
  NPP_CHECK_LAUNCH("node_bwd_apply_kernel");
  return NPP_OK;
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_node_fwd(const npp_view4* a, const float* scale_a, const float* shift_a, const npp_view4* b,
                 const float* scale_b, const float* shift_b, const npp_view4* y_raw, const npp_view4* y_relu, int dtype,
                 npp_stream_t s) {
  if (!view_ok(a, dtype) || (!y_raw && !y_relu)) return NPP_E_INVALID;
  if ((scale_a == nullptr) != (shift_a == nullptr) || (scale_b == nullptr) != (shift_b == nullptr)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(a, b))) return NPP_E_INVALID;
  if (!b && scale_b) return NPP_E_INVALID;
  if (y_raw && (!view_ok(y_raw, dtype) || !same_shape(a, y_raw))) return NPP_E_INVALID;
  if (y_relu && (!view_ok(y_relu, dtype) || !same_shape(a, y_relu))) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return node_fwd_t<T>(a, scale_a, shift_a, b, scale_b, shift_b, y_raw, y_relu, as_stream(s)););
}

static bool bn_fin_ok(const npp_bn_fin* f) { return f->stats && f->coef && f->c_run >= 0; }
static BnFin to_fin(const npp_bn_fin* f) {
  BnFin r;
  r.stats = f->stats; r.gamma = f->gamma; r.beta = f->beta; r.rmean = f->running_mean; r.rvar = f->running_var;
  r.coef = f->coef; r.momentum = f->momentum; r.eps = f->eps; r.c_run = f->c_run;
  return r;
}

int npp_node_fwd_bn(const npp_view4* a, const npp_bn_fin* fin_a, const float* scale_a, const float* shift_a,
                    const npp_view4* b, const npp_bn_fin* fin_b, const float* scale_b, const float* shift_b,
                    const npp_view4* y_raw, const npp_view4* y_relu, double count, int dtype, npp_stream_t s) {
  if (!view_ok(a, dtype) || (!y_raw && !y_relu) || count <= 0) return NPP_E_INVALID;
  if (fin_a && (!bn_fin_ok(fin_a) || scale_a || shift_a)) return NPP_E_INVALID;
  if (fin_b && (!bn_fin_ok(fin_b) || scale_b || shift_b || !b)) return NPP_E_INVALID;
  if ((scale_a == nullptr) != (shift_a == nullptr) || (scale_b == nullptr) != (shift_b == nullptr)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(a, b))) return NPP_E_INVALID;
  if (!b && scale_b) return NPP_E_INVALID;
  if (y_raw && (!view_ok(y_raw, dtype) || !same_shape(a, y_raw))) return NPP_E_INVALID;
  if (y_relu && (!view_ok(y_relu, dtype) || !same_shape(a, y_relu))) return NPP_E_INVALID;
  if ((fin_a && fin_a->c_run > a->c) || (fin_b && fin_b->c_run > a->c)) return NPP_E_INVALID;
  BnFin fa, fb;
  if (fin_a) fa = to_fin(fin_a);
  if (fin_b) fb = to_fin(fin_b);
  NPP_DISPATCH_DTYPE(dtype, return node_fwd_t<T>(a, scale_a, shift_a, b, scale_b, shift_b, y_raw, y_relu, as_stream(s),
                                                 fin_a ? &fa : nullptr, fin_b ? &fb : nullptr, count););
}

int npp_node_bwd_blocks(int n, int h, int w, int c, int dtype) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return NPP_E_INVALID;
  return node_bwd_blocks((int64_t)n * h * w, c, dtype == NPP_BF16 ? 8 : 4);
}

int npp_node_bwd_reduce(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out, const npp_view4* a,
                        const float* mean_a, const float* invstd_a, const npp_view4* b, const float* mean_b,
                        const float* invstd_b, const npp_view4* g_out, float* partials, int dtype, npp_stream_t s) {
  const npp_view4* ref = g_raw ? g_raw : g_relu;
  if (!ref || !view_ok(ref, dtype)) return NPP_E_INVALID;
  if (g_raw && g_relu && (!view_ok(g_relu, dtype) || !same_shape(ref, g_relu))) return NPP_E_INVALID;
  if (g_relu && (!relu_out || !view_ok(relu_out, dtype) || !same_shape(ref, relu_out))) return NPP_E_INVALID;
  if (a && (!view_ok(a, dtype) || !same_shape(ref, a) || !mean_a || !invstd_a)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(ref, b) || !mean_b || !invstd_b)) return NPP_E_INVALID;
  if (g_out && (!view_ok(g_out, dtype) || !same_shape(ref, g_out))) return NPP_E_INVALID;
  if ((a || b) && !partials) return NPP_E_INVALID;
  if (!a && !b && !g_out) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_reduce_t<T>(g_raw, g_relu, relu_out, a, mean_a, invstd_a, b, mean_b,
                                                        invstd_b, g_out, (a || b) ? partials : nullptr, nullptr,
                                                        nullptr, as_stream(s)););
}

int npp_node_bwd_reduce_atomic(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                               const npp_view4* a, const float* mean_a, const float* invstd_a, const npp_view4* b,
                               const float* mean_b, const float* invstd_b, const npp_view4* g_out, float* sums,
                               float* const* acc, const int* acc_valid, int dtype, npp_stream_t s) {
  const npp_view4* ref = g_raw ? g_raw : g_relu;
  if (!ref || !view_ok(ref, dtype)) return NPP_E_INVALID;
  if (g_raw && g_relu && (!view_ok(g_relu, dtype) || !same_shape(ref, g_relu))) return NPP_E_INVALID;
  if (g_relu && (!relu_out || !view_ok(relu_out, dtype) || !same_shape(ref, relu_out))) return NPP_E_INVALID;
  if (a && (!view_ok(a, dtype) || !same_shape(ref, a) || !mean_a || !invstd_a)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(ref, b) || !mean_b || !invstd_b)) return NPP_E_INVALID;
  if (g_out && (!view_ok(g_out, dtype) || !same_shape(ref, g_out))) return NPP_E_INVALID;
  if ((!a && !b) || !sums) return NPP_E_INVALID;
  AccSegs A;
  memset(&A, 0, sizeof A);
  const int nq = (a ? 2 : 0) + (b ? 2 : 0);
  if (acc) {
    if (!acc_valid) return NPP_E_INVALID;
    for (int i = 0; i < nq; ++i) {
      A.s[i].ptr = acc[i];
      A.s[i].valid = acc[i] ? acc_valid[i] : 0;
      if (A.s[i].valid < 0 || A.s[i].valid > ref->c) return NPP_E_INVALID;
    }
  }
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_reduce_t<T>(g_raw, g_relu, relu_out, a, mean_a, invstd_a, b, mean_b,
                                                        invstd_b, g_out, nullptr, sums, &A, as_stream(s)););
}

int npp_node_bwd_reduce_striped(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                                const npp_view4* a, const float* mean_a, const float* invstd_a, const npp_view4* b,
                                const float* mean_b, const float* invstd_b, const npp_view4* g_out, float* sums,
                                int stripes, int dtype, npp_stream_t s) {
  const npp_view4* ref = g_raw ? g_raw : g_relu;
  if (!ref || !view_ok(ref, dtype)) return NPP_E_INVALID;
  if (g_raw && g_relu && (!view_ok(g_relu, dtype) || !same_shape(ref, g_relu))) return NPP_E_INVALID;
  if (g_relu && (!relu_out || !view_ok(relu_out, dtype) || !same_shape(ref, relu_out))) return NPP_E_INVALID;
  if (a && (!view_ok(a, dtype) || !same_shape(ref, a) || !mean_a || !invstd_a)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(ref, b) || !mean_b || !invstd_b)) return NPP_E_INVALID;
  if (g_out && (!view_ok(g_out, dtype) || !same_shape(ref, g_out))) return NPP_E_INVALID;
  if ((!a && !b) || !sums || stripes < 1 || stripes > 64) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_reduce_t<T>(g_raw, g_relu, relu_out, a, mean_a, invstd_a, b, mean_b,
                                                        invstd_b, g_out, nullptr, sums, nullptr, as_stream(s), stripes););
}

int npp_node_bwd_reduce2(const npp_view4* g_raw, const npp_view4* g_raw2, const npp_view4* g_relu,
                         const npp_view4* g_relu2, const npp_view4* relu_out, const npp_view4* a, const float* mean_a,
                         const float* invstd_a, const npp_view4* b, const float* mean_b, const float* invstd_b,
                         const npp_view4* g_out, float* partials, float* sums, int stripes, int dtype, npp_stream_t s) {
  const npp_view4* ref = g_raw ? g_raw : g_relu;
  if (!ref || !view_ok(ref, dtype)) return NPP_E_INVALID;
  if (g_raw && g_relu && (!view_ok(g_relu, dtype) || !same_shape(ref, g_relu))) return NPP_E_INVALID;
  if (g_raw2 && (!g_raw || !view_ok(g_raw2, dtype) || !same_shape(ref, g_raw2))) return NPP_E_INVALID;
  if (g_relu2 && (!g_relu || !view_ok(g_relu2, dtype) || !same_shape(ref, g_relu2))) return NPP_E_INVALID;
  if (g_relu && (!relu_out || !view_ok(relu_out, dtype) || !same_shape(ref, relu_out))) return NPP_E_INVALID;
  if (a && (!view_ok(a, dtype) || !same_shape(ref, a) || !mean_a || !invstd_a)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(ref, b) || !mean_b || !invstd_b)) return NPP_E_INVALID;
  if (g_out && (!view_ok(g_out, dtype) || !same_shape(ref, g_out))) return NPP_E_INVALID;
  if ((g_relu || g_raw2) && !g_out) return NPP_E_INVALID;      // a combined gradient has to be written somewhere
  if ((a || b) && !partials && !sums) return NPP_E_INVALID;
  if (partials && sums) return NPP_E_INVALID;
  if (sums && (stripes < 1 || stripes > 64)) return NPP_E_INVALID;
  if (!a && !b && !g_out) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_reduce_t<T>(g_raw, g_relu, relu_out, a, mean_a, invstd_a, b, mean_b,
                                                        invstd_b, g_out, (a || b) ? partials : nullptr,
                                                        (a || b) ? sums : nullptr, nullptr, as_stream(s),
                                                        sums ? stripes : 1, g_raw2, g_relu2););
}

int npp_node_bwd_reduce3(const npp_view4* g_raw, const npp_view4* g_relu, const npp_view4* relu_out,
                         const npp_view4* extra0, int extra0_is_relu, const npp_view4* extra1, int extra1_is_relu,
                         const npp_view4* a, const float* mean_a, const float* invstd_a, const npp_view4* b,
                         const float* mean_b, const float* invstd_b, const npp_view4* g_out, float* partials, int dtype,
                         npp_stream_t s) {
  const npp_view4* ref = g_raw ? g_raw : g_relu;
  if (!ref || !view_ok(ref, dtype)) return NPP_E_INVALID;
  if (g_raw && g_relu && (!view_ok(g_relu, dtype) || !same_shape(ref, g_relu))) return NPP_E_INVALID;
  if (g_relu && (!relu_out || !view_ok(relu_out, dtype) || !same_shape(ref, relu_out))) return NPP_E_INVALID;
  if (!extra0 && extra1) return NPP_E_INVALID;                  // slots are filled in order
  const npp_view4* ex[2] = {extra0, extra1};
  const int kind[2] = {extra0_is_relu, extra1_is_relu};
  for (int i = 0; i < 2; ++i) {
    if (!ex[i]) continue;
    if (!view_ok(ex[i], dtype) || !same_shape(ref, ex[i])) return NPP_E_INVALID;
    if (kind[i] ? !g_relu : !g_raw) return NPP_E_INVALID;       // an extra gradient needs the primary one of its kind
  }
  if (a && (!view_ok(a, dtype) || !same_shape(ref, a) || !mean_a || !invstd_a)) return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(ref, b) || !mean_b || !invstd_b)) return NPP_E_INVALID;
  if (g_out && (!view_ok(g_out, dtype) || !same_shape(ref, g_out))) return NPP_E_INVALID;
  if ((g_relu || extra0) && !g_out) return NPP_E_INVALID;       // a combined gradient has to be written somewhere
  if ((a || b) && !partials) return NPP_E_INVALID;
  if (!a && !b && !g_out) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_reduce_t<T>(g_raw, g_relu, relu_out, a, mean_a, invstd_a, b, mean_b,
                                                        invstd_b, g_out, (a || b) ? partials : nullptr, nullptr, nullptr,
                                                        as_stream(s), 1, extra0, extra1, extra0_is_relu ? 1 : 0,
                                                        extra1_is_relu ? 1 : 0););
}

int npp_node_bwd_apply_striped(const npp_view4* g, const npp_view4* a, const float* gamma_a, const float* mean_a,
                               const float* invstd_a, const npp_view4* da, const npp_view4* b, const float* gamma_b,
                               const float* mean_b, const float* invstd_b, const npp_view4* db, const float* sums,
                               int stripes, float* const* acc, const int* acc_valid, double count, int dtype,
                               npp_stream_t s) {
  if (!view_ok(g, dtype) || count <= 0 || (!a && !b) || !sums || stripes < 1 || stripes > 64) return NPP_E_INVALID;
  if (a && (!view_ok(a, dtype) || !same_shape(g, a) || !mean_a || !invstd_a || !da || !view_ok(da, dtype) ||
            !same_shape(g, da)))
    return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(g, b) || !mean_b || !invstd_b || !db || !view_ok(db, dtype) ||
            !same_shape(g, db)))
    return NPP_E_INVALID;
  const int C = g->c;
  const int nq = (a ? 2 : 0) + (b ? 2 : 0);
  AccSegs A;
  memset(&A, 0, sizeof A);
  if (acc) {
    if (!acc_valid) return NPP_E_INVALID;
    // acc rows follow the sums rows that exist: [d beta_a, d gamma_a] (if a) then [d beta_b, d gamma_b] (if b);
    // the kernel indexes sides a -> s[0..1], b -> s[2..3]
    int i = 0;
    if (a) { A.s[0].ptr = acc[i]; A.s[0].valid = acc[i] ? acc_valid[i] : 0; ++i; A.s[1].ptr = acc[i]; A.s[1].valid = acc[i] ? acc_valid[i] : 0; ++i; }
    if (b) { A.s[2].ptr = acc[i]; A.s[2].valid = acc[i] ? acc_valid[i] : 0; ++i; A.s[3].ptr = acc[i]; A.s[3].valid = acc[i] ? acc_valid[i] : 0; ++i; }
    for (int k = 0; k < 4; ++k)
      if (A.s[k].valid < 0 || A.s[k].valid > C) return NPP_E_INVALID;
  }
  const float* sums_a = a ? sums : nullptr;
  const float* sums_b = b ? sums + (a ? 2 : 0) * (int64_t)C : nullptr;
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_apply_t<T>(g, a, gamma_a, mean_a, invstd_a, sums_a, da, b, gamma_b, mean_b,
                                                       invstd_b, sums_b, db, count, as_stream(s), stripes, nq * C, &A););
}

int npp_reduce_partials(const float* partials, int rows, int len, float* out, npp_stream_t s) {
  if (!partials || !out || rows <= 0 || len <= 0) return NPP_E_INVALID;
  NPP_LAUNCH((reduce_partials_kernel), (len + 31) / 32, 256, 0, as_stream(s), partials, rows, len, out);
  NPP_CHECK_LAUNCH("reduce_partials_kernel");
  return NPP_OK;
}

int npp_reduce_partials_acc(const float* partials, int rows, int len, float* out, int seg_len, float* const* acc,
                            const int* acc_valid, npp_stream_t s) {
  if (!partials || !out || rows <= 0 || len <= 0 || seg_len <= 0 || len % seg_len || len / seg_len > NPP_ACC_MAX || !acc ||
      !acc_valid)
    return NPP_E_INVALID;
  AccSegs A;
  for (int i = 0; i < NPP_ACC_MAX; ++i) {
    const bool on = i < len / seg_len;
    A.s[i].ptr = on ? acc[i] : nullptr;
    A.s[i].valid = on ? acc_valid[i] : 0;
    if (on && (acc_valid[i] < 0 || acc_valid[i] > seg_len)) return NPP_E_INVALID;
  }
  NPP_LAUNCH((reduce_partials_acc_kernel), (len + 31) / 32, 256, 0, as_stream(s), partials, rows, len, out, seg_len, A);
  NPP_CHECK_LAUNCH("reduce_partials_acc_kernel");
  return NPP_OK;
}

int npp_node_bwd_apply(const npp_view4* g, const npp_view4* a, const float* gamma_a, const float* mean_a,
                       const float* invstd_a, const float* sums_a, const npp_view4* da, const npp_view4* b,
                       const float* gamma_b, const float* mean_b, const float* invstd_b, const float* sums_b,
                       const npp_view4* db, double count, int dtype, npp_stream_t s) {
  if (!view_ok(g, dtype) || count <= 0 || (!a && !b)) return NPP_E_INVALID;
  if (a && (!view_ok(a, dtype) || !same_shape(g, a) || !mean_a || !invstd_a || !sums_a || !da || !view_ok(da, dtype) ||
            !same_shape(g, da)))
    return NPP_E_INVALID;
  if (b && (!view_ok(b, dtype) || !same_shape(g, b) || !mean_b || !invstd_b || !sums_b || !db || !view_ok(db, dtype) ||
            !same_shape(g, db)))
    return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(dtype, return node_bwd_apply_t<T>(g, a, gamma_a, mean_a, invstd_a, sums_a, da, b, gamma_b, mean_b,
                                                       invstd_b, sums_b, db, count, as_stream(s)););
}

}  // extern "C"
