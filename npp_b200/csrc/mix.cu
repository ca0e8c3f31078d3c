// MixedOp / weighted-node kernels of the search supernet (models/model_search_interact.py).
//
//   MixedOp.forward (:56-74):  temp1 = sum_k w_k * op_k(x[:, :C/2]);  ans = channel_shuffle(cat(temp1, x[:, C/2:]), 2)
//   supernet node  (:648-654, :352-356): s = base + sum_j beta_j * MixedOp_j(h_j, alpha_j)
//
// Every candidate primitive ends in a training-mode BatchNorm2d(affine=False) (operations.py:61,79,215,240 and the
// extra nn.BatchNorm2d of :48-49), so the reference runs K BN-apply passes, K scalar multiplies, K-1 adds, a cat and a
// channel shuffle (and their backward passes) per MixedOp.  Here:
//
//   fwd     : out[.., 2c] = sum_k w_k * (y_k*scale_k + shift_k)[c],  out[.., 2c+1] = pass[c]     (interleave = the
//             cat + channel_shuffle(groups=2) of :70-71), or out[.., c] = sum_k ... without the interleave (weighted
//             node sums, cross-scale MixedOps that are resampled before the shuffle).  ONE pass over the K branch
//             outputs; w_k are read from device memory (softmax(alpha) rows: no host sync).
//   bwd 1/2 : g = dout[.., 2c] (or dout[.., c]); per-block partial sums of  sum g  and, per branch,  sum g*xhat_k
//             (BatchNorm branch) or  sum g*y_k  (plain branch): they give both the BatchNorm backward terms and the
//             architecture gradient  dw_k = <g, branch_k>  — per-block partials, folded by npp_reduce_partials
//             (deterministic, no atomics).
//   bwd 2/2 : d y_k = w_k * gamma_k*invstd_k*(g - s0/N - xhat_k*s_k/N)  (BatchNorm) or  w_k * g  (plain) for all
//             branches from ONE read of g, plus d pass = dout[.., 2c+1].
//
// HBM-bound: 16-byte vectors, consecutive threads on consecutive channel vectors of a pixel.
#include "view.cuh"

namespace npp {

constexpr int kMixMax = NPP_MIX_MAX;

template <typename T>
struct MView {  // pixel-indexed view (dense fast path), as node.cu's PView
  T* p;
  int64_t sn, sh, sw;
  int H, W;
  int dense;
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
static inline MView<T> mview(const npp_view4* v) {
  MView<T> d;
  if (!v || !v->ptr) {
    d.p = nullptr; d.sn = d.sh = d.sw = 0; d.H = d.W = 1; d.dense = 1;
    return d;
  }
  d.p = static_cast<T*>(v->ptr);
  d.sn = v->sn; d.sh = v->sh; d.sw = v->sw;
  d.H = v->h; d.W = v->w;
  d.dense = ((v->sh == (int64_t)v->w * v->sw) && (v->sn == (int64_t)v->h * v->sh)) ? 1 : 0;
  return d;
}

template <int V>
__device__ __forceinline__ void ld_coef(const float* p, int c0, float (&v)[V]) {
#pragma unroll
  for (int i = 0; i < V; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c0 + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
}

template <typename T>
struct MixArgs {
  int k, interleave, npix, C;  // C = channels of the mixed half
  MView<const T> y[kMixMax];
  const float* c0p[kMixMax];   // fwd: scale   | bwd: mean    (nullptr = plain branch)
  const float* c1p[kMixMax];   // fwd: shift   | bwd: invstd
  const float* gam[kMixMax];   // bwd apply: gamma (nullptr = 1)
  MView<T> dy[kMixMax];        // bwd apply outputs (p == nullptr = not needed)
  const float* wts;            // device [k]
  MView<const T> pass;         // fwd pass-through half
  MView<T> out;                // fwd output
  MView<const T> g;            // bwd: gradient of the output
  MView<T> dpass;              // bwd apply: gradient of the pass-through half
  float* partials;             // bwd reduce: [gridDim.x][k+1][C]
  const float* sums;           // bwd apply: [k+1][C]
  float inv_count;
  VecGeom geom;
};

// loads the V gradient values of mixed channels c0..c0+V-1 of pixel p (even channels of the interleaved output)
template <typename T, int V>
__device__ __forceinline__ void load_g(const MView<const T>& g, int interleave, int p, int c0, float (&v)[V],
                                       float (&odd)[V]) {
  if (!interleave) {
    Pack<T>::load(g.at(p, c0), v);
    return;
  }
  float a[V], b[V];
  const T* q = g.at(p, 2 * c0);
  Pack<T>::load(q, a);
  Pack<T>::load(q + V, b);
#pragma unroll
  for (int i = 0; i < V / 2; ++i) {
    v[i] = a[2 * i]; odd[i] = a[2 * i + 1];
    v[V / 2 + i] = b[2 * i]; odd[V / 2 + i] = b[2 * i + 1];
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(256) mix_fwd_kernel(const MixArgs<T> A) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  const int tcv = threadIdx.x % A.geom.cvb;
  const int trow = threadIdx.x / A.geom.cvb;
  const int mycv = blockIdx.y * A.geom.cvb + tcv;
  if (trow >= A.geom.rows || mycv >= A.geom.cv) return;
  const int c0 = mycv * V;
  // per-thread channel coefficients: out = sum_k (y_k * ws_k + wt_k) with ws = w*scale, wt = w*shift
  float ws[kMixMax][V];
  float bias[V];
#pragma unroll
  for (int i = 0; i < V; ++i) bias[i] = 0.f;
#pragma unroll
  for (int j = 0; j < kMixMax; ++j) {
    if (j < A.k) {
      const float w = A.wts ? A.wts[j] : 1.f;
      if (A.c0p[j]) {
        float sc[V], sh[V];
        ld_coef<V>(A.c0p[j], c0, sc);
        ld_coef<V>(A.c1p[j], c0, sh);
#pragma unroll
        for (int i = 0; i < V; ++i) { ws[j][i] = w * sc[i]; bias[i] = fmaf(w, sh[i], bias[i]); }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) ws[j][i] = w;
      }
    }
  }
  const int step = gridDim.x * A.geom.rows;
  for (int p = blockIdx.x * A.geom.rows + trow; p < A.npix; p += step) {
    uint4 q[kMixMax];
#pragma unroll
    for (int j = 0; j < kMixMax; ++j)
      if (j < A.k) q[j] = ldraw(A.y[j].at(p, c0));
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = bias[i];
#pragma unroll
    for (int j = 0; j < kMixMax; ++j) {
      if (j < A.k) {
        float v[V];
        Pack<T>::unpack(q[j], v);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(v[i], ws[j][i], acc[i]);
      }
    }
    if (!A.interleave) {
      Pack<T>::store(A.out.at(p, c0), acc);
    } else {
      float ps[V], lo[V], hi[V];
      Pack<T>::load(A.pass.at(p, c0), ps);
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        lo[2 * i] = acc[i]; lo[2 * i + 1] = ps[i];
        hi[2 * i] = acc[V / 2 + i]; hi[2 * i + 1] = ps[V / 2 + i];
      }
      T* o = A.out.at(p, 2 * c0);
      Pack<T>::store(o, lo);
      Pack<T>::store(o + V, hi);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward 1/2
template <typename T>
__global__ void __launch_bounds__(256) mix_bwd_reduce_kernel(const MixArgs<T> A) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  __shared__ float red[256 * V];
  const int tcv = threadIdx.x % A.geom.cvb;
  const int trow = threadIdx.x / A.geom.cvb;
  const int mycv = tcv;  // gy == 1, checked on the host
  const bool active = trow < A.geom.rows && mycv < A.geom.cv;
  const int c0 = mycv * V;
  float s0[V], s[kMixMax][V], mu[kMixMax][V];
#pragma unroll
  for (int i = 0; i < V; ++i) s0[i] = 0.f;
#pragma unroll
  for (int j = 0; j < kMixMax; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) { s[j][i] = 0.f; mu[j][i] = 0.f; }
  if (active) {
#pragma unroll
    for (int j = 0; j < kMixMax; ++j)
      if (j < A.k && A.c0p[j]) ld_coef<V>(A.c0p[j], c0, mu[j]);
    const int step = gridDim.x * A.geom.rows;
    for (int p = blockIdx.x * A.geom.rows + trow; p < A.npix; p += step) {
      uint4 q[kMixMax];
#pragma unroll
      for (int j = 0; j < kMixMax; ++j)
        if (j < A.k) q[j] = ldraw(A.y[j].at(p, c0));
      float g[V], odd[V];
      load_g<T, V>(A.g, A.interleave, p, c0, g, odd);
#pragma unroll
      for (int i = 0; i < V; ++i) s0[i] += g[i];
#pragma unroll
      for (int j = 0; j < kMixMax; ++j) {
        if (j < A.k) {
          float v[V];
          Pack<T>::unpack(q[j], v);
#pragma unroll
          for (int i = 0; i < V; ++i) s[j][i] = fmaf(g[i], v[i] - mu[j][i], s[j][i]);
        }
      }
    }
  }
  float* out = A.partials + (int64_t)blockIdx.x * (A.k + 1) * A.C;
#pragma unroll
  for (int r = 0; r <= kMixMax; ++r) {
    if (r <= A.k) {  // uniform across the block
      constexpr int kLast = kMixMax - 1;
      const int j = r == 0 ? 0 : (r - 1 > kLast ? kLast : r - 1);
      __syncthreads();
#pragma unroll
      for (int v = 0; v < V; ++v) red[threadIdx.x * V + v] = (r == 0) ? s0[v] : s[j][v];
      __syncthreads();
      block_colsum(red, A.geom.rows, A.geom.cvb * V);
      if (active && trow == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float t = red[tcv * V + v];
          if (r > 0 && A.c0p[j]) t *= A.c1p[j][c0 + v];  // centred sum * invstd = sum g*xhat
          out[r * A.C + c0 + v] = t;
        }
      }
    }
  }
}

// dw[j] = sum_c (gamma_j[c] * S_j[c] + beta_j[c] * S_0[c])   (BatchNorm branch; gamma = 1 / beta = 0 when NULL)
//       = sum_c S_j[c]                                         (plain branch: S_j = sum g*y_j)
struct MixDwArgs {
  int k, C;
  const float* sums;
  const float* gamma[kMixMax];
  const float* beta[kMixMax];
  float* dw;
};
__global__ void mix_dw_kernel(const MixDwArgs A) {
  pdl_wait();
  const int j = blockIdx.x;
  if (j >= A.k) return;
  float t = 0.f;
  for (int c = threadIdx.x; c < A.C; c += blockDim.x) {
    float v = A.sums[(j + 1) * A.C + c];
    if (A.gamma[j]) v *= A.gamma[j][c];
    if (A.beta[j]) v = fmaf(A.beta[j][c], A.sums[c], v);
    t += v;
  }
  __shared__ float red[32];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    float u = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    u = warp_sum(u);
    if (threadIdx.x == 0) A.dw[j] = u;
  }
}

// ------------------------------------------------------------------------------------------------ backward 2/2
template <typename T>
__global__ void __launch_bounds__(256) mix_bwd_apply_kernel(const MixArgs<T> A) {
  pdl_wait();
  constexpr int V = Pack<T>::N;
  const int tcv = threadIdx.x % A.geom.cvb;
  const int trow = threadIdx.x / A.geom.cvb;
  const int mycv = blockIdx.y * A.geom.cvb + tcv;
  if (trow >= A.geom.rows || mycv >= A.geom.cv) return;
  const int c0 = mycv * V;
  // d y_j = ca*g + cb*y_j + ck  (BatchNorm) or ca*g (plain)
  float ca[kMixMax][V], cb[kMixMax][V], ck[kMixMax][V];
  float s0[V];
  ld_coef<V>(A.sums, c0, s0);
#pragma unroll
  for (int j = 0; j < kMixMax; ++j) {
    if (j < A.k && A.dy[j].p) {
      const float w = A.wts ? A.wts[j] : 1.f;
      if (A.c0p[j]) {
        float mu[V], is[V], sj[V], ga[V];
        ld_coef<V>(A.c0p[j], c0, mu);
        ld_coef<V>(A.c1p[j], c0, is);
        ld_coef<V>(A.sums + (int64_t)(j + 1) * A.C, c0, sj);
        if (A.gam[j]) ld_coef<V>(A.gam[j], c0, ga);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float a = w * (A.gam[j] ? ga[i] : 1.f) * is[i];
          ca[j][i] = a;
          cb[j][i] = -a * is[i] * sj[i] * A.inv_count;
          ck[j][i] = -a * s0[i] * A.inv_count - cb[j][i] * mu[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) { ca[j][i] = w; cb[j][i] = 0.f; ck[j][i] = 0.f; }
      }
    }
  }
  const int step = gridDim.x * A.geom.rows;
  for (int p = blockIdx.x * A.geom.rows + trow; p < A.npix; p += step) {
    uint4 q[kMixMax];
#pragma unroll
    for (int j = 0; j < kMixMax; ++j)
      if (j < A.k && A.dy[j].p && A.c0p[j]) q[j] = ldraw(A.y[j].at(p, c0));
    float g[V], odd[V];
    load_g<T, V>(A.g, A.interleave, p, c0, g, odd);
    if (A.interleave && A.dpass.p) Pack<T>::store(A.dpass.at(p, c0), odd);
#pragma unroll
    for (int j = 0; j < kMixMax; ++j) {
      if (j < A.k && A.dy[j].p) {
        float v[V];
        if (A.c0p[j]) {
          Pack<T>::unpack(q[j], v);
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = fmaf(ca[j][i], g[i], fmaf(cb[j][i], v[i], ck[j][i]));
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = ca[j][i] * g[i];
        }
        Pack<T>::store(A.dy[j].at(p, c0), v);
      }
    }
  }
}

static inline int mix_grid(int64_t npix, const VecGeom& g, int per_thread, int blocks_per_sm) {
  int64_t gx = (npix + (int64_t)g.rows * per_thread - 1) / ((int64_t)g.rows * per_thread);
  const int64_t cap = ((int64_t)sm_count() * blocks_per_sm + g.gy - 1) / g.gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return (int)gx;
}

static int mix_check(const npp_mix_desc* d, int dtype) {
  if (!d || d->k < 1 || d->k > kMixMax) return NPP_E_INVALID;
  for (int j = 0; j < d->k; ++j) {
    if (!view_ok(&d->y[j], dtype) || !same_shape(&d->y[0], &d->y[j])) return NPP_E_INVALID;
  }
  return NPP_OK;
}

template <typename T>
static void mix_fill(MixArgs<T>& A, const npp_mix_desc* d, bool fwd) {
  constexpr int V = Pack<T>::N;
  A.k = d->k;
  A.interleave = d->interleave ? 1 : 0;
  A.C = d->y[0].c;
  A.npix = (int)((int64_t)d->y[0].n * d->y[0].h * d->y[0].w);
  for (int j = 0; j < kMixMax; ++j) {
    const bool on = j < d->k;
    A.y[j] = mview<const T>(on ? &d->y[j] : nullptr);
    A.c0p[j] = on ? (fwd ? d->scale[j] : d->mean[j]) : nullptr;
    A.c1p[j] = on ? (fwd ? d->shift[j] : d->invstd[j]) : nullptr;
    A.gam[j] = on ? d->gamma[j] : nullptr;
    A.dy[j] = mview<T>(on && !fwd ? &d->dy[j] : nullptr);
  }
  A.wts = nullptr; A.partials = nullptr; A.sums = nullptr; A.inv_count = 0.f;
  A.pass = mview<const T>(nullptr); A.out = mview<T>(nullptr); A.g = mview<const T>(nullptr);
  A.dpass = mview<T>(nullptr);
  A.geom = vec_geom(A.C, V);
}

// the other side of an (optionally interleaved) mix: 2*C channels when interleaved, C otherwise
static bool mix_side_ok(const npp_mix_desc* d, const npp_view4* v, int dtype) {
  if (!view_ok(v, dtype)) return false;
  const npp_view4& y = d->y[0];
  return v->n == y.n && v->h == y.h && v->w == y.w && v->c == (d->interleave ? 2 * y.c : y.c);
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_mix_fwd(const npp_mix_desc* d, const float* wts, const npp_view4* pass, const npp_view4* out, int dtype,
                npp_stream_t s) {
  int rc = mix_check(d, dtype);
  if (rc) return rc;
  if (!mix_side_ok(d, out, dtype)) return NPP_E_INVALID;
  if (d->interleave && (!pass || !view_ok(pass, dtype) || !same_shape(&d->y[0], pass))) return NPP_E_INVALID;
  for (int j = 0; j < d->k; ++j)
    if ((d->scale[j] == nullptr) != (d->shift[j] == nullptr)) return NPP_E_INVALID;
  if ((int64_t)d->y[0].n * d->y[0].h * d->y[0].w > 0x3fffffff) return NPP_E_UNSUPPORTED;
  NPP_DISPATCH_DTYPE(
      dtype, MixArgs<T> A; mix_fill<T>(A, d, true); A.wts = wts; A.pass = mview<const T>(d->interleave ? pass : nullptr);
      A.out = mview<T>(out); dim3 grid((unsigned)mix_grid(A.npix, A.geom, 2, 8), (unsigned)A.geom.gy);
      NPP_LAUNCH((mix_fwd_kernel<T>), grid, 256, 0, as_stream(s), A); NPP_CHECK_LAUNCH("mix_fwd_kernel"); return NPP_OK;);
}

int npp_mix_bwd_reduce(const npp_mix_desc* d, const npp_view4* g, float* partials, int dtype, npp_stream_t s) {
  int rc = mix_check(d, dtype);
  if (rc) return rc;
  if (!mix_side_ok(d, g, dtype) || !partials) return NPP_E_INVALID;
  for (int j = 0; j < d->k; ++j)
    if ((d->mean[j] == nullptr) != (d->invstd[j] == nullptr)) return NPP_E_INVALID;
  const int64_t npix = (int64_t)d->y[0].n * d->y[0].h * d->y[0].w;
  if (npix > 0x3fffffff) return NPP_E_UNSUPPORTED;
  const int blocks = npp_node_bwd_blocks(d->y[0].n, d->y[0].h, d->y[0].w, d->y[0].c, dtype);
  if (blocks <= 0) return NPP_E_INVALID;
  NPP_DISPATCH_DTYPE(
      dtype, MixArgs<T> A; mix_fill<T>(A, d, false); if (A.geom.gy != 1) return NPP_E_UNSUPPORTED;
      A.g = mview<const T>(g); A.partials = partials; NPP_LAUNCH((mix_bwd_reduce_kernel<T>), dim3((unsigned)blocks, 1), 256, 0, as_stream(s), A);
      NPP_CHECK_LAUNCH("mix_bwd_reduce_kernel"); return NPP_OK;);
}

int npp_mix_dw(const npp_mix_desc* d, const float* sums, const float* const* beta, float* dw, npp_stream_t s) {
  if (!d || d->k < 1 || d->k > kMixMax || !sums || !dw) return NPP_E_INVALID;
  MixDwArgs A;
  A.k = d->k; A.C = d->y[0].c; A.sums = sums; A.dw = dw;
  for (int j = 0; j < kMixMax; ++j) {
    const bool bn = j < d->k && d->mean[j] != nullptr;
    A.gamma[j] = bn ? d->gamma[j] : nullptr;
    A.beta[j] = (bn && beta) ? beta[j] : nullptr;
  }
  NPP_LAUNCH((mix_dw_kernel), d->k, 256, 0, as_stream(s), A);
  NPP_CHECK_LAUNCH("mix_dw_kernel");
  return NPP_OK;
}

int npp_mix_bwd_apply(const npp_mix_desc* d, const npp_view4* g, const float* wts, const float* sums, double count,
                      const npp_view4* dpass, int dtype, npp_stream_t s) {
  int rc = mix_check(d, dtype);
  if (rc) return rc;
  if (!mix_side_ok(d, g, dtype) || !sums || count <= 0) return NPP_E_INVALID;
  if (dpass && (!d->interleave || !view_ok(dpass, dtype) || !same_shape(&d->y[0], dpass))) return NPP_E_INVALID;
  bool any = dpass != nullptr;
  for (int j = 0; j < d->k; ++j) {
    if ((d->mean[j] == nullptr) != (d->invstd[j] == nullptr)) return NPP_E_INVALID;
    if (d->dy[j].ptr) {
      if (!view_ok(&d->dy[j], dtype) || !same_shape(&d->y[0], &d->dy[j])) return NPP_E_INVALID;
      any = true;
    }
  }
  if (!any) return NPP_OK;
  if ((int64_t)d->y[0].n * d->y[0].h * d->y[0].w > 0x3fffffff) return NPP_E_UNSUPPORTED;
  NPP_DISPATCH_DTYPE(
      dtype, MixArgs<T> A; mix_fill<T>(A, d, false); A.g = mview<const T>(g); A.wts = wts; A.sums = sums;
      A.inv_count = (float)(1.0 / count); A.dpass = mview<T>(dpass);
      dim3 grid((unsigned)mix_grid(A.npix, A.geom, 2, 8), (unsigned)A.geom.gy);
      NPP_LAUNCH((mix_bwd_apply_kernel<T>), grid, 256, 0, as_stream(s), A); NPP_CHECK_LAUNCH("mix_bwd_apply_kernel");
      return NPP_OK;);
}

}  // extern "C"
