// Standalone device test: tcgen05 implicit-GEMM convolution (fprop / dgrad / wgrad / BN stats)
// against the CUDA-core direct convolution of the same library, on seeded random bf16 data.
// Usage: test_conv <case index | -1 for count>     (each case in its own process: a deadlocked
// mbarrier pipeline must not take the other cases down with it).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <dlfcn.h>
#include "../../include/npp_b200.h"


#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

struct Case {
  const char* name;
  int n, h, w, cin, cout, k, stride, pad, dil, hoff, woff;
  int xpad, ypad;  // extra channels in the underlying buffers (views are channel slices)
  int bias;
};

static const Case cases[] = {
    {"1x1 128->128 96x96 n2", 2, 96, 96, 128, 128, 1, 1, 0, 1, 0, 0, 0, 0, 0},
    {"1x1 64->64 32x32 n1", 1, 32, 32, 64, 64, 1, 1, 0, 1, 0, 0, 0, 0, 0},
    {"3x3 64->64 32x32 n2", 2, 32, 32, 64, 64, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 128->128 96x96 n2", 2, 96, 96, 128, 128, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 32->32 24x24 n3 (ragged)", 3, 24, 24, 32, 32, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 256->256 12x12 n4", 4, 12, 12, 256, 256, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"1x1 1024->384 bias 24x24 n2", 2, 24, 24, 1024, 384, 1, 1, 0, 1, 0, 0, 0, 0, 1},
    {"1x1 256->24 bias 48x48 n2 (padded cout)", 2, 48, 48, 256, 24, 1, 1, 0, 1, 0, 0, 0, 0, 1},
    {"3x3 s2 64->128 48x48 n2", 2, 48, 48, 64, 128, 3, 2, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 s2 8->64 64x64 n2 (stem)", 2, 64, 64, 8, 64, 3, 2, 1, 1, 0, 0, 0, 0, 0},
    {"1x1 s2 128->32 48x48 n2 (factorized a)", 2, 48, 48, 128, 32, 1, 2, 0, 1, 0, 0, 0, 32, 0},
    {"1x1 s2 off1 128->32 48x48 n2 (factorized b)", 2, 48, 48, 128, 32, 1, 2, 0, 1, 1, 1, 0, 32, 0},
    {"3x3 slice-in 64(+64)->64(+32) 24x24 n2", 2, 24, 24, 64, 64, 3, 1, 1, 1, 0, 0, 64, 32, 0},
    {"3x3 384->8 bias-free 24x24 n2 (edge head)", 2, 24, 24, 384, 8, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"1x1 8->8 bias 24x24 n2 (tiny K)", 2, 24, 24, 8, 8, 1, 1, 0, 1, 0, 0, 0, 0, 1},
    {"3x3 512->512 24x24 n2", 2, 24, 24, 512, 512, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 128->128 96x96 n8 (perf shape)", 8, 96, 96, 128, 128, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 128->128 96x96 n32 (bench shape)", 32, 96, 96, 128, 128, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"1x1 128->128 96x96 n32 (bench shape)", 32, 96, 96, 128, 128, 1, 1, 0, 1, 0, 0, 0, 0, 0},
    {"1x1 512->128 96x96 n32 (bench shape)", 32, 96, 96, 512, 128, 1, 1, 0, 1, 0, 0, 0, 0, 0},
    {"1x1 1024->512 bias 96x96 n32 (bench shape)", 32, 96, 96, 1024, 512, 1, 1, 0, 1, 0, 0, 0, 0, 1},
    {"3x3 32->32 96x96 n32 (bench shape)", 32, 96, 96, 32, 32, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 256->256 48x48 n32 (bench shape)", 32, 48, 48, 256, 256, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 64->64 48x48 n32 (bench shape)", 32, 48, 48, 64, 64, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 128->128 24x24 n32 (bench shape)", 32, 24, 24, 128, 128, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 256->256 12x12 n32 (bench shape)", 32, 12, 12, 256, 256, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 384->128 96x96 n4 (three cin blocks)", 4, 96, 96, 384, 128, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 40->24 20x36 n3 (ragged, odd channels)", 3, 20, 36, 40, 24, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 1024->1024 12x12 n8 (wide)", 8, 12, 12, 1024, 1024, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 64->64 40x24 n3 (ragged 256-pixel tiles)", 3, 40, 24, 64, 64, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"3x3 128->72 bias 48x48 n4 (two slabs, second partial)", 4, 48, 48, 128, 72, 3, 1, 1, 1, 0, 0, 0, 0, 1},
    {"3x3 slice-in 128(+64)->128(+32) 32x32 n5", 5, 32, 32, 128, 128, 3, 1, 1, 1, 0, 0, 64, 32, 0},
    {"3x3 64->64 96x96 n32 (bench shape)", 32, 96, 96, 64, 64, 3, 1, 1, 1, 0, 0, 0, 0, 0},
    {"1x1 32->32 96x96 n32 (bench shape)", 32, 96, 96, 32, 32, 1, 1, 0, 1, 0, 0, 0, 0, 0},
    {"1x1 64->64 48x48 n32 (bench shape)", 32, 48, 48, 64, 64, 1, 1, 0, 1, 0, 0, 0, 0, 0},
};

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xffff) / 32768.0f - 1.0f;
}

static void fill_bf16(std::vector<__nv_bfloat16>& v) {
  for (auto& e : v) e = __float2bfloat16(frand());
}

struct Cmp { double max_abs, max_ref, mean_abs; };
static Cmp compare_bf16(const std::vector<__nv_bfloat16>& a, const std::vector<__nv_bfloat16>& b) {
  Cmp c{0, 0, 0};
  for (size_t i = 0; i < a.size(); ++i) {
    const double x = __bfloat162float(a[i]), y = __bfloat162float(b[i]);
    const double d = fabs(x - y);
    if (d > c.max_abs) c.max_abs = d;
    if (fabs(y) > c.max_ref) c.max_ref = fabs(y);
    c.mean_abs += d;
  }
  c.mean_abs /= (double)a.size();
  return c;
}
static Cmp compare_f32(const std::vector<float>& a, const std::vector<float>& b) {
  Cmp c{0, 0, 0};
  for (size_t i = 0; i < a.size(); ++i) {
    const double d = fabs((double)a[i] - b[i]);
    if (d > c.max_abs) c.max_abs = d;
    if (fabs(b[i]) > c.max_ref) c.max_ref = fabs(b[i]);
    c.mean_abs += d;
  }
  c.mean_abs /= (double)a.size();
  return c;
}

int main(int argc, char** argv) {
  const int ncases = (int)(sizeof(cases) / sizeof(cases[0]));
  const int idx = argc > 1 ? atoi(argv[1]) : -1;
  if (idx < 0) { printf("%d\n", ncases); return 0; }
  if (idx >= ncases) return 1;
  const Case& c = cases[idx];
  printf("[case %d] %s\n", idx, c.name);
  const int taps = c.k * c.k;
  const int ho = (c.h - c.hoff + 2 * c.pad - c.dil * (c.k - 1) - 1) / c.stride + 1;
  const int wo = (c.w - c.woff + 2 * c.pad - c.dil * (c.k - 1) - 1) / c.stride + 1;
  const int xc = c.cin + c.xpad, yc = c.cout + c.ypad;
  const size_t xn = (size_t)c.n * c.h * c.w * xc, yn = (size_t)c.n * ho * wo * yc;
  std::vector<__nv_bfloat16> hx(xn), hdy(yn);
  fill_bf16(hx);
  fill_bf16(hdy);
  std::vector<float> hw((size_t)c.cout * taps * c.cin), hb(c.cout);
  for (auto& e : hw) e = frand() * 0.1f;
  for (auto& e : hb) e = frand();

  __nv_bfloat16 *dx, *dy_tc, *dy_ref, *ddy, *ddx_tc, *ddx_ref, *w, *wt;
  float *w32, *bias, *stats, *dw_tc, *dw_ref;
  CK(cudaMalloc(&dx, xn * 2)); CK(cudaMalloc(&dy_tc, yn * 2)); CK(cudaMalloc(&dy_ref, yn * 2));
  CK(cudaMalloc(&ddy, yn * 2)); CK(cudaMalloc(&ddx_tc, xn * 2)); CK(cudaMalloc(&ddx_ref, xn * 2));
  CK(cudaMalloc(&w, hw.size() * 2)); CK(cudaMalloc(&wt, hw.size() * 2)); CK(cudaMalloc(&w32, hw.size() * 4));
  CK(cudaMalloc(&bias, c.cout * 4)); CK(cudaMalloc(&stats, 2 * c.cout * 4));
  CK(cudaMalloc(&dw_tc, hw.size() * 4)); CK(cudaMalloc(&dw_ref, hw.size() * 4));
  CK(cudaMemcpy(dx, hx.data(), xn * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy, hdy.data(), yn * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w32, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(bias, hb.data(), c.cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy_tc, 0x7f, yn * 2)); CK(cudaMemset(dy_ref, 0x7f, yn * 2));
  CK(cudaMemset(ddx_tc, 0x7f, xn * 2)); CK(cudaMemset(ddx_ref, 0x7f, xn * 2));
  CK(cudaMemset(stats, 0, 2 * c.cout * 4)); CK(cudaMemset(dw_tc, 0, hw.size() * 4)); CK(cudaMemset(dw_ref, 0, hw.size() * 4));

  int rc = npp_pack_weight(w32, w, wt, c.cout, taps, c.cin, c.cout, c.cin, NPP_BF16, 0);
  if (rc) { printf("pack rc=%d %s\n", rc, npp_last_error()); return 1; }

  npp_view4 vx{dx, c.n, c.h, c.w, c.cin, (int64_t)c.h * c.w * xc, (int64_t)c.w * xc, xc};
  npp_view4 vy_tc{dy_tc, c.n, ho, wo, c.cout, (int64_t)ho * wo * yc, (int64_t)wo * yc, yc};
  npp_view4 vy_ref = vy_tc; vy_ref.ptr = dy_ref;
  npp_view4 vdy = vy_tc; vdy.ptr = ddy;
  npp_view4 vdx_tc = vx; vdx_tc.ptr = ddx_tc;
  npp_view4 vdx_ref = vx; vdx_ref.ptr = ddx_ref;
  const float* bptr = c.bias ? bias : nullptr;
  int fails = 0;

  // ---------------- fprop
  rc = npp_conv2d_fwd(&vx, w, bptr, &vy_tc, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, stats, 0);
  if (rc) { printf("  fwd rc=%d %s\n", rc, npp_last_error()); return 1; }
  rc = npp_conv2d_direct_fwd(&vx, w, bptr, &vy_ref, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, NPP_BF16, 0);
  if (rc) { printf("  direct fwd rc=%d %s\n", rc, npp_last_error()); return 1; }
  CK(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> a(yn), b(yn);
  CK(cudaMemcpy(a.data(), dy_tc, yn * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), dy_ref, yn * 2, cudaMemcpyDeviceToHost));
  {
    // compare only the view's channels; the padding channels must be untouched (0x7f7f)
    std::vector<__nv_bfloat16> av, bv;
    size_t untouched_bad = 0;
    for (size_t p = 0; p < (size_t)c.n * ho * wo; ++p)
      for (int ch = 0; ch < yc; ++ch) {
        if (ch < c.cout) { av.push_back(a[p * yc + ch]); bv.push_back(b[p * yc + ch]); }
        else if (*reinterpret_cast<uint16_t*>(&a[p * yc + ch]) != 0x7f7f) ++untouched_bad;
      }
    Cmp m = compare_bf16(av, bv);
    const bool ok = m.max_abs <= 0.02 * (m.max_ref + 1e-6) && untouched_bad == 0;
    printf("  fprop : max_abs=%.5f max_ref=%.4f mean_abs=%.6f clobbered=%zu %s\n", m.max_abs, m.max_ref, m.mean_abs,
           untouched_bad, ok ? "PASS" : "FAIL");
    fails += !ok;
    // statistics vs host sums over the tcgen05 output
    std::vector<float> hs(2 * c.cout);
    CK(cudaMemcpy(hs.data(), stats, 2 * c.cout * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int ch = 0; ch < c.cout; ++ch) {
      double s = 0, s2 = 0;
      for (size_t p = 0; p < (size_t)c.n * ho * wo; ++p) {
        const double v = __bfloat162float(a[p * yc + ch]);
        s += v; s2 += v * v;
      }
      worst = fmax(worst, fabs(s - hs[ch]) / (fabs(s) + 1.0));
      worst = fmax(worst, fabs(s2 - hs[c.cout + ch]) / (fabs(s2) + 1.0));
    }
    const bool sok = worst < 1e-3;
    printf("  stats : worst rel err=%.2e %s\n", worst, sok ? "PASS" : "FAIL");
    fails += !sok;
  }

  // ---------------- dgrad
  rc = npp_conv2d_dgrad(&vdy, wt, &vdx_tc, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, 0);
  if (rc) { printf("  dgrad rc=%d %s\n", rc, npp_last_error()); return 1; }
  rc = npp_conv2d_direct_dgrad(&vdy, w, &vdx_ref, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, NPP_BF16, 0);
  if (rc) { printf("  direct dgrad rc=%d %s\n", rc, npp_last_error()); return 1; }
  CK(cudaDeviceSynchronize());
  {
    std::vector<__nv_bfloat16> ga(xn), gb(xn), av, bv;
    CK(cudaMemcpy(ga.data(), ddx_tc, xn * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gb.data(), ddx_ref, xn * 2, cudaMemcpyDeviceToHost));
    size_t untouched_bad = 0;
    for (size_t p = 0; p < (size_t)c.n * c.h * c.w; ++p)
      for (int ch = 0; ch < xc; ++ch) {
        if (ch < c.cin) { av.push_back(ga[p * xc + ch]); bv.push_back(gb[p * xc + ch]); }
        else if (*reinterpret_cast<uint16_t*>(&ga[p * xc + ch]) != 0x7f7f) ++untouched_bad;
      }
    Cmp m = compare_bf16(av, bv);
    const bool ok = m.max_abs <= 0.02 * (m.max_ref + 1e-6) && untouched_bad == 0;
    printf("  dgrad : max_abs=%.5f max_ref=%.4f mean_abs=%.6f clobbered=%zu %s\n", m.max_abs, m.max_ref, m.mean_abs,
           untouched_bad, ok ? "PASS" : "FAIL");
    fails += !ok;
  }

  // ---------------- wgrad (workspace path: split-K partials + reduce kernel; the atomic path is checked below)
  void* wsp = nullptr;
  const int64_t ws_bytes = npp_conv2d_wgrad_workspace_bytes();
  CK(cudaMalloc(&wsp, ws_bytes));
  rc = npp_conv2d_wgrad_ws(&vx, &vdy, dw_tc, c.cout, c.cin, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, wsp,
                           ws_bytes, 0);
  if (rc) { printf("  wgrad rc=%d %s\n", rc, npp_last_error()); return 1; }
  rc = npp_conv2d_direct_wgrad(&vx, &vdy, dw_ref, c.cout, c.cin, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff,
                               NPP_BF16, 0);
  if (rc) { printf("  direct wgrad rc=%d %s\n", rc, npp_last_error()); return 1; }
  CK(cudaDeviceSynchronize());
  {
    std::vector<float> ga(hw.size()), gb(hw.size());
    CK(cudaMemcpy(ga.data(), dw_tc, hw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gb.data(), dw_ref, hw.size() * 4, cudaMemcpyDeviceToHost));
    Cmp m = compare_f32(ga, gb);
    const bool ok = m.max_abs <= 2e-3 * (m.max_ref + 1e-6);
    printf("  wgrad : max_abs=%.5f max_ref=%.4f mean_abs=%.6f %s\n", m.max_abs, m.max_ref, m.mean_abs, ok ? "PASS" : "FAIL");
    fails += !ok;
    // the atomic path (no workspace) must agree too; and a second accumulating call must double the result (+=)
    CK(cudaMemset(dw_tc, 0, hw.size() * 4));
    rc = npp_conv2d_wgrad(&vx, &vdy, dw_tc, c.cout, c.cin, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, 0);
    if (rc) { printf("  wgrad(atomic) rc=%d %s\n", rc, npp_last_error()); return 1; }
    rc = npp_conv2d_wgrad_ws(&vx, &vdy, dw_tc, c.cout, c.cin, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, wsp,
                             ws_bytes, 0);
    if (rc) { printf("  wgrad(ws, accumulate) rc=%d %s\n", rc, npp_last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ga.data(), dw_tc, hw.size() * 4, cudaMemcpyDeviceToHost));
    for (auto& e : gb) e *= 2.f;
    Cmp m2 = compare_f32(ga, gb);
    const bool ok2 = m2.max_abs <= 2e-3 * (m2.max_ref + 1e-6);
    printf("  wgrad : atomic + workspace accumulate == 2x: max_abs=%.5f max_ref=%.4f %s\n", m2.max_abs, m2.max_ref,
           ok2 ? "PASS" : "FAIL");
    fails += !ok2;
  }

  // ---------------- timing of the tcgen05 kernels (warm, 10 iterations each)
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double flops = 2.0 * c.n * ho * wo * (double)c.cout * c.cin * taps;
    float ms;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) npp_conv2d_fwd(&vx, w, bptr, &vy_tc, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, nullptr, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  time  : fprop %.1f us (%.1f TFLOP/s)", ms * 100, flops / (ms * 1e-4) / 1e12);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) npp_conv2d_fwd(&vx, w, bptr, &vy_tc, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, stats, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  +stats %.1f us (%.1f TFLOP/s)", ms * 100, flops / (ms * 1e-4) / 1e12);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) npp_conv2d_dgrad(&vdy, wt, &vdx_tc, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  dgrad %.1f us (%.1f TFLOP/s)", ms * 100, flops / (ms * 1e-4) / 1e12);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) npp_conv2d_wgrad_ws(&vx, &vdy, dw_tc, c.cout, c.cin, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, wsp, ws_bytes, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  wgrad %.1f us (%.1f TFLOP/s)", ms * 100, flops / (ms * 1e-4) / 1e12);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) npp_conv2d_wgrad(&vx, &vdy, dw_tc, c.cout, c.cin, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  wgrad(atomic) %.1f us\n", ms * 100);
    // test-only profile build of the library (-DNPP_C3_PROF): where conv3_kernel's three roles wait
    typedef int (*prof_fn)(unsigned long long*, int);
    prof_fn prof = (prof_fn)dlsym(RTLD_DEFAULT, "npp_debug_c3_prof");
    if (prof != nullptr) {
      unsigned long long v[32];
      prof(nullptr, 1);
      for (int i = 0; i < 10; ++i) npp_conv2d_fwd(&vx, w, bptr, &vy_tc, c.k, c.k, c.stride, c.pad, c.dil, c.hoff, c.woff, stats, 0);
      CK(cudaDeviceSynchronize());
      prof(v, 1);
      if (v[16]) {
        const double u = (double)v[16];
        printf("  epiprof (cycles per 128-row epilogue unit, fprop+stats, %.0f units/launch): staging wait %.0f | barrier 1 %.0f | "
               "ld wait + convert + st.shared %.0f | fence + barrier 2 %.0f | store issue %.0f | statistics %.0f\n",
               u / 10, v[17] / u, v[18] / u, v[19] / u, v[20] / u, v[21] / u, v[22] / u);
      }
      if (v[8]) {
        const double t = (double)v[8];  // tiles over 10 launches
        printf("  c3prof (cycles per 256-pixel tile, fprop+stats): producer total %.0f wait-empty %.0f | weights total %.0f wait %.0f | "
               "mma total %.0f wait A %.0f wait B %.0f wait tmem %.0f issue %.0f commit %.0f | epilogue total %.0f wait acc %.0f | "
               "tiles/launch %.0f\n",
               v[0] / t, v[1] / t, v[9] / t, v[10] / t, v[2] / t, v[3] / t, v[4] / t, v[5] / t, v[11] / t, v[12] / t,
               v[6] / t, v[7] / t, t / 10);
      }
    }
  }
  printf("[case %d] %s\n", idx, fails ? "FAILED" : "OK");
  return fails ? 1 : 0;
}
