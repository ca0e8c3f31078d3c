// Shared bilinear index arithmetic (ATen area_pixel_compute_source_index, fp32) for resample.cu and loss.cu.
#pragma once
#include "common.cuh"

namespace npp {

struct Axis {
  int in, out;
  float scale;  // source index per output index
  int align;
};

static Axis make_axis(int in, int out, int align, double scale_factor) {
  Axis a;
  a.in = in; a.out = out; a.align = align;
  if (align)
    a.scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  else
    a.scale = scale_factor > 0 ? (float)(1.0 / scale_factor) : (float)in / (float)out;
  return a;
}

__device__ __forceinline__ void bilinear_taps(const Axis& a, int o, int& i0, int& i1, float& l0, float& l1) {
  float src;
  if (a.align) {
    src = a.scale * (float)o;
  } else {
    src = a.scale * ((float)o + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
  }
  i0 = (int)src;
  if (i0 > a.in - 1) i0 = a.in - 1;
  i1 = i0 + ((i0 < a.in - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// candidate output range [lo, hi] whose taps can include input index i (conservative)
__device__ __forceinline__ void bilinear_range(const Axis& a, int i, int& lo, int& hi) {
  if (a.scale <= 0.f) { lo = 0; hi = a.out - 1; return; }
  const float inv = 1.f / a.scale;
  const float off = a.align ? 0.f : 0.5f;
  lo = (int)floorf(((float)i - 1.f + off) * inv - off) - 2;
  hi = (int)ceilf(((float)i + 1.f + off) * inv - off) + 2;
  if (lo < 0) lo = 0;
  if (hi > a.out - 1) hi = a.out - 1;
}

}  // namespace npp
