// Shared device / host helpers for libnpp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/npp_b200.h"

namespace npp {

void set_error(const char* what, cudaError_t e);
void set_error_str(const char* what);
void count_launch(int n);  // kernel-launch counter behind npp_launch_count()

#define NPP_CHECK_LAUNCH(name)                          \
  do {                                                  \
    cudaError_t e__ = cudaGetLastError();               \
    if (e__ != cudaSuccess) {                           \
      ::npp::set_error(name, e__);                      \
      return NPP_E_CUDA;                                \
    }                                                   \
    ::npp::count_launch(1);                             \
  } while (0)

static inline cudaStream_t as_stream(npp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A training step is ~3 600 mostly short kernels in one stream / one CUDA graph; between two dependent kernels the
// GPU otherwise drains, flushes and only then launches the next grid (~2-2.5 us per launch, measured on the tiny
// fold / fill kernels).  Every kernel of this library is launched with programmatic stream serialization: the next
// grid may be scheduled while the previous one is still finishing, and blocks at `pdl_wait()` — the first statement
// of every kernel (after the barrier / TMEM / tensor-map set-up in the tcgen05 kernels) — until the previous grid has
// completed and its memory is visible.  Every thread of every block executes the wait before it touches global
// memory or exits, so completion stays transitive along the stream (C after B after A).  The launch attribute is
// captured into CUDA graphs as a programmatic edge.  NPP_PDL=0 launches everything fully serialized (A/B, bisecting).
// (An early `griddepcontrol.launch_dependents` right after the wait — next grid's blocks resident while this grid's last
// wave drains — was measured SLOWER on B200: 101.7 vs 99.7 ms per training step, profiles/r02_pdl_ab.txt; the waiting
// blocks take registers / shared memory away from the draining grid.  Wait-only PDL: 99.9 vs 100.15 ms.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in NPP_CHECK_LAUNCH
}
#define NPP_LAUNCH(kernel, grid, block, smem, st, ...) ::npp::launch_k(kernel, grid, block, smem, st, ##__VA_ARGS__)

// ---- dtype traits: storage T, math in fp32 ---------------------------------------------------
template <typename T> struct VecOf;  // 16-byte vector of T
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte load/store of VecOf<T>::N elements converted to/from fp32 registers
template <typename T> struct Pack;
template <> struct Pack<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
};
template <> struct Pack<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      u[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  }
};

// 16-byte raw load (kept packed in 4 registers until it is consumed: more loads in flight per thread)
template <typename T>
__device__ __forceinline__ uint4 ldraw(const T* p) { return *reinterpret_cast<const uint4*>(p); }

// ---- view checks -------------------------------------------------------------------------------
static inline bool view_ok(const npp_view4* v, int dtype) {
  if (!v || !v->ptr || v->n <= 0 || v->h <= 0 || v->w <= 0 || v->c <= 0) return false;
  const int vec = dtype == NPP_BF16 ? 8 : 4;
  const int esz = dtype == NPP_BF16 ? 2 : 4;
  if (v->c % vec) return false;
  if (v->sw % vec || v->sh % vec || v->sn % vec) return false;
  if (reinterpret_cast<uintptr_t>(v->ptr) % 16) return false;
  (void)esz;
  return true;
}
static inline bool same_shape(const npp_view4* a, const npp_view4* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();

}  // namespace npp
