// Multi-tensor Adam: one launch updates every parameter of the model (the reference's torch.optim.Adam call,
// augment_lip_sync.py:210, issues thousands of small foreach kernels per step; this is the same arithmetic —
// torch/optim/adam.py single-tensor path, no amsgrad — over a device table of tensor descriptors).
#include "common.cuh"

namespace npp {

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
  float lr;
  float wd;
  int64_t* step;  // this tensor's own step counter (torch keeps one per parameter: a parameter whose gradient is
                  // None in some step is skipped and its bias correction lags behind)
};
static_assert(sizeof(AdamTensor) == 56, "table layout is part of the ABI (npp_b200/optim.py mirrors it)");

__global__ void __launch_bounds__(256)
adam_kernel(const AdamTensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
            const int* __restrict__ chunk_index, int chunk_elems, float beta1, float beta2, float eps) {
  pdl_wait();
  const AdamTensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t begin = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
  int64_t end = begin + chunk_elems;
  if (end > t.n) end = t.n;
  const double k = (double)(*t.step + 1);
  const float bc1 = (float)(1.0 - pow((double)beta1, k));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, k));
  const float step_size = t.lr / bc1;
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    float g = t.g[i];
    const float p = t.p[i];
    if (t.wd != 0.f) g = fmaf(t.wd, p, g);
    const float m = fmaf(beta1, t.m[i], (1.f - beta1) * g);
    const float v = fmaf(beta2, t.v[i], (1.f - beta2) * g * g);
    t.m[i] = m;
    t.v[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    t.p[i] = p - step_size * (m / denom);
  }
}

__global__ void adam_bump_step_kernel(const AdamTensor* __restrict__ tensors, int ntensors) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ntensors) *tensors[i].step += 1;
}

}  // namespace npp

using namespace npp;

extern "C" {

int npp_adam_step(const void* tensor_table, int ntensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                  int nchunks, int chunk_elems, float beta1, float beta2, float eps, npp_stream_t s) {
  if (!tensor_table || !chunk_tensor || !chunk_index || ntensors < 0 || nchunks < 0 || chunk_elems <= 0)
    return NPP_E_INVALID;
  if (nchunks == 0) return NPP_OK;
  cudaStream_t st = as_stream(s);
  NPP_LAUNCH((adam_kernel), nchunks, 256, 0, st, static_cast<const AdamTensor*>(tensor_table), chunk_tensor, chunk_index,
                                       chunk_elems, beta1, beta2, eps);
  NPP_CHECK_LAUNCH("adam_kernel");
  NPP_LAUNCH((adam_bump_step_kernel), (ntensors + 255) / 256, 256, 0, st, static_cast<const AdamTensor*>(tensor_table), ntensors);
  NPP_CHECK_LAUNCH("adam_bump_step_kernel");
  return NPP_OK;
}

}  // extern "C"
