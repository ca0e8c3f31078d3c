// SyncBN statistics exchange over NVLink peer memory (replaces the all_gather / all_reduce that
// torch's SyncBatchNorm issues per BatchNorm and direction, augment_lip_sync.py:191).
//
// The messages are tiny (2C..4C floats, <= 16 KB) and there are ~860 of them per training step, so what matters is
// latency, not bandwidth: an NCCL all-reduce costs 12 us (2 GPUs) to 27 us (8 GPUs) each — 23 ms of a 104 ms step at
// 8 GPUs.  Here every rank owns one "communication buffer" that all its peers map (CUDA IPC; NVSwitch gives every GPU
// a direct path to every peer) and one single-block kernel per exchange does a ONE-SHOT all-reduce:
//
//   1. copy the local vector(s) into my staging slot            (local stores + __threadfence_system)
//   2. write my sequence number into every peer's flag word     (st.release.sys over NVLink, one thread per peer)
//   3. spin until every peer's sequence number arrived in MY flag words (ld.acquire.sys on local memory)
//   4. read every peer's staging slot over NVLink and add the vectors in RANK ORDER (bit-identical result on
//      every rank, run-to-run deterministic), write the sums to the destination(s)
//
// Ordering / reuse: all ranks issue the same sequence of exchanges (data-parallel replicas of one program), so a
// device-side counter gives every exchange the same sequence number s on every rank.  Staging slots and flag rows are
// used round-robin (s % RING).  A rank can only pass step 3 of exchange s once every peer has STARTED exchange s,
// i.e. finished reading in exchange s-1; hence nobody is more than one exchange ahead of anybody else and RING >= 2
// slots are enough (4 are kept).  Works unchanged inside CUDA graphs: pointers are static, the counter lives in
// device memory.  A spin that lasts longer than `timeout_ns` (a peer died / ranks diverged) sets the error word of the
// local buffer and falls through instead of hanging the GPU; npp_peer_status() reports it.
#include "common.cuh"

namespace npp {

constexpr int kRing = NPP_PEER_RING;
constexpr int kMaxFloats = NPP_PEER_MAX_FLOATS;
constexpr int kMaxRanks = NPP_PEER_MAX_RANKS;

// Layout of one rank's communication buffer (npp_peer_buffer_bytes()).
struct PeerBuf {
  unsigned int flags[kRing][kMaxRanks];  // flags[slot][p] = last sequence number rank p announced for this slot
  unsigned int seq;                      // exchanges issued by THIS rank (device-side counter)
  unsigned int error;                    // != 0: an exchange timed out (value = its sequence number)
  unsigned int pad[2];
  float staging[kRing][kMaxFloats];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {  // system-scope relaxed load: never served from L1
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerArgs {
  PeerBuf* bufs[kMaxRanks];  // bufs[p] = rank p's buffer as mapped into this process (bufs[rank] = the local one)
  int rank, world;
  const float* src0; float* dst0; int n0;   // first vector
  const float* src1; float* dst1; int n1;   // optional second vector (n1 == 0: none); n0, n1 multiples of 4
  unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(512, 1) peer_allreduce_kernel(const PeerArgs A) {
  __shared__ unsigned int s_seq;
  PeerBuf* me = A.bufs[A.rank];
  if (threadIdx.x == 0) s_seq = me->seq + 1u;
  __syncthreads();
  const unsigned int seq = s_seq;
  const int slot = seq % kRing;
  float* stage = me->staging[slot];
  const int n = A.n0 + A.n1;
  // 1. local vector(s) -> my staging slot
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
    const float4 v = i < A.n0 ? *reinterpret_cast<const float4*>(A.src0 + i) : *reinterpret_cast<const float4*>(A.src1 + (i - A.n0));
    *reinterpret_cast<float4*>(stage + i) = v;
  }
  __threadfence_system();
  __syncthreads();
  // 2. announce, 3. wait — one thread per peer
  if (threadIdx.x < A.world) {
    const int p = threadIdx.x;
    if (p != A.rank) st_release_sys(&A.bufs[p]->flags[slot][A.rank], seq);
    if (p == A.rank) me->seq = seq;
  }
  if (threadIdx.x < A.world && threadIdx.x != A.rank) {
    const unsigned int* f = &me->flags[slot][threadIdx.x];
    const unsigned long long t0 = globaltimer_ns();
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(f) - seq) < 0) {
      if (++spins > 64u) __nanosleep(20);      // the common case (peers a few microseconds apart) never sleeps
      if ((spins & 1023u) == 0 && globaltimer_ns() - t0 > A.timeout_ns) {
        atomicCAS(&me->error, 0u, seq);
        break;
      }
    }
  }
  __syncthreads();
  // 4. sum over ranks in rank order (same order everywhere: identical bits on every rank).  All peer loads of a
  //    thread are issued before the first one is consumed: one NVLink round trip per vector, not one per peer.
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < A.world) v[p] = p == A.rank ? *reinterpret_cast<const float4*>(stage + i) : ld_peer4(A.bufs[p]->staging[slot] + i);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < A.world) { acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w; }
    if (i < A.n0) *reinterpret_cast<float4*>(A.dst0 + i) = acc;
    else *reinterpret_cast<float4*>(A.dst1 + (i - A.n0)) = acc;
  }
}

}  // namespace npp

extern "C" {

int64_t npp_peer_buffer_bytes(void) { return (int64_t)sizeof(npp::PeerBuf); }

int npp_peer_alloc(void** ptr) {
  if (!ptr) return NPP_E_INVALID;
  cudaError_t e = cudaMalloc(ptr, sizeof(npp::PeerBuf));
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, sizeof(npp::PeerBuf));
  if (e != cudaSuccess) { npp::set_error("npp_peer_alloc", e); return NPP_E_CUDA; }
  return NPP_OK;
}

int npp_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { npp::set_error("npp_peer_free", e); return NPP_E_CUDA; }
  return NPP_OK;
}

int npp_peer_export(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == NPP_PEER_HANDLE_BYTES, "IPC handle size");
  if (!ptr || !handle64) return NPP_E_INVALID;
  cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), ptr);
  if (e != cudaSuccess) { npp::set_error("cudaIpcGetMemHandle", e); return NPP_E_CUDA; }
  return NPP_OK;
}

int npp_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return NPP_E_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { npp::set_error("cudaIpcOpenMemHandle", e); return NPP_E_CUDA; }
  return NPP_OK;
}

int npp_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { npp::set_error("cudaIpcCloseMemHandle", e); return NPP_E_CUDA; }
  return NPP_OK;
}

int npp_peer_allreduce(const npp_peer_comm* comm, const float* src0, float* dst0, int n0, const float* src1,
                       float* dst1, int n1, npp_stream_t stream) {
  if (!comm || comm->world < 1 || comm->world > NPP_PEER_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world)
    return NPP_E_INVALID;
  if (!src0 || !dst0 || n0 <= 0 || n0 % 4 || n1 < 0 || n1 % 4 || (n1 && (!src1 || !dst1))) return NPP_E_INVALID;
  if (n0 + n1 > NPP_PEER_MAX_FLOATS) return NPP_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(dst0) | reinterpret_cast<uintptr_t>(src1) |
       reinterpret_cast<uintptr_t>(dst1)) % 16)
    return NPP_E_INVALID;
  npp::PeerArgs A;
  for (int p = 0; p < NPP_PEER_MAX_RANKS; ++p) A.bufs[p] = p < comm->world ? static_cast<npp::PeerBuf*>(comm->bufs[p]) : nullptr;
  for (int p = 0; p < comm->world; ++p)
    if (!A.bufs[p]) return NPP_E_INVALID;
  A.rank = comm->rank; A.world = comm->world;
  A.src0 = src0; A.dst0 = dst0; A.n0 = n0;
  A.src1 = src1; A.dst1 = dst1; A.n1 = n1;
  A.timeout_ns = comm->timeout_ms > 0 ? (unsigned long long)comm->timeout_ms * 1000000ull : 30000000000ull;
  npp::peer_allreduce_kernel<<<1, 512, 0, npp::as_stream(stream)>>>(A);
  NPP_CHECK_LAUNCH("peer_allreduce_kernel");
  return NPP_OK;
}

int npp_peer_status(const npp_peer_comm* comm, unsigned int* seq, unsigned int* error) {
  if (!comm || !comm->bufs[comm->rank]) return NPP_E_INVALID;
  unsigned int v[2] = {0, 0};
  const npp::PeerBuf* me = static_cast<const npp::PeerBuf*>(comm->bufs[comm->rank]);
  cudaError_t e = cudaMemcpy(v, &me->seq, sizeof v, cudaMemcpyDeviceToHost);  // synchronises: a host-side health check
  if (e != cudaSuccess) { npp::set_error("npp_peer_status", e); return NPP_E_CUDA; }
  if (seq) *seq = v[0];
  if (error) *error = v[1];
  return NPP_OK;
}

}  // extern "C"
