// SyncBN statistics exchange over NVLink peer memory (replaces the all_gather / all_reduce that
// torch's SyncBatchNorm issues per BatchNorm and direction, augment_lip_sync.py:191).
//
// The messages are tiny (2C..4C floats, <= 16 KB) and there are ~860 of them per training step, so what matters is
// latency, not bandwidth: an NCCL all-reduce costs 12 us (2 GPUs) to 27 us (8 GPUs) each — 23 ms of a 104 ms step at
// 8 GPUs.  Here every rank owns one "communication buffer" that all its peers map (CUDA IPC; NVSwitch gives every GPU
// a direct path to every peer) and one single-block kernel per exchange does a ONE-SHOT all-reduce in push form:
//
//   1. every rank writes its vector, element by element as 8-byte {value, sequence number} pairs, into its own row
//      of EVERY peer's buffer (NVLink stores; value and tag are one store, so they become visible together);
//   2. every rank polls the rows of its OWN buffer until all elements carry the current sequence number and adds the
//      rows in RANK ORDER (bit-identical result on every rank, run-to-run deterministic).
//   No fence, no flag word, no remote read: one NVLink one-way latency after the slowest peer has started.
//
// Ordering / reuse: all ranks issue the same sequence of exchanges (data-parallel replicas of one program), so a
// device-side counter gives every exchange the same sequence number s on every rank.  Rows are used round-robin
// (s % RING).  A rank finishes exchange s only after every peer has STARTED exchange s (it needs their data), and a
// peer starts s only after it finished reading the rows of s-1; hence nobody is more than one exchange ahead of anybody
// else and RING >= 2 slots are enough (4 are kept).  Works unchanged inside CUDA graphs: pointers are static, the
// counter lives in device memory.  A poll that lasts longer than `timeout_ns` (a peer died / ranks diverged) sets the
// error word of the local buffer and falls through instead of hanging the GPU; npp_peer_status() reports it.
#include "common.cuh"

namespace npp {

constexpr int kRing = NPP_PEER_RING;
constexpr int kMaxFloats = NPP_PEER_MAX_FLOATS;
constexpr int kMaxRanks = NPP_PEER_MAX_RANKS;

// Layout of one rank's communication buffer (npp_peer_buffer_bytes()): for every ring slot and every SENDER rank a row of
// (value, sequence) pairs.  Rank r writes its vector into row [slot][r] of EVERY peer's buffer; nobody ever reads
// remote memory.
struct PeerBuf {
  unsigned int seq;      // exchanges issued by THIS rank (device-side counter)
  unsigned int error;    // != 0: an exchange timed out (value = its sequence number)
  unsigned int pad[2];
  uint2 slots[kRing][kMaxRanks][kMaxFloats];   // .x = fp32 bits, .y = sequence number of the exchange that wrote it
};

__device__ __forceinline__ void st_pair_sys(uint2* p, unsigned int v, unsigned int tag) {
  // one 8-byte store: value and tag become visible together (the "LL" trick: no fence, no separate flag)
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_pair_sys(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
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
  const float* src1; float* dst1; int n1;   // optional second vector (n1 == 0: none)
  unsigned long long timeout_ns;
};

// One-shot all-reduce, PUSH form (v2; v1 pulled the peers' staging buffers after a flag handshake: 10.6 us per exchange
// at 8 ranks, profiles/r02_peer_allreduce_latency_8gpu_v1_pull.txt — a fence with remote acknowledgement on the sender
// plus a remote read round trip on the receiver).  Here every element travels as ONE 8-byte store {value, sequence
// number} into every peer's buffer: no fence, no flag, no remote read — the receiver polls its LOCAL rows until every
// element carries the current sequence number (one NVLink one-way latency after the slowest peer started), then adds
// the rows in rank order.
__global__ void __launch_bounds__(512, 1) peer_allreduce_kernel(const PeerArgs A) {
  pdl_wait();
  __shared__ unsigned int s_seq;
  PeerBuf* me = A.bufs[A.rank];
  if (threadIdx.x == 0) s_seq = me->seq + 1u;
  __syncthreads();
  const unsigned int seq = s_seq;
  const int slot = seq % kRing;
  const int n = A.n0 + A.n1;
  if (threadIdx.x == 0) me->seq = seq;
  // 1. push my vector(s) into row [slot][rank] of every peer (my own contribution is read from src directly)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = i < A.n0 ? A.src0[i] : A.src1[i - A.n0];
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < A.world && p != A.rank) st_pair_sys(&A.bufs[p]->slots[slot][A.rank][i], __float_as_uint(v), seq);
  }
  // 2. collect: rows of my own buffer.  All rows of an element are loaded before the first tag is inspected (the
  //    loads overlap); only rows whose tag is still old are polled again.  Summed in rank order (same order on every
  //    rank: bit-identical sums).
  const unsigned long long t0 = globaltimer_ns();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint2 e[kMaxRanks];
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < A.world && p != A.rank) e[p] = ld_pair_sys(&me->slots[slot][p][i]);
    unsigned int spins = 0;
    bool missing = true;
    while (missing) {
      missing = false;
#pragma unroll
      for (int p = 0; p < kMaxRanks; ++p) {
        if (p < A.world && p != A.rank && e[p].y != seq) {
          e[p] = ld_pair_sys(&me->slots[slot][p][i]);
          if (e[p].y != seq) missing = true;
        }
      }
      if (missing) {
        if (++spins > 64u) __nanosleep(20);
        if ((spins & 1023u) == 0 && globaltimer_ns() - t0 > A.timeout_ns) {
          atomicCAS(&me->error, 0u, seq);
          break;
        }
      }
    }
    const float mine = i < A.n0 ? A.src0[i] : A.src1[i - A.n0];
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < A.world) acc += (p == A.rank) ? mine : __uint_as_float(e[p].x);
    // dst may alias src: element i of src is only ever read by this thread, before this store
    if (i < A.n0) A.dst0[i] = acc; else A.dst1[i - A.n0] = acc;
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
  if (!src0 || !dst0 || n0 <= 0 || n1 < 0 || (n1 && (!src1 || !dst1))) return NPP_E_INVALID;
  if (n0 + n1 > NPP_PEER_MAX_FLOATS) return NPP_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(dst0) | reinterpret_cast<uintptr_t>(src1) |
       reinterpret_cast<uintptr_t>(dst1)) % 4)
    return NPP_E_INVALID;
  npp::PeerArgs A;
  for (int p = 0; p < NPP_PEER_MAX_RANKS; ++p) A.bufs[p] = p < comm->world ? static_cast<npp::PeerBuf*>(comm->bufs[p]) : nullptr;
  for (int p = 0; p < comm->world; ++p)
    if (!A.bufs[p]) return NPP_E_INVALID;
  A.rank = comm->rank; A.world = comm->world;
  A.src0 = src0; A.dst0 = dst0; A.n0 = n0;
  A.src1 = src1; A.dst1 = dst1; A.n1 = n1;
  A.timeout_ns = comm->timeout_ms > 0 ? (unsigned long long)comm->timeout_ms * 1000000ull : 30000000000ull;
  NPP_LAUNCH((npp::peer_allreduce_kernel), 1, 512, 0, npp::as_stream(stream), A);
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
