// Dense convolution (fprop / dgrad / wgrad) as implicit GEMM on the sm_100a tensor cores.
//
//   fprop / dgrad : D[pixels, Cout] = sum_taps A_tap[pixels, Cin] * W_tap[Cout, Cin]^T
//       A_tap is the activation box shifted by the tap offset, fetched by ONE 4-D TMA per
//       (tap, 64-channel chunk); out-of-bounds coordinates are zero-filled by the TMA unit, which
//       is exactly the convolution's zero padding.  Stride-2 convolutions address four "phase"
//       tensor maps (even/odd rows x even/odd columns) so every tap is still a dense box.
//       Both operands are K-major SWIZZLE_128B tiles consumed by tcgen05.mma (M=128, N=BN, K=16)
//       with the fp32 accumulator in TMEM; a 4-warp epilogue drains TMEM (tcgen05.ld), adds the
//       bias, rounds to bf16, accumulates BatchNorm batch statistics and TMA-stores the tile.
//       Persistent CTAs (one per SM), STAGES-deep smem ring, two TMEM accumulator stages so the
//       epilogue of tile i overlaps the MMAs of tile i+1.
//   wgrad : dW_tap[Cout, Cin] = sum_pixels dY[pixels, Cout]^T * X_tap[pixels, Cin]   (written in torch OIHW order)
//       Same boxes, but now the pixel axis is the GEMM K axis, i.e. both operands are MN-major
//       SWIZZLE_128B tiles (tcgen05 handles the transpose in the descriptor).  Split-K over
//       pixels; partial sums are combined with red.global.add.f32.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (TMEM lane quarter = warp & 3).
#include "common.cuh"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace npp {
namespace tc {

using namespace ptx;

constexpr int kMaxTaps = 9;
constexpr int BM = 128;       // pixels per tile (UMMA M)
constexpr int BK = 64;        // channels per k-block: 64 bf16 = one 128-byte swizzle row
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

struct TapTable {
  int8_t map[kMaxTaps];   // which activation tensor map (phase) the tap reads
  int8_t dh[kMaxTaps];    // row / column offset of the box in that map's coordinates
  int8_t dw[kMaxTaps];
  int8_t btap[kMaxTaps];  // index into the weight tensor's tap dimension
};

struct Maps {
  CUtensorMap a[4];  // activation (phase) maps: dims (C, W, H, N)
  CUtensorMap b;     // weights: dims (K, taps, rows)
  CUtensorMap d;     // output: dims (C, W, H, N)
};

struct Geom {
  int num_taps, kc_blocks;
  int tw, th, tn;                 // pixel tile box, tw*th*tn == 128 (fprop) or 64 (wgrad K block)
  int tiles_w, tiles_h, tiles_n;  // pixel tiles per dimension
  int W, H, N;                    // output (fprop) pixel-space extents, for masking
  int n_blocks;                   // Cout blocks of BN
  int cout;
  int num_ptiles;                 // pixel tiles; total work items = num_ptiles * n_blocks
  int k_last;                     // K16 steps of the last 64-channel block that hold real channels (1..4): the
                                  // zero-filled rest of a partial block is not multiplied at all
};

template <int BN>
struct FpropCfg {
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = 2 * BM * 128;  // two bf16 staging slabs (128 rows x 64 ch)
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int BAR_BYTES = 1024;
  static constexpr int SMEM = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024 /*align slack*/;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {32..256}
};

// =================================================================================================
// fprop / dgrad kernel
// =================================================================================================
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ Maps maps, const Geom g, const TapTable taps,
                 const float* __restrict__ bias, float* __restrict__ stats) {
  using Cfg = FpropCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t smem_out = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = smem_out + Cfg::OUT_BYTES;
  const uint32_t full_bar = bar_base;                   // STAGES x 8 B
  const uint32_t empty_bar = bar_base + 8 * STAGES;     // STAGES x 8 B
  const uint32_t tfull_bar = bar_base + 16 * STAGES;    // 2 x 8 B
  const uint32_t tempty_bar = tfull_bar + 16;           // 2 x 8 B
  const uint32_t tmem_slot = tempty_bar + 16;           // 4 B
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic alias of smem_base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.a[0]);
    prefetch_tensormap(&maps.b);
    prefetch_tensormap(&maps.d);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above (tensor-map prefetch, mbarrier init, TMEM allocation) overlaps the tail of the previous kernel
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int num_k = g.num_taps * g.kc_blocks;
  // Static schedule: CTA b owns output-channel block nb = b % n_blocks and every pt_step-th pixel tile; CTAs
  // b..b+n_blocks-1 work on the same pixel tile at the same time, so its activation boxes are L2 hits.
  const int nb = blockIdx.x % g.n_blocks;
  const int pt_start = blockIdx.x / g.n_blocks;
  const int pt_step = gridDim.x / g.n_blocks;
  constexpr int SLABS = BN >= 64 ? BN / 64 : 1;
  constexpr int SLAB_COLS = BN >= 64 ? 64 : BN;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (the whole warp runs the
    // loop converged and one elected lane issues: with `if (lane == 0)` around the loop the compiler wraps every
    // TMA / MMA instruction, whose operands live in uniform registers, in a divergence "waterfall" loop)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int ptile = pt_start; ptile < g.num_ptiles; ptile += pt_step) {
        int pt = ptile;
        const int w0 = (pt % g.tiles_w) * g.tw;
        pt /= g.tiles_w;
        const int h0 = (pt % g.tiles_h) * g.th;
        const int n0 = (pt / g.tiles_h) * g.tn;
        for (int t = 0; t < g.num_taps; ++t) {
          const CUtensorMap* ma = &maps.a[taps.map[t]];
          const int cw = w0 + taps.dw[t], ch = h0 + taps.dh[t], bt = taps.btap[t];
          for (int kc = 0; kc < g.kc_blocks; ++kc) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
              tma_load_4d(smem_a + stage * Cfg::A_BYTES, ma, full_bar + 8 * stage, kc * BK, cw, ch, n0);
              tma_load_3d(smem_b + stage * Cfg::B_BYTES, &maps.b, full_bar + 8 * stage, kc * BK, bt, nb * BN);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane)
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int kc = 0;  // 64-channel block index of the current k-block (the tap loop is the outer one)
      for (int ptile = pt_start; ptile < g.num_ptiles; ptile += pt_step) {
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_a + stage * Cfg::A_BYTES, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(smem_b + stage * Cfg::B_BYTES, 16, 1024);
          const int ks = (kc + 1 == g.kc_blocks) ? g.k_last : BK / 16;
          if (++kc == g.kc_blocks) kc = 0;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr>>4)
              if (k < ks) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar + 8 * stage);
            if (kb == num_k - 1) umma_commit(tfull_bar + 8 * acc);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;  // tile row == TMEM lane
    const int etid = threadIdx.x - kEpiWarp0 * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sbuf = 0;
    // BatchNorm statistics are accumulated in registers over all tiles of this CTA (same channel block nb for
    // every tile) and flushed once at the end: per-channel atomics drop from one per tile to one per CTA —
    // same-address atomics serialise in L2 and were the bottleneck of the first version.
    float acc_s[SLABS][8], acc_q[SLABS][8];
#pragma unroll
    for (int sl = 0; sl < SLABS; ++sl)
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc_s[sl][k] = 0.f; acc_q[sl][k] = 0.f; }
    for (int ptile = pt_start; ptile < g.num_ptiles; ptile += pt_step) {
      int pt = ptile;
      const int w0 = (pt % g.tiles_w) * g.tw;
      pt /= g.tiles_w;
      const int h0 = (pt % g.tiles_h) * g.th;
      const int n0 = (pt / g.tiles_h) * g.tn;
      // ragged tiles (rows that fall outside the tensor) are masked out of the statistics
      const bool tile_full = (w0 + g.tw <= g.W) && (h0 + g.th <= g.H) && (n0 + g.tn <= g.N);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll
      for (int slab = 0; slab < SLABS; ++slab) {
        const int co_base = nb * BN + slab * 64;
        if (co_base >= g.cout) continue;  // uniform across the CTA
        // the TMA store that last read this staging buffer must have finished reading it
        if (etid == 0) tma_store_wait_read<1>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint8_t* stg = smem_gen + (smem_out - smem_base) + sbuf * (BM * 128);
#pragma unroll
        for (int half = 0; half < SLAB_COLS / 32; ++half) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(acc * BN + slab * 64 + half * 32);
          tmem_ld_32x32(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // four 16-byte chunks (8 channels each)
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = half * 32 + j * 8 + e * 2;
              float v0 = __uint_as_float(r[j * 8 + e * 2]);
              float v1 = __uint_as_float(r[j * 8 + e * 2 + 1]);
              if (bias != nullptr) {
                const int c0 = co_base + col;
                v0 += (c0 < g.cout) ? __ldg(bias + c0) : 0.f;
                v1 += (c0 + 1 < g.cout) ? __ldg(bias + c0 + 1) : 0.f;
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              pk[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            const int chunk = half * 4 + j;
            *reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          tma_store_4d(&maps.d, smem_out + sbuf * (BM * 128), co_base, w0, h0, n0);
          tma_store_commit();
        }
        if (stats != nullptr) {
          // Per-channel sum / sum of squares of the bf16-rounded tile.  Epilogue warp q owns the 16-byte chunks
          // {2q, 2q+1} (8 channels each); its 32 lanes are (chunk, 8-row group) pairs, each lane adds up 8 rows with
          // LDS.128 (rows visited in a lane-skewed order so the 128B swizzle keeps the 8 lanes of a phase on
          // different banks) into its register accumulators.
          const int c2 = lane >> 4, rg = lane & 15;
          const int chunk = 2 * q + c2;
          if (chunk * 8 < SLAB_COLS) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = rg * 8 + ((i + rg) & 7);
              bool ok = true;
              if (!tile_full) {
                const int rw = rr % g.tw, rh = (rr / g.tw) % g.th, rn = rr / (g.tw * g.th);
                ok = (w0 + rw < g.W) && (h0 + rh < g.H) && (n0 + rn < g.N);
              }
              const uint4 u = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((chunk ^ (rr & 7)) << 4));
              if (ok) {
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float v0 = __uint_as_float(uu[e] << 16), v1 = __uint_as_float(uu[e] & 0xffff0000u);
                  acc_s[slab][2 * e] += v0; acc_q[slab][2 * e] = fmaf(v0, v0, acc_q[slab][2 * e]);
                  acc_s[slab][2 * e + 1] += v1; acc_q[slab][2 * e + 1] = fmaf(v1, v1, acc_q[slab][2 * e + 1]);
                }
              }
            }
          }
        }
        sbuf ^= 1;
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (stats != nullptr) {
      // fold the 16 row groups of each half-warp, then one atomic per (CTA, channel, moment)
      const int c2 = lane >> 4, rg = lane & 15;
      const int chunk = 2 * q + c2;
#pragma unroll
      for (int sl = 0; sl < SLABS; ++sl) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float a = acc_s[sl][k], b = acc_q[sl][k];
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          const int co = nb * BN + sl * 64 + chunk * 8 + k;
          if (rg == 0 && chunk * 8 < SLAB_COLS && co < g.cout) {
            atomicAdd(stats + co, a);
            atomicAdd(stats + g.cout + co, b);
          }
        }
      }
    }
    if (etid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// Second-generation epilogue: EIGHT warps (256 threads) drain the accumulator.
//
// ncu on the first kernel above: for the small-K layers (1x1 convs, C = 32/64 cells — most of the 822 launches of a
// step) the MMA loop of a tile takes a few hundred cycles and the tile time is the epilogue's: four warps, one per
// SM sub-partition, each a serial chain tcgen05.ld -> wait -> convert -> st.shared -> barrier -> TMA store ->
// statistics with nothing to overlap it (1x1 128->128 @96^2 x32: 35 us without / 66 us with statistics against a
// 23 us HBM floor).  Here two warps share every TMEM lane quarter: warp (q, hf) owns rows 32q..32q+31 and the
// 32-column half hf of each 64-column slab, the TMEM load is issued before the staging-buffer barrier so its latency
// overlaps the wait, the accumulator is handed back to the MMA warp as soon as the last load has landed (before the
// statistics of the last slab), and the statistics pass is spread over 256 threads (4 rows each instead of 8).
// =================================================================================================
#ifdef NPP_C3_PROF
// Test-only build (tests/csrc/_bin/prof): where does conv3_kernel wait?  Cycle counters summed over all CTAs:
// 0 producer total, 1 producer wait A-empty, 2 MMA total, 3 MMA wait A-full, 4 MMA wait B-full, 5 MMA wait TMEM-empty,
// 6 epilogue total (thread 128), 7 epilogue wait TMEM-full, 8 tiles, 9 weight producer total, 10 weight producer wait,
// 11 MMA issue blocks (8 tcgen05.mma each), 12 tcgen05.commit;
// epilogue thread 128, per 128-row unit (epi_unit): 16 units, 17 wait for the staging slab (TMA store read), 18 first
// barrier, 19 tcgen05.ld wait + convert + st.shared, 20 proxy fence + second barrier, 21 store issue, 22 statistics
__device__ unsigned long long g_c3_prof[32];
#define C3P_DECL(v) long long v = 0
#define C3P_WAIT(v, stmt) do { const long long _t = clock64(); stmt; v += clock64() - _t; } while (0)
#define C3P_ADD(i, v) atomicAdd(&g_c3_prof[i], (unsigned long long)(v))
#define C3P_STAMP(t) const long long t = clock64()
#else
#define C3P_DECL(v)
#define C3P_WAIT(v, stmt) stmt
#define C3P_ADD(i, v)
#define C3P_STAMP(t)
#endif

constexpr int kThreads2 = 384;  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue

struct EpiMask {   // which rows of a 128-row unit are real pixels (statistics must skip padding rows)
  int tw, th;      // pixel box of the whole tile
  int W, H, N;     // tensor extents
  int w0, h0, n0;  // tile origin
  int p_off;       // linear index (in the tile's box) of the unit's first row
  bool full;
};

// BatchNorm sums of one 16-byte chunk (8 channels) of a staging slab: lane l adds up rows 4l..4l+3, visited in a
// lane-skewed order so that the 8 lanes of an LDS.128 phase hit 8 different swizzle slots.  All four loads are issued
// before the first use and rows outside the tensor are masked by a select, not a branch (the first version branched
// per row, which serialised the four shared-memory round trips: ~620 cycles per unit).
template <bool MASKED>
__device__ __forceinline__ void epi_chunk_sums_t(const uint8_t* __restrict__ stg, const int chunk, const int lane,
                                                 const EpiMask& mk, float* __restrict__ as, float* __restrict__ aq) {
  uint4 u[4];
  bool ok[4];
  uint32_t ad[4];
  const uint32_t stg_a = smem_u32(stg);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = lane * 4 + ((i + (lane >> 1)) & 3);
    ad[i] = stg_a + rr * 128 + ((chunk ^ (rr & 7)) << 4);
    ok[i] = true;
    if (MASKED) {
      const int p = mk.p_off + rr;
      const int pw = p % mk.tw, ph = (p / mk.tw) % mk.th, pn = p / (mk.tw * mk.th);
      ok[i] = (mk.w0 + pw < mk.W) && (mk.h0 + ph < mk.H) && (mk.n0 + pn < mk.N);
    }
  }
  // one statement: the four loads are issued back to back into four distinct register quads (left to itself the
  // compiler recycles one quad and serialises load -> use -> load)
  asm volatile(
      "ld.shared.v4.u32 {%0, %1, %2, %3}, [%16];\n\t"
      "ld.shared.v4.u32 {%4, %5, %6, %7}, [%17];\n\t"
      "ld.shared.v4.u32 {%8, %9, %10, %11}, [%18];\n\t"
      "ld.shared.v4.u32 {%12, %13, %14, %15}, [%19];"
      : "=r"(u[0].x), "=r"(u[0].y), "=r"(u[0].z), "=r"(u[0].w), "=r"(u[1].x), "=r"(u[1].y), "=r"(u[1].z), "=r"(u[1].w),
        "=r"(u[2].x), "=r"(u[2].y), "=r"(u[2].z), "=r"(u[2].w), "=r"(u[3].x), "=r"(u[3].y), "=r"(u[3].z), "=r"(u[3].w)
      : "r"(ad[0]), "r"(ad[1]), "r"(ad[2]), "r"(ad[3])
      : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t uu[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 v = make_float2(__uint_as_float(uu[e] << 16), __uint_as_float(uu[e] & 0xffff0000u));
      if (MASKED && !ok[i]) v = make_float2(0.f, 0.f);
      const float2 s2 = __fadd2_rn(make_float2(as[2 * e], as[2 * e + 1]), v);
      const float2 q2 = __ffma2_rn(v, v, make_float2(aq[2 * e], aq[2 * e + 1]));
      as[2 * e] = s2.x; as[2 * e + 1] = s2.y;
      aq[2 * e] = q2.x; aq[2 * e + 1] = q2.y;
    }
  }
}

__device__ __forceinline__ void epi_chunk_sums(const uint8_t* __restrict__ stg, const int chunk, const int lane,
                                               const EpiMask& mk, float* __restrict__ as, float* __restrict__ aq) {
  if (mk.full) epi_chunk_sums_t<false>(stg, chunk, lane, mk, as, aq);   // uniform across the CTA
  else epi_chunk_sums_t<true>(stg, chunk, lane, mk, as, aq);
}

// One unit = 128 rows (TMEM lanes) x one 64-column slab (SLAB_COLS = 32 for the BN = 32 kernels) of the accumulator:
// TMEM -> (+bias) -> bf16 -> swizzled staging slab -> TMA store; BatchNorm sums of the rounded values.
template <int SLAB_COLS>
__device__ __forceinline__ void epi_unit(const uint32_t taddr, uint8_t* stg, const uint32_t stg_s,
                                         const CUtensorMap* md, const int co_base, const int c_w, const int c_h,
                                         const int c_n, const float* __restrict__ bias, const int cout,
                                         const uint32_t release_bar, const bool want_stats, const EpiMask& mk,
                                         float (&as)[8], float (&aq)[8], const int q, const int hf, const int ew,
                                         const int lane, const int etid) {
  const bool has_cols = hf * 32 < SLAB_COLS;  // warp-uniform
  const int row = q * 32 + lane;
  uint32_t r[32];
  C3P_STAMP(e_t0);
  if (has_cols) tmem_ld_32x32(taddr, r);
  // the TMA store that last read this staging slab (two units ago) must have finished reading it
  if (etid == 0) tma_store_wait_read<1>();
  C3P_STAMP(e_t1);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  C3P_STAMP(e_t2);
  if (has_cols) {
    tmem_ld_wait_regs(r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // four 16-byte chunks (8 channels each)
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v0 = __uint_as_float(r[j * 8 + e * 2]);
        float v1 = __uint_as_float(r[j * 8 + e * 2 + 1]);
        if (bias != nullptr) {
          const int c0 = co_base + hf * 32 + j * 8 + e * 2;
          v0 += (c0 < cout) ? __ldg(bias + c0) : 0.f;
          v1 += (c0 + 1 < cout) ? __ldg(bias + c0 + 1) : 0.f;
        }
        __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        pk[e] = *reinterpret_cast<uint32_t*>(&h);
      }
      const int chunk = hf * 4 + j;
      *reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  if (release_bar != 0u) {  // last unit of the tile: every TMEM read of this thread has completed
    tc_fence_before();
    mbar_arrive(release_bar);
  }
  C3P_STAMP(e_t3);
  fence_proxy_async_smem();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  C3P_STAMP(e_t4);
  if (etid == 0) {
    tma_store_4d(md, stg_s, co_base, c_w, c_h, c_n);
    tma_store_commit();
  }
  C3P_STAMP(e_t5);
  if (want_stats && ew * 8 < SLAB_COLS) {
    // epilogue warp ew owns the 16-byte chunk ew (8 channels) of every row
    epi_chunk_sums(stg, ew, lane, mk, as, aq);
  }
#ifdef NPP_C3_PROF
  if (etid == 0) {
    const long long e_t6 = clock64();
    C3P_ADD(16, 1); C3P_ADD(17, e_t1 - e_t0); C3P_ADD(18, e_t2 - e_t1); C3P_ADD(19, e_t3 - e_t2);
    C3P_ADD(20, e_t4 - e_t3); C3P_ADD(21, e_t5 - e_t4); C3P_ADD(22, e_t6 - e_t5);
  }
#endif
}

// fold the 32 row groups of a warp, then one atomic per (CTA, channel, moment)
template <int SLABS, int SLAB_COLS>
__device__ __forceinline__ void epi_flush_stats(float (&acc_s)[SLABS][8], float (&acc_q)[SLABS][8],
                                                float* __restrict__ stats, const int co0, const int cout, const int ew,
                                                const int lane) {
  if (ew * 8 >= SLAB_COLS) return;  // warp-uniform
  // Same-address float atomics from all CTAs arrive at the same moment and serialise in L2 (they were ~4 us of a
  // 20 us launch): 16-byte vector reductions carry four channels per operation.
  const bool vec = ((reinterpret_cast<uintptr_t>(stats) & 15u) == 0) && ((cout & 3) == 0);
#pragma unroll
  for (int sl = 0; sl < SLABS; ++sl) {
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = warp_sum(acc_s[sl][k]); b[k] = warp_sum(acc_q[sl][k]); }
    const int co = co0 + sl * 64 + ew * 8;
    if (lane == 0) {
      if (vec && co + 8 <= cout) {
        red_add_v4(stats + co, a[0], a[1], a[2], a[3]);
        red_add_v4(stats + co + 4, a[4], a[5], a[6], a[7]);
        red_add_v4(stats + cout + co, b[0], b[1], b[2], b[3]);
        red_add_v4(stats + cout + co + 4, b[4], b[5], b[6], b[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (co + k < cout) {
            atomicAdd(stats + co + k, a[k]);
            atomicAdd(stats + cout + co + k, b[k]);
          }
      }
    }
  }
}

// conv3_kernel<32>: both 128-pixel halves of a tile in ONE round (a 32-column accumulator keeps only four of the
// eight epilogue warps busy in epi_unit, and a tile would need two rounds of barriers / store waits).  Warp (q, hf)
// drains rows q*32.. of half hf into staging buffer hf; two TMA stores; statistics: warp ew owns the 16-byte chunk
// ew & 3 (8 channels) of buffer ew >> 2.
__device__ __forceinline__ void epi_pair32(const uint32_t taddr, uint8_t* stg, const uint32_t stg_s,
                                           const CUtensorMap* md, const int c_w, const int c_h, const int h_step,
                                           const int c_n, const float* __restrict__ bias, const int cout,
                                           const uint32_t release_bar, const bool want_stats, EpiMask mk,
                                           float (&as)[8], float (&aq)[8], const int q, const int hf, const int ew,
                                           const int lane, const int etid) {
  const int row = q * 32 + lane;
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  // both staging buffers are rewritten: every earlier TMA store must have finished reading them
  if (etid == 0) tma_store_wait_read<0>();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  tmem_ld_wait_regs(r);
  uint8_t* my = stg + hf * (128 * 128);
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // four 16-byte chunks (8 channels each)
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = __uint_as_float(r[j * 8 + e * 2]);
      float v1 = __uint_as_float(r[j * 8 + e * 2 + 1]);
      if (bias != nullptr) {
        const int c0 = j * 8 + e * 2;
        v0 += (c0 < cout) ? __ldg(bias + c0) : 0.f;
        v1 += (c0 + 1 < cout) ? __ldg(bias + c0 + 1) : 0.f;
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      pk[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(my + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  tc_fence_before();  // every TMEM read of this thread has completed: the accumulator stage is free
  mbar_arrive(release_bar);
  fence_proxy_async_smem();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (etid == 0) {
    tma_store_4d(md, stg_s, 0, c_w, c_h, c_n);
    tma_store_4d(md, stg_s + 128 * 128, 0, c_w, c_h + h_step, c_n);
    tma_store_commit();
  }
  if (want_stats) {
    const int half = ew >> 2, chunk = ew & 3;
    mk.p_off = half * 128;
    epi_chunk_sums(stg + half * (128 * 128), chunk, lane, mk, as, aq);
  }
}

// ---- two-group epilogue (EPI2) ------------------------------------------------------------------------------
// epi_unit marches all eight epilogue warps through one 128 x 64 unit in lockstep: per unit two 256-thread barriers,
// the wait for the staging slab, the tcgen05.ld latency and the proxy fence sit on everybody's critical path (~540 of
// the ~1470 cycles of a unit, profiles/r02_conv3_wait_profile.txt) while the issue slots are 75 % idle.  Here the warps
// form two independent groups of four (one staging slab and one named barrier each) that work on DIFFERENT units —
// alternate 64-channel slabs of a tile, the two pixel halves of a conv3 tile, or alternate tiles — so one group's
// bubbles are filled by the other group's conversions and sums.  Warp q of a group drains rows q*32.. of the whole
// slab (two tcgen05.ld of 32 columns) and owns the 16-byte chunks 2q, 2q+1 (16 channels) for the BatchNorm sums.
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int SLAB_COLS>
__device__ __forceinline__ void epi_unit_grp(const uint32_t taddr, uint8_t* stg, const uint32_t stg_s,
                                             const CUtensorMap* md, const int co_base, const int c_w, const int c_h,
                                             const int c_n, const float* __restrict__ bias, const int cout,
                                             const uint32_t release_bar, const bool want_stats, const EpiMask& mk,
                                             float (&as)[SLAB_COLS / 4], float (&aq)[SLAB_COLS / 4], const int q,
                                             const int lane, const int gtid, const int bar_id) {
  constexpr int HALVES = SLAB_COLS / 32;  // tcgen05.ld rounds of 32 columns
  const int row = q * 32 + lane;
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  // this group's previous TMA store must have finished reading the staging slab
  if (gtid == 0) tma_store_wait_read<0>();
  bar_sync_named(bar_id, 128);
#pragma unroll
  for (int hf = 0; hf < HALVES; ++hf) {
    tmem_ld_wait_regs(r);
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float v0 = __uint_as_float(r[2 * e]);
      float v1 = __uint_as_float(r[2 * e + 1]);
      if (bias != nullptr) {
        const int c0 = co_base + hf * 32 + 2 * e;
        v0 += (c0 < cout) ? __ldg(bias + c0) : 0.f;
        v1 += (c0 + 1 < cout) ? __ldg(bias + c0 + 1) : 0.f;
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      pk[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    if (hf + 1 < HALVES) tmem_ld_32x32(taddr + 32, r);  // the second half is in flight while the first is stored
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int chunk = hf * 4 + j;
      *reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4)) =
          make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
  }
  if (release_bar != 0u) {  // this thread's last TMEM read of the tile has completed
    tc_fence_before();
    mbar_arrive(release_bar);
  }
  fence_proxy_async_smem();
  bar_sync_named(bar_id, 128);
  if (gtid == 0) {
    tma_store_4d(md, stg_s, co_base, c_w, c_h, c_n);
    tma_store_commit();
  }
  if (want_stats) {
    constexpr int NCH = SLAB_COLS / 32;  // 16-byte chunks per warp: 2 (64-column slab) or 1
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) epi_chunk_sums(stg, NCH * q + cc, lane, mk, as + cc * 8, aq + cc * 8);
  }
}

// fold the 32 row groups of a warp; one 16-byte reduction per four channels (see epi_flush_stats)
template <int NCH>
__device__ __forceinline__ void epi_flush_grp(float (&as)[NCH * 8], float (&aq)[NCH * 8], float* __restrict__ stats,
                                              const int co_slab, const int cout, const int q, const int lane) {
  const bool vec = ((reinterpret_cast<uintptr_t>(stats) & 15u) == 0) && ((cout & 3) == 0);
#pragma unroll
  for (int cc = 0; cc < NCH; ++cc) {
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = warp_sum(as[cc * 8 + k]); b[k] = warp_sum(aq[cc * 8 + k]); }
    const int co = co_slab + (NCH * q + cc) * 8;
    if (lane == 0) {
      if (vec && co + 8 <= cout) {
        red_add_v4(stats + co, a[0], a[1], a[2], a[3]);
        red_add_v4(stats + co + 4, a[4], a[5], a[6], a[7]);
        red_add_v4(stats + cout + co, b[0], b[1], b[2], b[3]);
        red_add_v4(stats + cout + co + 4, b[4], b[5], b[6], b[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (co + k < cout) {
            atomicAdd(stats + co + k, a[k]);
            atomicAdd(stats + cout + co + k, b[k]);
          }
      }
    }
  }
}

// Same pipeline as conv_gemm_kernel (TMA producer warp, MMA warp, two TMEM accumulator stages), eight-warp epilogue.
template <int BN, bool EPI2>
__global__ void __launch_bounds__(kThreads2, 1)
conv_gemm2_kernel(const __grid_constant__ Maps maps, const Geom g, const TapTable taps,
                  const float* __restrict__ bias, float* __restrict__ stats) {
  using Cfg = FpropCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t smem_out = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = smem_out + Cfg::OUT_BYTES;
  const uint32_t full_bar = bar_base;                 // 2 sets (one per MMA issuer warp, see conv3_kernel) x STAGES
  const uint32_t empty_bar = bar_base + 16 * STAGES;
  const uint32_t tfull_bar = bar_base + 24 * STAGES;
  const uint32_t tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.a[0]);
    prefetch_tensormap(&maps.b);
    prefetch_tensormap(&maps.d);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(full_bar + 8 * (STAGES + i), 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      // arrivals per tile: all eight epilogue warps, or (EPI2 with a single slab: the two groups alternate tiles) four
      mbar_init(tempty_bar + 8 * i, (EPI2 && BN <= 64) ? 128 : 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above (tensor-map prefetch, mbarrier init, TMEM allocation) overlaps the tail of the previous kernel
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int num_k = g.num_taps * g.kc_blocks;
  const int nb = blockIdx.x % g.n_blocks;
  const int pt_start = blockIdx.x / g.n_blocks;
  const int pt_step = gridDim.x / g.n_blocks;
  constexpr int SLABS = BN >= 64 ? BN / 64 : 1;
  constexpr int SLAB_COLS = BN >= 64 ? 64 : BN;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;  // local tile counter: tile `it` belongs to MMA issuer it & 1
      for (int ptile = pt_start; ptile < g.num_ptiles; ptile += pt_step, ++it) {
        const uint32_t full_t = full_bar + 8 * (it & 1) * STAGES;
        int pt = ptile;
        const int w0 = (pt % g.tiles_w) * g.tw;
        pt /= g.tiles_w;
        const int h0 = (pt % g.tiles_h) * g.th;
        const int n0 = (pt / g.tiles_h) * g.tn;
        for (int t = 0; t < g.num_taps; ++t) {
          const CUtensorMap* ma = &maps.a[taps.map[t]];
          const int cw = w0 + taps.dw[t], ch = h0 + taps.dh[t], bt = taps.btap[t];
          for (int kc = 0; kc < g.kc_blocks; ++kc) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(full_t + 8 * stage, Cfg::STAGE_BYTES);
              tma_load_4d(smem_a + stage * Cfg::A_BYTES, ma, full_t + 8 * stage, kc * BK, cw, ch, n0);
              tma_load_3d(smem_b + stage * Cfg::B_BYTES, &maps.b, full_t + 8 * stage, kc * BK, bt, nb * BN);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // two MMA issuer warps alternating tiles (see conv3_kernel): warp 1 -> accumulator stage 0, warp 2 -> stage 1
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      const int which = warp - 1;
      const int acc = which;
      const uint32_t full_t = full_bar + 8 * which * STAGES;
      const uint32_t tmem_d = tmem_base + acc * BN;
      int stage = 0;
      uint32_t bits = 0;  // parity of this issuer's next fill, one bit per ring stage
      uint32_t acc_phase = 0;
      auto skip_tile = [&]() { stage = (stage + num_k) % STAGES; };
      if (which) skip_tile();
      for (int ptile = pt_start + which * pt_step; ptile < g.num_ptiles; ptile += 2 * pt_step) {
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        int kc = 0;  // 64-channel block index of the current k-block (the tap loop is the outer one)
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_t + 8 * stage, (bits >> stage) & 1u);
          bits ^= 1u << stage;
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_a + stage * Cfg::A_BYTES, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(smem_b + stage * Cfg::B_BYTES, 16, 1024);
          const int ks = (kc + 1 == g.kc_blocks) ? g.k_last : BK / 16;
          if (++kc == g.kc_blocks) kc = 0;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              if (k < ks) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar + 8 * stage);
            if (kb == num_k - 1) umma_commit(tfull_bar + 8 * acc);
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0;
        }
        skip_tile();  // the other issuer's tile
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4 && EPI2) {
    // ------------------------------------------------------------------ two-group epilogue (see epi_unit_grp)
    const int ew = warp - 4, q = warp & 3, gi = ew >> 2;
    const int gtid = threadIdx.x - 128 - gi * 128;
    int nslab = (g.cout - nb * BN + 63) / 64;  // slabs of this CTA's channel block that hold real channels
    if (nslab > SLABS) nslab = SLABS;
    constexpr int GS = SLABS >= 2 ? SLABS / 2 : 1;  // slabs of a tile per group
    constexpr int NCH = SLAB_COLS / 32;
    uint8_t* stg = smem_gen + (smem_out - smem_base) + gi * (BM * 128);
    const uint32_t stg_s = smem_out + gi * (BM * 128);
    float gs_[GS][NCH * 8], gq_[GS][NCH * 8];
#pragma unroll
    for (int sl = 0; sl < GS; ++sl)
#pragma unroll
      for (int k = 0; k < NCH * 8; ++k) { gs_[sl][k] = 0.f; gq_[sl][k] = 0.f; }
    // SLABS >= 2: both groups work on every tile (group gi: slabs gi, gi + 2); one slab: the groups alternate tiles
    // (group gi <-> accumulator stage gi, like the MMA issuer warps)
    constexpr bool ALT = SLABS < 2;
    int acc = ALT ? gi : 0;
    uint32_t acc_phase = 0;
    for (int ptile = pt_start + (ALT ? gi * pt_step : 0); ptile < g.num_ptiles; ptile += (ALT ? 2 : 1) * pt_step) {
      int pt = ptile;
      EpiMask mk;
      mk.tw = g.tw; mk.th = g.th; mk.W = g.W; mk.H = g.H; mk.N = g.N; mk.p_off = 0;
      mk.w0 = (pt % g.tiles_w) * g.tw;
      pt /= g.tiles_w;
      mk.h0 = (pt % g.tiles_h) * g.th;
      mk.n0 = (pt / g.tiles_h) * g.tn;
      mk.full = (mk.w0 + g.tw <= g.W) && (mk.h0 + g.th <= g.H) && (mk.n0 + g.tn <= g.N);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      bool released = false;
#pragma unroll
      for (int s2 = 0; s2 < GS; ++s2) {
        const int slab = ALT ? 0 : gi + 2 * s2;
        if (slab < nslab) {  // uniform across the group
          const bool last = ALT || slab + 2 >= nslab;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(acc * BN + slab * 64);
          epi_unit_grp<SLAB_COLS>(taddr, stg, stg_s, &maps.d, nb * BN + slab * 64, mk.w0, mk.h0, mk.n0, bias, g.cout,
                                  last ? tempty_bar + 8 * acc : 0u, stats != nullptr, mk, gs_[s2], gq_[s2], q, lane, gtid,
                                  1 + gi);
          released = released || last;
        }
      }
      if (!released) {  // no slab of this tile belongs to the group: the accumulator stage is free as far as it goes
        tc_fence_before();
        mbar_arrive(tempty_bar + 8 * acc);
      }
      if (ALT) {
        acc_phase ^= 1;
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    if (stats != nullptr) {
#pragma unroll
      for (int s2 = 0; s2 < GS; ++s2)
        epi_flush_grp<NCH>(gs_[s2], gq_[s2], stats, nb * BN + (ALT ? 0 : gi + 2 * s2) * 64, g.cout, q, lane);
    }
    if (gtid == 0) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    const int ew = warp - 4, q = warp & 3, hf = ew >> 2;
    const int etid = threadIdx.x - 128;
    int nslab = (g.cout - nb * BN + 63) / 64;  // slabs of this CTA's channel block that hold real channels
    if (nslab > SLABS) nslab = SLABS;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sbuf = 0;
    float acc_s[SLABS][8], acc_q[SLABS][8];
#pragma unroll
    for (int sl = 0; sl < SLABS; ++sl)
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc_s[sl][k] = 0.f; acc_q[sl][k] = 0.f; }
    for (int ptile = pt_start; ptile < g.num_ptiles; ptile += pt_step) {
      int pt = ptile;
      EpiMask mk;
      mk.tw = g.tw; mk.th = g.th; mk.W = g.W; mk.H = g.H; mk.N = g.N; mk.p_off = 0;
      mk.w0 = (pt % g.tiles_w) * g.tw;
      pt /= g.tiles_w;
      mk.h0 = (pt % g.tiles_h) * g.th;
      mk.n0 = (pt / g.tiles_h) * g.tn;
      mk.full = (mk.w0 + g.tw <= g.W) && (mk.h0 + g.th <= g.H) && (mk.n0 + g.tn <= g.N);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll
      for (int slab = 0; slab < SLABS; ++slab) {
        if (slab < nslab) {  // uniform across the CTA
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(acc * BN + slab * 64 + hf * 32);
          epi_unit<SLAB_COLS>(taddr, smem_gen + (smem_out - smem_base) + sbuf * (BM * 128), smem_out + sbuf * (BM * 128),
                              &maps.d, nb * BN + slab * 64, mk.w0, mk.h0, mk.n0, bias, g.cout,
                              slab == nslab - 1 ? tempty_bar + 8 * acc : 0u, stats != nullptr, mk, acc_s[slab],
                              acc_q[slab], q, hf, ew, lane, etid);
          sbuf ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (stats != nullptr) epi_flush_stats<SLABS, SLAB_COLS>(acc_s, acc_q, stats, nb * BN, g.cout, ew, lane);
    if (etid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// 3x3 / stride 1 / pad 1 fprop and dgrad with Cout <= 128: 256-pixel tiles and vertical-tap sharing.
//
// The kernels above fetch one activation box and one weight tile per (tap, 64-channel block): 32 KB of L2 -> SM
// traffic per 128 x 128 x 64 MMA block = 64 flop/B, and the measured time of the dominant layer (128 -> 128 @ 96^2
// x 32 images, 96 us) is exactly the L2 -> SM limit of the chip (~6300 B/clk, B300_MICROARCH.md) — the tensor
// pipe idles a third of the time.  Two changes bring the tile to ~150 flop/B:
//   * a CTA owns a TW x TH = 256-pixel tile of one image = two M=128 accumulators that share every weight tile;
//   * the activation box carries one halo row above and below (TH + 2 rows of TW pixels, TW % 8 == 0): the three
//     vertical taps of a kernel column are the SAME shared-memory box read 0 / TW / 2*TW pixels further down —
//     whole 1024-byte SWIZZLE_128B atoms, so only the UMMA descriptor's start address moves.
// Per (kernel column, 64-channel block): one box of (TH+2)*TW*128 B and three weight tiles feed 24 MMAs.
// Activation boxes and weight tiles travel in separate mbarrier rings (different sizes, different reuse).
// =================================================================================================
struct C3Maps {
  CUtensorMap a;  // input: dims (C, W, H, N), box (64, tw, th + 2, 1)
  CUtensorMap b;  // weights: dims (K, 9, rows), box (64, 1, BN)
  CUtensorMap d;  // output: dims (C, W, H, N), box (64, tw, th / 2, 1)
};
struct C3Geom {
  int tw, th;            // tw * th == 256, tw % 8 == 0, th even
  int tiles_w, tiles_h;  // per image
  int num_ptiles;
  int W, H;
  int kc_blocks, cout;
  int a_bytes;           // (th + 2) * tw * 128
  int sa, sb;            // ring depths (activation boxes, weight tiles)
  int flip;              // 0: fprop taps (x[h + r - 1][w + s - 1]); 1: dgrad taps (dy[h + 1 - r][w + 1 - s])
  int two_producers;     // weight tiles issued by a second producer thread (warp 3)
  int k_last;            // K16 steps of the last 64-channel block that hold real channels (see Geom)
  int wres;              // weights resident: all nine tap tiles of the (single) 64-channel block are loaded once per CTA
};
constexpr int kC3MaxSA = 4, kC3MaxSB = 8;
constexpr int kC3OutBytes = 2 * 128 * 128;

template <int BN, bool EPI2>
__global__ void __launch_bounds__(kThreads2, 1)
conv3_kernel(const __grid_constant__ C3Maps maps, const C3Geom g, const float* __restrict__ bias,
             float* __restrict__ stats) {
  constexpr int B_BYTES = BN * 128;
  constexpr int TMEM_COLS = 4 * BN;  // 2 accumulator stages x 2 pixel halves; 128 / 256 / 512
  constexpr int SLABS = BN >= 64 ? BN / 64 : 1;
  constexpr int SLAB_COLS = BN >= 64 ? 64 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + g.sa * g.a_bytes;
  const uint32_t smem_out = smem_b + g.sb * B_BYTES;
  const uint32_t bar_base = smem_out + kC3OutBytes;
  // "full" barriers exist twice, one set per MMA issuer warp (tiles alternate between two issuers, see below): every
  // barrier then has ONE waiter that visits its phases in order — with a shared set, the issuer of tile i + 1 could
  // poll a stage whose previous fill (tile i) has not landed yet and mistake the older completed phase of the same
  // parity for its own
  const uint32_t afull_bar = bar_base;                          // 2 x kC3MaxSA x 8 B
  const uint32_t aempty_bar = afull_bar + 16 * kC3MaxSA;
  const uint32_t bfull_bar = aempty_bar + 8 * kC3MaxSA;         // 2 x kC3MaxSB x 8 B
  const uint32_t bempty_bar = bfull_bar + 16 * kC3MaxSB;
  const uint32_t tfull_bar = bempty_bar + 8 * kC3MaxSB;         // 2 x 8 B
  const uint32_t tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.a);
    prefetch_tensormap(&maps.d);
  }
  if (warp == 3 && lane == 0) prefetch_tensormap(&maps.b);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.sa; ++i) {
      mbar_init(afull_bar + 8 * i, 1);
      mbar_init(afull_bar + 8 * (kC3MaxSA + i), 1);
      mbar_init(aempty_bar + 8 * i, 1);
    }
    for (int i = 0; i < (g.wres ? 1 : g.sb); ++i) {
      mbar_init(bfull_bar + 8 * i, 1);
      mbar_init(bfull_bar + 8 * (kC3MaxSB + i), 1);
      mbar_init(bempty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above (tensor-map prefetch, mbarrier init, TMEM allocation) overlaps the tail of the previous kernel
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane)
    {
      int as = 0, bs = 0;
      uint32_t ap = 0, bp = 0;
      C3P_DECL(p_wait); C3P_DECL(p_tiles);
#ifdef NPP_C3_PROF
      const long long p_t0 = clock64();
#endif
      int it = 0;  // local tile counter: tile `it` belongs to MMA issuer it & 1
      for (int ptile = blockIdx.x; ptile < g.num_ptiles; ptile += gridDim.x, ++it) {
#ifdef NPP_C3_PROF
        ++p_tiles;
#endif
        const uint32_t afull_t = afull_bar + 8 * (it & 1) * kC3MaxSA;
        const uint32_t bfull_t = bfull_bar + 8 * (it & 1) * kC3MaxSB;
        int pt = ptile;
        const int w0 = (pt % g.tiles_w) * g.tw;
        pt /= g.tiles_w;
        const int h0 = (pt % g.tiles_h) * g.th;
        const int n0 = pt / g.tiles_h;
        for (int s = 0; s < 3; ++s) {
          const int cw = w0 + (g.flip ? 1 - s : s - 1);
          for (int kc = 0; kc < g.kc_blocks; ++kc) {
            C3P_WAIT(p_wait, mbar_wait(aempty_bar + 8 * as, ap ^ 1));
            if (elect_one()) {
              mbar_arrive_expect_tx(afull_t + 8 * as, g.a_bytes);
              tma_load_4d(smem_a + as * g.a_bytes, &maps.a, afull_t + 8 * as, kc * BK, cw, h0 - 1, n0);
            }
            __syncwarp();
            if (++as == g.sa) { as = 0; ap ^= 1; }
            if (g.two_producers || g.wres) continue;  // the weight tiles are fed by warp 3
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
              mbar_wait(bempty_bar + 8 * bs, bp ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(bfull_t + 8 * bs, B_BYTES);
                tma_load_3d(smem_b + bs * B_BYTES, &maps.b, bfull_t + 8 * bs, kc * BK, r * 3 + s, 0);
              }
              __syncwarp();
              if (++bs == g.sb) { bs = 0; bp ^= 1; }
            }
          }
        }
      }
#ifdef NPP_C3_PROF
      if (lane == 0) { C3P_ADD(0, clock64() - p_t0); C3P_ADD(1, p_wait); C3P_ADD(8, p_tiles); }
#endif
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ second TMA producer: weight tiles only, so
    // that the activation ring runs its full depth ahead instead of being throttled by the (shorter) weight ring
    if (g.wres) {
      // resident weights (Cin <= 64, small Cout): nine tap tiles, loaded once, one barrier
      if (elect_one()) {
        mbar_arrive_expect_tx(bfull_bar, 9 * B_BYTES);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) tma_load_3d(smem_b + t * B_BYTES, &maps.b, bfull_bar, 0, t, 0);
      }
      __syncwarp();
    } else if (g.two_producers) {
      int bs = 0;
      uint32_t bp = 0;
      C3P_DECL(b_wait);
#ifdef NPP_C3_PROF
      const long long b_t0 = clock64();
#endif
      int it = 0;
      for (int ptile = blockIdx.x; ptile < g.num_ptiles; ptile += gridDim.x, ++it) {
        const uint32_t bfull_t = bfull_bar + 8 * (it & 1) * kC3MaxSB;
        for (int s = 0; s < 3; ++s) {
          for (int kc = 0; kc < g.kc_blocks; ++kc) {
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
              C3P_WAIT(b_wait, mbar_wait(bempty_bar + 8 * bs, bp ^ 1));
              if (elect_one()) {
                mbar_arrive_expect_tx(bfull_t + 8 * bs, B_BYTES);
                tma_load_3d(smem_b + bs * B_BYTES, &maps.b, bfull_t + 8 * bs, kc * BK, r * 3 + s, 0);
              }
              __syncwarp();
              if (++bs == g.sb) { bs = 0; bp ^= 1; }
            }
          }
        }
      }
#ifdef NPP_C3_PROF
      if (lane == 0) { C3P_ADD(9, clock64() - b_t0); C3P_ADD(10, b_wait); }
#endif
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: TWO warps, alternating tiles
    // (warp 1: tiles 0, 2, ... into accumulator stage 0; warp 2 — the TMEM allocator, idle otherwise — tiles 1, 3, ...
    // into stage 1).  Each warp runs its loop converged and one elected lane issues.  Why two: the wait-cycle profile
    // (tests/csrc/_bin/prof, profiles/r02_conv3_wait_profile.txt) showed the issuing thread itself as the bottleneck —
    // ~330 cycles of scalar bookkeeping (mbarrier polls, descriptor arithmetic, elect, commits) per group of 8 MMAs
    // that occupy the tensor pipe for 128-512 cycles; with two issuers one warp's bookkeeping hides behind the other
    // warp's MMAs.  The rings are consumed strictly in tile order, so a warp simply skips the stages of the other
    // warp's tile; every stage use still has exactly one committing thread.
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      const int which = warp - 1;
      const int acc = which;
      const int steps = 3 * g.kc_blocks;
      int as = 0, bs = 0;
      uint32_t abits = 0, bbits = 0;  // parity of this issuer's next fill, one bit per ring stage
      uint32_t acc_phase = 0;
      const uint32_t afull_t = afull_bar + 8 * which * kC3MaxSA;
      const uint32_t bfull_t = bfull_bar + 8 * which * kC3MaxSB;
      auto skip_tile = [&]() {
        as += steps;
        while (as >= g.sa) as -= g.sa;
        if (!g.wres) {
          bs += 3 * steps;
          while (bs >= g.sb) bs -= g.sb;
        }
      };
      if (which) skip_tile();
      C3P_DECL(m_wa); C3P_DECL(m_wb); C3P_DECL(m_wt); C3P_DECL(m_issue); C3P_DECL(m_commit);
#ifdef NPP_C3_PROF
      const long long m_t0 = clock64();
#endif
      if (g.wres) {
        mbar_wait(bfull_bar, 0);
        tc_fence_after();
      }
      const uint32_t tmem_d = tmem_base + acc * 2 * BN;
      const uint32_t rstep = static_cast<uint32_t>(g.tw * 128) >> 4;  // one image row of the box, in descriptor units
      for (int ptile = blockIdx.x + which * gridDim.x; ptile < g.num_ptiles; ptile += 2 * gridDim.x) {
        C3P_WAIT(m_wt, mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1));
        tc_fence_after();
        for (int step = 0; step < steps; ++step) {
          C3P_WAIT(m_wa, mbar_wait(afull_t + 8 * as, (abits >> as) & 1u));
          abits ^= 1u << as;
          tc_fence_after();
          const uint64_t da00 = make_smem_desc_sw128(smem_a + as * g.a_bytes, 16, 1024);
          if (g.wres) {
            // one activation box (kernel column `step`) against the three resident weight tiles of that column:
            // up to 24 MMAs back to back, one commit
            const uint64_t db00 = make_smem_desc_sw128(smem_b + step * B_BYTES, 16, 1024);
            if (elect_one()) {
#ifdef NPP_C3_PROF
              const long long i_t0 = clock64();
#endif
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const uint64_t db = db00 + r * (3 * B_BYTES / 16);          // tap index r * 3 + step
                // vertical tap r reads the box `ro` image rows (ro * tw pixels = ro * tw * 128 bytes) further down
                const uint64_t dar = da00 + (g.flip ? 2 - r : r) * rstep;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                  const uint64_t da = dar + m * (128 * 128 / 16);  // second pixel half: 128 rows of 128 bytes further
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    if (k < g.k_last) umma_f16(tmem_d + m * BN, da + 2 * k, db + 2 * k, idesc, (step | r | k) != 0 ? 1u : 0u);
                }
              }
#ifdef NPP_C3_PROF
              const long long i_t1 = clock64();
#endif
              umma_commit(aempty_bar + 8 * as);
              if (step == steps - 1) umma_commit(tfull_bar + 8 * acc);
#ifdef NPP_C3_PROF
              m_issue += i_t1 - i_t0;
              m_commit += clock64() - i_t1;
#endif
            }
            __syncwarp();
          } else {
            // the three weight tiles (vertical taps) of this (kernel column, 64-channel block)
            int bi[3];
            {
              int t = bs;
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                bi[r] = t;
                C3P_WAIT(m_wb, mbar_wait(bfull_t + 8 * t, (bbits >> t) & 1u));
                bbits ^= 1u << t;
                if (++t == g.sb) t = 0;
              }
              bs = t;
            }
            tc_fence_after();
            const int ks = (step % g.kc_blocks == g.kc_blocks - 1) ? g.k_last : BK / 16;
            if (elect_one()) {
#ifdef NPP_C3_PROF
              const long long i_t0 = clock64();
              long long i_c = 0;
#endif
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const uint64_t db = make_smem_desc_sw128(smem_b + bi[r] * B_BYTES, 16, 1024);
                const uint64_t dar = da00 + (g.flip ? 2 - r : r) * rstep;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                  const uint64_t da = dar + m * (128 * 128 / 16);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    if (k < ks) umma_f16(tmem_d + m * BN, da + 2 * k, db + 2 * k, idesc, (step | r | k) != 0 ? 1u : 0u);
                }
#ifdef NPP_C3_PROF
                const long long c_t0 = clock64();
#endif
                umma_commit(bempty_bar + 8 * bi[r]);   // the weight producer may refill this tile
#ifdef NPP_C3_PROF
                i_c += clock64() - c_t0;
#endif
              }
#ifdef NPP_C3_PROF
              const long long i_t1 = clock64();
#endif
              umma_commit(aempty_bar + 8 * as);
              if (step == steps - 1) umma_commit(tfull_bar + 8 * acc);
#ifdef NPP_C3_PROF
              m_issue += i_t1 - i_t0 - i_c;
              m_commit += clock64() - i_t1 + i_c;
#endif
            }
            __syncwarp();
          }
          if (++as == g.sa) as = 0;
        }
        skip_tile();  // the other issuer's tile
        acc_phase ^= 1;
      }
#ifdef NPP_C3_PROF
      if (lane == 0) { C3P_ADD(2, clock64() - m_t0); C3P_ADD(3, m_wa); C3P_ADD(4, m_wb); C3P_ADD(5, m_wt); }
      if (m_issue) { C3P_ADD(11, m_issue); C3P_ADD(12, m_commit); }
#endif
    }
  } else if (warp >= 4 && EPI2 && BN >= 64) {
    // ------------------------------------------------------------------ two-group epilogue (see epi_unit_grp):
    // BN = 64: group gi drains pixel half gi of every tile; BN = 128: group gi drains slab gi of both pixel halves
    const int ew = warp - 4, q = warp & 3, gi = ew >> 2;
    const int gtid = threadIdx.x - 128 - gi * 128;
    int nslab = (g.cout + 63) / 64;
    if (nslab > SLABS) nslab = SLABS;
    uint8_t* stg = smem_gen + (smem_out - smem_base) + gi * (128 * 128);
    const uint32_t stg_s = smem_out + gi * (128 * 128);
    float gs_[16], gq_[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { gs_[k] = 0.f; gq_[k] = 0.f; }
    int acc = 0;
    uint32_t acc_phase = 0;
    C3P_DECL(e_wait);
#ifdef NPP_C3_PROF
    const long long e_t0 = clock64();
#endif
    for (int ptile = blockIdx.x; ptile < g.num_ptiles; ptile += gridDim.x) {
      int pt = ptile;
      EpiMask mk;
      mk.tw = g.tw; mk.th = g.th; mk.W = g.W; mk.H = g.H; mk.N = 1; mk.n0 = 0;
      mk.w0 = (pt % g.tiles_w) * g.tw;
      pt /= g.tiles_w;
      mk.h0 = (pt % g.tiles_h) * g.th;
      const int n0 = pt / g.tiles_h;
      mk.full = (mk.w0 + g.tw <= g.W) && (mk.h0 + g.th <= g.H);
      C3P_WAIT(e_wait, mbar_wait(tfull_bar + 8 * acc, acc_phase));
      tc_fence_after();
      if constexpr (SLABS == 1) {
        const int m = gi;
        mk.p_off = m * 128;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>((acc * 2 + m) * BN);
        epi_unit_grp<64>(taddr, stg, stg_s, &maps.d, 0, mk.w0, mk.h0 + m * (g.th >> 1), n0, bias, g.cout,
                         tempty_bar + 8 * acc, stats != nullptr, mk, gs_, gq_, q, lane, gtid, 1 + gi);
      } else {
        const int slab = gi;
        if (slab < nslab) {  // uniform across the group
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            mk.p_off = m * 128;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                   static_cast<uint32_t>((acc * 2 + m) * BN + slab * 64);
            epi_unit_grp<64>(taddr, stg, stg_s, &maps.d, slab * 64, mk.w0, mk.h0 + m * (g.th >> 1), n0, bias, g.cout,
                             m == 1 ? tempty_bar + 8 * acc : 0u, stats != nullptr, mk, gs_, gq_, q, lane, gtid, 1 + gi);
          }
        } else {
          tc_fence_before();
          mbar_arrive(tempty_bar + 8 * acc);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
#ifdef NPP_C3_PROF
    if (gtid == 0 && gi == 0) { C3P_ADD(6, clock64() - e_t0); C3P_ADD(7, e_wait); }
#endif
    if (stats != nullptr) epi_flush_grp<2>(gs_, gq_, stats, (SLABS == 1 ? 0 : gi) * 64, g.cout, q, lane);
    if (gtid == 0) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (eight warps)
    const int ew = warp - 4, q = warp & 3, hf = ew >> 2;
    const int etid = threadIdx.x - 128;
    int nslab = (g.cout + 63) / 64;
    if (nslab > SLABS) nslab = SLABS;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sbuf = 0;
    float acc_s[SLABS][8], acc_q[SLABS][8];
#pragma unroll
    for (int sl = 0; sl < SLABS; ++sl)
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc_s[sl][k] = 0.f; acc_q[sl][k] = 0.f; }
    C3P_DECL(e_wait);
#ifdef NPP_C3_PROF
    const long long e_t0 = clock64();
#endif
    for (int ptile = blockIdx.x; ptile < g.num_ptiles; ptile += gridDim.x) {
      int pt = ptile;
      EpiMask mk;
      mk.tw = g.tw; mk.th = g.th; mk.W = g.W; mk.H = g.H; mk.N = 1; mk.n0 = 0;
      mk.w0 = (pt % g.tiles_w) * g.tw;
      pt /= g.tiles_w;
      mk.h0 = (pt % g.tiles_h) * g.th;
      const int n0 = pt / g.tiles_h;
      mk.full = (mk.w0 + g.tw <= g.W) && (mk.h0 + g.th <= g.H);
      C3P_WAIT(e_wait, mbar_wait(tfull_bar + 8 * acc, acc_phase));
      tc_fence_after();
      if constexpr (BN == 32) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>((acc * 2 + hf) * BN);
        epi_pair32(taddr, smem_gen + (smem_out - smem_base), smem_out, &maps.d, mk.w0, mk.h0, g.th >> 1, n0, bias,
                   g.cout, tempty_bar + 8 * acc, stats != nullptr, mk, acc_s[0], acc_q[0], q, hf, ew, lane, etid);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mk.p_off = m * 128;
#pragma unroll
        for (int slab = 0; slab < SLABS; ++slab) {
          if (slab < nslab) {  // uniform across the CTA
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                   static_cast<uint32_t>((acc * 2 + m) * BN + slab * 64 + hf * 32);
            epi_unit<SLAB_COLS>(taddr, smem_gen + (smem_out - smem_base) + sbuf * (128 * 128),
                                smem_out + sbuf * (128 * 128), &maps.d, slab * 64, mk.w0, mk.h0 + m * (g.th >> 1), n0,
                                bias, g.cout, (m == 1 && slab == nslab - 1) ? tempty_bar + 8 * acc : 0u,
                                stats != nullptr, mk, acc_s[slab], acc_q[slab], q, hf, ew, lane, etid);
            sbuf ^= 1;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
#ifdef NPP_C3_PROF
    if (etid == 0) { C3P_ADD(6, clock64() - e_t0); C3P_ADD(7, e_wait); }
#endif
    if constexpr (BN == 32) {
      if (stats != nullptr) epi_flush_stats<SLABS, 64>(acc_s, acc_q, stats, 0, g.cout, ew & 3, lane);  // all 8 warps
    } else {
      if (stats != nullptr) epi_flush_stats<SLABS, SLAB_COLS>(acc_s, acc_q, stats, 0, g.cout, ew, lane);
    }
    if (etid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =================================================================================================
// wgrad kernel: dW[co, tap, ci] += sum_{pixels in this CTA's K range} dY[p, co] * X[p + tap, ci]
// =================================================================================================
struct WMaps {
  CUtensorMap x[4];  // activation (phase) maps
  CUtensorMap dy;
};
struct WGeom {
  int num_taps;
  int tw, th, tn;  // K block = 64 pixels
  int tiles_w, tiles_h, tiles_n;
  int total_ptiles;
  int co_blocks, ci_blocks;  // of 128 / BN
  int cout, cin, dw_ld;      // dw_ld = leading dimension (cin of the fp32 gradient)
  int ksplit;
};

template <int BN>
struct WgradCfg {
  static constexpr int KP = 64;                 // pixels per k-block
  static constexpr int A_BYTES = 2 * KP * 128;  // two 64-channel slabs of dY
  static constexpr int B_BYTES = (BN / 64) * KP * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulators: one per MMA issuer warp (k-blocks alternate)
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WMaps maps, const WGeom g, const TapTable taps,
                  float* __restrict__ dw, float* __restrict__ ws) {
  using Cfg = WgradCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int SLAB = Cfg::KP * 128;  // bytes of one [64 pixels x 64 channels] slab
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bar_base;                 // 2 sets (one per MMA issuer warp) x STAGES
  const uint32_t empty_bar = bar_base + 16 * STAGES;
  const uint32_t tfull_bar = bar_base + 24 * STAGES;
  const uint32_t tmem_slot = tfull_bar + 8;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // output tile of this CTA
  int ot = blockIdx.x;
  const int ci_blk = ot % g.ci_blocks;
  ot /= g.ci_blocks;
  const int co_blk = ot % g.co_blocks;
  const int tap = ot / g.co_blocks;
  // K range (pixel tiles)
  const int p_begin = static_cast<int>((static_cast<int64_t>(g.total_ptiles) * blockIdx.y) / g.ksplit);
  const int p_end = static_cast<int>((static_cast<int64_t>(g.total_ptiles) * (blockIdx.y + 1)) / g.ksplit);
  const int num_k = p_end - p_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.dy);
    prefetch_tensormap(&maps.x[taps.map[tap]]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(full_bar + 8 * (STAGES + i), 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    mbar_init(tfull_bar, num_k >= 2 ? 2 : 1);  // one arrival per issuer warp that has work
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above (tensor-map prefetch, mbarrier init, TMEM allocation) overlaps the tail of the previous kernel
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (num_k > 0) {
    if (warp == 0) {
      {  // converged warp, one elected lane issues (see conv_gemm_kernel)
        const CUtensorMap* mx = &maps.x[taps.map[tap]];
        const int dwx = taps.dw[tap], dhx = taps.dh[tap];
        // narrow layers (<= 64 output channels in this block): the second 64-channel dY slab would be all padding —
        // it is not fetched (TMA cost is per box row), the MMA reads stale shared memory there and the accumulator
        // rows it produces (co >= cout) are dropped by the epilogue / the reduce kernel
        const int a_slabs = (g.cout - co_blk * 128 > 64) ? 2 : 1;
        int stage = 0;
        uint32_t phase = 0;
        for (int p = p_begin; p < p_end; ++p) {
          int pt = p;
          const int w0 = (pt % g.tiles_w) * g.tw;
          pt /= g.tiles_w;
          const int h0 = (pt % g.tiles_h) * g.th;
          const int n0 = (pt / g.tiles_h) * g.tn;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_a + stage * Cfg::A_BYTES;
          const uint32_t sb = smem_b + stage * Cfg::B_BYTES;
          const uint32_t fb = full_bar + 8 * (((p - p_begin) & 1) * STAGES + stage);  // k-block -> issuer (p - p_begin) & 1
          if (elect_one()) {
            mbar_arrive_expect_tx(fb, a_slabs * SLAB + Cfg::B_BYTES);
            for (int s = 0; s < a_slabs; ++s)
              tma_load_4d(sa + s * SLAB, &maps.dy, fb, co_blk * 128 + s * 64, w0, h0, n0);
#pragma unroll
            for (int s = 0; s < BN / 64; ++s)
              tma_load_4d(sb + s * SLAB, mx, fb, ci_blk * BN + s * 64, w0 + dwx, h0 + dhx, n0);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // Two MMA issuer warps (see conv3_kernel): k-blocks alternate, each issuer accumulates into its own TMEM
      // accumulator (columns which * BN ..), the epilogue adds the two.  One full-barrier set per issuer.
      {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        const int which = warp - 1;
        const uint32_t tmem_d = tmem_base + which * BN;
        const uint32_t full_t = full_bar + 8 * which * STAGES;
        int stage = which % STAGES;
        uint32_t bits = 0;  // parity of this issuer's next fill, one bit per ring stage
        const int last = ((num_k - 1 - which) & ~1) + which;  // this issuer's last k-block
        for (int kb = which; kb < num_k; kb += 2) {
          mbar_wait(full_t + 8 * stage, (bits >> stage) & 1u);
          bits ^= 1u << stage;
          tc_fence_after();
          // MN-major SW128: 64-element MN slabs LBO apart, 8-row K groups SBO = 1024 B apart
          const uint64_t da = make_smem_desc_sw128(smem_a + stage * Cfg::A_BYTES, SLAB, 1024);
          const uint64_t db = make_smem_desc_sw128(smem_b + stage * Cfg::B_BYTES, SLAB, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < Cfg::KP / 16; ++k) {
              // 16 pixels (K) further = 16 rows x 128 B = 2048 B -> +128 in (addr>>4)
              umma_f16(tmem_d, da + 128 * k, db + 128 * k, idesc, (kb >= 2 || k != 0) ? 1u : 0u);
            }
            umma_commit(empty_bar + 8 * stage);
            if (kb == last) umma_commit(tfull_bar);
          }
          __syncwarp();
          stage = (stage + 2) % STAGES;
        }
      }
    } else if (warp >= kEpiWarp0) {
      const int q = warp & 3;
      const int co = co_blk * 128 + q * 32 + lane;
      const int bt = taps.btap[tap];
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      // split-K partial of this CTA: plain coalesced stores into the workspace [k slice][tile][128][BN] (folded by
      // wgrad_reduce_kernel) — or, without a workspace, fp32 atomics straight into dW
      float* part = ws ? ws + ((static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 128 + (q * 32 + lane)) * BN
                       : nullptr;
#pragma unroll 1
      for (int c32 = 0; c32 < BN / 32; ++c32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c32 * 32, r);
        tmem_ld_wait();
        if (num_k >= 2) {  // the second issuer's accumulator
          uint32_t r2[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN + c32 * 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        if (part) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(part + c32 * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ci = ci_blk * BN + c32 * 32 + j;
            if (co < g.cout && ci < g.cin)
              atomicAdd(dw + (static_cast<int64_t>(co) * g.cin + ci) * g.num_taps + bt, __uint_as_float(r[j]));
          }
        }
      }
    }
  } else if (ws && warp >= kEpiWarp0) {
    // no pixels for this k slice: its partial is zero
    float* part = ws + ((static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 128 + ((warp & 3) * 32 + lane)) * BN;
    for (int j = 0; j < BN; j += 4) *reinterpret_cast<uint4*>(part + j) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}



// =================================================================================================
// wgrad for 3x3 / stride 1 / dilation 1 convolutions (62 % of the network's FLOPs): one CTA owns a COLUMN of three
// vertical taps.  The per-tap kernel above re-reads dY and X once per tap — 18 operand tiles per 9 taps, which made
// it L2->SM-bandwidth bound (ncu: 620 TFLOP/s at 64 flop/byte).  Here the X box carries one halo row above and
// below (th+2 rows of tw pixels), so the three taps are the SAME shared-memory tile read at row offsets
// 0 / tw / 2*tw pixels — multiples of 1024 bytes, i.e. whole SWIZZLE_128B atoms, so the UMMA descriptor just starts
// later — and dY is loaded once for all three: 3 MMAs groups per (dY + X-with-halo) load, three 128 x BN fp32
// accumulators in TMEM.  Grid: (3 tap columns x Cout blocks x Cin blocks, split-K over pixel tiles).
// =================================================================================================
struct W3Geom {
  int tw, th;                // pixel tile of one k-block: tw * th == 64 pixels of one image
  int tiles_w, tiles_h;      // tiles per image
  int total_ptiles;          // tiles_w * tiles_h * N
  int co_blocks, ci_blocks;  // of 128 / BN
  int cout, cin;
  int ksplit, stages;
  int bslab;                 // bytes of one 64-channel X slab: (th + 2) * tw * 128
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                   const W3Geom g, float* __restrict__ dw, float* __restrict__ ws) {
  constexpr int A_BYTES = 2 * 64 * 128;  // two 64-channel slabs of dY, 64 pixels each
  constexpr int ASLAB = 64 * 128;
  // BN = 64: two MMA issuer warps (k-blocks alternate), each with its own set of three accumulators (2 x 192 columns);
  // BN = 128: three accumulators already take 384 of the 512 columns -> one issuer
  constexpr bool DUAL = (BN == 64);
  constexpr int TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full_bar = smem_base;            // 2 sets x 8 x 8 B
  const uint32_t empty_bar = smem_base + 128;     // 8 x 8 B
  const uint32_t tfull_bar = smem_base + 192;
  const uint32_t tmem_slot = smem_base + 200;
  const uint32_t ring = smem_base + 1024;
  const uint32_t b_bytes = (BN / 64) * g.bslab;
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int ot = blockIdx.x;
  const int ci_blk = ot % g.ci_blocks;
  ot /= g.ci_blocks;
  const int co_blk = ot % g.co_blocks;
  const int q = ot / g.co_blocks;  // horizontal tap index 0..2 of this CTA's tap column
  const int p_begin = static_cast<int>((static_cast<int64_t>(g.total_ptiles) * blockIdx.y) / g.ksplit);
  const int p_end = static_cast<int>((static_cast<int64_t>(g.total_ptiles) * (blockIdx.y + 1)) / g.ksplit);
  const int num_k = p_end - p_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_dy);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(full_bar + 8 * (8 + i), 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    mbar_init(tfull_bar, (DUAL && num_k >= 2) ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above (tensor-map prefetch, mbarrier init, TMEM allocation) overlaps the tail of the previous kernel
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (num_k > 0) {
    if (warp == 0) {
      {  // converged warp, one elected lane issues (see conv_gemm_kernel)
        const int a_slabs = (g.cout - co_blk * 128 > 64) ? 2 : 1;  // see conv_wgrad_kernel
        int stage = 0;
        uint32_t phase = 0;
        for (int p = p_begin; p < p_end; ++p) {
          int pt = p;
          const int w0 = (pt % g.tiles_w) * g.tw;
          pt /= g.tiles_w;
          const int h0 = (pt % g.tiles_h) * g.th;
          const int n0 = pt / g.tiles_h;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t sa = ring + stage * stage_bytes;
          const uint32_t sb = sa + A_BYTES;
          const uint32_t fb = full_bar + 8 * ((DUAL ? ((p - p_begin) & 1) : 0) * 8 + stage);
          if (elect_one()) {
            mbar_arrive_expect_tx(fb, a_slabs * ASLAB + b_bytes);
            for (int s = 0; s < a_slabs; ++s)
              tma_load_4d(sa + s * ASLAB, &map_dy, fb, co_blk * 128 + s * 64, w0, h0, n0);
#pragma unroll
            for (int s = 0; s < BN / 64; ++s)
              tma_load_4d(sb + s * g.bslab, &map_x, fb, ci_blk * BN + s * 64, w0 + q - 1, h0 - 1, n0);
          }
          __syncwarp();
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 || (DUAL && warp == 2)) {
      {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        constexpr int NI = DUAL ? 2 : 1;  // issuer warps
        const int which = warp - 1;
        const uint32_t tmem_d = tmem_base + which * 3 * BN;
        const uint32_t full_t = full_bar + 8 * which * 8;
        int stage = which % g.stages;
        uint32_t bits = 0;  // parity of this issuer's next fill, one bit per ring stage
        const int last = NI == 2 ? ((num_k - 1 - which) & ~1) + which : num_k - 1;
        for (int kb = which; kb < num_k; kb += NI) {
          mbar_wait(full_t + 8 * stage, (bits >> stage) & 1u);
          bits ^= 1u << stage;
          tc_fence_after();
          const uint32_t sa = ring + stage * stage_bytes;
          const uint32_t sb = sa + A_BYTES;
          // MN-major SW128: 64-channel slabs LBO apart, 8-pixel K groups SBO = 1024 B apart
          const uint64_t da = make_smem_desc_sw128(sa, ASLAB, 1024);
          const uint64_t db0 = make_smem_desc_sw128(sb, g.bslab, 1024);
          const uint32_t rstep = static_cast<uint32_t>(g.tw * 128) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              // vertical tap r reads the X tile r image rows (r * tw pixels = r * tw * 128 bytes) further down
              const uint64_t db = db0 + r * rstep;
#pragma unroll
              for (int k = 0; k < 4; ++k)  // 64 pixels = 4 x K16; 16 pixels = 2048 B -> +128 in (addr >> 4)
                umma_f16(tmem_d + r * BN, da + 128 * k, db + 128 * k, idesc, (kb >= NI || k != 0) ? 1u : 0u);
            }
            umma_commit(empty_bar + 8 * stage);
            if (kb == last) umma_commit(tfull_bar);
          }
          __syncwarp();
          stage += NI;
          while (stage >= g.stages) stage -= g.stages;
        }
      }
    } else if (warp >= kEpiWarp0) {
      const int qw = warp & 3;
      const int co = co_blk * 128 + qw * 32 + lane;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      // workspace layout [k slice][tile][3 vertical taps][128][BN]
      float* part = ws ? ws + ((static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 3 * 128 + (qw * 32 + lane)) * BN
                       : nullptr;
#pragma unroll 1
      for (int r = 0; r < 3; ++r) {
#pragma unroll 1
        for (int c32 = 0; c32 < BN / 32; ++c32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + r * BN + c32 * 32, v);
          tmem_ld_wait();
          if (DUAL && num_k >= 2) {  // the second issuer's accumulators
            uint32_t v2[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + 3 * BN + r * BN + c32 * 32, v2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          }
          if (part) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(part + static_cast<size_t>(r) * 128 * BN + c32 * 32 + j) =
                  make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int ci = ci_blk * BN + c32 * 32 + j;
              if (co < g.cout && ci < g.cin)
                atomicAdd(dw + (static_cast<int64_t>(co) * g.cin + ci) * 9 + r * 3 + q, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  } else if (ws && warp >= kEpiWarp0) {
    float* part = ws + ((static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 3 * 128 + ((warp & 3) * 32 + lane)) * BN;
    for (int r = 0; r < 3; ++r)
      for (int j = 0; j < BN; j += 4)
        *reinterpret_cast<uint4*>(part + static_cast<size_t>(r) * 128 * BN + j) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// Second phase of the split-K weight gradient: dW[co, ci, tap] += sum over k slices of the partial tiles the wgrad
// CTAs stored.  ws layout [k slice][tile][R][128][BN], tile = (tap group * co_blocks + co_blk) * ci_blocks + ci_blk.
// The first version combined partials with fp32 atomics from every CTA's epilogue: 2-7 M scattered red.global
// operations landing at the same moment cost 30-45 us per launch, more than the MMA main loop itself.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int ksplit, int tiles, int R, int BN,
                    int co_blocks, int ci_blocks, int cout, int cin, int num_taps, const TapTable taps, int mode3,
                    int kgroups) {
  pdl_wait();
  // block = E consecutive elements x kgroups k groups (E * kgroups == 256): with many k slices (50-150, each
  // per_slice floats apart) one thread per element would walk them serially, so up to 8 threads share an element
  // and fold through shared memory; with few slices kgroups == 1 and every thread owns an element.
  __shared__ float red[256];
  const int E = 256 / kgroups;
  const int el = threadIdx.x % E, kg = threadIdx.x / E;
  const int64_t per_slice = static_cast<int64_t>(tiles) * R * 128 * BN;  // multiple of 256
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * E; base < per_slice; base += static_cast<int64_t>(gridDim.x) * E) {
    const int64_t idx = base + el;
    const float* p = ws + idx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = kg;
    for (; k + 3 * kgroups < ksplit; k += 4 * kgroups) {
      s0 += p[k * per_slice];
      s1 += p[(k + kgroups) * per_slice];
      s2 += p[(k + 2 * kgroups) * per_slice];
      s3 += p[(k + 3 * kgroups) * per_slice];
    }
    for (; k < ksplit; k += kgroups) s0 += p[k * per_slice];
    float t = (s0 + s1) + (s2 + s3);
    if (kgroups > 1) {
      red[threadIdx.x] = t;
      __syncthreads();
      if (kg == 0)
        for (int i = 1; i < kgroups; ++i) t += red[i * E + el];
    }
    if (kg == 0) {
      const int ci_l = static_cast<int>(idx % BN);
      int64_t u = idx / BN;
      const int co_l = static_cast<int>(u % 128);
      u /= 128;
      const int r = static_cast<int>(u % R);
      const int tile = static_cast<int>(u / R);
      const int ci_blk = tile % ci_blocks;
      const int co_blk = (tile / ci_blocks) % co_blocks;
      const int tg = tile / (ci_blocks * co_blocks);
      const int co = co_blk * 128 + co_l, ci = ci_blk * BN + ci_l;
      if (co < cout && ci < cin) {
        const int tap = mode3 ? r * 3 + tg : taps.btap[tg];
        dw[(static_cast<int64_t>(co) * cin + ci) * num_taps + tap] += t;
      }
    }
    if (kgroups > 1) __syncthreads();
  }
}

static int launch_wgrad_reduce(const float* ws, float* dw, int ksplit, int tiles, int R, int BN, int co_blocks,
                               int ci_blocks, int cout, int cin, int num_taps, const TapTable& taps, int mode3,
                               cudaStream_t st) {
  const int64_t per_slice = (int64_t)tiles * R * 128 * BN;
  // one thread per element (1 KB contiguous per block and slice) unless that leaves most SMs without a block:
  // then up to 8 threads share an element's k slices
  int kgroups = 1;
  while (kgroups < 8 && (per_slice / 256) * kgroups < 2 * 148 && kgroups * 4 <= ksplit) kgroups <<= 1;
  const int E = 256 / kgroups;
  int64_t blocks = per_slice / E;
  if (blocks > 148 * 32) blocks = 148 * 32;
  NPP_LAUNCH((wgrad_reduce_kernel), (unsigned)blocks, 256, 0, st, ws, dw, ksplit, tiles, R, BN, co_blocks, ci_blocks, cout, cin,
                                                        num_taps, taps, mode3, kgroups);
  NPP_CHECK_LAUNCH("wgrad_reduce_kernel");
  return NPP_OK;
}

// =================================================================================================
// Host side: tensor maps, tile geometry, launches
// =================================================================================================
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// 4-D bf16 map over (C, W, H, N) with element strides (1, sw, sh, sn); box (64, bw, bh, bn).
static int encode_act_map(CUtensorMap* m, const void* ptr, int64_t C, int64_t W, int64_t H, int64_t N,
                          int64_t sw, int64_t sh, int64_t sn, int bw, int bh, int bn) {
  auto fn = get_encode_fn();
  if (!fn) return NPP_E_NODRIVER;
  if (W <= 0 || H <= 0 || N <= 0 || C <= 0) return NPP_E_INVALID;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled(act) failed: %d dims=(%lld,%lld,%lld,%lld) strides=(%lld,%lld,%lld) "
             "box=(64,%d,%d,%d) ptr=%p",
             (int)r, (long long)C, (long long)W, (long long)H, (long long)N, (long long)sw, (long long)sh,
             (long long)sn, bw, bh, bn, ptr);
    set_error_str(buf);
    return NPP_E_CUDA;
  }
  return NPP_OK;
}

// 3-D bf16 weight map over (K, taps, rows) of a dense [rows, taps, K] array; box (64, 1, brows).
static int encode_w_map(CUtensorMap* m, const void* ptr, int64_t K, int64_t taps, int64_t rows, int brows) {
  auto fn = get_encode_fn();
  if (!fn) return NPP_E_NODRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)taps, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * taps * 2};
  cuuint32_t box[3] = {64u, 1u, (cuuint32_t)brows};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(weight) failed: %d K=%lld taps=%lld rows=%lld", (int)r,
             (long long)K, (long long)taps, (long long)rows);
    set_error_str(buf);
    return NPP_E_CUDA;
  }
  return NPP_OK;
}

// Kernel-selection switches, read once from the environment (diagnostics / A-B timing; defaults = newest kernels):
//   NPP_CONV_EPI8=0  four-warp epilogue (conv_gemm_kernel) instead of conv_gemm2_kernel
//   NPP_CONV3=0      no 256-pixel halo-sharing kernel for 3x3 / stride-1 fprop + dgrad
//   NPP_CONV3_MIN_TILES=n  smallest number of 256-pixel tiles for which conv3_kernel is used (default 74)
//   NPP_CONV3_PAD_PCT=p    largest padded / real pixel ratio (percent) conv3_kernel accepts (default 107)
//   NPP_CONV3_2PROD=0      one TMA producer thread for both rings of conv3_kernel (default: two)
//   NPP_CONV3_TILE=i       only the i-th tile candidate {16x16, 32x8, 8x32} (experiments)
static int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

struct PixTile { int tw, th, tn; };
// Factor `prod` (a power of two) into a (tw, th, tn) box minimising padded work for a (W,H,N) grid.
static PixTile choose_tile(int64_t W, int64_t H, int64_t N, int prod) {
  PixTile best{prod, 1, 1};
  double best_cost = 1e300;
  for (int tw = 1; tw <= prod; tw <<= 1)
    for (int th = 1; tw * th <= prod; th <<= 1) {
      const int tn = prod / (tw * th);
      if (tw > 256 || th > 256 || tn > 256) continue;
      const double cost = (double)(cdiv64(W, tw) * tw) * (double)(cdiv64(H, th) * th) * (double)(cdiv64(N, tn) * tn);
      // prefer wide boxes on ties (longer contiguous runs per TMA row)
      if (cost < best_cost || (cost == best_cost && tw > best.tw)) {
        best_cost = cost;
        best = PixTile{tw, th, tn};
      }
    }
  return best;
}

struct PixSpace {  // the (W,H,N) pixel grid a launch tiles, possibly flattened to one dimension
  int64_t W, H, N;
};

static bool dense_pixels(const npp_view4* v) { return v->sh == (int64_t)v->w * v->sw && v->sn == (int64_t)v->h * v->sh; }

static bool bf16_view_ok(const npp_view4* v) { return view_ok(v, NPP_BF16); }

template <int BN>
static int launch_fprop(const Maps& maps, const Geom& g, const TapTable& taps, const float* bias, float* stats,
                        cudaStream_t st) {
  using Cfg = FpropCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_gemm)", e); return NPP_E_CUDA; }
    attr_set = true;
  }
  // grid = (CTAs per channel block) x n_blocks, one persistent CTA per SM
  int per_nb = sm_count() / g.n_blocks;
  if (per_nb > g.num_ptiles) per_nb = g.num_ptiles;
  const int grid = per_nb * g.n_blocks;
  static const int epi8 = env_flag("NPP_CONV_EPI8", 1);
  static const int epi2 = env_flag("NPP_CONV_EPI2", 1);
  if (epi8) {
    static bool attr2_set = false;
    if (!attr2_set) {
      cudaError_t e = cudaFuncSetAttribute(conv_gemm2_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv_gemm2_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_gemm2)", e); return NPP_E_CUDA; }
      attr2_set = true;
    }
    if (epi2) {
      NPP_LAUNCH((conv_gemm2_kernel<BN, true>), grid, kThreads2, Cfg::SMEM, st, maps, g, taps, bias, stats);
    } else {
      NPP_LAUNCH((conv_gemm2_kernel<BN, false>), grid, kThreads2, Cfg::SMEM, st, maps, g, taps, bias, stats);
    }
    NPP_CHECK_LAUNCH("conv_gemm2_kernel");
    return NPP_OK;
  }
  NPP_LAUNCH((conv_gemm_kernel<BN>), grid, kThreads, Cfg::SMEM, st, maps, g, taps, bias, stats);
  NPP_CHECK_LAUNCH("conv_gemm_kernel");
  return NPP_OK;
}

static int pick_bn(int cout) {
  if (cout <= 32) return 32;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  const int64_t w256 = cdiv64(cout, 256) * 256, w128 = cdiv64(cout, 128) * 128;
  return w256 <= w128 ? 256 : 128;
}

// One implicit-GEMM launch: D(pixel space of `d`) = sum_taps A(map, offset) * B(tap)
//   a_maps[i]: activation views addressed by the taps (already phase-decomposed), d: output view.
static int run_gemm(const npp_view4* a_views, int n_a, const npp_view4* d, const void* wmat, int wK, int wTaps,
                    int wRows, const TapTable& taps, int num_taps, bool allow_flatten, const float* bias,
                    float* stats, cudaStream_t st) {
  Maps maps;
  memset(&maps, 0, sizeof maps);
  Geom g;
  memset(&g, 0, sizeof g);
  bool flat = allow_flatten && n_a == 1 && num_taps == 1 && taps.dh[0] == 0 && taps.dw[0] == 0 &&
              dense_pixels(&a_views[0]) && dense_pixels(d) && a_views[0].n == d->n && a_views[0].h == d->h &&
              a_views[0].w == d->w;
  PixSpace ps{d->w, d->h, d->n};
  if (flat) ps = PixSpace{(int64_t)d->w * d->h * d->n, 1, 1};
  const PixTile t = choose_tile(ps.W, ps.H, ps.N, BM);
  int rc;
  for (int i = 0; i < n_a; ++i) {
    const npp_view4& v = a_views[i];
    if (!v.ptr) continue;
    if (flat)
      rc = encode_act_map(&maps.a[i], v.ptr, v.c, ps.W, 1, 1, v.sw, v.sw * ps.W, v.sw * ps.W, t.tw, t.th, t.tn);
    else
      rc = encode_act_map(&maps.a[i], v.ptr, v.c, v.w, v.h, v.n, v.sw, v.sh, v.sn, t.tw, t.th, t.tn);
    if (rc) return rc;
  }
  if (flat)
    rc = encode_act_map(&maps.d, d->ptr, d->c, ps.W, 1, 1, d->sw, d->sw * ps.W, d->sw * ps.W, t.tw, t.th, t.tn);
  else
    rc = encode_act_map(&maps.d, d->ptr, d->c, d->w, d->h, d->n, d->sw, d->sh, d->sn, t.tw, t.th, t.tn);
  if (rc) return rc;
  int bn = pick_bn(wRows);
  {
    // few pixel tiles (12x12 maps): a wide BN leaves most SMs without a CTA; narrower channel blocks share the pixel
    // tile through L2 and fill the machine
    const int64_t pt = cdiv64(ps.W, t.tw) * cdiv64(ps.H, t.th) * cdiv64(ps.N, t.tn);
    while (bn > 64 && pt * cdiv64(wRows, bn) * 2 <= sm_count()) bn >>= 1;
  }
  rc = encode_w_map(&maps.b, wmat, wK, wTaps, wRows, bn);
  if (rc) return rc;
  g.num_taps = num_taps;
  g.kc_blocks = (int)cdiv64(wK, BK);
  g.k_last = (int)cdiv64(wK - (int64_t)(g.kc_blocks - 1) * BK, 16);
  g.tw = t.tw; g.th = t.th; g.tn = t.tn;
  g.tiles_w = (int)cdiv64(ps.W, t.tw);
  g.tiles_h = (int)cdiv64(ps.H, t.th);
  g.tiles_n = (int)cdiv64(ps.N, t.tn);
  g.W = (int)ps.W; g.H = (int)ps.H; g.N = (int)ps.N;
  g.n_blocks = (int)cdiv64(wRows, bn);
  g.cout = wRows;
  const int64_t ptiles = (int64_t)g.tiles_w * g.tiles_h * g.tiles_n;
  if (ptiles * g.n_blocks > 0x7fffffff || g.n_blocks > sm_count()) return NPP_E_UNSUPPORTED;
  g.num_ptiles = (int)ptiles;
  switch (bn) {
    case 32: return launch_fprop<32>(maps, g, taps, bias, stats, st);
    case 64: return launch_fprop<64>(maps, g, taps, bias, stats, st);
    case 128: return launch_fprop<128>(maps, g, taps, bias, stats, st);
    default: return launch_fprop<256>(maps, g, taps, bias, stats, st);
  }
}

template <int BN>
static int launch_conv3(const C3Maps& maps, const C3Geom& g, int smem, const float* bias, float* stats, cudaStream_t st) {
  static bool attr_set = false;
  // two-group epilogue: measured a gain for the generic kernel with statistics (1x1 128 -> 128 @ 96^2: 41.3 -> 37.2 us,
  // 1024 -> 512: 292 -> 268 us) but not here (3x3 128 -> 128: 68.5 -> 70.5 us), profiles/r02_epi2_ab.txt: off for conv3
  static const int epi2 = env_flag("NPP_CONV3_EPI2", 0);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv3)", e); return NPP_E_CUDA; }
    attr_set = true;
  }
  int grid = sm_count();
  if (grid > g.num_ptiles) grid = g.num_ptiles;
  if (epi2) {
    NPP_LAUNCH((conv3_kernel<BN, true>), grid, kThreads2, smem, st, maps, g, bias, stats);
  } else {
    NPP_LAUNCH((conv3_kernel<BN, false>), grid, kThreads2, smem, st, maps, g, bias, stats);
  }
  NPP_CHECK_LAUNCH("conv3_kernel");
  return NPP_OK;
}

// 3x3 / stride 1 / pad 1 / dilation 1 fprop (flip = 0) or dgrad (flip = 1) through conv3_kernel.
// a: the tensor the taps read, d: the tensor produced (same N, H, W).  NPP_E_UNSUPPORTED = use the generic kernel.
static int conv3_try(const npp_view4* a, const npp_view4* d, const void* wmat, int wK, int wRows, int flip,
                     const float* bias, float* stats, cudaStream_t st) {
  static const int enabled = env_flag("NPP_CONV3", 1);
  static const int min_tiles = env_flag("NPP_CONV3_MIN_TILES", 74);
  static const int pad_pct = env_flag("NPP_CONV3_PAD_PCT", 107);  // tests raise it to reach the ragged-tile code
  if (!enabled) return NPP_E_UNSUPPORTED;
  if (a->n != d->n || a->h != d->h || a->w != d->w || wRows > 128) return NPP_E_UNSUPPORTED;
  // 256-pixel tile of one image; a one-row shift must be a whole number of 1024-byte swizzle atoms (tw % 8 == 0)
  static const int cand[3][2] = {{16, 16}, {32, 8}, {8, 32}};
  static const int force_tile = env_flag("NPP_CONV3_TILE", -1);  // experiments: 0 / 1 / 2 = only that candidate
  static const int two_prod = env_flag("NPP_CONV3_2PROD", 1);
  int tw = 0, th = 0;
  int64_t best = -1;
  for (int i = 0; i < 3; ++i) {
    if (force_tile >= 0 && i != force_tile) continue;
    const int64_t cost = cdiv64(d->w, cand[i][0]) * cand[i][0] * cdiv64(d->h, cand[i][1]) * cand[i][1];
    if (best < 0 || cost < best) { best = cost; tw = cand[i][0]; th = cand[i][1]; }
  }
  if (best * 100 > (int64_t)d->w * d->h * pad_pct) return NPP_E_UNSUPPORTED;  // padded work would eat the gain
  C3Geom g;
  memset(&g, 0, sizeof g);
  g.tw = tw; g.th = th;
  g.tiles_w = (int)cdiv64(d->w, tw);
  g.tiles_h = (int)cdiv64(d->h, th);
  const int64_t ptiles = (int64_t)g.tiles_w * g.tiles_h * d->n;
  if (ptiles < min_tiles || ptiles > 0x7fffffff) return NPP_E_UNSUPPORTED;
  g.num_ptiles = (int)ptiles;
  g.W = d->w; g.H = d->h;
  g.kc_blocks = (int)cdiv64(wK, BK);
  g.k_last = (int)cdiv64(wK - (int64_t)(g.kc_blocks - 1) * BK, 16);
  g.cout = wRows;
  g.flip = flip;
  g.two_producers = two_prod;
  g.a_bytes = (th + 2) * tw * 128;
  const int bn = pick_bn(wRows);
  const int b_bytes = bn * 128;
  const int avail = 227 * 1024 - kC3OutBytes - 1024 /*barriers*/ - 1024 /*alignment slack*/;
  static const int wres_on = env_flag("NPP_CONV3_WRES", 1);
  g.sa = 3;
  if (wres_on && g.kc_blocks == 1 && 9 * b_bytes + 3 * g.a_bytes <= avail) {
    // Cin <= 64 and the nine tap tiles fit beside the activation ring: weights stay resident for the whole kernel
    // (no weight ring, no per-tap barrier round trip, 4 instead of 13 tcgen05.commit per tile)
    g.wres = 1;
    g.sb = 9;
    g.sa = (avail - 9 * b_bytes) / g.a_bytes;
    if (g.sa > kC3MaxSA) g.sa = kC3MaxSA;
  } else {
    static const int sa_env = env_flag("NPP_CONV3_SA", 0);  // experiments: activation ring depth (the weight ring gets the rest)
    if (sa_env >= 2 && sa_env <= kC3MaxSA && sa_env * g.a_bytes + 3 * b_bytes <= avail) g.sa = sa_env;
    else if (g.sa * g.a_bytes + 4 * b_bytes > avail) g.sa = 2;
    g.sb = (avail - g.sa * g.a_bytes) / b_bytes;
    if (g.sb > kC3MaxSB) g.sb = kC3MaxSB;
    if (g.sb < 3) return NPP_E_UNSUPPORTED;
  }
  const int smem = g.sa * g.a_bytes + g.sb * b_bytes + kC3OutBytes + 1024 + 1024;
  C3Maps maps;
  memset(&maps, 0, sizeof maps);
  int rc = encode_act_map(&maps.a, a->ptr, a->c, a->w, a->h, a->n, a->sw, a->sh, a->sn, tw, th + 2, 1);
  if (rc) return rc;
  rc = encode_act_map(&maps.d, d->ptr, d->c, d->w, d->h, d->n, d->sw, d->sh, d->sn, tw, th / 2, 1);
  if (rc) return rc;
  rc = encode_w_map(&maps.b, wmat, wK, 9, wRows, bn);
  if (rc) return rc;
  switch (bn) {
    case 32: return launch_conv3<32>(maps, g, smem, bias, stats, st);
    case 64: return launch_conv3<64>(maps, g, smem, bias, stats, st);
    default: return launch_conv3<128>(maps, g, smem, bias, stats, st);
  }
}

static inline int floordiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

// phase (ph,pw) sub-view of v for stride-2 addressing
static npp_view4 phase_view(const npp_view4& v, int ph, int pw) {
  npp_view4 r = v;
  r.ptr = static_cast<char*>(v.ptr) + (ph * v.sh + pw * v.sw) * 2;
  r.h = (v.h - ph + 1) / 2;
  r.w = (v.w - pw + 1) / 2;
  r.sh = 2 * v.sh;
  r.sw = 2 * v.sw;
  if (r.h <= 0 || r.w <= 0) r.ptr = nullptr;
  return r;
}

static int check_conv_args(const npp_view4* x, const npp_view4* y, int kh, int kw, int stride, int pad, int dil) {
  if (!bf16_view_ok(x) || !bf16_view_ok(y)) return NPP_E_INVALID;
  if (kh <= 0 || kw <= 0 || kh * kw > kMaxTaps || (stride != 1 && stride != 2) || pad < 0 || dil < 1)
    return NPP_E_UNSUPPORTED;
  if (x->n != y->n) return NPP_E_INVALID;
  if (pad + dil * (kh > kw ? kh : kw) > 100) return NPP_E_UNSUPPORTED;  // int8 tap offsets
  return NPP_OK;
}

int conv_fwd(const npp_view4* x, const void* w, const float* bias, const npp_view4* y, int kh, int kw, int stride,
             int pad, int dil, int hoff, int woff, float* stats, cudaStream_t st) {
  int rc = check_conv_args(x, y, kh, kw, stride, pad, dil);
  if (rc) return rc;
  if (!w) return NPP_E_INVALID;
  if (kh == 3 && kw == 3 && stride == 1 && dil == 1 && pad == 1 && hoff == 0 && woff == 0) {
    rc = conv3_try(x, y, w, x->c, y->c, 0, bias, stats, st);
    if (rc != NPP_E_UNSUPPORTED) return rc;
  }
  TapTable taps;
  memset(&taps, 0, sizeof taps);
  npp_view4 av[4];
  memset(av, 0, sizeof av);
  int n_a = 1;
  if (stride == 1) {
    av[0] = *x;
  } else {
    n_a = 4;
  }
  int t = 0;
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s, ++t) {
      const int ah = -pad + r * dil + hoff, aw = -pad + s * dil + woff;
      if (stride == 1) {
        taps.map[t] = 0; taps.dh[t] = (int8_t)ah; taps.dw[t] = (int8_t)aw;
      } else {
        const int ph = ((ah % 2) + 2) % 2, pw = ((aw % 2) + 2) % 2;
        taps.map[t] = (int8_t)(ph * 2 + pw);
        taps.dh[t] = (int8_t)floordiv2(ah - ph);
        taps.dw[t] = (int8_t)floordiv2(aw - pw);
        if (!av[ph * 2 + pw].ptr) av[ph * 2 + pw] = phase_view(*x, ph, pw);
        if (!av[ph * 2 + pw].ptr) return NPP_E_UNSUPPORTED;
      }
      taps.btap[t] = (int8_t)t;
    }
  return run_gemm(av, n_a, y, w, x->c, kh * kw, y->c, taps, kh * kw, stride == 1, bias, stats, st);
}

int fill_zero_view_bf16(const npp_view4* v, cudaStream_t st);  // elementwise.cu

int conv_dgrad(const npp_view4* dy, const void* wt, const npp_view4* dx, int kh, int kw, int stride, int pad,
               int dil, int hoff, int woff, cudaStream_t st) {
  int rc = check_conv_args(dx, dy, kh, kw, stride, pad, dil);
  if (rc) return rc;
  if (!wt) return NPP_E_INVALID;
  if (kh == 3 && kw == 3 && stride == 1 && dil == 1 && pad == 1 && hoff == 0 && woff == 0) {
    rc = conv3_try(dy, dx, wt, dy->c, dx->c, 1, nullptr, nullptr, st);
    if (rc != NPP_E_UNSUPPORTED) return rc;
  }
  if (stride == 1) {
    TapTable taps;
    memset(&taps, 0, sizeof taps);
    int t = 0;
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s, ++t) {
        taps.map[t] = 0;
        taps.dh[t] = (int8_t)(pad - r * dil - hoff);
        taps.dw[t] = (int8_t)(pad - s * dil - woff);
        taps.btap[t] = (int8_t)t;
      }
    return run_gemm(dy, 1, dx, wt, dy->c, kh * kw, dx->c, taps, kh * kw, true, nullptr, nullptr, st);
  }
  // stride 2: each output phase of dx is its own dense convolution over dy with a tap subset
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      npp_view4 dv = phase_view(*dx, ph, pw);
      if (!dv.ptr) continue;
      TapTable taps;
      memset(&taps, 0, sizeof taps);
      int nt = 0;
      for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s) {
          const int eh = ph + pad - r * dil - hoff, ew = pw + pad - s * dil - woff;
          if ((eh & 1) || (ew & 1)) continue;
          taps.map[nt] = 0;
          taps.dh[nt] = (int8_t)(eh / 2);   // eh even: exact also for negatives
          taps.dw[nt] = (int8_t)(ew / 2);
          taps.btap[nt] = (int8_t)(r * kw + s);
          ++nt;
        }
      if (nt == 0) {
        rc = fill_zero_view_bf16(&dv, st);
      } else {
        rc = run_gemm(dy, 1, &dv, wt, dy->c, kh * kw, dx->c, taps, nt, false, nullptr, nullptr, st);
      }
      if (rc) return rc;
    }
  return NPP_OK;
}

template <int BN>
static int launch_wgrad(const WMaps& maps, const WGeom& g, const TapTable& taps, float* dw, float* ws, size_t ws_bytes,
                        cudaStream_t st) {
  using Cfg = WgradCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_wgrad)", e); return NPP_E_CUDA; }
    attr_set = true;
  }
  dim3 grid(g.num_taps * g.co_blocks * g.ci_blocks, g.ksplit);
  const size_t need = (size_t)grid.x * grid.y * 128 * BN * sizeof(float);
  // one k slice: its atomics are the only writers anyway; narrow layers (<= 64 x 64): most of the 128 x BN tile is
  // masked out, the few atomics are cheaper than storing and re-reading whole tiles
  if (g.ksplit == 1 || need > ws_bytes || (g.cout <= 64 && g.cin <= 64)) ws = nullptr;
  NPP_LAUNCH((conv_wgrad_kernel<BN>), grid, kThreads, Cfg::SMEM, st, maps, g, taps, dw, ws);
  NPP_CHECK_LAUNCH("conv_wgrad_kernel");
  if (ws)
    return launch_wgrad_reduce(ws, dw, g.ksplit, (int)grid.x, 1, BN, g.co_blocks, g.ci_blocks, g.cout, g.cin,
                               g.num_taps, taps, 0, st);
  return NPP_OK;
}

constexpr int kW3SmemMax = 225 * 1024;

template <int BN>
static int launch_wgrad3(const CUtensorMap& mx, const CUtensorMap& mdy, const W3Geom& g, float* dw, float* ws,
                         size_t ws_bytes, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kW3SmemMax);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_wgrad3)", e); return NPP_E_CUDA; }
    attr_set = true;
  }
  const int stage_bytes = 2 * 64 * 128 + (BN / 64) * g.bslab;
  const int smem = 2048 + g.stages * stage_bytes;
  dim3 grid(3 * g.co_blocks * g.ci_blocks, g.ksplit);
  const size_t need = (size_t)grid.x * grid.y * 3 * 128 * BN * sizeof(float);
  if (g.ksplit == 1 || need > ws_bytes || (g.cout <= 64 && g.cin <= 64)) ws = nullptr;
  NPP_LAUNCH((conv_wgrad3_kernel<BN>), grid, kThreads, smem, st, mx, mdy, g, dw, ws);
  NPP_CHECK_LAUNCH("conv_wgrad3_kernel");
  if (ws) {
    TapTable none;
    memset(&none, 0, sizeof none);
    return launch_wgrad_reduce(ws, dw, g.ksplit, (int)grid.x, 3, BN, g.co_blocks, g.ci_blocks, g.cout, g.cin, 9, none,
                               1, st);
  }
  return NPP_OK;
}

// 3x3 / stride 1 / dilation 1 / pad 1 weight gradient with vertical-tap sharing; NPP_E_UNSUPPORTED = not this shape.
static int conv_wgrad3(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, float* ws,
                       size_t ws_bytes, cudaStream_t st) {
  if (x->h != dy->h || x->w != dy->w) return NPP_E_UNSUPPORTED;
  // pixel tile (tw, th), tw * th == 64, tw >= 8 so that a one-row shift is a whole number of 1024-byte swizzle atoms
  int best_tw = 0, best_th = 0;
  double best_cost = 1e300;
  for (int tw = 8; tw <= 64; tw <<= 1) {
    const int th = 64 / tw;
    const double cost = (double)(cdiv64(dy->w, tw) * tw) * (double)(cdiv64(dy->h, th) * th);
    if (cost < best_cost || (cost == best_cost && th > best_th)) { best_cost = cost; best_tw = tw; best_th = th; }
  }
  W3Geom g;
  memset(&g, 0, sizeof g);
  g.tw = best_tw; g.th = best_th;
  g.tiles_w = (int)cdiv64(dy->w, g.tw);
  g.tiles_h = (int)cdiv64(dy->h, g.th);
  const int64_t ptiles = (int64_t)g.tiles_w * g.tiles_h * dy->n;
  if (ptiles > 0x7fffffff) return NPP_E_UNSUPPORTED;
  g.total_ptiles = (int)ptiles;
  const int bn = dw_cin <= 64 ? 64 : 128;
  g.co_blocks = (int)cdiv64(dw_cout, 128);
  g.ci_blocks = (int)cdiv64(dw_cin, bn);
  g.cout = dw_cout; g.cin = dw_cin;
  g.bslab = (g.th + 2) * g.tw * 128;
  const int stage_bytes = 2 * 64 * 128 + (bn / 64) * g.bslab;
  int stages = (kW3SmemMax - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return NPP_E_UNSUPPORTED;
  g.stages = stages;
  const int out_tiles = 3 * g.co_blocks * g.ci_blocks;
  int ks = sm_count() / out_tiles;
  if (ks > g.total_ptiles / 2) ks = g.total_ptiles / 2;
  if (ks < 1) ks = 1;
  g.ksplit = ks;
  CUtensorMap mx, mdy;
  int rc = encode_act_map(&mx, x->ptr, x->c, x->w, x->h, x->n, x->sw, x->sh, x->sn, g.tw, g.th + 2, 1);
  if (rc) return rc;
  rc = encode_act_map(&mdy, dy->ptr, dy->c, dy->w, dy->h, dy->n, dy->sw, dy->sh, dy->sn, g.tw, g.th, 1);
  if (rc) return rc;
  return bn == 64 ? launch_wgrad3<64>(mx, mdy, g, dw, ws, ws_bytes, st)
                  : launch_wgrad3<128>(mx, mdy, g, dw, ws, ws_bytes, st);
}

int conv_wgrad(const npp_view4* x, const npp_view4* dy, float* dw, int dw_cout, int dw_cin, int kh, int kw,
               int stride, int pad, int dil, int hoff, int woff, float* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = check_conv_args(x, dy, kh, kw, stride, pad, dil);
  if (rc) return rc;
  if (!dw || dw_cout <= 0 || dw_cin <= 0 || dw_cout > dy->c || dw_cin > x->c) return NPP_E_INVALID;
  if (ws && (reinterpret_cast<uintptr_t>(ws) & 15)) return NPP_E_INVALID;
  if (kh == 3 && kw == 3 && stride == 1 && dil == 1 && pad == 1 && hoff == 0 && woff == 0) {
    rc = conv_wgrad3(x, dy, dw, dw_cout, dw_cin, ws, ws_bytes, st);
    if (rc != NPP_E_UNSUPPORTED) return rc;
  }
  TapTable taps;
  memset(&taps, 0, sizeof taps);
  npp_view4 av[4];
  memset(av, 0, sizeof av);
  if (stride == 1) av[0] = *x;
  int t = 0;
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s, ++t) {
      const int ah = -pad + r * dil + hoff, aw = -pad + s * dil + woff;
      if (stride == 1) {
        taps.map[t] = 0; taps.dh[t] = (int8_t)ah; taps.dw[t] = (int8_t)aw;
      } else {
        const int ph = ((ah % 2) + 2) % 2, pw = ((aw % 2) + 2) % 2;
        taps.map[t] = (int8_t)(ph * 2 + pw);
        taps.dh[t] = (int8_t)floordiv2(ah - ph);
        taps.dw[t] = (int8_t)floordiv2(aw - pw);
        if (!av[ph * 2 + pw].ptr) av[ph * 2 + pw] = phase_view(*x, ph, pw);
        if (!av[ph * 2 + pw].ptr) return NPP_E_UNSUPPORTED;
      }
      taps.btap[t] = (int8_t)t;
    }
  const bool flat = stride == 1 && kh * kw == 1 && taps.dh[0] == 0 && taps.dw[0] == 0 && dense_pixels(x) &&
                    dense_pixels(dy) && x->h == dy->h && x->w == dy->w;
  PixSpace ps{dy->w, dy->h, dy->n};
  if (flat) ps = PixSpace{(int64_t)dy->w * dy->h * dy->n, 1, 1};
  const PixTile pt = choose_tile(ps.W, ps.H, ps.N, 64);
  WMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < 4; ++i) {
    const npp_view4& v = av[i];
    if (!v.ptr) continue;
    if (flat)
      rc = encode_act_map(&maps.x[i], v.ptr, v.c, ps.W, 1, 1, v.sw, v.sw * ps.W, v.sw * ps.W, pt.tw, pt.th, pt.tn);
    else
      rc = encode_act_map(&maps.x[i], v.ptr, v.c, v.w, v.h, v.n, v.sw, v.sh, v.sn, pt.tw, pt.th, pt.tn);
    if (rc) return rc;
  }
  if (flat)
    rc = encode_act_map(&maps.dy, dy->ptr, dy->c, ps.W, 1, 1, dy->sw, dy->sw * ps.W, dy->sw * ps.W, pt.tw, pt.th, pt.tn);
  else
    rc = encode_act_map(&maps.dy, dy->ptr, dy->c, dy->w, dy->h, dy->n, dy->sw, dy->sh, dy->sn, pt.tw, pt.th, pt.tn);
  if (rc) return rc;
  WGeom g;
  memset(&g, 0, sizeof g);
  g.num_taps = kh * kw;
  g.tw = pt.tw; g.th = pt.th; g.tn = pt.tn;
  g.tiles_w = (int)cdiv64(ps.W, pt.tw);
  g.tiles_h = (int)cdiv64(ps.H, pt.th);
  g.tiles_n = (int)cdiv64(ps.N, pt.tn);
  const int64_t ptiles = (int64_t)g.tiles_w * g.tiles_h * g.tiles_n;
  if (ptiles > 0x7fffffff) return NPP_E_UNSUPPORTED;
  g.total_ptiles = (int)ptiles;
  const int bn = dw_cin <= 64 ? 64 : (dw_cin <= 128 ? 128 : 256);
  g.co_blocks = (int)cdiv64(dw_cout, 128);
  g.ci_blocks = (int)cdiv64(dw_cin, bn);
  g.cout = dw_cout; g.cin = dw_cin; g.dw_ld = dw_cin;
  const int out_tiles = g.num_taps * g.co_blocks * g.ci_blocks;
  int ks = sm_count() / out_tiles;
  if (ks < 1) ks = 1;
  if (ks > g.total_ptiles / 2) ks = g.total_ptiles / 2;
  if (ks < 1) ks = 1;
  g.ksplit = ks;
  switch (bn) {
    case 64: return launch_wgrad<64>(maps, g, taps, dw, ws, ws_bytes, st);
    case 128: return launch_wgrad<128>(maps, g, taps, dw, ws, ws_bytes, st);
    default: return launch_wgrad<256>(maps, g, taps, dw, ws, ws_bytes, st);
  }
}

// fp32 master weight in torch's OIHW layout [cout, cin, taps] -> packed OHWI [cout_pad, taps, cin_pad] and/or its
// transpose [cin_pad, taps, cout_pad] (zero padded), as bf16 (tcgen05 path) or fp32 (validation mode).
template <typename D>
__global__ void pack_weight_kernel(const float* __restrict__ w32, D* __restrict__ w, D* __restrict__ wt, int cout,
                                   int taps, int cin, int cout_pad, int cin_pad) {
  pdl_wait();
  const int64_t total = (int64_t)cout_pad * taps * cin_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int t = (int)((i / cin_pad) % taps);
    const int co = (int)(i / ((int64_t)cin_pad * taps));
    const float v = (co < cout && ci < cin) ? w32[((int64_t)co * cin + ci) * taps + t] : 0.f;
    const D h = from_f<D>(v);
    if (w) w[i] = h;
    if (wt) wt[((int64_t)ci * taps + t) * cout_pad + co] = h;
  }
}

// All dense-conv weights of a model in ONE launch (the per-conv pack_weight launches were 412 tiny kernels per
// step): a device table of tensor descriptors, one block per chunk of chunk_elems packed elements.
struct PackTensor {
  const float* w32;
  __nv_bfloat16* w;
  __nv_bfloat16* wt;
  int cout, taps, cin, cout_pad, cin_pad, pad_;
};
static_assert(sizeof(PackTensor) == 48, "table layout is part of the ABI (npp_b200/engine.py mirrors it)");

// Pixel-pair layout (PackTensor.pad_ == 1; round-2 candidate, see functional._ConvPairFn): a 3x3 / stride-1 / pad-1
// convolution with 32 input and 32 output channels on an NHWC tensor [N, H, W, 32] is the same arithmetic as a
// 3x3 convolution with 64 -> 64 channels on the SAME memory read as [N, H, W/2, 64] (two neighbouring pixels = one
// "super-pixel"): output (po, co) of super-pixel j reads input (pi, ci) of super-pixel j + ds through the original
// horizontal tap s = 2*ds + pi - po + 1 when 0 <= s <= 2 (zero otherwise); vertical taps are unchanged.  Rows
// of the packed matrix = po*32 + co, columns = pi*32 + ci, horizontal tap index = ds + 1.
__device__ __forceinline__ float pair_weight(const float* __restrict__ w32, int row, int tap, int col) {
  const int po = row >> 5, co = row & 31, pi = col >> 5, ci = col & 31;
  const int r = tap / 3, s = 2 * (tap % 3 - 1) + pi - po + 1;
  return (s >= 0 && s <= 2) ? w32[(co * 32 + ci) * 9 + r * 3 + s] : 0.f;
}

__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const PackTensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                          const int* __restrict__ chunk_index, int chunk_elems) {
  pdl_wait();
  const PackTensor t = tensors[chunk_tensor[blockIdx.x]];
  if (t.pad_ == 1) {
    const int total = 64 * 9 * 64;
    const int begin = chunk_index[blockIdx.x] * chunk_elems;
    const int end = begin + chunk_elems < total ? begin + chunk_elems : total;
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
      const int col = i & 63, tap = (i >> 6) % 9, row = i / (64 * 9);
      const __nv_bfloat16 h = __float2bfloat16_rn(pair_weight(t.w32, row, tap, col));
      if (t.w) t.w[i] = h;
      if (t.wt) t.wt[(col * 9 + tap) * 64 + row] = h;
    }
    return;
  }
  const int64_t total = (int64_t)t.cout_pad * t.taps * t.cin_pad;
  const int64_t begin = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
  int64_t end = begin + chunk_elems;
  if (end > total) end = total;
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int ci = (int)(i % t.cin_pad);
    const int tp = (int)((i / t.cin_pad) % t.taps);
    const int co = (int)(i / ((int64_t)t.cin_pad * t.taps));
    const float v = (co < t.cout && ci < t.cin) ? t.w32[((int64_t)co * t.cin + ci) * t.taps + tp] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    if (t.w) t.w[i] = h;
    if (t.wt) t.wt[((int64_t)ci * t.taps + tp) * t.cout_pad + co] = h;
  }
}

// Tile form of the multi-tensor pack (round-2 candidate, functional._state["pack_tiles"]): the element-wise kernel
// above reads the OIHW master with a stride of `taps` floats between threads and writes the transposed copy as
// scattered 2-byte stores (0.79 ms per step for 77 M weights against ~0.1 ms of HBM time).  Here a block owns a
// 32 (Cout) x 32 (Cin) x taps tile: the rows are contiguous runs of the master (coalesced fp32 loads into shared
// memory), both packed layouts are then written as 64-byte runs; the row stride 32*taps + 1 keeps both shared-memory
// read patterns conflict-free for odd tap counts (1, 9).
constexpr int kPackTapsMax = 9;
__global__ void __launch_bounds__(256)
pack_weights_tiles_kernel(const PackTensor* __restrict__ tensors, const int* __restrict__ tile_tensor,
                          const int* __restrict__ tile_index) {
  pdl_wait();
  __shared__ float tile[32 * (32 * kPackTapsMax + 1)];
  const PackTensor t = tensors[tile_tensor[blockIdx.x]];
  const int ci_tiles = (t.cin_pad + 31) / 32;
  const int co0 = (tile_index[blockIdx.x] / ci_tiles) * 32, ci0 = (tile_index[blockIdx.x] % ci_tiles) * 32;
  const int taps = t.taps, ld = 32 * taps + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int nci = t.cin - ci0;            // real input channels in this tile (<= 0: all padding)
  if (nci > 32) nci = 32;
  for (int r = warp; r < 32; r += 8) {
    const int co = co0 + r;
    const float* src = t.w32 + ((int64_t)co * t.cin + ci0) * taps;   // nci * taps contiguous floats
    for (int k = lane; k < 32 * taps; k += 32)
      tile[r * ld + k] = (co < t.cout && k < nci * taps) ? src[k] : 0.f;
  }
  __syncthreads();
  // w[co][tap][ci]: one 32-channel run per (row, tap)
  if (t.w)
    for (int q = warp; q < 32 * taps; q += 8) {
      const int r = q / taps, tp = q % taps;
      if (co0 + r < t.cout_pad && ci0 + lane < t.cin_pad)
        t.w[((int64_t)(co0 + r) * taps + tp) * t.cin_pad + ci0 + lane] = __float2bfloat16_rn(tile[r * ld + lane * taps + tp]);
    }
  // wt[ci][tap][co]: one 32-row run per (channel, tap)
  if (t.wt)
    for (int q = warp; q < 32 * taps; q += 8) {
      const int c = q / taps, tp = q % taps;
      if (ci0 + c < t.cin_pad && co0 + lane < t.cout_pad)
        t.wt[((int64_t)(ci0 + c) * taps + tp) * t.cout_pad + co0 + lane] = __float2bfloat16_rn(tile[lane * ld + c * taps + tp]);
    }
}

int pack_weights_tiles(const void* table, int ntensors, const int* tile_tensor, const int* tile_index, int ntiles,
                       cudaStream_t st) {
  if (!table || !tile_tensor || !tile_index || ntensors <= 0 || ntiles < 0) return NPP_E_INVALID;
  if (ntiles == 0) return NPP_OK;
  NPP_LAUNCH((pack_weights_tiles_kernel), ntiles, 256, 0, st, static_cast<const PackTensor*>(table), tile_tensor, tile_index);
  NPP_CHECK_LAUNCH("pack_weights_tiles_kernel");
  return NPP_OK;
}

int pack_weights_multi(const void* table, int ntensors, const int* chunk_tensor, const int* chunk_index, int nchunks,
                       int chunk_elems, cudaStream_t st) {
  if (!table || !chunk_tensor || !chunk_index || ntensors <= 0 || nchunks < 0 || chunk_elems <= 0) return NPP_E_INVALID;
  if (nchunks == 0) return NPP_OK;
  NPP_LAUNCH((pack_weights_multi_kernel), nchunks, 256, 0, st, static_cast<const PackTensor*>(table), chunk_tensor, chunk_index,
                                                     chunk_elems);
  NPP_CHECK_LAUNCH("pack_weights_multi_kernel");
  return NPP_OK;
}

// dW[co, ci, r, s] += sum over the pair-layout entries that carry that tap (inverse of pair_weight): the weight
// gradient of the 64 -> 64 super-pixel convolution folded back onto the 32 x 32 x 3 x 3 master gradient.
__global__ void __launch_bounds__(256) fold_pair_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (co, ci, r, s)
  if (i >= 32 * 32 * 9) return;
  const int s = i % 3, r = (i / 3) % 3, ci = (i / 9) & 31, co = i / (9 * 32);
  float acc = 0.f;
#pragma unroll
  for (int po = 0; po < 2; ++po)
#pragma unroll
    for (int pi = 0; pi < 2; ++pi) {
      const int t2 = s - 1 - pi + po;       // 2 * ds
      if (t2 & 1) continue;
      const int ds = t2 / 2;                // exact: t2 even (also for negatives)
      if (ds < -1 || ds > 1) continue;
      // dwp is torch OIHW of the paired conv: [64 rows][64 cols][3][3]
      acc += dwp[((po * 32 + co) * 64 + (pi * 32 + ci)) * 9 + r * 3 + (ds + 1)];
    }
  dw[i] += acc;
}

int fold_pair_wgrad(const float* dwp, float* dw, cudaStream_t st) {
  if (!dwp || !dw) return NPP_E_INVALID;
  NPP_LAUNCH((fold_pair_wgrad_kernel), (32 * 32 * 9 + 255) / 256, 256, 0, st, dwp, dw);
  NPP_CHECK_LAUNCH("fold_pair_wgrad_kernel");
  return NPP_OK;
}

// one-off packing of a single pair-layout weight (tests; the per-step path is pack_weights_multi with pad_ = 1)
__global__ void __launch_bounds__(256) pack_pair_kernel(const float* __restrict__ w32, __nv_bfloat16* __restrict__ w,
                                                        __nv_bfloat16* __restrict__ wt) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 9 * 64) return;
  const int col = i & 63, tap = (i >> 6) % 9, row = i / (64 * 9);
  const __nv_bfloat16 h = __float2bfloat16_rn(pair_weight(w32, row, tap, col));
  if (w) w[i] = h;
  if (wt) wt[(col * 9 + tap) * 64 + row] = h;
}

int pack_weight_pair(const float* w32, void* w, void* wt, cudaStream_t st) {
  if (!w32 || (!w && !wt)) return NPP_E_INVALID;
  NPP_LAUNCH((pack_pair_kernel), (64 * 9 * 64 + 255) / 256, 256, 0, st, w32, static_cast<__nv_bfloat16*>(w),
                                                              static_cast<__nv_bfloat16*>(wt));
  NPP_CHECK_LAUNCH("pack_pair_kernel");
  return NPP_OK;
}

int pack_weight(const float* w32, void* w, void* wt, int cout, int taps, int cin, int cout_pad, int cin_pad,
                int out_dtype, cudaStream_t st) {
  if (!w32 || (!w && !wt) || cout <= 0 || taps <= 0 || cin <= 0 || cout_pad < cout || cin_pad < cin)
    return NPP_E_INVALID;
  const int64_t total = (int64_t)cout_pad * taps * cin_pad;
  int grid = (int)((total + 255) / 256);
  if (grid > 4096) grid = 4096;
  if (out_dtype == NPP_BF16)
    NPP_LAUNCH((pack_weight_kernel<__nv_bfloat16>), grid, 256, 0, st, w32, static_cast<__nv_bfloat16*>(w),
                                                             static_cast<__nv_bfloat16*>(wt), cout, taps, cin, cout_pad,
                                                             cin_pad);
  else if (out_dtype == NPP_F32)
    NPP_LAUNCH((pack_weight_kernel<float>), grid, 256, 0, st, w32, static_cast<float*>(w), static_cast<float*>(wt), cout, taps,
                                                    cin, cout_pad, cin_pad);
  else
    return NPP_E_UNSUPPORTED;
  NPP_CHECK_LAUNCH("pack_weight_kernel");
  return NPP_OK;
}

}  // namespace tc
}  // namespace npp

#ifdef NPP_C3_PROF
extern "C" int npp_debug_c3_prof(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, npp::tc::g_c3_prof, sizeof(unsigned long long) * 32);
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(npp::tc::g_c3_prof, z, sizeof z);
  }
  return 0;
}
#endif
