// Prefill-regime kernel: warp-specialised TMA -> dequant -> tcgen05.mma pipeline with TMEM accumulators.
//
// Replaces the reference's `_gemm_kernel` main loop (conch/kernels/quantization/gemm.py:337-401: register dequant in the
// critical path of every `tl.dot`, 128x64x64 tile, 2 stages).  Orientation is swap-AB: the MMA computes
//     D^T[n, m] = sum_k W^T[n, k] * x[m, k]
// so the 128 TMEM lanes are 128 weight columns n and the MMA-N dimension is the activation tile m_tile (16..256).
// That makes (a) every dequant thread own ONE weight column -- scale / zero-point live in registers, a packed word
// (8 consecutive k of one column, conch/third_party/vllm/quant_utils.py:119-124) turns into one 16-byte K-major chunk,
// (b) small M cheap (MMA-N = 16 instead of 112 wasted rows of a 128-row tile), (c) x a plain K-major TMA tile.
//
// Roles (640 threads):  warp 0         TMA producer   (packed W tile, 8|16 rows x 128 int32, deep ring)
//                       warp 3         TMA producer   (x tile, 128B-swizzled; multicast across a cluster if enabled)
//                       warp 1         MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, M=128, N=m_tile, K=16)
//                       warp 2         TMEM allocator; second MMA issuer (own accumulator) when m_tile <= 128
//                       warps 4..19    dequant        (smem packed -> registers -> dequant_word() -> operand A): four
//                                      SETS of four warps; set j owns stages j, j+4, ... (a thread = one weight column,
//                                      all 64 k of the stage), so four stages are being dequantised concurrently and a
//                                      set's barrier / shared-load / tcgen05.st latencies hide behind the other sets;
//                                      then epilogue  (tcgen05.ld -> cast -> global)
// Operand A (dequantised W) goes either to shared memory in the canonical K-major SWIZZLE_128B layout (variant SS) or
// straight into tensor memory with tcgen05.st (variant TS: the weights never touch shared memory again, which halves
// the smem traffic per MMA -- the limiter of this kernel).
// Barriers: x ring x_full (TMA tx); W ring w_full (TMA tx) / w_free (the 4 warps of a set); A ring a_ready (the 4 warps
// of a set) / mma_done (tcgen05.commit: releases the A slot AND the x slot of that K block); d_full.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "dequant.cuh"
#include "host_util.cuh"
#include "tc_ptx.cuh"

namespace cb200 {

TcTuning g_tc_tuning = {0, 0};
int g_tc_cluster = 0;  // 0 = default (no cluster)
#ifdef CB200_TC_TRACE
long long* g_tc_trace = nullptr;
extern "C" __attribute__((visibility("default"))) void cb200_debug_set_tc_trace(void* p) { g_tc_trace = reinterpret_cast<long long*>(p); }
#endif

// ------------------------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------------------------
struct TcParams {
  Problem p;
  int kblocks_total;   // K / 64
  // 1-D grid of work items: items [0, t_full) are whole tiles (n fastest); the remaining t_tail tiles -- the partial last
  // wave, or all tiles when there are too few to fill the GPU -- are split `splits` ways along K and come last, split-major
  // (item t_full + z * t_tail + i = K slice z of tile t_full + i), so the last wave keeps every SM busy.
  int n_tiles;         // tiles along N
  int t_full, t_tail, splits;
  int kb_per_split;    // 64-wide K blocks per K slice
  float* ws_partial;   // [splits][t_tail][m_tile][128] fp32, tile-local
  int cluster;         // CTAs per cluster along n (1, 2 or 4): they share one x tile through TMA multicast
#ifdef CB200_TC_TRACE
  long long* trace;    // [K/64][16] clock64 stamps of CTA (0,0,0) (debug builds only)
#endif
};

template <int MT, int BITS, bool TS>
struct TcCfg {
  static constexpr int NTILE = 128;
  static constexpr int KB = 64;
  static constexpr int PF = 32 / BITS;
  static constexpr int WROWS = KB / PF;                 // packed rows per stage: 8 (int4) / 16 (int8)
  static constexpr int X_BYTES = MT * 128;              // MT rows x 64 halfs, 128B swizzle
  static constexpr int W_BYTES = WROWS * NTILE * 4;
  static constexpr int A_BYTES = TS ? 0 : NTILE * 128;  // dequantised W tile, K-major 128B swizzle
  static constexpr int SMEM_BUDGET = 227 * 1024 - 2048;
  // Three independent rings (a slot is released as soon as ITS consumer is done with it, not when the MMA retires):
  //   W ring  (packed weights, 4-8 KB per slot): TMA -> dequant warps; deep, it hides the HBM latency of the weights
  //   A ring  (dequantised tile: 32 TMEM columns per slot, or 16 KB of shared memory in SS mode): dequant -> MMA
  //   x ring  (activation tile, MT x 128 B per slot): TMA -> MMA
  // SW and SA must be multiples of the number of dequant sets (4): then slot s is always served by set s % 4, which
  // therefore observes EVERY phase of that slot's barriers in order.  (A set that skipped a phase could pass a parity
  // wait one phase early -- mbarrier parity only distinguishes adjacent phases.)  The x ring is only touched by
  // single threads that visit every iteration, so any depth <= SA works (its slots are released by the same
  // tcgen05.commit as the A slots: ONE commit per K block keeps the MMA issue loop short).
  static constexpr int SW = BITS == 4 ? (MT <= 128 ? 16 : 8) : (MT <= 128 ? 8 : 4);
  static constexpr int SA = TS ? 8 : 4;
  static constexpr int SX_FIT = (SMEM_BUDGET - SW * W_BYTES - SA * A_BYTES) / X_BYTES;
  static constexpr int SX = SX_FIT > SA ? SA : SX_FIT;   // <= SA: the x producer reuses the A ring's mma_done barriers
  // Two MMA issuers on two accumulators for the small tiles: below m_tile ~256 the kernel is bound by the ~55 cycles
  // one thread needs per tcgen05.mma issue, not by the tensor pipe, so warps 1 and 2 take alternate rounds of K blocks
  // into D0 / D1 (summed in the epilogue) and the issue rate doubles.
  static constexpr bool DUAL = TS && MT <= 128;
  static constexpr int A_COL0 = DUAL ? 2 * MT : MT;     // first TMEM column of the A ring (TS)
  static constexpr int TMEM_NEED = A_COL0 + (TS ? SA * 32 : 0);
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = SX * X_BYTES + SW * W_BYTES + SA * A_BYTES + 1024 /*alignment slack*/ + 1024 /*barriers*/;
  static_assert(SX >= 3, "x ring too shallow");
  static_assert(TMEM_NEED <= 512, "TMEM over-subscribed");
};

constexpr int kTcSets = 4;                       // dequant warp sets (4 warps each)
constexpr int kTcThreads = 128 + 128 * kTcSets;  // 640

template <class A, int BITS, int MT, bool TS, int ZK>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, TcParams prm) {
  using Cfg = TcCfg<MT, BITS, TS>;
  constexpr int SX = Cfg::SX, SW = Cfg::SW, SA = Cfg::SA;
  constexpr bool FZP = ZK == CB200_ZP_ACT;
  constexpr int PAIRS = 16 / BITS;     // pair registers per packed word
  constexpr int WPT = 64 / Cfg::PF;    // packed words per dequant thread per stage (all 64 k of one column): 8 | 16
  constexpr int SETS = kTcSets;
  static_assert(SW % kTcSets == 0 && SA % kTcSets == 0, "ring slot <-> dequant set mapping must be fixed (mbarrier phase order)");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = smem_base;
  const uint32_t a_base = x_base + SX * Cfg::X_BYTES;
  const uint32_t w_base = a_base + SA * Cfg::A_BYTES;
  const uint32_t bar_base = w_base + SW * Cfg::W_BYTES;
  auto x_full = [&](int s) { return bar_base + 8u * s; };
  auto w_full = [&](int s) { return bar_base + 8u * (SX + s); };
  auto w_free = [&](int s) { return bar_base + 8u * (SX + SW + s); };
  auto a_ready = [&](int s) { return bar_base + 8u * (SX + 2 * SW + s); };
  auto mma_done = [&](int s) { return bar_base + 8u * (SX + 2 * SW + SA + s); };  // frees A slot s and its x slot
  const uint32_t d_full = bar_base + 8u * (SX + 2 * SW + 2 * SA);
  const uint32_t tmem_slot = d_full + 8u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base

  const Problem& p = prm.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = int(blockIdx.x), zsplit = 0;
  const bool split = int(blockIdx.x) >= prm.t_full && prm.splits > 1;
  if (int(blockIdx.x) >= prm.t_full) {
    const int r = int(blockIdx.x) - prm.t_full;
    zsplit = r / prm.t_tail;
    tile = prm.t_full + (r - zsplit * prm.t_tail);
  }
  const int n0 = (tile % prm.n_tiles) * Cfg::NTILE;
  const int m0 = (tile / prm.n_tiles) * MT;
  const int kb0 = split ? zsplit * prm.kb_per_split : 0;
  const int kb1 = split ? min(kb0 + prm.kb_per_split, prm.kblocks_total) : prm.kblocks_total;
  const int nk = kb1 - kb0;

  // ---- one-time setup ------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_w);
    for (int s = 0; s < SX; ++s) {
      mbar_init(x_full(s), 1);
    }
    for (int s = 0; s < SW; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_free(s), 4);  // the four warps of the owning set
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_ready(s), 4);
      mbar_init(mma_done(s), prm.cluster);  // every CTA of the cluster must have released the x slot (x is multicast)
    }
    mbar_init(d_full, Cfg::DUAL ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  if (prm.cluster > 1) {
    cluster_sync_all();  // peers multicast into this CTA's shared memory and barriers: they must be initialised first
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t crank = prm.cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = uint16_t((1u << prm.cluster) - 1u);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
#ifdef CB200_TC_TRACE
  const bool tracing = prm.trace && int(blockIdx.x) == 7 + prm.n_tiles;
#define TRACE(slot, it) do { if (tracing) prm.trace[(it) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot, it) do { } while (0)
#endif

  // The three single-thread roles below all follow one pattern: the WHOLE warp runs the loop convergently, lanes poll
  // the barriers of several K blocks in ONE try_wait (different addresses per lane; a try_wait on an already completed
  // barrier still costs ~140 cycles, so they must not be serialised), then one elected lane issues for all of them.
  // (`elect.sync` instead of `lane == 0`: the compiler then emits straight-line UTMALDG / UTCHMMA code rather than a
  // per-active-thread uniformisation loop around every such instruction.)
  if (warp == 0) {
    // ================================ TMA producer: packed weights ==================================================
    constexpr int UW = 4;
    for (int it = 0; it < nk; it += UW) {
      const int nb = min(UW, nk - it);
      {
        int i = it + (lane & (UW - 1));
        if (i >= nk) i = it;
        mbar_wait(w_free(i % SW), ((i / SW) & 1) ^ 1);
      }
      __syncwarp();
      TRACE(0, it);
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < UW; ++u) {
          if (u < nb) {
            const int s = (it + u) % SW;
            mbar_expect_tx(w_full(s), Cfg::W_BYTES);
            tma_load_2d(w_base + s * Cfg::W_BYTES, &tmap_w, n0, (kb0 + it + u) * Cfg::WROWS, w_full(s));
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ================================ TMA producer: activations =====================================================
    constexpr int UX = 2;
    for (int it = 0; it < nk; it += UX) {
      const int nb = min(UX, nk - it);
      {
        int i = it + (lane & (UX - 1));
        if (i >= nk) i = it;
        if (i >= SX) {  // slot last read by the MMAs of K block i - SX
          const int j = i - SX;
          mbar_wait(mma_done(j % SA), (j / SA) & 1);
        }
      }
      __syncwarp();
      TRACE(1, it);
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < UX; ++u) {
          if (u < nb) {
            const int s = (it + u) % SX;
            mbar_expect_tx(x_full(s), Cfg::X_BYTES);
            if (prm.cluster > 1) {
              // this CTA fetches rows [crank * MT/cluster, +MT/cluster) of the shared x tile and multicasts them to all
              const uint32_t rows = MT / prm.cluster;
              tma_load_2d_mc(x_base + s * Cfg::X_BYTES + crank * rows * 128, &tmap_x, (kb0 + it + u) * Cfg::KB,
                             m0 + int(crank * rows), x_full(s), cmask);
            } else {
              tma_load_2d(x_base + s * Cfg::X_BYTES, &tmap_x, (kb0 + it + u) * Cfg::KB, m0, x_full(s));
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 || (Cfg::DUAL && warp == 2)) {
    // ================================ MMA issuer(s) ===============================================================
    // The whole warp runs the loop convergently and ONE elected lane issues: with `elect.sync` the compiler emits
    // straight-line UTCHMMA code (a plain `lane == 0` branch makes it wrap every tcgen05 instruction in a
    // per-active-thread uniformisation loop, which made this issue loop -- not the tensor pipe -- the limiter).
    constexpr uint32_t fmt = std::is_same<A, __half>::value ? 0u : 1u;  // F16 / BF16
    // instruction descriptor: D=f32 [4,6), A fmt [7,10), B fmt [10,13), A/B K-major (bits 15,16 = 0),
    // N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(MT >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    // UM K blocks per round: lanes poll the 2 UM barriers {x_full, a_ready} x {it .. it + UM - 1} concurrently.
    constexpr int UM = 2;                        // (with two issuers, two rounds are in flight: 4 of the 8 A slots)
    constexpr int NISS = Cfg::DUAL ? 2 : 1;      // issuers; issuer `who` takes rounds who, who + NISS, ...
    static_assert(UM * NISS <= SX && UM * NISS < SA || !TS, "MMA rounds larger than the rings");
    const int who = Cfg::DUAL ? warp - 1 : 0;
    const uint32_t d_tmem = tmem_base + uint32_t(who * MT);  // this issuer's accumulator
    const int it_first = who * UM;
    for (int it = it_first; it < nk; it += UM * NISS) {
      const int nb = min(UM, nk - it);
      TRACE(2, it);
      {
        const int l = lane & (2 * UM - 1);
        int i = it + (l >> 1);
        if (i >= nk) i = it;
        const uint32_t bar = (l & 1) ? a_ready(i % SA) : x_full(i % SX);
        const uint32_t par = uint32_t((l & 1) ? (i / SA) : (i / SX)) & 1u;
        mbar_wait(bar, par);
      }
      __syncwarp();
      TRACE(4, it);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < UM; ++u) {
          if (u < nb) {
            const int sx = (it + u) % SX, sa = (it + u) % SA;
            const uint64_t bdesc = make_sw128_desc(x_base + sx * Cfg::X_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {  // four K=16 MMAs per 64-wide K block: +32 bytes inside the swizzle atom
              const uint32_t acc = ((it - it_first) | u | k4) != 0;
              if constexpr (TS) {
                umma_ts(d_tmem, tmem_base + Cfg::A_COL0 + sa * 32 + k4 * 8, bdesc + uint64_t(k4 * 2), idesc, acc);
              } else {
                const uint64_t adesc = make_sw128_desc(a_base + sa * Cfg::A_BYTES);
                umma_ss(d_tmem, adesc + uint64_t(k4 * 2), bdesc + uint64_t(k4 * 2), idesc, acc);
              }
            }
            // once these MMAs have read them: release the A slot, and the x slot in every CTA that multicasts into it
            if (prm.cluster > 1) {
              umma_commit_mc(mma_done(sa), cmask);
            } else {
              umma_commit(mma_done(sa));
            }
          }
        }
      }
      __syncwarp();
      TRACE(5, it);
      TRACE(12, it);
    }
    // all of this issuer's MMAs retired -> its accumulator is final (arrives at once if it never issued any)
    if (elect_one()) umma_commit(d_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ dequant warps, then epilogue ==================================================
    const int q = warp & 3;            // TMEM lane quarter this warp may touch
    const int set = (warp - 4) >> 2;   // this warp's set: owns stages set, set + SETS, ...
    const int nl = q * 32 + lane;      // weight column inside the tile == TMEM lane
    const int n = n0 + nl;
    const bool n_ok = n < p.N;
    const A* sp = reinterpret_cast<const A*>(p.s);
    A raw_s = A(0.f), raw_zf = A(0.f);
    int raw_z = 0;  // integer zero-point kept RAW: converting at the load would stall on it one stage too early
    int zscalar = 0;
    A zfscalar = A(0.f);
    if (ZK != CB200_ZP_NONE && p.zp_scalar) {
      if constexpr (ZK == CB200_ZP_I32) zscalar = __ldg(reinterpret_cast<const int32_t*>(p.zp));
      if constexpr (ZK == CB200_ZP_U8) zscalar = __ldg(reinterpret_cast<const uint8_t*>(p.zp));
      if constexpr (ZK == CB200_ZP_ACT) zfscalar = *reinterpret_cast<const A*>(p.zp);
    }
    const bool zp_per_group = ZK != CB200_ZP_NONE && !p.zp_scalar;
    // scale / zero-point of this thread's column for the group of K block `kb` (one prefetch ahead, raw)
    // (group index kept incrementally: this set advances SETS K blocks per stage; no integer division in the loop)
    const int gb = p.g / Cfg::KB;  // K blocks per quantisation group
    int grp = (kb0 + set) / gb, grp_rem = (kb0 + set) % gb;
    auto fetch_meta = [&](int kb) {
      raw_z = zscalar;
      raw_zf = zfscalar;
      if (n_ok && kb < kb1) {
        const size_t idx = size_t(grp) * p.N + n;
        raw_s = sp[idx];
        if (zp_per_group) {
          if constexpr (ZK == CB200_ZP_I32) raw_z = __ldg(reinterpret_cast<const int32_t*>(p.zp) + idx);
          if constexpr (ZK == CB200_ZP_U8) raw_z = __ldg(reinterpret_cast<const uint8_t*>(p.zp) + idx);
          if constexpr (ZK == CB200_ZP_ACT) raw_zf = reinterpret_cast<const A*>(p.zp)[idx];
        }
      }
      grp_rem += SETS;
      while (grp_rem >= gb) {
        grp_rem -= gb;
        ++grp;
      }
    };
    fetch_meta(kb0 + set);

    for (int it = set; it < nk; it += SETS) {
      const int sw = it % SW, sa = it % SA;
      const ColQ cq = make_colq<A, BITS, Order::Natural, FZP>(raw_s, float(raw_z + p.bias), raw_zf);
      if (q == 0 && lane == 0) TRACE(6, it);
      fetch_meta(kb0 + it + SETS);  // for this set's next stage; consumed one full stage later
      if (q == 0 && lane == 0) TRACE(7, it);
      mbar_wait(w_full(sw), (it / SW) & 1);
      if (q == 0 && lane == 0) TRACE(8, it);
      const uint32_t* wst = reinterpret_cast<const uint32_t*>(smem_gen + (w_base - smem_base) + sw * Cfg::W_BYTES);
      uint32_t regs[32];
#pragma unroll
      for (int j = 0; j < WPT; ++j) {
        const uint32_t word = wst[j * Cfg::NTILE + nl];
        dequant_word<A, BITS, Order::Natural, FZP>(word, cq, *reinterpret_cast<uint32_t(*)[PAIRS]>(&regs[j * PAIRS]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(w_free(sw));  // the packed tile is in registers: the producer may refill the slot
      if (q == 0 && lane == 0) TRACE(9, it);
      mbar_wait(mma_done(sa), ((it / SA) & 1) ^ 1);
      if (q == 0 && lane == 0) TRACE(10, it);
      tc_fence_after();
      if constexpr (TS) {
        // lane nl of TMEM, columns [A_COL0 + 32 sa, +32): the 64 k of this column, 2 per 32-bit cell
        tmem_st32(tmem_base + (uint32_t(q * 32) << 16) + Cfg::A_COL0 + sa * 32, regs);
        tmem_st_wait();
        tc_fence_before();
      } else {
        // row nl of the K-major tile: 16-byte chunk c holds k = 8c..8c+7; 128B swizzle: chunk ^= row & 7
        const uint32_t row = a_base + sa * Cfg::A_BYTES + nl * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = row + ((uint32_t(c) ^ (nl & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(regs[c * 4 + 0]), "r"(regs[c * 4 + 1]),
                       "r"(regs[c * 4 + 2]), "r"(regs[c * 4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(sa));
      if (q == 0 && lane == 0) TRACE(11, it);
    }

    // ---- epilogue: D^T[n, m] (lane = n, column = m) -> out[m, n] ------------------------------------------------------
    mbar_wait(d_full, 0);
    tc_fence_after();
    constexpr int ESETS = MT >= 64 ? SETS : (MT / 16);  // sets taking part in the epilogue (>= 16 columns per warp)
    constexpr int EC = MT / ESETS;                      // columns (m) handled by this warp
    constexpr int CH = EC >= 32 ? 32 : EC;              // columns per tcgen05.ld
    float* part = split ? prm.ws_partial + (size_t(zsplit) * prm.t_tail + size_t(tile - prm.t_full)) * (MT * Cfg::NTILE) : nullptr;
#pragma unroll 1
    for (int cb = 0; cb < (set < ESETS ? EC : 0); cb += CH) {
      const int col0 = set * EC + cb;
      uint32_t r[32];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + col0;
      if constexpr (CH == 32) tmem_ld32(taddr, r);
      if constexpr (CH == 16) tmem_ld16(taddr, r);
      uint32_t r2[32];
      const bool second = Cfg::DUAL && nk > 2;  // the second issuer's accumulator holds rounds 1, 3, ... (none if nk <= 2)
      if (second) {
        if constexpr (CH == 32) tmem_ld32(taddr + MT, r2);
        if constexpr (CH == 16) tmem_ld16(taddr + MT, r2);
      }
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int m = m0 + col0 + c;
        if (m < p.M && n_ok) {
          float v = __uint_as_float(r[c]);
          if (second) v += __uint_as_float(r2[c]);
          if (split) {
            __stcg(part + (col0 + c) * Cfg::NTILE + nl, v);  // tile-local [m][n]: 128 contiguous bytes per warp store
          } else {
            store_out<A>(p.out, p.out_dtype, size_t(m) * p.N + n, v + bias_at(p, n));
          }
        }
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------------------------
  tc_fence_before();
  if (prm.cluster > 1) {
    cluster_sync_all();  // a peer may still be multicasting into / signalling this CTA
  } else {
    __syncthreads();
  }
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant (CB200_TC2_TS) for large M: m_tile 128, one CTA per SM walks the tiles t = blockIdx.x, + gridDim.x, ...
// The W / A / x rings run on a K-block counter that does not stop at tile boundaries, so producers, dequant warps and the
// MMA issuer roll straight into the next tile; the accumulator is double-buffered in TMEM (2 x 128 columns + the 8 x 32
// column A ring = 512) and four dedicated epilogue warps drain tile j while the MMAs of tile j + 1 run.  What this removes
// from the non-persistent kernel: TMEM allocation / barrier setup / pipeline fill per tile, the epilogue that nothing
// overlapped, and most of the partial last wave (tiles are 128 x 128 instead of 128 x 256).
// ------------------------------------------------------------------------------------------------------------------
template <int BITS>
struct Tc2Cfg {
  static constexpr int MT = 128, NTILE = 128, KB = 64;
  static constexpr int PF = 32 / BITS;
  static constexpr int WROWS = KB / PF;
  static constexpr int X_BYTES = MT * 128;
  static constexpr int W_BYTES = WROWS * NTILE * 4;
  static constexpr int SW = BITS == 4 ? 16 : 8;  // 64 KB
  static constexpr int SA = 8;                   // 256 TMEM columns
  static constexpr int SX = 8;                   // 128 KB
  static constexpr int A_COL0 = 2 * MT;          // two accumulators first
  static constexpr int SMEM_BYTES = SX * X_BYTES + SW * W_BYTES + 1024 /*alignment slack*/ + 1024 /*barriers*/;
};
constexpr int kTc2Threads = 128 + 128 * kTcSets + 128;  // control warps, dequant warps, epilogue warps: 768

struct Tc2Params {
  Problem p;
  int nk;          // K blocks per tile (K / 64)
  int n_tiles;     // tiles along N
  int total_tiles;
  int gb;          // K blocks per quantisation group
  int gb_shift;    // log2(gb) if gb is a power of two, else -1
};

template <class A, int BITS, int ZK>
__global__ void __launch_bounds__(kTc2Threads, 1)
    tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, Tc2Params prm) {
  using Cfg = Tc2Cfg<BITS>;
  constexpr int MT = Cfg::MT;
  constexpr int SX = Cfg::SX, SW = Cfg::SW, SA = Cfg::SA;
  constexpr bool FZP = ZK == CB200_ZP_ACT;
  constexpr int PAIRS = 16 / BITS;
  constexpr int WPT = 64 / Cfg::PF;
  constexpr int SETS = kTcSets;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = smem_base;
  const uint32_t w_base = x_base + SX * Cfg::X_BYTES;
  const uint32_t bar_base = w_base + SW * Cfg::W_BYTES;
  auto x_full = [&](int s) { return bar_base + 8u * s; };
  auto w_full = [&](int s) { return bar_base + 8u * (SX + s); };
  auto w_free = [&](int s) { return bar_base + 8u * (SX + SW + s); };
  auto a_ready = [&](int s) { return bar_base + 8u * (SX + 2 * SW + s); };
  auto mma_done = [&](int s) { return bar_base + 8u * (SX + 2 * SW + SA + s); };
  auto d_full = [&](int b) { return bar_base + 8u * (SX + 2 * SW + 2 * SA + b); };
  auto d_empty = [&](int b) { return bar_base + 8u * (SX + 2 * SW + 2 * SA + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (SX + 2 * SW + 2 * SA + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const Problem& p = prm.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = prm.nk;
  const int my_tiles = (prm.total_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int G = my_tiles * nk;  // K blocks this CTA walks through, tile after tile

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_w);
    for (int s = 0; s < SX; ++s) mbar_init(x_full(s), 1);
    for (int s = 0; s < SW; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_free(s), 4);
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_ready(s), 4);
      mbar_init(mma_done(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d_full(b), 1);
      mbar_init(d_empty(b), 4);  // the four epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // (tile, K block) cursor: K blocks are visited g = 0, 1, ... G-1; tile j = g / nk is t = blockIdx.x + j * gridDim.x
  struct Cursor {
    int j, kb, t;
  };
  auto advance = [&](Cursor& c, int by) {
    c.kb += by;
    while (c.kb >= nk) {
      c.kb -= nk;
      ++c.j;
      c.t += int(gridDim.x);
    }
  };

  if (warp == 0) {
    // ================================ TMA producer: packed weights ==================================================
    constexpr int UW = 4;
    Cursor c{0, 0, int(blockIdx.x)};
    for (int g = 0; g < G; g += UW) {
      const int nb = min(UW, G - g);
      {
        int i = g + (lane & (UW - 1));
        if (i >= G) i = g;
        mbar_wait(w_free(i % SW), ((i / SW) & 1) ^ 1);
      }
      __syncwarp();
      if (elect_one()) {
        Cursor cc = c;
        for (int u = 0; u < nb; ++u) {
          const int s = (g + u) % SW;
          mbar_expect_tx(w_full(s), Cfg::W_BYTES);
          tma_load_2d(w_base + s * Cfg::W_BYTES, &tmap_w, (cc.t % prm.n_tiles) * Cfg::NTILE, cc.kb * Cfg::WROWS, w_full(s));
          advance(cc, 1);
        }
      }
      __syncwarp();
      advance(c, nb);
    }
  } else if (warp == 3) {
    // ================================ TMA producer: activations =====================================================
    constexpr int UX = 2;
    Cursor c{0, 0, int(blockIdx.x)};
    for (int g = 0; g < G; g += UX) {
      const int nb = min(UX, G - g);
      {
        int i = g + (lane & (UX - 1));
        if (i >= G) i = g;
        if (i >= SX) {
          const int q = i - SX;
          mbar_wait(mma_done(q % SA), (q / SA) & 1);
        }
      }
      __syncwarp();
      if (elect_one()) {
        Cursor cc = c;
        for (int u = 0; u < nb; ++u) {
          const int s = (g + u) % SX;
          mbar_expect_tx(x_full(s), Cfg::X_BYTES);
          tma_load_2d(x_base + s * Cfg::X_BYTES, &tmap_x, cc.kb * Cfg::KB, (cc.t / prm.n_tiles) * MT, x_full(s));
          advance(cc, 1);
        }
      }
      __syncwarp();
      advance(c, nb);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================================================
    constexpr uint32_t fmt = std::is_same<A, __half>::value ? 0u : 1u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(MT >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    constexpr int UM = 4;
    Cursor c{0, 0, int(blockIdx.x)};
    for (int g = 0; g < G; g += UM) {
      const int nb = min(UM, G - g);
      {
        const int l = lane & (2 * UM - 1);
        int i = g + (l >> 1);
        if (i >= G) i = g;
        const uint32_t bar = (l & 1) ? a_ready(i % SA) : x_full(i % SX);
        const uint32_t par = uint32_t((l & 1) ? (i / SA) : (i / SX)) & 1u;
        mbar_wait(bar, par);
      }
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        Cursor cc = c;
        for (int u = 0; u < nb; ++u) {
          const int sx = (g + u) % SX, sa = (g + u) % SA;
          const int buf = cc.j & 1;
          if (cc.kb == 0) {  // first K block of a tile: its accumulator must have been drained (tile j - 2)
            mbar_wait(d_empty(buf), ((cc.j >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + uint32_t(buf * MT);
          const uint64_t bdesc = make_sw128_desc(x_base + sx * Cfg::X_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (cc.kb | k4) != 0;
            umma_ts(d_tmem, tmem_base + Cfg::A_COL0 + sa * 32 + k4 * 8, bdesc + uint64_t(k4 * 2), idesc, acc);
          }
          umma_commit(mma_done(sa));
          if (cc.kb == nk - 1) umma_commit(d_full(buf));  // the tile's accumulator is complete once these retire
          advance(cc, 1);
        }
      }
      __syncwarp();
      advance(c, nb);
    }
  } else if (warp >= 4 && warp < 4 + 4 * SETS) {
    // ================================ dequant warps ==================================================================
    const int q = warp & 3;
    const int set = (warp - 4) >> 2;
    const int nl = q * 32 + lane;
    const A* sp = reinterpret_cast<const A*>(p.s);
    A raw_s = A(0.f), raw_zf = A(0.f);
    int raw_z = 0;
    int zscalar = 0;
    A zfscalar = A(0.f);
    if (ZK != CB200_ZP_NONE && p.zp_scalar) {
      if constexpr (ZK == CB200_ZP_I32) zscalar = __ldg(reinterpret_cast<const int32_t*>(p.zp));
      if constexpr (ZK == CB200_ZP_U8) zscalar = __ldg(reinterpret_cast<const uint8_t*>(p.zp));
      if constexpr (ZK == CB200_ZP_ACT) zfscalar = *reinterpret_cast<const A*>(p.zp);
    }
    const bool zp_per_group = ZK != CB200_ZP_NONE && !p.zp_scalar;
    int n_col = 0;  // this thread's weight column in the cursor's tile (recomputed only when the tile changes)
    auto fetch_meta = [&](const Cursor& c) {  // scale / zero-point of this thread's column for K block c.kb of tile c.t
      raw_z = zscalar;
      raw_zf = zfscalar;
      const int n = n_col;
      if (c.j < my_tiles && n < p.N) {
        const int grp = prm.gb_shift >= 0 ? (c.kb >> prm.gb_shift) : (c.kb / prm.gb);
        const size_t idx = size_t(grp) * p.N + n;
        raw_s = sp[idx];
        if (zp_per_group) {
          if constexpr (ZK == CB200_ZP_I32) raw_z = __ldg(reinterpret_cast<const int32_t*>(p.zp) + idx);
          if constexpr (ZK == CB200_ZP_U8) raw_z = __ldg(reinterpret_cast<const uint8_t*>(p.zp) + idx);
          if constexpr (ZK == CB200_ZP_ACT) raw_zf = reinterpret_cast<const A*>(p.zp)[idx];
        }
      }
    };
    Cursor c{0, 0, int(blockIdx.x)};
    advance(c, set);
    n_col = (c.t % prm.n_tiles) * Cfg::NTILE + nl;
    fetch_meta(c);
    for (int g = set; g < G; g += SETS) {
      const int sw = g % SW, sa = g % SA;
      const ColQ cq = make_colq<A, BITS, Order::Natural, FZP>(raw_s, float(raw_z + p.bias), raw_zf);
      const int j_before = c.j;
      advance(c, SETS);
      if (c.j != j_before) n_col = (c.t % prm.n_tiles) * Cfg::NTILE + nl;
      fetch_meta(c);  // for this set's next K block; consumed one full iteration later
      mbar_wait(w_full(sw), (g / SW) & 1);
      const uint32_t* wst = reinterpret_cast<const uint32_t*>(smem_gen + (w_base - smem_base) + sw * Cfg::W_BYTES);
      uint32_t regs[32];
#pragma unroll
      for (int j = 0; j < WPT; ++j) {
        const uint32_t word = wst[j * Cfg::NTILE + nl];
        dequant_word<A, BITS, Order::Natural, FZP>(word, cq, *reinterpret_cast<uint32_t(*)[PAIRS]>(&regs[j * PAIRS]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(w_free(sw));
      mbar_wait(mma_done(sa), ((g / SA) & 1) ^ 1);
      tc_fence_after();
      tmem_st32(tmem_base + (uint32_t(q * 32) << 16) + Cfg::A_COL0 + sa * 32, regs);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(sa));
    }
  } else if (warp >= 4 + 4 * SETS) {
    // ================================ epilogue warps: drain tile j while tile j + 1 is being multiplied ==============
    const int q = warp & 3;  // TMEM lane quarter (warp 20 -> 0)
    const int nl = q * 32 + lane;
    int t = int(blockIdx.x);
    for (int j = 0; j < my_tiles; ++j, t += int(gridDim.x)) {
      const int buf = j & 1;
      const int n = (t % prm.n_tiles) * Cfg::NTILE + nl;
      const int m0 = (t / prm.n_tiles) * MT;
      const bool n_ok = n < p.N;
      mbar_wait(d_full(buf), (j >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int col0 = 0; col0 < MT; col0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * MT + col0), r);
        tmem_ld_wait();
#pragma unroll
        for (int cidx = 0; cidx < 32; ++cidx) {
          const int m = m0 + col0 + cidx;
          if (m < p.M && n_ok) store_out<A>(p.out, p.out_dtype, size_t(m) * p.N + n, __uint_as_float(r[cidx]) + bias_at(p, n));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// out tile = cast(sum_z partial[z][tile][m][n]) for the split tiles -- all SMs in parallel (measured faster than a last-CTA
// fix-up: the partial planes are far more than one CTA can sum quickly)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, Problem p, int mt, int n_tiles, int t_full, int t_tail,
                                     int splits) {
  const int per_tile4 = mt * 32;  // float4 per tile (128 columns)
  const int64_t total4 = int64_t(t_tail) * per_tile4;
  const float4* p4 = reinterpret_cast<const float4*>(part);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 v = __ldcg(p4 + i);
    for (int z = 1; z < splits; ++z) {
      const float4 t = __ldcg(p4 + size_t(z) * total4 + i);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    const int tl = int(i / per_tile4), rem = int(i - int64_t(tl) * per_tile4);
    const int tile = t_full + tl;
    const int m = (tile / n_tiles) * mt + (rem >> 5);
    const int n = (tile % n_tiles) * 128 + (rem & 31) * 4;
    if (m < p.M && n < p.N) {  // N % 64 == 0: a group of four columns is inside or outside as a whole
      const size_t o = size_t(m) * p.N + n;
      store_out<float>(p.out, p.out_dtype, o + 0, v.x + bias_at(p, n));
      store_out<float>(p.out, p.out_dtype, o + 1, v.y + bias_at(p, n + 1));
      store_out<float>(p.out, p.out_dtype, o + 2, v.z + bias_at(p, n + 2));
      store_out<float>(p.out, p.out_dtype, o + 3, v.w + bias_at(p, n + 3));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
struct TcPlan {
  int m_tile, n_tiles, tiles;
  int t_full, t_tail, splits, kb_per_split;
};

static int tc_sm_count() {
  const DeviceInfo* di = current_device_info();
  return di && di->sms > 0 ? di->sms : 148;
}

static TcPlan tc_plan(int M, int K, int N) {
  TcPlan pl;
  int mt = g_tc_tuning.m_tile;
  if (mt != 16 && mt != 32 && mt != 64 && mt != 128 && mt != 256) {
    mt = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256;
  }
  pl.m_tile = mt;
  const int kblocks = K / 64;
  const int sms = tc_sm_count();
  pl.n_tiles = (N + 127) / 128;
  pl.tiles = pl.n_tiles * ((M + mt - 1) / mt);
  pl.t_full = pl.tiles;  // default: no split
  pl.t_tail = 0;
  int ks = 1;
  if (g_tc_tuning.split_k > 0) {  // forced: every tile split
    ks = g_tc_tuning.split_k;
    pl.t_full = 0;
  } else if (pl.tiles <= sms / 2) {  // too few tiles to fill the GPU: split all of them, exactly one wave
    ks = sms / pl.tiles;
    while (ks > 1 && kblocks / ks < 4) --ks;  // keep at least four K blocks per CTA
    pl.t_full = 0;
  } else if (pl.tiles > sms && g_tc_tuning.split_k == 0) {
    // several waves: split only the partial LAST wave so that it keeps every SM busy (a quarter-full last wave of
    // whole tiles costs a whole tile time)
    const int tail = pl.tiles % sms;
    if (tail != 0 && tail * 2 <= sms) {
      ks = sms / tail;
      while (ks > 1 && kblocks / ks < 8) --ks;
      if (ks > 8) ks = 8;
      if (ks > 1) pl.t_full = pl.tiles - tail;
    }
  }
  if (ks > kblocks) ks = kblocks;
  if (ks > 32) ks = 32;
  if (ks <= 1) pl.t_full = pl.tiles;
  pl.t_tail = pl.tiles - pl.t_full;
  pl.kb_per_split = (kblocks + ks - 1) / ks;
  pl.splits = pl.t_tail > 0 ? (kblocks + pl.kb_per_split - 1) / pl.kb_per_split : 1;
  if (pl.splits <= 1) {
    pl.splits = 1;
    pl.t_full = pl.tiles;
    pl.t_tail = 0;
  }
  return pl;
}

bool tc_supported(const Problem& p, int bits) {
  (void)bits;
  return encode_fn() != nullptr && p.K % 64 == 0 && p.N % 64 == 0 && p.g % 64 == 0;
}

size_t tc_workspace_bytes(int64_t M, int64_t K, int64_t N) {
  const TcPlan pl = tc_plan(int(M), int(K), int(N));
  return pl.splits > 1 ? size_t(CB200_WORKSPACE_COUNTER_BYTES) + size_t(pl.splits) * pl.t_tail * pl.m_tile * 128 * sizeof(float) : 0;
}

template <class A, int BITS, int MT, bool TS, int ZK>
static int launch_cfg(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& prm, dim3 grid, cudaStream_t stream) {
  using Cfg = TcCfg<MT, BITS, TS>;
  auto kern = tc_kernel<A, BITS, MT, TS, ZK>;
  static unsigned long long attr_mask = 0;  // per instantiation, one bit per device
  if (int rc = set_max_smem_once(kern, Cfg::SMEM_BYTES, attr_mask)) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = prm.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = prm.cluster > 1 ? 1 : 0;
  CB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mx, mw, prm));
  return 0;
}

template <class A, int BITS, int MT, bool TS>
static int launch_zk(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& prm, dim3 grid, cudaStream_t stream) {
  switch (prm.p.zp_kind) {
    case CB200_ZP_NONE: return launch_cfg<A, BITS, MT, TS, CB200_ZP_NONE>(mx, mw, prm, grid, stream);
    case CB200_ZP_I32: return launch_cfg<A, BITS, MT, TS, CB200_ZP_I32>(mx, mw, prm, grid, stream);
    case CB200_ZP_U8: return launch_cfg<A, BITS, MT, TS, CB200_ZP_U8>(mx, mw, prm, grid, stream);
    default: return launch_cfg<A, BITS, MT, TS, CB200_ZP_ACT>(mx, mw, prm, grid, stream);
  }
}

template <class A, int BITS, bool TS>
static int launch_mt(int mt, const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& prm, dim3 grid, cudaStream_t stream) {
  switch (mt) {
    case 16: return launch_zk<A, BITS, 16, TS>(mx, mw, prm, grid, stream);
    case 32: return launch_zk<A, BITS, 32, TS>(mx, mw, prm, grid, stream);
    case 64: return launch_zk<A, BITS, 64, TS>(mx, mw, prm, grid, stream);
    case 128: return launch_zk<A, BITS, 128, TS>(mx, mw, prm, grid, stream);
    default: return launch_zk<A, BITS, 256, TS>(mx, mw, prm, grid, stream);
  }
}

// x: [M, K] row-major, box {64, x_box_rows}, 128-byte swizzle; packed W: [K/pf, N] int32, box {128, 64/pf}
static int make_tmaps(EncodeTiledFn enc, const Problem& p, int bits, int act_dtype, int x_box_rows, CUtensorMap& mx,
                      CUtensorMap& mw) {
  const int pf = 32 / bits;
  // x: [M, K] row-major -> dims {K, M}, box {64, m_tile}, 128-byte swizzle; rows beyond M are zero-filled by TMA
  {
    const cuuint64_t dims[2] = {cuuint64_t(p.K), cuuint64_t(p.M)};
    const cuuint64_t strides[1] = {cuuint64_t(p.K) * 2};
    const cuuint32_t box[2] = {64, cuuint32_t(x_box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&mx, act_dtype == CB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(x) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  // packed W: [K/pf, N] int32 -> dims {N, K/pf}, box {128, 64/pf}, no swizzle; columns beyond N zero-filled
  {
    const cuuint64_t dims[2] = {cuuint64_t(p.N), cuuint64_t(p.K / pf)};
    const cuuint64_t strides[1] = {cuuint64_t(p.N) * 4};
    const cuuint32_t box[2] = {128, cuuint32_t(64 / pf)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<uint32_t*>(p.wq), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(w_q_packed) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }

  return 0;
}

int launch_tc(const Problem& p, int bits, int act_dtype, int variant, void* ws, size_t ws_bytes, cudaStream_t stream,
              int* launches) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return CB200_ERR_DEVICE;
  }
  const TcPlan pl = tc_plan(p.M, p.K, p.N);
  const int mt = pl.m_tile;
  // Optional (cb200_set_tc_cluster): CTAs with neighbouring n tiles and the same m tile share the x tile -- each fetches
  // 1/cluster of it and multicasts.  Measured neutral (2) to slightly slower (4) on every prefill shape: the x traffic
  // from L2 is not what limits this kernel, so the default is no cluster.
  const int n_tiles = pl.n_tiles;
  int cluster = g_tc_cluster > 0 ? g_tc_cluster : 1;
  // (clusters are runs of `cluster` consecutive work items: they must be n-neighbours of one m tile in both item ranges)
  while (cluster > 1 && (n_tiles % cluster != 0 || mt / cluster < 8 || pl.t_full % cluster != 0 || pl.t_tail % cluster != 0)) cluster /= 2;

  CUtensorMap mx, mw;
  if (int rc = make_tmaps(enc, p, bits, act_dtype, mt / cluster, mx, mw)) return rc;

  TcParams prm;
  prm.p = p;
  prm.kblocks_total = p.K / 64;
  prm.n_tiles = pl.n_tiles;
  prm.t_full = pl.t_full;
  prm.t_tail = pl.t_tail;
  prm.splits = pl.splits;
  prm.kb_per_split = pl.kb_per_split;
  prm.ws_partial = nullptr;
  prm.cluster = cluster;
#ifdef CB200_TC_TRACE
  prm.trace = g_tc_trace;
#endif
  if (pl.splits > 1) {
    const size_t need = size_t(CB200_WORKSPACE_COUNTER_BYTES) + size_t(pl.splits) * pl.t_tail * mt * 128 * sizeof(float);
    if (!ws || ws_bytes < need) {
      set_error("tcgen05 split-K=%d over %d tiles needs %zu workspace bytes, got %zu", pl.splits, pl.t_tail, need, ws_bytes);
      return CB200_ERR_WORKSPACE;
    }
    prm.ws_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES);
  }
  dim3 grid(pl.t_full + pl.t_tail * pl.splits, 1, 1);

  const bool ts = variant != CB200_TC_SS;
  int rc;
  if (act_dtype == CB200_F16) {
    rc = bits == 4 ? (ts ? launch_mt<__half, 4, true>(mt, mx, mw, prm, grid, stream) : launch_mt<__half, 4, false>(mt, mx, mw, prm, grid, stream))
                   : (ts ? launch_mt<__half, 8, true>(mt, mx, mw, prm, grid, stream) : launch_mt<__half, 8, false>(mt, mx, mw, prm, grid, stream));
  } else {
    rc = bits == 4 ? (ts ? launch_mt<__nv_bfloat16, 4, true>(mt, mx, mw, prm, grid, stream)
                         : launch_mt<__nv_bfloat16, 4, false>(mt, mx, mw, prm, grid, stream))
                   : (ts ? launch_mt<__nv_bfloat16, 8, true>(mt, mx, mw, prm, grid, stream)
                         : launch_mt<__nv_bfloat16, 8, false>(mt, mx, mw, prm, grid, stream));
  }
  if (rc) return rc;
  if (launches) ++*launches;
  if (pl.splits > 1) {
    const int64_t total4 = int64_t(pl.t_tail) * mt * 32;
    int blocks = int((total4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(prm.ws_partial, p, mt, pl.n_tiles, pl.t_full, pl.t_tail, pl.splits);
    CB200_CUDA_CHECK(cudaGetLastError());
    if (launches) ++*launches;
  }
  return 0;
}

// ---- persistent variant ---------------------------------------------------------------------------------------------
template <class A, int BITS>
static int launch_tc2_zk(const CUtensorMap& mx, const CUtensorMap& mw, const Tc2Params& prm, int ctas, cudaStream_t stream) {
  void (*kern)(const CUtensorMap, const CUtensorMap, Tc2Params);
  switch (prm.p.zp_kind) {
    case CB200_ZP_NONE: kern = tc2_kernel<A, BITS, CB200_ZP_NONE>; break;
    case CB200_ZP_I32: kern = tc2_kernel<A, BITS, CB200_ZP_I32>; break;
    case CB200_ZP_U8: kern = tc2_kernel<A, BITS, CB200_ZP_U8>; break;
    default: kern = tc2_kernel<A, BITS, CB200_ZP_ACT>; break;
  }
  CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg<BITS>::SMEM_BYTES));
  kern<<<ctas, kTc2Threads, Tc2Cfg<BITS>::SMEM_BYTES, stream>>>(mx, mw, prm);
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_tc2(const Problem& p, int bits, int act_dtype, cudaStream_t stream, int* launches) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return CB200_ERR_DEVICE;
  }
  const int sm_count = tc_sm_count();
  CUtensorMap mx, mw;
  if (int rc = make_tmaps(enc, p, bits, act_dtype, 128, mx, mw)) return rc;
  Tc2Params prm;
  prm.p = p;
  prm.nk = p.K / 64;
  prm.n_tiles = (p.N + 127) / 128;
  prm.total_tiles = prm.n_tiles * ((p.M + 127) / 128);
  prm.gb = p.g / 64;
  prm.gb_shift = -1;
  for (int sh = 0; sh < 31; ++sh) {
    if ((1 << sh) == prm.gb) prm.gb_shift = sh;
  }
  const int ctas = prm.total_tiles < sm_count ? prm.total_tiles : sm_count;  // one persistent CTA per SM
  int rc;
  if (act_dtype == CB200_F16) {
    rc = bits == 4 ? launch_tc2_zk<__half, 4>(mx, mw, prm, ctas, stream) : launch_tc2_zk<__half, 8>(mx, mw, prm, ctas, stream);
  } else {
    rc = bits == 4 ? launch_tc2_zk<__nv_bfloat16, 4>(mx, mw, prm, ctas, stream)
                   : launch_tc2_zk<__nv_bfloat16, 8>(mx, mw, prm, ctas, stream);
  }
  if (rc) return rc;
  if (launches) ++*launches;
  return 0;
}

}  // namespace cb200
