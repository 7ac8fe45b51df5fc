// Streaming decode kernel (CB200_DECODE_STREAM): the M <= 16 regime as a PERSISTENT, TMA-fed, warp-specialised kernel that is
// built to overlap with its neighbours in the stream.  Design notes are at the top of stream_kernel.cu (host plan + dispatch);
// this header holds the kernel template and the launch ladder shared by stream_f16_w4.cu / stream_f16_w8.cu / stream_bf16_w4.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dequant.cuh"
#include "mma_ptx.cuh"
#include "tc_ptx.cuh"

namespace cb200 {

constexpr int kStreamGroups = 2;                          // consumer groups (4 warps each) per CTA
constexpr int kStreamThreads = 128 * kStreamGroups;       // 8 warps: with two CTAs per SM every scheduler holds 4 warps -> 128 registers
constexpr int kStreamStageW = 8192;                       // packed weight bytes per stage: 16 packed rows x 128 columns

struct StreamGeom {
  int m_rows;          // rows of x in this pass (<= 16)
  int spc;             // stages per 128-column block = K / KS (KS = 16 packed rows = 128 k for int4, 64 k for int8)
  int ncb;             // 128-column blocks = ceil(N / 128)
  int q, r;            // run partition: CTA c owns stages [c*q + min(c,r), (c+1)*q + min(c+1,r)) of the k-fastest stage list
  int ring;            // ring depth in stages (>= kStreamGroups)
  int stage_bytes;     // multiple of 1024
  int x_off;           // byte offset of the x tile inside a stage
  int x_bytes;         // bytes one x box delivers (m_rows * KS * 2)
  int meta_off;        // byte offset of the group rows inside a stage
  int meta_row_bytes;  // 256 B of scales + the zero-point row
  int gps;             // group rows per stage: 1, or 2 when a group boundary can fall in the middle of a stage (int4, g % 128 != 0)
  int hpg;             // half-stages (2 steps = 8 packed rows) per scale group = g / (8 * PF)
  int hpg_shift;       // log2(hpg) when it is a power of two, else -1
  int maxslots;        // split-K partial slots per (column block, 32-column quarter)
  int w3d;             // the packed weights are described by ONE 3-D tensor map (box 32 x 16 x 4): one TMA per stage instead of four
  int pdl;             // launched with programmatic stream serialization
  int dbg;             // debug switches (cb200_debug_set_stream_switches): 1 consumers skip the math, 2 no W loads, 4 no x loads, 8 no scale / zp loads
  long long* trace;    // [CTAs][16] globaltimer stamps when non-null (benchmarks/stream_trace.py)
};

// stage index -> owning CTA under the (q, r) partition
__device__ __forceinline__ int stream_cta_of(int s, int q, int r) {
  const int big = r * (q + 1);
  return s < big ? s / (q + 1) : r + (s - big) / q;
}

__device__ __forceinline__ void stream_stamp(long long* trace, int slot) {
  if (trace) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 32 + slot] = t;
  }
}

template <bool FRESH, class A>
__device__ __forceinline__ void mma_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (FRESH) {
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816_nv<A>(d, zero4, a0, a1, a2, a3, b0, b1);
  } else {
    mma16816_nv<A>(d, d, a0, a1, a2, a3, b0, b1);
  }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// one probe first (the common case: the data is already there), the bounded spin only when it fails
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!done) mbar_wait(bar, parity);
}

// A: activation type, BITS: 4|8, MT: 8-row m tiles (1: M <= 8, 2: M <= 16), ZK: cb200_zp_kind.
// (bf16 x int8 has no magic-number operand -- 128 + q does not fit a bf16 significand -- and stays on decode_kernel.)
// 8 warps = 2 groups x 4 quarter-warps; there is NO producer warp (a 9th warp would cap the kernel at 96 registers: the
// register file is partitioned per scheduler, and two 9-warp CTAs put 5 warps on one of them).  Warp 0 requests the first
// `ring` stages; after that, the LAST of the four warps to finish a stage refills its slot with the stage `ring` ahead.
template <class A, int BITS, int MT, int ZK>
__global__ void __launch_bounds__(kStreamThreads, MT == 1 ? 2 : 1)
    stream_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Problem p,
                  const StreamGeom geo, float* __restrict__ ws_partial, unsigned int* __restrict__ ws_counter) {
  constexpr int G = kStreamGroups;
  constexpr int PF = 32 / BITS;        // k per packed word
  constexpr int KS = 16 * PF;          // k per stage
  constexpr int PAIRS = 16 / BITS;     // operand registers per packed word
  constexpr bool HALF = std::is_same<A, __half>::value;
  constexpr bool EO = HALF && BITS == 4;  // fp16 int4: nibbles at bits 4-7 arrive as 16 q -> their own accumulator, scaled by 1/16 in fp32
  constexpr bool FZP = ZK == CB200_ZP_ACT;
  constexpr uint32_t ZES = ZK == CB200_ZP_I32 ? 4u : ZK == CB200_ZP_U8 ? 1u : ZK == CB200_ZP_ACT ? 2u : 0u;
  constexpr int MAXPEND = 4;  // column blocks whose split-K arrival is deferred to the end of the run
  static_assert(HALF || BITS == 4, "bf16 x int8 is not a stream_kernel configuration");

  // the next kernel in the stream may start now: it prefetches its own (static) weights while this one runs
  if (geo.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) stream_stamp(geo.trace, 0);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;  // 128-byte swizzle atoms need 1024-byte alignment
  const uint32_t smem0 = smem_u32(smem_raw) + smem_pad;
  // control page (1 KB): full[ring] mbarriers @0, per-slot departure counters @512, 16 zero bytes @768
  const uint32_t full0 = smem0;
  unsigned int* const depart = reinterpret_cast<unsigned int*>(smem_raw + smem_pad + 512);
  const uint32_t ring0 = smem0 + 1024u;
  const unsigned char* const ring_ptr = smem_raw + smem_pad + 1024;  // C++ view of the ring (plain ld.shared the compiler may schedule)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = geo.m_rows, N = p.N;
  const int spc = geo.spc;
  const int cta = blockIdx.x;
  const int run_a = cta * geo.q + min(cta, geo.r);
  const int n_run = geo.q + (cta < geo.r ? 1 : 0);
  const bool zp_per_group = ZK != CB200_ZP_NONE && !p.zp_scalar;
  const int first_cb = run_a / spc;  // (the only divisions outside the publish path)
  const int first_kt = run_a - first_cb * spc;
  const int ring = geo.ring;
  const int stage_bytes = geo.stage_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) mbar_init(full0 + 8u * i, 1);
    fence_barrier_init();
    prefetch_tensormap(&map_w);
    prefetch_tensormap(&map_x);
  }
  if (threadIdx.x < 32) depart[threadIdx.x] = 0u;
  // 16 zero bytes: what lanes whose x row does not exist (m >= M) read
  if (threadIdx.x >= 32 && threadIdx.x < 36) reinterpret_cast<uint32_t*>(smem_raw + smem_pad + 768)[threadIdx.x - 32] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) stream_stamp(geo.trace, 1);

  // ---- requesting one stage (executed by ONE lane): weights + scale / zero-point rows (static), and the x slice ---------------
  // (cb, kt): the stage's 128-column block and 16-packed-row slab; slot: its ring slot
  auto issue_static = [&](int slot, int cb, int kt) {
    const uint32_t sbase = ring0 + uint32_t(slot) * uint32_t(stage_bytes);
    const uint32_t bar = full0 + 8u * slot;
    const int cols = min(128, N - cb * 128);  // 64 or 128
    const uint32_t meta_bytes = uint32_t(cols) * (2u + (zp_per_group ? ZES : 0u));
    const uint32_t w_bytes = geo.w3d ? uint32_t(kStreamStageW) : uint32_t(cols / 32) * 2048u;
    mbar_expect_tx(bar, ((geo.dbg & 2) ? 0u : w_bytes) + ((geo.dbg & 4) ? 0u : uint32_t(geo.x_bytes)) +
                            ((geo.dbg & 8) ? 0u : uint32_t(geo.gps) * meta_bytes));
    if (!(geo.dbg & 2)) {
      if (geo.w3d) {
        tma_load_3d(sbase, &map_w, 0, kt * 16, cb * 4, bar);  // columns beyond N are zero-filled
      } else {
        for (int j = 0; j < cols / 32; ++j) tma_load_2d(sbase + 2048u * j, &map_w, cb * 128 + 32 * j, kt * 16, bar);
      }
    }
    if (!(geo.dbg & 8)) {
      const int hs = 2 * kt;
      const int grp = geo.hpg_shift >= 0 ? hs >> geo.hpg_shift : hs / geo.hpg;
      for (int h = 0; h < geo.gps; ++h) {
        const int gi = h == 0 ? grp : (geo.hpg_shift >= 0 ? (hs + 1) >> geo.hpg_shift : (hs + 1) / geo.hpg);
        const size_t idx = size_t(gi) * N + size_t(cb) * 128;
        const uint32_t dst = sbase + uint32_t(geo.meta_off) + uint32_t(h) * uint32_t(geo.meta_row_bytes);
        bulk_g2s(dst, reinterpret_cast<const A*>(p.s) + idx, uint32_t(cols) * 2u, bar);
        if (zp_per_group) bulk_g2s(dst + 256u, reinterpret_cast<const char*>(p.zp) + idx * ZES, uint32_t(cols) * ZES, bar);
      }
    }
  };
  auto issue_x = [&](int slot, int kt) {
    if (!(geo.dbg & 4)) {
      tma_load_2d(ring0 + uint32_t(slot) * uint32_t(stage_bytes) + uint32_t(geo.x_off), &map_x, kt * KS, 0, full0 + 8u * slot);
    }
  };

  // ---- prologue: the first `ring` stages.  Weights, scales and zero-points are static, so they are requested BEFORE
  // griddepcontrol.wait: with programmatic dependent launch that happens while the previous kernel in the stream is still
  // running, and the HBM never idles between kernels.  x belongs to the stream order: after the wait. ---------------------------
  const int pre = min(ring, n_run);
  if (warp == 0 && elect_one()) {
    int cb = first_cb, kt = first_kt;
    for (int i = 0; i < pre; ++i) {
      issue_static(i, cb, kt);
      if (++kt == spc) {
        kt = 0;
        ++cb;
      }
    }
    stream_stamp(geo.trace, 2);
  }
  if (geo.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");  // x / out / workspace belong to the stream order
  if (warp == 0 && elect_one()) {
    stream_stamp(geo.trace, 3);
    int kt = first_kt;
    for (int i = 0; i < pre; ++i) {
      issue_x(i, kt);
      if (++kt == spc) kt = 0;
    }
  }
  __syncwarp();

  // ===================================================== consumers =====================================================
  const int cg = warp >> 2;  // group: takes stages cg, cg + G, ... of the run
  const int cq = warp & 3;   // 32-column quarter of the stage
  const int g8 = lane >> 2, tig = lane & 3;
  const bool tracer = geo.trace != nullptr && warp == 0 && lane == 0;
  // column chunk of this lane: chunk c = 4 consecutive columns.  f() spreads the two g8 of a shared-memory phase over both
  // halves of the 128-byte row so that, with the TMA 128-byte swizzle, the 8 lanes of a phase hit 8 different 16-byte bank groups.
  const int chunk = (g8 >> 1) + 4 * (g8 & 1);
  const int wlane0 = cq * 2048 + tig * 128 + ((chunk ^ tig) << 4);        // even steps: row & 7 = tig
  const int wlane1 = cq * 2048 + tig * 128 + ((chunk ^ (4 + tig)) << 4);  // odd steps: row & 7 = 4 + tig
  const int lcol = 32 * cq + 4 * chunk;  // first of this lane's four columns inside the stage
  // x operand addressing.  Lane (g8, tig) feeds row m = 8 mt + g8 of x into the MMA's 8-wide operand.  Rows that do not exist
  // (m >= M) read a 16-byte zero chunk in the control page instead -- same instruction stream, no predicates: per lane a base
  // offset relative to the ring and a stage / step stride that are 0 for such lanes.
  int xbase[MT], xstage[MT], xstep[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int m = mt * 8 + g8;
    const bool ok = m < M;
    xbase[mt] = ok ? geo.x_off + m * (KS * 2) + tig * (PF * 2) : 768 - 1024;  // (the zero chunk sits 256 B before the ring)
    xstage[mt] = ok ? stage_bytes : 0;
    xstep[mt] = ok ? 4 * PF * 2 : 0;
  }

  float zscalar = 0.f;
  if (ZK != CB200_ZP_NONE && p.zp_scalar) {  // one zero-point for the whole matrix (static data)
    if constexpr (ZK == CB200_ZP_I32) zscalar = float(__ldg(reinterpret_cast<const int32_t*>(p.zp)));
    if constexpr (ZK == CB200_ZP_U8) zscalar = float(__ldg(reinterpret_cast<const uint8_t*>(p.zp)));
    if constexpr (ZK == CB200_ZP_ACT) zscalar = ActOps<A>::to_float(*reinterpret_cast<const A*>(p.zp));
  }
  const float fbias = float(p.bias);
  constexpr float MAGIC = HALF ? 0.f : 128.f;         // bf16 operands are 128 + q; fp16 operands are q * 2^-24 (subnormals)
  constexpr float OPSCALE = HALF ? 16777216.f : 1.f;  // 2^24 undoes the subnormal scaling
  constexpr uint32_t ONE2 = HALF ? 0x3C003C00u : 0x3F803F80u;

  float sa[4], sb[4];    // per column: G multiplier, X multiplier
  float acc[2][MT][4];   // running sum of this group of warps over its stages of the current column block
  float ge[2][MT][4];    // G of the open scale group: sum_k x * operand (exact products, fp32 sums)
  float go[2][MT][4];    // EO: the same for the nibbles that arrive as 16 q
  float ax[MT][4];       // X[m] = sum_k x[m,k] of the open scale group, in the accumulator layout
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][mt][i] = ge[t][mt][i] = go[t][mt][i] = ax[mt][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) sa[j] = sb[j] = 0.f;

  auto close_group = [&]() {  // acc += sa[n] * G[m,n] + sb[n] * X[m]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 2 * t + (i >> 1);
          float gsum = ge[t][mt][i];
          if constexpr (EO) gsum = fmaf(go[t][mt][i], 0.0625f, gsum);  // the odd nibbles were 16 q: exact power-of-two scaling
          acc[t][mt][i] = fmaf(sa[c], gsum, fmaf(sb[c], ax[mt][i], acc[t][mt][i]));
        }
  };

  auto refresh = [&](const unsigned char* row) {  // scales / zero-points of this lane's four columns from the staged group row
    A s4[4];
    *reinterpret_cast<uint2*>(s4) = *reinterpret_cast<const uint2*>(row + lcol * 2);
    float z[4] = {zscalar, zscalar, zscalar, zscalar};
    if (zp_per_group) {
      if constexpr (ZK == CB200_ZP_I32) {
        const int4 zr = *reinterpret_cast<const int4*>(row + 256 + lcol * 4);
        z[0] = float(zr.x); z[1] = float(zr.y); z[2] = float(zr.z); z[3] = float(zr.w);
      } else if constexpr (ZK == CB200_ZP_U8) {
        const uint32_t zr = *reinterpret_cast<const uint32_t*>(row + 256 + lcol);
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = float((zr >> (8 * j)) & 0xffu);
      } else if constexpr (FZP) {
        A zf[4];
        *reinterpret_cast<uint2*>(zf) = *reinterpret_cast<const uint2*>(row + 256 + lcol * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = ActOps<A>::to_float(zf[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sj = ActOps<A>::to_float(s4[j]);
      sa[j] = sj * OPSCALE;
      sb[j] = -sj * (MAGIC + z[j] + fbias);
    }
  };

  // one step: 4 packed rows x this warp's 32 columns (32 k for int4, 16 k for int8) against every row of x
  auto step = [&](auto fresh_tag, const uint4 wv, const auto& xp) {
    constexpr bool FRESH = decltype(fresh_tag)::value;  // first step of a scale group: the MMAs start from zero
    uint32_t xb[MT][PAIRS];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if constexpr (BITS == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(xp[mt]);
        xb[mt][0] = __byte_perm(v.x, v.z, 0x5410);  // (k0,k4): the pair order the mask-only nibble extraction produces
        xb[mt][1] = __byte_perm(v.x, v.z, 0x7632);  // (k1,k5)
        xb[mt][2] = __byte_perm(v.y, v.w, 0x5410);  // (k2,k6)
        xb[mt][3] = __byte_perm(v.y, v.w, 0x7632);  // (k3,k7)
      } else {
        const uint2 v = *reinterpret_cast<const uint2*>(xp[mt]);
        xb[mt][0] = v.x;  // (k0,k1)
        xb[mt][1] = v.y;  // (k2,k3)
      }
    }
    const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t d[4][PAIRS];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if constexpr (EO) {  // fp16: 0x0000 | v read as fp16 is the subnormal v * 2^-24, exactly
        const uint32_t t8 = words[j] >> 8;
        d[j][0] = words[j] & 0x000f000fu;  // (k0, k4)
        d[j][1] = words[j] & 0x00f000f0u;  // (16 k1, 16 k5)
        d[j][2] = t8 & 0x000f000fu;        // (k2, k6)
        d[j][3] = t8 & 0x00f000f0u;        // (16 k3, 16 k7)
      } else if constexpr (BITS == 4) {    // bf16: 0x4300 | q = 128 + q
#pragma unroll
        for (int i = 0; i < 4; ++i) d[j][i] = lop3_and_or(words[j] >> (4 * i), 0x000f000fu, 0x43004300u);
      } else {  // fp16 int8: bytes as subnormals
        d[j][0] = prmt(words[j], 0u, 0x4140u);
        d[j][1] = prmt(words[j], 0u, 0x4342u);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if constexpr (BITS == 4) {
        // MMA "e": operand pairs 0 and 2 (plain nibbles) with x pairs 0 and 2; MMA "o": pairs 1 and 3 (the x16 kind for fp16)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mma_acc<FRESH, A>(ge[t][mt], d[2 * t][0], d[2 * t + 1][0], d[2 * t][2], d[2 * t + 1][2], xb[mt][0], xb[mt][2]);
          if constexpr (EO) {
            mma_acc<FRESH, A>(go[t][mt], d[2 * t][1], d[2 * t + 1][1], d[2 * t][3], d[2 * t + 1][3], xb[mt][1], xb[mt][3]);
          } else {
            mma_acc<false, A>(ge[t][mt], d[2 * t][1], d[2 * t + 1][1], d[2 * t][3], d[2 * t + 1][3], xb[mt][1], xb[mt][3]);
          }
        }
        mma_acc<FRESH, A>(ax[mt], ONE2, ONE2, ONE2, ONE2, xb[mt][0], xb[mt][2]);
        mma_acc<false, A>(ax[mt], ONE2, ONE2, ONE2, ONE2, xb[mt][1], xb[mt][3]);
      } else {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mma_acc<FRESH, A>(ge[t][mt], d[2 * t][0], d[2 * t + 1][0], d[2 * t][1], d[2 * t + 1][1], xb[mt][0], xb[mt][1]);
        }
        mma_acc<FRESH, A>(ax[mt], ONE2, ONE2, ONE2, ONE2, xb[mt][0], xb[mt][1]);
      }
    }
  };

  // ---- split-K publication of one 128-column block --------------------------------------------------------------------------
  // Every group of every CTA whose run touches the block writes an fp32 partial slot (zeros if it had no stage there).  The
  // ARRIVAL (fence + counter bump, a full round trip to L2) is deferred: a warp remembers the blocks it has written and
  // announces all of them at the end of its run with one fence and overlapping atomics; the last arriver of a (block,
  // quarter) sums the slots in slot order -- deterministic, self-resetting counters, no second launch.
  int pend_cb[MAXPEND];
  int npend = 0;
  auto slot_geometry = [&](int cb, int& nslots, int& myslot) {
    const int first_cta = stream_cta_of(cb * spc, geo.q, geo.r);
    const int last_cta = stream_cta_of(cb * spc + spc - 1, geo.q, geo.r);
    nslots = (last_cta - first_cta + 1) * G;
    myslot = (cta - first_cta) * G + cg;
  };
  auto write_partial = [&](int cb) {
    int nslots, myslot;
    slot_geometry(cb, nslots, myslot);
    const int unit = M * 32;
    float* mine = ws_partial + ((size_t(cb) * 4 + cq) * size_t(geo.maxslots) + size_t(myslot)) * unit;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int m = mt * 8 + 2 * tig;
        const int lc = 4 * chunk + 2 * t;
        if (m < M) __stcg(reinterpret_cast<float2*>(mine + m * 32 + lc), make_float2(acc[t][mt][0], acc[t][mt][2]));
        if (m + 1 < M) __stcg(reinterpret_cast<float2*>(mine + (m + 1) * 32 + lc), make_float2(acc[t][mt][1], acc[t][mt][3]));
      }
  };
  auto announce = [&]() {  // fence once, bump every pending block's counter (lane j handles pending block j), reduce where last
    if (npend == 0) return;
    __threadfence();
    __syncwarp();
    unsigned prev = 0;
    int my_nslots = 0;
    if (lane < npend) {
      int myslot;
      slot_geometry(pend_cb[0], my_nslots, myslot);
#pragma unroll
      for (int j = 1; j < MAXPEND; ++j) {
        if (lane == j) slot_geometry(pend_cb[j], my_nslots, myslot);
      }
      int cbj = pend_cb[0];
#pragma unroll
      for (int j = 1; j < MAXPEND; ++j) cbj = lane == j ? pend_cb[j] : cbj;
      prev = atomicAdd(ws_counter + cbj * 4 + cq, 1u);
    }
#pragma unroll
    for (int j = 0; j < MAXPEND; ++j) {
      if (j >= npend) break;
      const unsigned pj = __shfl_sync(0xffffffffu, prev, j);
      const int nsl = __shfl_sync(0xffffffffu, my_nslots, j);
      if (pj != unsigned(nsl - 1)) continue;
      __threadfence();  // last arriver: every other partial is visible now
      const int cb = pend_cb[j];
      const int col0 = cb * 128 + 32 * cq;
      const int unit = M * 32;
      const float* base = ws_partial + (size_t(cb) * 4 + cq) * size_t(geo.maxslots) * unit;
      // each lane owns 4 consecutive columns of one row: 8 lanes per row, 16-byte partial loads, 8-byte (fp16) stores
      for (int e = lane; e < M * 8; e += 32) {
        const int m = e >> 3, c4 = (e & 7) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < nsl; ++s) {
          const float4 q4 = __ldcg(reinterpret_cast<const float4*>(base + size_t(s) * unit + m * 32 + c4));
          v.x += q4.x; v.y += q4.y; v.z += q4.z; v.w += q4.w;
        }
        store_out4_all(p, size_t(m) * N + col0 + c4, v);
      }
      if (lane == 0) ws_counter[cb * 4 + cq] = 0u;  // leave the counter clean for the next launch
    }
    npend = 0;
  };
  auto publish = [&](int cb) {
    if (cb * 128 + 32 * cq >= N) return;  // quarter beyond N (N % 128 == 64): never computed, never counted
    write_partial(cb);
    pend_cb[npend < MAXPEND ? npend : MAXPEND - 1] = cb;  // (npend < MAXPEND here: announce() below keeps it so)
    ++npend;
    if (npend == MAXPEND) announce();
  };

  if (tracer) stream_stamp(geo.trace, 5);

  // cursors of this group's next stage (incremental: no divisions in the loop); ring >= G
  int cb = first_cb, kt = first_kt + cg;
  while (kt >= spc) {
    kt -= spc;
    ++cb;
  }
  int slot = cg;
  uint32_t par = 0;
  int cur_cb = first_cb;
  const int last_cb = (run_a + n_run - 1) / spc;
  int cur_grp = -1;
  bool open = false;  // ge / go / ax hold a partial scale group
  const bool gps2 = BITS == 4 && geo.gps == 2;
  const bool compute = !(geo.dbg & 1);
  for (int i = cg;; i += G) {
    const bool have = i < n_run;
    const int tcb = have ? cb : last_cb + 1;
    while (cur_cb < tcb) {  // this group is done with column block cur_cb
      if (open) close_group();
      open = false;
      cur_grp = -1;
      if (tracer && cur_cb == last_cb) stream_stamp(geo.trace, 7);
      publish(cur_cb);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[t][mt][j] = 0.f;
      ++cur_cb;
    }
    if (!have) break;
    mbar_wait_fast(full0 + 8u * slot, par);
    if (tracer && i == cg) stream_stamp(geo.trace, 6);
    if (cb * 128 + 32 * cq < N && compute) {
      const unsigned char* const sp = ring_ptr + slot * stage_bytes;
      // the stage's packed words (4 vectors) and activations (4 per m tile): independent shared-memory loads up front
      uint4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[u] = *reinterpret_cast<const uint4*>(sp + u * 512 + ((u & 1) ? wlane1 : wlane0));
      const unsigned char* xp[4][MT];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) xp[u][mt] = ring_ptr + (xbase[mt] + slot * xstage[mt] + u * xstep[mt]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // half-stage = 2 steps: a scale group can only start at a half-stage boundary
        if (h == 0 || gps2) {
          const int hs = 2 * kt + h;
          const int grp = geo.hpg_shift >= 0 ? hs >> geo.hpg_shift : hs / geo.hpg;
          if (grp != cur_grp) {
            if (open) close_group();
            open = false;
            cur_grp = grp;
            refresh(sp + geo.meta_off + (gps2 ? h * geo.meta_row_bytes : 0));
          }
        }
        if (!open) {
          step(std::true_type{}, wv[2 * h], xp[2 * h]);
        } else {
          step(std::false_type{}, wv[2 * h], xp[2 * h]);
        }
        open = true;
        step(std::false_type{}, wv[2 * h + 1], xp[2 * h + 1]);
      }
    }
    // ---- leave the stage; the last of the group's four warps to leave refills the slot with the stage `ring` ahead ------------
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();  // this warp's reads of the slot are done before anybody may overwrite it
      const unsigned old = atomicAdd(depart + slot, 1u);
      if (old == 3u) {
        depart[slot] = 0u;
        const int nxt = i + ring;
        if (nxt < n_run) {
          int ncb = cb, nkt = kt + ring;
          while (nkt >= spc) {
            nkt -= spc;
            ++ncb;
          }
          issue_static(slot, ncb, nkt);
          issue_x(slot, nkt);
        }
      }
    }
    __syncwarp();
    // advance to this group's next stage
    slot += G;
    if (slot >= ring) {
      slot -= ring;
      par ^= 1u;
    }
    kt += G;
    while (kt >= spc) {
      kt -= spc;
      ++cb;
    }
  }
  announce();
  if (tracer) {
    stream_stamp(geo.trace, 8);
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    geo.trace[blockIdx.x * 32 + 9] = sm;
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <class A, int BITS, int MT, int ZK>
static int launch_stream_cfg(const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const StreamGeom& geo, int ctas,
                             size_t smem, float* wsp, unsigned* wsc, cudaStream_t stream) {
  auto kern = stream_kernel<A, BITS, MT, ZK>;
  CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  // two of these kernels (this one and its successor in the stream) are meant to share an SM: ask for the full carve-out
  CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = geo.pdl ? 1 : 0;
  CB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mw, mx, p, geo, wsp, wsc));
  return 0;
}

template <class A, int BITS, int MT>
static int launch_stream_zk(const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const StreamGeom& geo, int ctas,
                            size_t smem, float* wsp, unsigned* wsc, cudaStream_t stream) {
  switch (p.zp_kind) {
    case CB200_ZP_NONE: return launch_stream_cfg<A, BITS, MT, CB200_ZP_NONE>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
    case CB200_ZP_I32: return launch_stream_cfg<A, BITS, MT, CB200_ZP_I32>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
    case CB200_ZP_U8: return launch_stream_cfg<A, BITS, MT, CB200_ZP_U8>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
    default: return launch_stream_cfg<A, BITS, MT, CB200_ZP_ACT>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
  }
}

template <class A, int BITS>
static int launch_stream_typed(const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const StreamGeom& geo, int ctas,
                               size_t smem, float* wsp, unsigned* wsc, cudaStream_t stream) {
  return geo.m_rows > 8 ? launch_stream_zk<A, BITS, 2>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream)
                        : launch_stream_zk<A, BITS, 1>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
}

// one entry per translation unit (stream_f16_w4.cu, stream_f16_w8.cu, stream_bf16_w4.cu)
#define CB200_STREAM_ENTRY(name)                                                                                          \
  int name(const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const StreamGeom& geo, int ctas, size_t smem, \
           float* wsp, unsigned* wsc, cudaStream_t stream)
CB200_STREAM_ENTRY(launch_stream_f16_w4);
CB200_STREAM_ENTRY(launch_stream_f16_w8);
CB200_STREAM_ENTRY(launch_stream_bf16_w4);

}  // namespace cb200
