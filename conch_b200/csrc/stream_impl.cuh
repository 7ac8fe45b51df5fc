// Streaming ("ring") decode kernel: device code + typed launcher, shared by the per-(activation dtype, weight width)
// translation units stream_f16_w4.cu / stream_f16_w8.cu / stream_bf16_w4.cu.  The design notes, the host-side plan and the
// dispatch are in stream_kernel.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dequant.cuh"
#include "host_util.cuh"
#include "mma_ptx.cuh"

namespace cb200 {

constexpr int kStreamMaxRing = 16;      // stages of the weight ring, at most
constexpr int kStreamMaxWarps = 16;     // consumer warps per CTA, at most (M <= 8; 14 for the 16-row variant)
constexpr int kStreamHdr = 1024;        // shared memory in front of everything: the mbarriers

struct StreamGeom {
  int m_rows;       // rows of x in this pass (<= 8 * MT)
  int nw;           // consumer warps per group = 32-column slices of a column block
  int ngrp;         // consumer groups: group h takes the stages h, h + ngrp, ... of the CTA's run
  int S;            // K slabs = cluster size: the S CTAs of a cluster share a column block and reduce through distributed smem
  int QPB;          // quads (4 packed rows) per column = K / (4 * pf)
  int qpg;          // quads per scale group (QPB when there is one group)
  int one_group;    // per-channel scales (group_size == K)
  int gps;          // scale rows per stage: 4 / qpg for qpg < 4, else 1
  int ngroups;      // K / group_size
  int zes;          // bytes per staged zero-point (4 int32, 1 uint8, 2 activation dtype, 0 none / scalar)
  int w_bytes;      // packed weights per stage: nw * 2048
  int stage_bytes;
  int ring;         // stages
  int x_pitch;      // bytes per staged x row
  int x_bytes;      // staged x slab
  int red_bytes;    // cluster reduction buffers + the group sums of x (at its end)
  int xsum_bytes;
  int win_stages;   // 16-row variant: stages of x staged at a time (the slab of x is windowed so that the weight ring keeps its depth)
  int no_hint;      // debug (cb200_set_stream_tuning ring_kb < 0): weight tiles without the L2 evict_first policy
  unsigned int* done;  // cb200_mpgemm_push_signal: self-resetting counter of finished 32-column slices (workspace), else unused
  long long* trace; // debug builds: [CTA][8] globaltimer stamps
};

namespace ring {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
// same, with an L2 cache policy (the weights are read exactly once: evict_first keeps them from displacing the activations,
// scales and zero-points that ARE re-read, by this kernel's other CTAs and by the next kernels)
__device__ __forceinline__ void tma_3d_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 16 bytes into a peer CTA's shared memory; the peer's mbarrier counts them
__device__ __forceinline__ void st_async_v4(uint32_t raddr, float4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"(rbar)
               : "memory");
}

// m16n8k16 with the accumulator as an in/out operand, and with a zero accumulator input (first MMA of a scale group)
template <class A>
__device__ __forceinline__ void mma_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<A, __half>::value) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}
template <class A>
__device__ __forceinline__ void mma_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<A, __half>::value) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
  } else {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
  }
}

// Per-(activation type, weight width) constants.
//   fp16 x int4 : nibbles at bits 0-3 of a byte become q * 2^-24 (fp16 subnormals: the value bits ARE the operand), nibbles at bits
//                 4-7 become 16 q * 2^-24 and accumulate separately (Q): G = (P + Q / 16) * 2^24, no instruction touches a weight
//   bf16 x int4 : 0x4300 | q = 128 + q (a bf16 significand holds 8 bits), one shift per nibble position; the 128 is taken out
//                 with the zero-point: G = P + Q, correction -(128 + zb) * X
//   fp16 x int8 : bytes become q * 2^-24; one MMA per 16-column tile and quad (a quad is 16 k), P only
//   bf16 x int8 : a byte does not fit a bf16 significand, its two NIBBLES do: q = lo + 16 hi, both planes as 0x4300 | nibble
//                 (the bf16 x int4 extraction applied to a word of 4 bytes yields (lo k0, lo k2), (hi k0, hi k2), (lo k1, lo k3),
//                 (hi k1, hi k3)); P takes the low plane, Q the high one against the SAME activations: G = P + 16 Q, and the
//                 128 (1 + 16) = 2176 X that the magic operands add goes out with the zero-point.  Two MMAs per byte pair -- the
//                 int4 rate per byte of weights
template <class A, int BITS>
struct Fmt {
  static constexpr bool HALF = std::is_same<A, __half>::value;
  static constexpr int PF = 32 / BITS;                 // k per packed word
  static constexpr int XQ = 4 * PF * 2;                // bytes of one x row per quad (4 packed rows)
  static constexpr int CPQ = XQ / 16;                  // 16-byte chunks of x per quad and row
  static constexpr bool NIB8 = !HALF && BITS == 8;     // bf16 x int8: two nibble planes
  static constexpr bool HAS_Q = BITS == 4 || NIB8;     // second accumulator set
  static constexpr float QSCALE = HALF ? 0.0625f : NIB8 ? 16.f : 1.f;
  static constexpr float MAGIC = HALF ? 0.f : NIB8 ? 2176.f : 128.f;
  static constexpr float OPSCALE = HALF ? 16777216.f : 1.f;
};

template <int MT, bool HAS_Q>
struct Accum {
  float P[MT][2][4];
  float Q[HAS_Q ? MT : 1][2][4];
  float XS[MT][4];  // 16-row variant only: sums of x over the open scale group, by an MMA against an all-ones operand
  float XT[MT][4];  //   (int4: the second half of the lane's activations accumulates here -- two chains of half the length)
};

// One quad: this lane's 4 packed words (4 columns, packed row 4 * quad + t) against its activations.
//   int4: xv[mt] = the lane's 8 activations permuted to {(k0,k4), (k2,k6), (k1,k5), (k3,k7)}
//   int8: xv[mt].x/.y = (k0,k1), (k2,k3) -- bf16 x int8: (k0,k2), (k1,k3)
template <class A, int BITS, int MT, bool FRESH>
__device__ __forceinline__ void quad_math(Accum<MT, Fmt<A, BITS>::HAS_Q>& a, const uint4 wv, const uint4 (&xv)[MT]) {
  using F = Fmt<A, BITS>;
  const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
  if constexpr (MT > 1) {
    // X[m] = sum_k x[m, k] for the zero-point term: the 16-row variant has no staging pass that could sum x on the side (its x
    // arrives by bulk copy and is used as it lies), so an all-ones 16-row operand makes every row of D the column sums of x^T --
    // c0 / c1 are X[2t], X[2t + 1], the layout the accumulators want.  Exact: products by 1.0, fp32 accumulation.
    constexpr uint32_t kOnes = F::HALF ? 0x3c003c00u : 0x3f803f80u;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if constexpr (FRESH) mma_zero<A>(a.XS[mt], kOnes, kOnes, kOnes, kOnes, xv[mt].x, xv[mt].y);
      else mma_acc<A>(a.XS[mt], kOnes, kOnes, kOnes, kOnes, xv[mt].x, xv[mt].y);
      if constexpr (BITS == 4) {
        if constexpr (FRESH) mma_zero<A>(a.XT[mt], kOnes, kOnes, kOnes, kOnes, xv[mt].z, xv[mt].w);
        else mma_acc<A>(a.XT[mt], kOnes, kOnes, kOnes, kOnes, xv[mt].z, xv[mt].w);
      }
    }
  }
  if constexpr (BITS == 4 || F::NIB8) {
    uint32_t d[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if constexpr (F::HALF) {
        const uint32_t t8 = w[c] >> 8;
        d[c][0] = w[c] & 0x000f000fu;  // (k0, k4)
        d[c][1] = w[c] & 0x00f000f0u;  // 16 (k1, k5)
        d[c][2] = t8 & 0x000f000fu;    // (k2, k6)
        d[c][3] = t8 & 0x00f000f0u;    // 16 (k3, k7)
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) d[c][i] = lop3_and_or(w[c] >> (4 * i), 0x000f000fu, 0x43004300u);  // 128 + (k_i, k_i+4)
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) {
        // (int4: the odd nibbles belong to the second half of the lane's 8 activations; bf16 x int8: the high nibbles of the same 4)
        const uint32_t qx0 = F::NIB8 ? xv[mt].x : xv[mt].z, qx1 = F::NIB8 ? xv[mt].y : xv[mt].w;
        if constexpr (FRESH) {
          mma_zero<A>(a.P[mt][tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][2], d[2 * tl + 1][2], xv[mt].x, xv[mt].y);
          mma_zero<A>(a.Q[mt][tl], d[2 * tl][1], d[2 * tl + 1][1], d[2 * tl][3], d[2 * tl + 1][3], qx0, qx1);
        } else {
          mma_acc<A>(a.P[mt][tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][2], d[2 * tl + 1][2], xv[mt].x, xv[mt].y);
          mma_acc<A>(a.Q[mt][tl], d[2 * tl][1], d[2 * tl + 1][1], d[2 * tl][3], d[2 * tl + 1][3], qx0, qx1);
        }
      }
  } else {
    uint32_t d[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      d[c][0] = prmt(w[c], 0u, 0x4140u);  // (k0, k1) * 2^-24
      d[c][1] = prmt(w[c], 0u, 0x4342u);  // (k2, k3) * 2^-24
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) {
        if constexpr (FRESH) {
          mma_zero<A>(a.P[mt][tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][1], d[2 * tl + 1][1], xv[mt].x, xv[mt].y);
        } else {
          mma_acc<A>(a.P[mt][tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][1], d[2 * tl + 1][1], xv[mt].x, xv[mt].y);
        }
      }
  }
}

#ifdef CB200_TC_TRACE
#define RING_STAMP(slot)                                                                             \
  do {                                                                                               \
    if (geo.trace && (threadIdx.x == 0 || (threadIdx.x == 32 * geo.ngrp * geo.nw && (slot) >= 6))) { \
      long long t_;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                         \
      geo.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = t_;                            \
    }                                                                                                \
  } while (0)
#else
#define RING_STAMP(slot) do { } while (0)
#endif

}  // namespace ring

// A: activation type, BITS: 4 | 8, MT: 8-row tiles of x (1: M <= 8, 2: M <= 16).
// CHAIN: cb200_mpgemm_chain -- the consumers wait for the completion counters of x inside the kernel (M <= 8 instantiations only).
template <class A, int BITS, int MT, bool CHAIN>
__global__ void __launch_bounds__(MT == 1 ? 32 * (kStreamMaxWarps + 1) : 32 * 15, 1)
    stream_kernel(const __grid_constant__ CUtensorMap tmw, const Problem p, const StreamGeom geo) {
  using namespace ring;
  using F = Fmt<A, BITS>;
  constexpr bool HAS_Q = F::HAS_Q;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // Programmatic dependent launch: release the NEXT kernel at once.  Its CTAs become resident as ours retire, set their rings up
  // and fetch their first stages of WEIGHTS (static data) before they wait for this grid to complete.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = geo.nw, S = geo.S, NGRP = geo.ngrp;
  const int rank = blockIdx.x, cb = blockIdx.y;  // gridDim.x == S == cluster size: blockIdx.x is the rank in the cluster
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t bar_full = sbase, bar_empty = sbase + 8 * kStreamMaxRing, bar_red = sbase + 16 * kStreamMaxRing;
  const uint32_t bar_x = sbase + 512;
  const uint32_t xs = sbase + kStreamHdr;
  const uint32_t recv = xs + geo.x_bytes;
  const uint32_t stages = (recv + geo.red_bytes + 1023u) & ~1023u;
  const int M = geo.m_rows;
  const int nthr = 32 * NGRP * nw;  // consumer threads
  RING_STAMP(0);
  if (warp == NGRP * nw) {  // the producer warp sets the barriers up, so that its first copies leave before anything else happens
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmw)) : "memory");
    if (lane < geo.ring) {
      mbar_init(bar_full + 8 * lane, 1);
      mbar_init(bar_empty + 8 * lane, nw);
    }
    if (MT > 1 && lane == 0) mbar_init(bar_x, 1);  // 16-row variant: the windows of x arrive by bulk copy
    if (S > 1 && rank + lane * S < nw) {  // the column slices (one per warp of a consumer group) this CTA reduces
      mbar_init(bar_red + 8 * lane, 1);
      mbar_expect_tx(bar_red + 8 * lane, uint32_t(S * NGRP) * uint32_t(M) * 128u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"r"(nthr + 32) : "memory");
  } else {
    asm volatile("bar.sync 2, %0;" ::"r"(nthr + 32) : "memory");
  }
  if (S > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // waited for just before the reduction
  const int QPB = geo.QPB, N = p.N, BN = 32 * nw;
  const int q_lo = rank * QPB / S, q_hi = (rank + 1) * QPB / S;
  const int qa0 = q_lo & ~3;  // stage i covers the quads [qa0 + 4i, qa0 + 4i + 4) that lie in [q_lo, q_hi)
  const int nstages = (q_hi - qa0 + 3) >> 2;
  const int cols = min(BN, N - cb * BN);
  const uint32_t s_off = geo.w_bytes, z_off = geo.w_bytes + geo.gps * BN * 2;
  const int qpg = geo.qpg;

  if (warp == NGRP * nw) {
    // ------------------------------------------------------------------ producer -------------------------------------------
    // The whole warp walks the stage list (every value below is warp-uniform, so it lives in uniform registers and the copy
    // instructions take their operands from there); one elected lane issues the copies.
    uint32_t leader = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t srow_bytes = uint32_t(cols) * 2u, zrow_bytes = uint32_t(cols) * uint32_t(geo.zes);
    const char* sp = reinterpret_cast<const char*>(p.s) + size_t(cb) * BN * 2;
    const char* zp = reinterpret_cast<const char*>(p.zp) + size_t(cb) * BN * geo.zes;
    const size_t s_pitch = size_t(N) * 2, z_pitch = size_t(N) * geo.zes;
    const uint64_t stream_policy = policy_evict_first();
    int grp0 = geo.one_group ? 0 : qa0 / qpg;  // first scale group of the stage; rem: quads of that group in front of the stage
    int rem = geo.one_group ? 0 : qa0 - grp0 * qpg;
    for (int i = 0; i < nstages; ++i) {
      if (i >= geo.ring) mbar_wait(bar_empty + 8 * slot, phase ^ 1);
      const uint32_t dst = stages + slot * geo.stage_bytes;
      const uint32_t fb = bar_full + 8 * slot;
      const int ng = min(geo.gps, geo.ngroups - grp0);
      if (leader) {
        mbar_expect_tx(fb, uint32_t(geo.w_bytes) + uint32_t(ng) * (srow_bytes + zrow_bytes));
        if (geo.no_hint) tma_3d(dst, &tmw, 0, (qa0 + 4 * i) * 4, cb * nw, fb);
        else tma_3d_hint(dst, &tmw, 0, (qa0 + 4 * i) * 4, cb * nw, fb, stream_policy);
#pragma unroll 1
        for (int j = 0; j < ng; ++j) {
          bulk_g2s(dst + s_off + j * BN * 2, sp + size_t(grp0 + j) * s_pitch, srow_bytes, fb);
          if (zrow_bytes) bulk_g2s(dst + z_off + j * BN * geo.zes, zp + size_t(grp0 + j) * z_pitch, zrow_bytes, fb);
        }
      }
      if (i == 0) RING_STAMP(6);
      if (!geo.one_group) {
        if (qpg >= 4) {
          rem += 4;
          if (rem >= qpg) {
            rem -= qpg;
            ++grp0;
          }
        } else {
          grp0 += 4 / qpg;
        }
      }
      if (++slot == geo.ring) {
        slot = 0;
        phase ^= 1;
      }
    }
    RING_STAMP(7);
    return;
  }

  // -------------------------------------------------------------------- consumers ------------------------------------------
  const int grp = warp / nw, wl = warp - grp * nw;  // consumer group, column slice [32 wl, 32 wl + 32) of the block
  const int gq = lane >> 2, t = lane & 3;
  const int ch = (gq >> 1) | ((gq & 1) << 2);  // this lane's 16-byte chunk (4 columns) of the warp's 128-byte rows
  const bool active = wl * 32 < cols;
  const int mshift = qpg == 1 ? 0 : qpg == 2 ? 1 : 2;  // scale row of a quad inside its stage: (quad & 3) >> mshift
  constexpr int XR = 8 * MT;                           // rows of the sums table
  const uint32_t xsum = recv + geo.red_bytes - geo.xsum_bytes;  // [stage][scale row][XR] fp32: sums of x over the quads of a scale row
  asm volatile("griddepcontrol.wait;" ::: "memory");  // x, out and the peers belong to the stream order
  if constexpr (CHAIN) {
    // cb200_mpgemm_chain: x is a gather buffer the other GPUs push into.  One warp spins until every rank's completion counter
    // has reached this call's number (system-scope acquire), the rest of the consumers wait at the barrier; the producer warp is
    // not involved -- the ring keeps filling with weights meanwhile.
    if (warp == 0) {
      uint32_t seen;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.wait_expected) : "memory");
      const uint32_t want = seen + 1u;
      if (lane < p.wait_count) {
        const uint32_t* f = p.wait_flags + lane;
        for (unsigned spin = 0;; ++spin) {
          uint32_t v;
          asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
          if (int32_t(v - want) >= 0) break;
          __nanosleep(64);
          if (spin > (1u << 24)) __trap();  // a lost signal becomes a launch error (after seconds), not a hung GPU
        }
      }
      __syncwarp();
      asm volatile("fence.acq_rel.sys;" ::: "memory");  // one acquire for the whole warp's polls
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  }
  RING_STAMP(1);
  // Stage the activations of the stages [wi0, wi1) of this slab (int4: every 8-k chunk permuted to the operand order) and sum them
  // per (stage, scale row): consecutive lanes hold the chunks of one stage of one row -> a fixed shuffle tree, the same whatever M
  // is.  M <= 8: one window = the whole slab.  16 rows: windows of geo.win_stages stages -- a whole slab of 16 rows (128 KB at
  // K = 4096) would leave the weight ring four stages, too few bytes in flight to cover the HBM latency.
  uint32_t x_phase = 0;
  auto stage_x = [&](int wi0, int wi1) {
    constexpr int CPS = 4 * F::CPQ;              // chunks per stage and row
    const int cw0 = wi0 * CPS;                   // first chunk of the window, counted from the aligned start qa0
    const int cpr = (wi1 - wi0) * CPS;           // chunks per row in the window
    const int c_lo = max((q_lo - qa0) * F::CPQ, cw0), c_hi = (q_hi - qa0) * F::CPQ;
    const int esh = (BITS == 4 ? 2 : 1) + mshift;  // chunks per scale row = 1 << esh
    const int e0 = wi0 * geo.gps;                  // first entry of the sums table in the window
    const char* xg = reinterpret_cast<const char*>(p.x) + size_t(qa0) * F::XQ;
    if constexpr (MT > 1) {
      // 16 rows: one bulk copy per row brings the window in, and that is all -- the register-staged pass below (a round trip per
      // four chunks and thread, a shuffle tree per chunk) cost 11 us of a 38 us kernel at 16 rows (profiles/r02_stream_m16_ncu.txt).
      const int c_end = min(c_hi, cw0 + cpr);
      const uint32_t row_bytes = uint32_t(c_end - c_lo) * 16u;
      if (warp == 0) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the previous window was read / written by ordinary accesses
          mbar_expect_tx(bar_x, uint32_t(M) * row_bytes);
        }
        __syncwarp();
        if (lane < M) bulk_g2s(xs + lane * geo.x_pitch, xg + size_t(lane) * p.K * 2 + size_t(c_lo) * 16, row_bytes, bar_x);
      }
      mbar_wait(bar_x, x_phase);
      x_phase ^= 1u;
      return;  // used as it lies: the consumers permute their chunks in registers and sum x with an MMA (quad_math)
    }
    constexpr int XU = 4;  // loads in flight per thread
    for (int base0 = threadIdx.x & ~31; base0 < M * cpr; base0 += XU * nthr) {
      uint4 v[XU];
      int mm[XU], cv[XU];
      bool ok[XU];
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        const int idx = base0 + u * nthr + lane;
        mm[u] = idx / cpr;
        cv[u] = cw0 + idx - mm[u] * cpr;
        ok[u] = idx < M * cpr && cv[u] >= c_lo && cv[u] < c_hi;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (ok[u]) {
          // (a pushed-into x is read past L1: these lines were written by other GPUs while this kernel was resident)
          const uint4* src = reinterpret_cast<const uint4*>(xg + size_t(mm[u]) * p.K * 2) + cv[u];
          if constexpr (CHAIN) v[u] = __ldcg(src);
          else v[u] = __ldg(src);
        }
      }
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        if (base0 + u * nthr >= M * cpr) break;  // (warp-uniform)
        float sum = 0.f;
        if (ok[u]) {
          uint4 o = v[u];
          if constexpr (BITS == 4) {
            o.x = prmt(v[u].x, v[u].z, 0x5410u);
            o.y = prmt(v[u].y, v[u].w, 0x5410u);
            o.z = prmt(v[u].x, v[u].z, 0x7632u);
            o.w = prmt(v[u].y, v[u].w, 0x7632u);
          } else if constexpr (F::NIB8) {  // two words' worth per chunk: (k0,k1),(k2,k3) -> (k0,k2),(k1,k3)
            o.x = prmt(v[u].x, v[u].y, 0x5410u);
            o.y = prmt(v[u].x, v[u].y, 0x7632u);
            o.z = prmt(v[u].z, v[u].w, 0x5410u);
            o.w = prmt(v[u].z, v[u].w, 0x7632u);
          }
          sts128(xs + mm[u] * geo.x_pitch + (cv[u] - c_lo) * 16, o);
          const uint32_t r[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 f;
            if constexpr (F::HALF) f = __half22float2(bits_h2(r[k]));
            else f = __bfloat1622float2(bits_b2(r[k]));
            sum += f.x + f.y;
          }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        if (esh > 1) sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (esh > 2) sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        if (esh > 3) sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        if (mm[u] < M && (cv[u] & ((1 << esh) - 1)) == 0)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum + (((cv[u] >> esh) - e0) * XR + mm[u]) * 4), "f"(sum) : "memory");
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  };

  float acc[MT][2][4];
  Accum<MT, HAS_Q> qa;
  float sa[4], sb[4], X[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[mt][0][i] = acc[mt][1][i] = 0.f;
      qa.P[mt][0][i] = qa.P[mt][1][i] = 0.f;
      if constexpr (HAS_Q) qa.Q[mt][0][i] = qa.Q[mt][1][i] = 0.f;
    }
    X[mt][0] = X[mt][1] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) sa[i] = sb[i] = 0.f;
  const uint32_t lane_w0 = wl * 2048 + t * 128 + (uint32_t(ch ^ t) << 4);      // quads 0 / 2 of a stage: + 0 / + 1024
  const uint32_t lane_w1 = wl * 2048 + t * 128 + (uint32_t(ch ^ t ^ 4) << 4);  // quads 1 / 3: + 512 / + 1536
  uint32_t lane_x[MT];                                                         // + quad * XQ (set per window of x)
  const uint32_t lane_s = s_off + (wl * 32 + ch * 4) * 2, lane_z = z_off + (wl * 32 + ch * 4) * geo.zes;
  uint32_t lane_xs = xsum + 8 * t;  // + entry * XR * 4 (+ 32 for the second row tile): X[m = 2t], X[m = 2t + 1] (per window)
  auto enter_window = [&](int wi0) {  // x and its sums are addressed by absolute quad / stage: rebase them on the staged window
    const int qbase = max(q_lo, qa0 + 4 * wi0);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int m = 8 * mt + gq;
      lane_x[mt] = xs + (m < M ? m : 0) * geo.x_pitch + t * (F::XQ / 4) - qbase * F::XQ;
    }
    lane_xs = xsum + 8 * t - uint32_t(wi0 * geo.gps) * (XR * 4);
  };
  const float fbias = float(p.bias);
  float zscalar = 0.f;
  if (p.zp_kind != CB200_ZP_NONE && p.zp_scalar) {  // one zero-point for the whole matrix: fetch it once
    if (p.zp_kind == CB200_ZP_I32) zscalar = float(__ldg(reinterpret_cast<const int32_t*>(p.zp)));
    else if (p.zp_kind == CB200_ZP_U8) zscalar = float(__ldg(reinterpret_cast<const uint8_t*>(p.zp)));
    else zscalar = ActOps<A>::to_float(*reinterpret_cast<const A*>(p.zp));
  }
  const int zes = geo.zes;

  // sa[c] = s * OPSCALE multiplies G = sum_k x * operand; sb[c] = -s * (MAGIC + bias + zero-point) multiplies X = sum_k x
  auto set_scales = [&](uint32_t st, uint32_t mrow) {
    const uint2 sv = lds64(st + lane_s + mrow * (BN * 2));
    A s4[4];
    *reinterpret_cast<uint2*>(s4) = sv;
    float z[4];
    if (zes == 4) {
      const uint4 zv = lds128(st + lane_z + mrow * (BN * 4));
      z[0] = float(int(zv.x)); z[1] = float(int(zv.y)); z[2] = float(int(zv.z)); z[3] = float(int(zv.w));
    } else if (zes == 1) {
      const uint32_t zv = lds32(st + lane_z + mrow * BN);
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] = float((zv >> (8 * c)) & 0xffu);
    } else if (zes == 2) {
      const uint2 zv = lds64(st + lane_z + mrow * (BN * 2));
      A z4[4];
      *reinterpret_cast<uint2*>(z4) = zv;
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] = ActOps<A>::to_float(z4[c]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] = zscalar;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float sf = ActOps<A>::to_float(s4[c]);
      sa[c] = sf * F::OPSCALE;
      sb[c] = -sf * (F::MAGIC + fbias + z[c]);
    }
  };
  auto close_group = [&]() {  // acc += sa[n] * G[m, n] + sb[n] * X[m]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int tl = 0; tl < 2; ++tl)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 2 * tl + (i >> 1);
          float G = qa.P[mt][tl][i];
          if constexpr (HAS_Q) G = fmaf(qa.Q[mt][tl][i], F::QSCALE, G);
          const float xsum_m = MT > 1 ? (BITS == 4 ? qa.XS[mt][i & 1] + qa.XT[mt][i & 1] : qa.XS[mt][i & 1]) : X[mt][i & 1];
          acc[mt][tl][i] = fmaf(sa[c], G, fmaf(sb[c], xsum_m, acc[mt][tl][i]));
        }
  };
  auto clear_pq = [&]() {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        qa.P[mt][0][k] = qa.P[mt][1][k] = 0.f;
        if constexpr (HAS_Q) qa.Q[mt][0][k] = qa.Q[mt][1][k] = 0.f;
        X[mt][k & 1] = 0.f;
        if constexpr (MT > 1) qa.XS[mt][k] = qa.XT[mt][k] = 0.f;
      }
  };
  auto add_xsum = [&](int entry) {
    if constexpr (MT > 1) return;  // (summed by quad_math)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint2 xr = lds64(lane_xs + entry * (XR * 4) + mt * 32);
      X[mt][0] += __uint_as_float(xr.x);
      X[mt][1] += __uint_as_float(xr.y);
    }
  };
  auto quad_addr = [&](uint32_t st, int j) { return st + ((j & 1) ? lane_w1 : lane_w0) + j * 512; };
  auto load_x = [&](uint4 (&xv)[MT], int quad) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if constexpr (BITS == 4) {
        const uint4 v = lds128(lane_x[mt] + quad * F::XQ);
        if constexpr (MT > 1) {  // raw chunk -> operand order {(k0,k4), (k2,k6), (k1,k5), (k3,k7)}
          xv[mt].x = prmt(v.x, v.z, 0x5410u);
          xv[mt].y = prmt(v.y, v.w, 0x5410u);
          xv[mt].z = prmt(v.x, v.z, 0x7632u);
          xv[mt].w = prmt(v.y, v.w, 0x7632u);
        } else {
          xv[mt] = v;
        }
      } else {
        const uint2 v = lds64(lane_x[mt] + quad * F::XQ);
        if constexpr (MT > 1 && F::NIB8) xv[mt] = make_uint4(prmt(v.x, v.y, 0x5410u), prmt(v.x, v.y, 0x7632u), 0u, 0u);  // raw window
        else xv[mt] = make_uint4(v.x, v.y, 0u, 0u);
      }
    }
  };

  // stages grp, grp + NGRP, ...: slot = i % ring, phase = (i / ring) & 1
  int slot = grp % geo.ring;
  uint32_t phase = (grp / geo.ring) & 1;
  auto advance = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
    slot += NGRP;  // (the ring is a multiple of the number of groups: at most one wrap)
    if (slot >= geo.ring) {
      slot -= geo.ring;
      phase ^= 1;
    }
  };
  bool open = false;  // P / Q / X hold a partial scale group
  int open_sg = 0;
  // any stage, one scale row at a time (first / last stage of an unaligned slab, small group sizes)
  auto generic_stage = [&](int i) {
    const int qb = qa0 + 4 * i;
    const int j0 = max(q_lo - qb, 0), j1 = min(q_hi - qb, 4);
    const uint32_t st = stages + slot * geo.stage_bytes;
    mbar_wait(bar_full + 8 * slot, phase);
    if (active) {
      const int qrow = 1 << mshift;
      for (int r = 0; r < (4 >> mshift); ++r) {
        const int ja = max(j0, r * qrow), jb = min(j1, (r + 1) * qrow);
        if (ja >= jb) continue;
        const int sg = geo.one_group ? 0 : (qb + ja) / qpg;
        if (open && sg != open_sg) {
          close_group();
          open = false;
        }
        if (!open) {
          set_scales(st, r);
          clear_pq();
          open = true;
          open_sg = sg;
        }
        add_xsum(i * geo.gps + r);
        for (int j = ja; j < jb; ++j) {
          uint4 xv[MT];
          load_x(xv, qb + j);
          quad_math<A, BITS, MT, false>(qa, lds128(quad_addr(st, j)), xv);
        }
      }
    }
    advance();
  };

  const int i_full0 = q_lo > qa0 ? 1 : 0, i_full1 = (q_hi - qa0) >> 2;  // stages [i_full0, i_full1) lie wholly inside the slab
  int i = grp;
  // (M <= 8: exactly one window -- written so that the compiler sees a single trip)
  for (int wi0 = 0; wi0 < (MT == 1 ? 1 : nstages); wi0 += (MT == 1 ? 1 : geo.win_stages)) {
  const int wi1 = MT == 1 ? nstages : min(nstages, wi0 + geo.win_stages);
  if (MT > 1 && wi0 > 0) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");  // every warp is done with the previous window of x
  stage_x(wi0, wi1);
  enter_window(wi0);
  if (wi0 == 0) RING_STAMP(2);
  const int w_full1 = min(i_full1, wi1);
  if (mshift == 2) {
    if (i < i_full0) {
      generic_stage(i);
      i += NGRP;
    }
    if (qpg == 4) {
      // ---- every stage is one scale group (int4: group size 128, int8: 64): no state carried between stages ------------
      if (open) {  // (an unaligned first stage is a group of its own)
        if (active) close_group();
        open = false;
      }
      for (; i < w_full1; i += NGRP) {
        const uint32_t st = stages + slot * geo.stage_bytes;
        mbar_wait(bar_full + 8 * slot, phase);
        if (i == grp) RING_STAMP(3);
        if (active) {
          uint4 wv[4], xv[4][MT];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            wv[j] = lds128(quad_addr(st, j));
            load_x(xv[j], qa0 + 4 * i + j);
          }
          set_scales(st, 0);
          if constexpr (MT == 1) {
            const uint2 xr = lds64(lane_xs + i * (XR * 4));
            X[0][0] = __uint_as_float(xr.x);
            X[0][1] = __uint_as_float(xr.y);
          }
          quad_math<A, BITS, MT, true>(qa, wv[0], xv[0]);
#pragma unroll
          for (int j = 1; j < 4; ++j) quad_math<A, BITS, MT, false>(qa, wv[j], xv[j]);
          close_group();
        }
        advance();
      }
    } else {
      // ---- scale groups span several stages (larger groups, per-channel) ------------------------------------------------
      for (; i < w_full1; i += NGRP) {
        const uint32_t st = stages + slot * geo.stage_bytes;
        mbar_wait(bar_full + 8 * slot, phase);
        if (i == grp) RING_STAMP(3);
        if (active) {
          uint4 wv[4], xv[4][MT];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            wv[j] = lds128(quad_addr(st, j));
            load_x(xv[j], qa0 + 4 * i + j);
          }
          const int sg = geo.one_group ? 0 : (qa0 + 4 * i) / qpg;
          if (open && sg != open_sg) {
            close_group();
            open = false;
          }
          if (!open) {
            set_scales(st, 0);
            clear_pq();
            open = true;
            open_sg = sg;
          }
          add_xsum(i);
#pragma unroll
          for (int j = 0; j < 4; ++j) quad_math<A, BITS, MT, false>(qa, wv[j], xv[j]);
        }
        advance();
      }
    }
    if (wi1 > i_full1)
      for (; i < wi1; i += NGRP) generic_stage(i);  // an unaligned last stage
  } else {
    for (; i < wi1; i += NGRP) generic_stage(i);
  }
  }  // windows of x
  if (open && active) close_group();
  RING_STAMP(4);

  // ---- reduce over the cluster's K slabs and the consumer groups, store ---------------------------------------------------
  // lane (gq, t) holds columns col..col+3 of rows m = 8 mt + 2t (c0 / c2 of the two 16-column tiles) and 8 mt + 2t + 1 (c1 / c3).
  // Every (rank, group) sends its partial to the CTA that owns the column slice: rank wl % S, buffer wl / S; that CTA's group-0
  // warp adds the S * NGRP partials in a fixed order (deterministic) and stores.
  if (S > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // every CTA of the cluster has its barriers up
  const uint32_t dest = uint32_t(wl % S), wslot = uint32_t(wl / S);
  const int nsrc = S * NGRP;
  const int col = cb * BN + wl * 32 + ch * 4;
  if (active) {
    const uint32_t lbuf = recv + ((wslot * nsrc + rank * NGRP + grp) * M) * 128 + ch * 16;  // [wslot][source][m][32 columns] fp32
    uint32_t rbuf = lbuf, rbar = 0;
    if (S > 1) {
      rbuf = mapa(lbuf, dest);
      rbar = mapa(bar_red + 8 * wslot, dest);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = 8 * mt + 2 * t + h;
        const float4 r = make_float4(acc[mt][0][h], acc[mt][0][2 + h], acc[mt][1][h], acc[mt][1][2 + h]);
        if (m < M) {
          if (S > 1) st_async_v4(rbuf + m * 128, r, rbar);
          else sts128f(lbuf + m * 128, r);
        }
      }
  }
  if (S == 1) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  if (active && grp == 0 && dest == uint32_t(rank)) {
    if (S > 1) mbar_wait(bar_red + 8 * wslot, 0);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = 8 * mt + 2 * t + h;
        if (m < M) {
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int src = 0; src < nsrc; ++src) {
            const float4 v = lds128f(recv + ((wslot * nsrc + src) * M + m) * 128 + ch * 16);
            r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
          }
          if (p.out_bias != nullptr) {
            r.x += bias_at(p, col); r.y += bias_at(p, col + 1); r.z += bias_at(p, col + 2); r.w += bias_at(p, col + 3);
          }
          store_out4_all(p, size_t(m) * N + col, r);
        }
      }
    push_signal_block_done(p, geo.done, unsigned(N / 32), lane);  // (no-op unless completion flags were asked for)
  }
  RING_STAMP(5);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <class A, int BITS, int MT, bool CHAIN>
static int launch_stream_mt(const CUtensorMap& mw, const Problem& p, const StreamGeom& geo, dim3 grid, size_t smem, bool pdl,
                            cudaStream_t stream) {
  auto kern = stream_kernel<A, BITS, MT, CHAIN>;
  static unsigned long long attr_mask = 0;  // per (instantiation, device)
  {
    const DeviceInfo* di = current_device_info();
    const unsigned long long bit = 1ull << ((di ? di->index : 0) & 63);
    if (!(attr_mask & bit)) {
      CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_mask |= bit;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(32 * (geo.ngrp * geo.nw + 1));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (geo.S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = geo.S;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mw, p, geo));
  return 0;
}

template <class A, int BITS>
static int launch_stream_typed(const CUtensorMap& mw, const Problem& p, const StreamGeom& geo, dim3 grid, size_t smem, bool pdl,
                               cudaStream_t stream) {
  if (p.wait_flags != nullptr) return launch_stream_mt<A, BITS, 1, true>(mw, p, geo, grid, smem, pdl, stream);  // (M <= 8: launch_stream)
  return geo.m_rows > 8 ? launch_stream_mt<A, BITS, 2, false>(mw, p, geo, grid, smem, pdl, stream)
                        : launch_stream_mt<A, BITS, 1, false>(mw, p, geo, grid, smem, pdl, stream);
}

// one entry per translation unit
#define CB200_STREAM_ENTRY(name) \
  int name(const CUtensorMap& mw, const Problem& p, const StreamGeom& geo, dim3 grid, size_t smem, bool pdl, cudaStream_t stream)
CB200_STREAM_ENTRY(launch_stream_f16_w4);
CB200_STREAM_ENTRY(launch_stream_f16_w8);
CB200_STREAM_ENTRY(launch_stream_bf16_w4);
CB200_STREAM_ENTRY(launch_stream_bf16_w8);

}  // namespace cb200
