// Prototype / micro-benchmark of the "ring" decode kernel (fp16 x uint4, int32 zero-points, group scales).
//
//   * one CTA per SM: 1 producer warp + NW consumer warps.  The producer moves packed-weight rows (and the scale / zero-point
//     row of their group) into a ring of shared-memory stages with cp.async.bulk -- it never waits for the stream order, so a
//     kernel launched with programmatic dependent launch fills its ring while its predecessor is still computing.
//   * half an SM per CTA (<= 112 registers, ~110 KB of shared memory): the successor is co-resident.
//   * work = k-fastest list of (256-column block, 4-packed-row quad) units cut into #CTAs equal runs (stream-K).
//   * consumers: warp w owns columns [32w, 32w+32) of the block; per quad one ld.shared.v4 of words + one of x, nibbles become
//     fp16 subnormal MMA operands (no arithmetic), mma.sync m16n8k16 with weights as the 16-row operand.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o ring_gemv.bin ring_gemv.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));     \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

constexpr int MAXRING = 16;
constexpr int MAXNW = 8;
constexpr int HDR = 1024;  // barriers: full[16], empty[16], red[8]

struct Geo {
  const __half* s;
  const int* zp;
  const __half* x;
  __half* out;
  long long* trace;
  int M, K, N;
  int QPB;      // quads (4 packed rows = 32 k) per column = K / 32
  int nw;       // consumer warps; a column block is 32 * nw columns
  int ngrp;     // consumer groups: group h takes the stages h, h + ngrp, ... of the CTA's run
  int S;        // K slabs = cluster size: the S CTAs of a cluster share a column block and reduce through distributed smem
  int qpg;      // quads per scale group
  int gshift;   // log2(qpg), or -1: one group (per-channel scales)
  int gps;      // scale groups per stage (a stage is an aligned block of 4 quads)
  int ngroups;
  int w_bytes;  // packed weights per stage: nw * 2048
  int stage_bytes;
  int ring;     // stages
  int x_pitch;  // bytes per staged x row
  int x_bytes;  // staged x slab
  int red_bytes;   // cluster reduction buffers + the group sums of x
  int xsum_bytes;
  int dbg;      // 1: no math, 2: no weight copies, 4: no scale / zero-point copies
};

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
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef TRACE
#define STAMP(slot)                                                                                    \
  do {                                                                                                 \
    if (g.trace && (threadIdx.x == 0 || (threadIdx.x == 32 * g.ngrp * g.nw && (slot) >= 6)))                    \
      g.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime();                           \
  } while (0)
#else
#define STAMP(slot) do { } while (0)
#endif

struct QuadAcc {
  float P[2][4], Q[2][4];
};

// one quad: this lane's 4 words (4 columns x 8 k of packed row 4 * quad + t) against its x chunk, pre-permuted to
// {(k0,k4), (k2,k6), (k1,k5), (k3,k7)}.  Nibbles at bits 0-3 of a byte become q * 2^-24 (fp16 subnormals), nibbles at bits 4-7
// become 16 q * 2^-24 and accumulate separately (Q), so no instruction touches the weights' values.
__device__ __forceinline__ void quad_math(QuadAcc& a, const uint4 wv, const uint4 xv) {
  const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
  uint32_t d[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#ifdef NO_DEQ
    d[c][0] = d[c][1] = d[c][2] = d[c][3] = w[c];
#else
    const uint32_t t8 = w[c] >> 8;
    d[c][0] = w[c] & 0x000f000fu;
    d[c][1] = w[c] & 0x00f000f0u;
    d[c][2] = t8 & 0x000f000fu;
    d[c][3] = t8 & 0x00f000f0u;
#endif
  }
#ifdef NO_MMA
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    a.P[tl][0] += __uint_as_float((d[2 * tl][0] ^ d[2 * tl + 1][0] ^ d[2 * tl][2] ^ d[2 * tl + 1][2]) & xv.x ^ xv.y);
    a.Q[tl][0] += __uint_as_float((d[2 * tl][1] ^ d[2 * tl + 1][1] ^ d[2 * tl][3] ^ d[2 * tl + 1][3]) & xv.z ^ xv.w);
  }
#else
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    mma16816(a.P[tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][2], d[2 * tl + 1][2], xv.x, xv.y);
    mma16816(a.Q[tl], d[2 * tl][1], d[2 * tl + 1][1], d[2 * tl][3], d[2 * tl + 1][3], xv.z, xv.w);
  }
#endif
}


// mma with a zero accumulator input (first MMA of a scale group: saves clearing the registers)
__device__ __forceinline__ void mma16816_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
template <bool FRESH>
__device__ __forceinline__ void quad_math2(QuadAcc& a, const uint4 wv, const uint4 xv) {
  const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
  uint32_t d[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t t8 = w[c] >> 8;
    d[c][0] = w[c] & 0x000f000fu;
    d[c][1] = w[c] & 0x00f000f0u;
    d[c][2] = t8 & 0x000f000fu;
    d[c][3] = t8 & 0x00f000f0u;
  }
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    if constexpr (FRESH) {
      mma16816_z(a.P[tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][2], d[2 * tl + 1][2], xv.x, xv.y);
      mma16816_z(a.Q[tl], d[2 * tl][1], d[2 * tl + 1][1], d[2 * tl][3], d[2 * tl + 1][3], xv.z, xv.w);
    } else {
      mma16816(a.P[tl], d[2 * tl][0], d[2 * tl + 1][0], d[2 * tl][2], d[2 * tl + 1][2], xv.x, xv.y);
      mma16816(a.Q[tl], d[2 * tl][1], d[2 * tl + 1][1], d[2 * tl][3], d[2 * tl + 1][3], xv.z, xv.w);
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(32 * (2 * MAXNW + 1), 1) ring_gemv(const __grid_constant__ CUtensorMap tmw, const Geo g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = g.nw, S = g.S, NGRP = g.ngrp;
  const int rank = blockIdx.x, cb = blockIdx.y;  // gridDim.x == S == cluster size: blockIdx.x is the rank in the cluster
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t bar_full = sbase, bar_empty = sbase + 8 * MAXRING, bar_red = sbase + 16 * MAXRING;
  const uint32_t xs = sbase + HDR;
  const uint32_t recv = xs + g.x_bytes;
  const uint32_t stages = (recv + g.red_bytes + 1023u) & ~1023u;
  const int M = g.M;
  STAMP(0);
  const int nthr = 32 * NGRP * nw;  // consumer threads
  if (warp == NGRP * nw) {          // the producer warp sets the barriers up, so that its first copies leave before anything else
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmw)) : "memory");
    if (lane < g.ring) {
      mbar_init(bar_full + 8 * lane, 1);
      mbar_init(bar_empty + 8 * lane, nw);
    }
    if (S > 1 && rank + lane * S < nw) {  // the column groups (one per warp of a consumer group) this CTA reduces
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
  const int QPB = g.QPB, N = g.N, BN = 32 * nw;
  const int q_lo = rank * QPB / S, q_hi = (rank + 1) * QPB / S;
  const int qa0 = q_lo & ~3;                         // stage i covers the quads [qa0 + 4i, qa0 + 4i + 4) that lie in [q_lo, q_hi)
  const int nstages = (q_hi - qa0 + 3) >> 2;
  const int cols = min(BN, N - cb * BN);
  const uint32_t s_off = g.w_bytes, z_off = g.w_bytes + g.gps * BN * 2;

  if (warp == NGRP * nw) {
    // ------------------------------------------------------------------ producer -----------------------------------------
    // The whole warp walks the stage list (every value below is warp-uniform, so it lives in uniform registers and the copy
    // instructions take their operands from there); one elected lane issues the copies.
    uint32_t leader = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t meta_bytes = uint32_t(cols) * 6u;
    const char* sp = reinterpret_cast<const char*>(g.s + cb * BN);
    const char* zp = reinterpret_cast<const char*>(g.zp + cb * BN);
    const size_t s_row = size_t(N) * 2, z_row = size_t(N) * 4;
    for (int i = 0; i < nstages; ++i) {
      if (i >= g.ring) mbar_wait(bar_empty + 8 * slot, phase ^ 1);
      const uint32_t dst = stages + slot * g.stage_bytes;
      const uint32_t fb = bar_full + 8 * slot;
      const int qa = qa0 + 4 * i;
      const int grp0 = g.gshift >= 0 ? qa >> g.gshift : 0;
      const int ng = min(g.gps, g.ngroups - grp0);
      if (leader) {
        mbar_expect_tx(fb, (g.dbg & 2 ? 0 : g.w_bytes) + (g.dbg & 4 ? 0 : ng * meta_bytes));
        if (!(g.dbg & 2)) tma_3d(dst, &tmw, 0, qa * 4, cb * nw, fb);
        if (!(g.dbg & 4)) {
#pragma unroll 1
          for (int j = 0; j < ng; ++j) {
            bulk_g2s(dst + s_off + j * BN * 2, sp + size_t(grp0 + j) * s_row, cols * 2, fb);
            bulk_g2s(dst + z_off + j * BN * 4, zp + size_t(grp0 + j) * z_row, cols * 4, fb);
          }
        }
      }
      if (i == 0) STAMP(6);
      if (++slot == g.ring) {
        slot = 0;
        phase ^= 1;
      }
    }
    STAMP(7);
    return;
  }

  // -------------------------------------------------------------------- consumers ---------------------------------------
  const int grp = warp / nw, wl = warp - grp * nw;  // consumer group, column slice [32 wl, 32 wl + 32) of the block
  const int gq = lane >> 2, t = lane & 3;
  const int ch = (gq >> 1) | ((gq & 1) << 2);  // this lane's 16-byte chunk (4 columns) of the warp's 128-byte rows
  const bool active = wl * 32 < cols;
  const int mshift = g.gshift >= 0 && g.gshift < 2 ? g.gshift : 2;  // scale row of a quad inside its stage: (q & 3) >> mshift
  const uint32_t xsum = recv + g.red_bytes - g.xsum_bytes;  // [stage][scale row][8] fp32: sums of x over the quads of a scale row
  asm volatile("griddepcontrol.wait;" ::: "memory");
  STAMP(1);
  {  // stage this slab's activations, every 8-k chunk permuted to the operand order {(k0,k4), (k2,k6), (k1,k5), (k3,k7)}, and sum
     // them per (stage, scale row): 16 consecutive lanes hold the 16 chunks of one stage of one row -> fixed shuffle tree
    const int cpr = nstages * 16;                 // chunks per row, counted from the aligned start qa0
    const int c_lo = (q_lo - qa0) * 4, c_hi = (q_hi - qa0) * 4;
    const int esh = 2 + mshift;                   // chunks per scale row = 1 << esh
    constexpr int XU = 4;  // loads in flight per thread
    for (int base0 = threadIdx.x & ~31; base0 < M * cpr; base0 += XU * nthr) {
      uint4 v[XU];
      int mm[XU], cv[XU];
      bool ok[XU];
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        const int idx = base0 + u * nthr + lane;
        mm[u] = idx / cpr;
        cv[u] = idx - mm[u] * cpr;
        ok[u] = idx < M * cpr && cv[u] >= c_lo && cv[u] < c_hi;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (ok[u]) v[u] = __ldg(reinterpret_cast<const uint4*>(g.x + size_t(mm[u]) * g.K + size_t(qa0) * 32) + cv[u]);
      }
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        if (base0 + u * nthr >= M * cpr) break;  // (warp-uniform)
        float sum = 0.f;
        if (ok[u]) {
          uint4 o;
          o.x = __byte_perm(v[u].x, v[u].z, 0x5410);
          o.y = __byte_perm(v[u].y, v[u].w, 0x5410);
          o.z = __byte_perm(v[u].x, v[u].z, 0x7632);
          o.w = __byte_perm(v[u].y, v[u].w, 0x7632);
          sts128(xs + mm[u] * g.x_pitch + (cv[u] - c_lo) * 16, o);
          const uint32_t r[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r[k]));
            sum += f.x + f.y;
          }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (esh > 2) sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        if (esh > 3) sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        if (base0 + u * nthr + lane < M * cpr && (cv[u] & ((1 << esh) - 1)) == 0)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum + ((cv[u] >> esh) * 8 + mm[u]) * 4), "f"(sum) : "memory");
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  }
  STAMP(2);

  float acc[2][4];
  QuadAcc qa;
  float sa[4], sb[4], X[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[0][i] = acc[1][i] = 0.f;
    sa[i] = sb[i] = 0.f;
    qa.P[0][i] = qa.P[1][i] = qa.Q[0][i] = qa.Q[1][i] = 0.f;
  }
  X[0] = X[1] = 0.f;
  const uint32_t lane_w0 = wl * 2048 + t * 128 + (uint32_t(ch ^ t) << 4);        // quads 0 / 2 of a stage: + 0 / + 1024
  const uint32_t lane_w1 = wl * 2048 + t * 128 + (uint32_t(ch ^ t ^ 4) << 4);    // quads 1 / 3: + 512 / + 1536
  const uint32_t lane_x = xs + (gq < M ? gq : 0) * g.x_pitch + t * 16 - q_lo * 64;  // + quad * 64
  const uint32_t lane_s = s_off + (wl * 32 + ch * 4) * 2, lane_z = z_off + (wl * 32 + ch * 4) * 4;
  const uint32_t lane_xs = xsum + 8 * t;  // + entry * 32: X[m = 2t], X[m = 2t + 1]

  auto set_scales = [&](uint32_t st, uint32_t mrow) {  // scales / zero-points of this lane's four columns
    const uint2 sv = lds64(st + lane_s + mrow * (BN * 2));
    const uint4 zv = lds128(st + lane_z + mrow * (BN * 4));
    const __half2 s01 = *reinterpret_cast<const __half2*>(&sv.x), s23 = *reinterpret_cast<const __half2*>(&sv.y);
    const float sf[4] = {__low2float(s01), __high2float(s01), __low2float(s23), __high2float(s23)};
    const int zi[4] = {int(zv.x), int(zv.y), int(zv.z), int(zv.w)};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sa[c] = sf[c] * 16777216.f;
      sb[c] = -sf[c] * float(zi[c]);
    }
  };
  auto close_group = [&]() {
#pragma unroll
    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = 2 * tl + (i >> 1);
        const float G = fmaf(qa.Q[tl][i], 0.0625f, qa.P[tl][i]);
        acc[tl][i] = fmaf(sa[c], G, fmaf(sb[c], X[i & 1], acc[tl][i]));
      }
  };
  auto quad_addr = [&](uint32_t st, int j) { return st + ((j & 1) ? lane_w1 : lane_w0) + j * 512; };
  auto clear_pq = [&]() {
#pragma unroll
    for (int k = 0; k < 4; ++k) qa.P[0][k] = qa.P[1][k] = qa.Q[0][k] = qa.Q[1][k] = 0.f;
  };

  // stages grp, grp + NGRP, ...: slot = i % ring, phase = (i / ring) & 1
  int slot = grp % g.ring;
  uint32_t phase = (grp / g.ring) & 1;
  auto advance = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
    slot += NGRP;
    if (slot >= g.ring) {
      slot -= g.ring;
      phase ^= 1;
    }
  };
  bool open = false;  // P / Q hold a partial scale group
  int open_sg = 0;
  // any stage, one scale row at a time (first / last stage of an unaligned slab, group sizes 32 / 64)
  auto generic_stage = [&](int i) {
    const int qb = qa0 + 4 * i;
    const int j0 = max(q_lo - qb, 0), j1 = min(q_hi - qb, 4);
    const uint32_t st = stages + slot * g.stage_bytes;
    mbar_wait(bar_full + 8 * slot, phase);
    if (active && !(g.dbg & 1)) {
      const int qrow = 1 << mshift;
      for (int r = 0; r < (4 >> mshift); ++r) {
        const int ja = max(j0, r * qrow), jb = min(j1, (r + 1) * qrow);
        if (ja >= jb) continue;
        const int sg = g.gshift >= 0 ? (qb + ja) >> g.gshift : 0;
        if (open && sg != open_sg) {
          close_group();
          open = false;
        }
        if (!open) {
          set_scales(st, r);
          clear_pq();
          X[0] = X[1] = 0.f;
          open = true;
          open_sg = sg;
        }
        const uint2 xr = lds64(lane_xs + (i * g.gps + r) * 32);
        X[0] += __uint_as_float(xr.x);
        X[1] += __uint_as_float(xr.y);
        for (int j = ja; j < jb; ++j) quad_math2<false>(qa, lds128(quad_addr(st, j)), lds128(lane_x + (qb + j) * 64));
      }
    }
    advance();
  };

  const int i_full0 = q_lo > qa0 ? 1 : 0, i_full1 = (q_hi - qa0) >> 2;  // stages [i_full0, i_full1) lie wholly inside the slab
  int i = grp;
  if (mshift == 2 && !(g.dbg & 1)) {
    if (i < i_full0) {
      generic_stage(i);
      i += NGRP;
    }
    if (g.qpg == 4) {
      // ---- every stage is one scale group (group size 128): no state carried between stages -------------------------------
      if (open) {  // (an unaligned first stage is a group of its own)
        if (active) close_group();
        open = false;
      }
      for (; i < i_full1; i += NGRP) {
        const uint32_t st = stages + slot * g.stage_bytes;
        mbar_wait(bar_full + 8 * slot, phase);
        if (i == grp) STAMP(3);
        if (active) {
          uint4 wv[4], xv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            wv[j] = lds128(quad_addr(st, j));
            xv[j] = lds128(lane_x + (qa0 + 4 * i + j) * 64);
          }
          set_scales(st, 0);
          const uint2 xr = lds64(lane_xs + i * 32);
          X[0] = __uint_as_float(xr.x);
          X[1] = __uint_as_float(xr.y);
          quad_math2<true>(qa, wv[0], xv[0]);
#pragma unroll
          for (int j = 1; j < 4; ++j) quad_math2<false>(qa, wv[j], xv[j]);
          close_group();
        }
        advance();
      }
    } else {
      // ---- scale groups span several stages (group size >= 256, per-channel) -----------------------------------------------
      for (; i < i_full1; i += NGRP) {
        const uint32_t st = stages + slot * g.stage_bytes;
        mbar_wait(bar_full + 8 * slot, phase);
        if (i == grp) STAMP(3);
        if (active) {
          uint4 wv[4], xv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            wv[j] = lds128(quad_addr(st, j));
            xv[j] = lds128(lane_x + (qa0 + 4 * i + j) * 64);
          }
          const int sg = g.gshift >= 0 ? (qa0 + 4 * i) >> g.gshift : 0;
          if (open && sg != open_sg) {
            close_group();
            open = false;
          }
          if (!open) {
            set_scales(st, 0);
            clear_pq();
            X[0] = X[1] = 0.f;
            open = true;
            open_sg = sg;
          }
          const uint2 xr = lds64(lane_xs + i * 32);
          X[0] += __uint_as_float(xr.x);
          X[1] += __uint_as_float(xr.y);
#pragma unroll
          for (int j = 0; j < 4; ++j) quad_math2<false>(qa, wv[j], xv[j]);
        }
        advance();
      }
    }
    for (; i < nstages; i += NGRP) generic_stage(i);  // an unaligned last stage
  } else {
    for (; i < nstages; i += NGRP) generic_stage(i);
  }
  if (open && active) close_group();
  STAMP(4);
  // ---- publish this warp's 32 columns ----------------------------------------------------------------------------------
  // lane (gq, t) holds columns col..col+3 of rows m = 2t (c0/c2) and 2t+1 (c1/c3)
  float4 r0 = make_float4(acc[0][0], acc[0][2], acc[1][0], acc[1][2]);
  float4 r1 = make_float4(acc[0][1], acc[0][3], acc[1][1], acc[1][3]);
  const int m0 = 2 * t, m1 = 2 * t + 1;
  const int col = cb * BN + wl * 32 + ch * 4;
  if (S > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // every CTA of the cluster has its barriers up
  // every (rank, consumer group) sends its partial to the CTA that reduces this column slice: rank wl % S, buffer wl / S
  const uint32_t dest = uint32_t(wl % S), wslot = uint32_t(wl / S);
  const int nsrc = S * NGRP;
  if (active) {
    const uint32_t lbuf = recv + ((wslot * nsrc + rank * NGRP + grp) * M) * 128 + ch * 16;  // [wslot][source][m][32 columns] fp32
    if (S > 1) {
      const uint32_t rbuf = mapa(lbuf, dest), rbar = mapa(bar_red + 8 * wslot, dest);
      if (m0 < M) st_async_v4(rbuf + m0 * 128, r0, rbar);
      if (m1 < M) st_async_v4(rbuf + m1 * 128, r1, rbar);
    } else {
      if (m0 < M) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lbuf + m0 * 128), "f"(r0.x), "f"(r0.y), "f"(r0.z), "f"(r0.w) : "memory");
      if (m1 < M) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lbuf + m1 * 128), "f"(r1.x), "f"(r1.y), "f"(r1.z), "f"(r1.w) : "memory");
    }
  }
  if (S == 1) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  if (active && grp == 0 && dest == uint32_t(rank)) {
    if (S > 1) mbar_wait(bar_red + 8 * wslot, 0);
    r0 = r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int src = 0; src < nsrc; ++src) {
      const uint32_t a = recv + ((wslot * nsrc + src) * M) * 128 + ch * 16;
      if (m0 < M) {
        const float4 v = lds128f(a + m0 * 128);
        r0.x += v.x; r0.y += v.y; r0.z += v.z; r0.w += v.w;
      }
      if (m1 < M) {
        const float4 v = lds128f(a + m1 * 128);
        r1.x += v.x; r1.y += v.y; r1.z += v.z; r1.w += v.w;
      }
    }
    if (m0 < M) {
      const __half2 lo = __floats2half2_rn(r0.x, r0.y), hi = __floats2half2_rn(r0.z, r0.w);
      uint2 pk = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      *reinterpret_cast<uint2*>(g.out + size_t(m0) * N + col) = pk;
    }
    if (m1 < M) {
      const __half2 lo = __floats2half2_rn(r1.x, r1.y), hi = __floats2half2_rn(r1.z, r1.w);
      uint2 pk = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      *reinterpret_cast<uint2*>(g.out + size_t(m1) * N + col) = pk;
    }
  }
  STAMP(5);
}

// L2 prefetch of the head of a GEMV's weights: the first `rows` packed rows of each of the S K-slabs (contiguous, rows * 4N bytes each).
// Launched with programmatic dependent launch BEFORE the GEMV it serves: it runs while the previous GEMV is still computing, and
// only completes once that GEMV has (griddepcontrol.wait), so the stream order seen by the next kernel is unchanged.
__global__ void __launch_bounds__(128) l2_prefetch(const uint32_t* wq, int K8, int N, int S, int rows) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const size_t slab_bytes = size_t(rows) * N * 4;
  const int chunks_per_slab = int((slab_bytes + 4095) / 4096);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks_per_slab * S; c += gridDim.x * blockDim.x) {
    const int r = c / chunks_per_slab, cc = c - r * chunks_per_slab;
    const char* base = reinterpret_cast<const char*>(wq + size_t((long long)r * (K8 / 4) / S * 4) * N);
    const size_t off = size_t(cc) * 4096;
    const uint32_t bytes = uint32_t(min(size_t(4096), slab_bytes - off));
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(bytes) : "memory");
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------------------------------------
// reference: out[m][n] = sum_k x[m][k] * fp16((q - z) * s), fp32 accumulation
__global__ void ref_gemv(const uint32_t* wq, const __half* s, const int* zp, const __half* x, float* out, int M, int K, int N, int gs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N) return;
  float acc = 0.f;
  for (int r = 0; r < K / 8; ++r) {
    const uint32_t w = wq[size_t(r) * N + n];
    const int grp = (r * 8) / gs;
    const __half sc = s[size_t(grp) * N + n];
    const int z = zp[size_t(grp) * N + n];
    for (int i = 0; i < 8; ++i) {
      const int qv = (w >> (4 * i)) & 15;
      const __half wv = __hmul(__int2half_rn(qv - z), sc);
      acc += __half2float(x[size_t(m) * K + r * 8 + i]) * __half2float(wv);
    }
  }
  out[size_t(m) * N + n] = acc;
}

__global__ void fill_u32(uint32_t* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned v = unsigned(i) * 2654435761u + seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; v *= 3266489917u; v ^= v >> 16;
    p[i] = v;
  }
}
__global__ void fill_zp(int* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned v = unsigned(i) * 2654435761u + seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
    p[i] = int(v & 15u);
  }
}
__global__ void fill_half(__half* p, size_t n, unsigned seed, float lo, float hi) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned v = unsigned(i) * 2654435761u + seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; v *= 3266489917u; v ^= v >> 16;
    p[i] = __float2half(lo + (hi - lo) * float(v >> 8) * (1.f / 16777216.f));
  }
}

struct Layer {
  int K, N;
  uint32_t* wq;
  __half* s;
  int* zp;
  __half* x;
  __half* out;
  CUtensorMap tmw;
  int nw, S;
};

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
  return reinterpret_cast<EncodeTiledFn>(sym);
}
static void make_tmap(Layer& l) {
  static EncodeTiledFn enc = encode_fn();
  const cuuint64_t dims[3] = {32, cuuint64_t(l.K / 8), cuuint64_t(l.N / 32)};
  const cuuint64_t strides[2] = {cuuint64_t(l.N) * 4, 128};
  const cuuint32_t box[3] = {32, 16, cuuint32_t(l.nw)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(&l.tmw, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, l.wq, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", int(r)); exit(1); }
}

static int g_sms = 148;
static int g_ring_kb = 200;
static int g_nw_override = 0, g_s_override = 0, g_ngrp_override = 0;
static long long* g_trace = nullptr;
static int g_pf_mb = 0;
static int g_dbg = 0;

// column-block width (consumer warps) and K slabs: as many CTAs as fit one wave of clusters
static void plan(int K, int N, int* nw_out, int* s_out) {
  const int cap[9] = {0, 148, 148, 0, 132, 0, 0, 0, 128};
  double best = -1;
  for (int nw = MAXNW; nw >= 4; --nw)
    for (int S : {1, 2, 4, 8}) {
      if (S == 8 && !g_s_override) continue;
      if (g_nw_override && nw != g_nw_override) continue;
      if (g_s_override && S != g_s_override) continue;
      const int ncb = (N + 32 * nw - 1) / (32 * nw);
      const int ctas = ncb * S;
      if ((ctas > cap[S] && !(g_nw_override || g_s_override)) || K / 32 / S < 4) continue;
      const double eff = double(N) * S / (32 * nw);  // SMs' worth of evenly loaded CTAs
      if (eff > best + 1e-9) { best = eff; *nw_out = nw; *s_out = S; }
    }
}

static Geo make_geo(const Layer& L, int M, int gs, size_t* smem, dim3* grid) {
  Geo g{};
  g.s = L.s; g.zp = L.zp; g.x = L.x; g.out = L.out; g.trace = g_trace;
  g.M = M; g.K = L.K; g.N = L.N;
  g.QPB = L.K / 32;
  g.nw = L.nw; g.S = L.S; g.ngrp = g_ngrp_override ? g_ngrp_override : std::max(1, 16 / L.nw);
  const int NGRP = g.ngrp;
  const int BN = 32 * g.nw;
  g.qpg = gs / 32;
  g.gshift = -1;
  if (gs < L.K) for (int sh = 0; sh < 20; ++sh) if ((1 << sh) == g.qpg) g.gshift = sh;
  if (g.gshift < 0) g.qpg = g.QPB;
  g.gps = g.gshift >= 0 && g.qpg < 4 ? 4 / g.qpg : 1;
  g.ngroups = L.K / gs;
  g.w_bytes = 2048 * g.nw;
  g.stage_bytes = ((g.w_bytes + g.gps * BN * 6 + 1023) / 1024) * 1024;
  const int nxq = (g.QPB + g.S - 1) / g.S;
  g.x_pitch = nxq * 64 + ((nxq * 64) % 128 == 0 ? 64 : 0);
  g.x_bytes = M * g.x_pitch;
  g.xsum_bytes = ((nxq + 3) / 4 + 2) * g.gps * 32;
  g.red_bytes = ((g.nw + g.S - 1) / g.S) * g.S * NGRP * M * 128 + g.xsum_bytes;
  const int nstages = (nxq + 3) / 4 + 1;
  int ring = (g_ring_kb * 1024 - HDR - g.x_bytes - g.red_bytes - 1024) / g.stage_bytes;
  ring = std::max(2, std::min(std::min(ring, nstages), MAXRING));
  g.ring = ring;
  g.dbg = g_dbg;
  *smem = size_t(HDR) + g.x_bytes + g.red_bytes + 1024 + size_t(ring) * g.stage_bytes;
  *grid = dim3(g.S, (L.N + BN - 1) / BN);
  return g;
}

static void launch(const Layer& L, int M, int gs, bool pdl, cudaStream_t st) {
  if (g_pf_mb > 0) {
    const size_t total = size_t(L.K / 8) * L.N * 4;
    const size_t want = std::min(total, size_t(g_pf_mb) << 20);
    int rows = int(want / L.S / (size_t(L.N) * 4));
    rows = std::min(rows, L.K / 8 / L.S);
    if (rows > 0) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(128);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl ? 1 : 0;
      CK(cudaLaunchKernelEx(&cfg, l2_prefetch, (const uint32_t*)L.wq, L.K / 8, L.N, L.S, rows));
    }
  }
  size_t smem;
  dim3 grid;
  Geo g = make_geo(L, M, gs, &smem, &grid);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(32 * (g.ngrp * g.nw + 1));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g.S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = g.S;
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
  CK(cudaLaunchKernelEx(&cfg, ring_gemv<1>, L.tmw, g));
}

int main(int argc, char** argv) {
  int M = 1, gs = 128, nsets = 8, layers = 32;
  bool do_trace = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--m")) M = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--ring-kb")) g_ring_kb = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--nw")) g_nw_override = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--s")) g_s_override = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--ngrp")) g_ngrp_override = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--pf")) g_pf_mb = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--sets")) nsets = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--dbg")) g_dbg = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--trace")) do_trace = true;
    else if (!strcmp(argv[i], "--gs")) gs = atoi(argv[++i]);
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device %s, %d SMs; ring_kb=%d M=%d gs=%d dbg=%d\n", prop.name, g_sms, g_ring_kb, M, gs, g_dbg);
  CK(cudaFuncSetAttribute(ring_gemv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, ring_gemv<1>));
  printf("ring_gemv<1>: %d regs, %zu B static smem\n", fa.numRegs, fa.sharedSizeBytes);

  CK(cudaFuncSetAttribute(ring_gemv<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const char* names[4] = {"qkv", "o", "gate_up", "down"};
  const int Ks[4] = {4096, 4096, 4096, 14336}, Ns[4] = {6144, 4096, 28672, 4096};
  std::vector<Layer> L(size_t(nsets) * 4);
  unsigned seed = 1;
  for (int sI = 0; sI < nsets; ++sI)
    for (int p = 0; p < 4; ++p) {
      Layer& l = L[sI * 4 + p];
      l.K = Ks[p]; l.N = Ns[p];
      const size_t nw = size_t(l.K / 8) * l.N, ng = size_t(l.K / gs) * l.N;
      CK(cudaMalloc(&l.wq, nw * 4)); CK(cudaMalloc(&l.s, ng * 2)); CK(cudaMalloc(&l.zp, ng * 4));
      CK(cudaMalloc(&l.x, size_t(8) * l.K * 2)); CK(cudaMalloc(&l.out, size_t(8) * l.N * 2));
      fill_u32<<<1184, 256>>>(l.wq, nw, seed++);
      fill_half<<<1184, 256>>>(l.s, ng, seed++, 0.005f, 0.02f);
      fill_zp<<<1184, 256>>>(l.zp, ng, seed++);
      fill_half<<<64, 256>>>(l.x, size_t(8) * l.K, seed++, -1.f, 1.f);
      CK(cudaMemset(l.out, 0, size_t(8) * l.N * 2));
      plan(l.K, l.N, &l.nw, &l.S);
      make_tmap(l);
    }
  for (int p = 0; p < 4; ++p) printf("plan %-8s nw=%d S=%d grid=%d CTAs\n", names[p], L[p].nw, L[p].S, L[p].S * ((L[p].N + 32 * L[p].nw - 1) / (32 * L[p].nw)));
  CK(cudaDeviceSynchronize());
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  // ---- correctness -----------------------------------------------------------------------------------------------------
  if (!g_dbg) {
    for (int p = 0; p < 4; ++p) {
      const Layer& l = L[p];
      float* ref;
      CK(cudaMalloc(&ref, size_t(M) * l.N * 4));
      ref_gemv<<<dim3((l.N + 127) / 128, M), 128, 0, st>>>(l.wq, l.s, l.zp, l.x, ref, M, l.K, l.N, gs);
      launch(l, M, gs, false, st);
      CK(cudaStreamSynchronize(st));
      std::vector<float> hr(size_t(M) * l.N);
      std::vector<__half> ho(size_t(M) * l.N);
      CK(cudaMemcpy(hr.data(), ref, hr.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ho.data(), l.out, ho.size() * 2, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxref = 0;
      size_t bad = 0;
      for (size_t i = 0; i < hr.size(); ++i) {
        const double d = fabs(double(__half2float(ho[i])) - hr[i]);
        maxerr = std::max(maxerr, d);
        maxref = std::max(maxref, fabs(double(hr[i])));
        if (d > 1e-2 + 1e-2 * fabs(hr[i])) ++bad;
      }
      printf("check %-8s M=%d: max |err| %.3e (max |ref| %.3f), %zu outside 1e-2/1e-2\n", names[p], M, maxerr, maxref, bad);
      CK(cudaFree(ref));
    }
  }

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time_graph = [&](auto body, int reps) {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    body();
    CK(cudaStreamEndCapture(st, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(exec, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(exec, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGraphExecDestroy(exec));
    CK(cudaGraphDestroy(graph));
    return double(ms) * 1e3 / reps;
  };
  const double peak = 6545.0;
  // ---- per shape: chain of the same projection over the weight sets -----------------------------------------------------------
  for (int pdl = 0; pdl < 2; ++pdl)
    for (int p = 0; p < 4; ++p) {
      const int chain = nsets * 4;
      const double us = time_graph([&] { for (int i = 0; i < chain; ++i) launch(L[(i % nsets) * 4 + p], M, gs, pdl, st); }, 10) / chain;
      const double bytes = double(Ks[p]) * Ns[p] / 2 + double(M) * Ks[p] * 2;
      printf("%-8s M=%d pdl=%d: %7.2f us  %7.1f GB/s  %5.1f%% of %.0f\n", names[p], M, pdl, us, bytes / us / 1e3, 100 * bytes / us / 1e3 / peak, peak);
    }
  // ---- the bench step: 32 layers x (qkv, o, gate_up, down) ----------------------------------------------------------------------
  for (int pdl = 0; pdl < 2; ++pdl) {
    const double us = time_graph([&] { for (int l = 0; l < layers; ++l) for (int p = 0; p < 4; ++p) launch(L[(l % nsets) * 4 + p], M, gs, pdl, st); }, 10);
    double bytes = 0;
    for (int p = 0; p < 4; ++p) bytes += double(Ks[p]) * Ns[p] / 2 + double(M) * Ks[p] * 2;
    bytes *= layers;
    printf("step (%d layers x 4) M=%d pdl=%d: %8.1f us  %7.1f GB/s  %5.1f%% of %.0f  (%.2f us per launch)\n", layers, M, pdl, us, bytes / us / 1e3,
           100 * bytes / us / 1e3 / peak, peak, us / (layers * 4));
  }
#ifdef TRACE
  if (do_trace) {
    const int nk = 12;
    CK(cudaMalloc(&g_trace, size_t(nk) * 256 * 8 * 8));
    CK(cudaMemset(g_trace, 0, size_t(nk) * 256 * 8 * 8));
    long long* base = g_trace;
    for (int i = 0; i < 8; ++i) launch(L[i], M, gs, true, st);  // warm
    for (int i = 0; i < nk; ++i) {
      g_trace = base + size_t(i) * 256 * 8;
      launch(L[8 + i], M, gs, true, st);
    }
    CK(cudaStreamSynchronize(st));
    std::vector<long long> h(size_t(nk) * 256 * 8);
    CK(cudaMemcpy(h.data(), base, h.size() * 8, cudaMemcpyDeviceToHost));
    long long t0 = 1LL << 62;
    for (auto v : h) if (v > 0) t0 = std::min(t0, v);
    const char* sn[8] = {"entry", "pdl-wait", "x-ready", "stage0", "loop-end", "done", "p-first", "p-last"};
    printf("trace (us since first stamp; min / median / max over CTAs)\n%-10s", "kernel");
    for (int s = 0; s < 8; ++s) printf(" %22s", sn[s]);
    printf("\n");
    for (int i = 0; i < nk; ++i) {
      printf("%-10s", names[(8 + i) % 4]);
      for (int s = 0; s < 8; ++s) {
        std::vector<double> v;
        for (int c = 0; c < 256; ++c) { long long x = h[(size_t(i) * 256 + c) * 8 + s]; if (x > 0) v.push_back((x - t0) * 1e-3); }
        std::sort(v.begin(), v.end());
        if (v.empty()) printf(" %22s", "-");
        else printf("  %6.1f/%6.1f/%6.1f", v.front(), v[v.size() / 2], v.back());
      }
      printf("\n");
    }
  }
#endif
  (void)do_trace;
  return 0;
}
