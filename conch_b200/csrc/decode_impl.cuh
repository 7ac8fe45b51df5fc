// Decode kernel template + its launch ladder, shared by the per-(activation dtype, weight width) translation units
// decode_f16_w4.cu / decode_f16_w8.cu / decode_bf16_w4.cu / decode_bf16_w8.cu (split so that they compile in parallel).
// The design notes are at the top of decode_kernel.cu, which holds the host-side plan and dispatch.
#pragma once
#include "common.cuh"
#include "dequant.cuh"
#include "mma_ptx.cuh"

namespace cb200 {

#ifdef CB200_TC_TRACE  // debug builds only (benchmarks/decode_trace.py): globaltimer stamps of warp 0 of every CTA
#define DEC_STAMP(slot) dec_stamp(geo.trace, slot)
__device__ __forceinline__ void dec_stamp(long long* trace, int slot) {
  if (trace && (threadIdx.x & 127) == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 8 + slot] = t;
    if (slot == 0) {
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      trace[blockIdx.x * 8 + 7] = sm;
    }
  }
}
#else
#define DEC_STAMP(slot) do { } while (0)
#endif

struct DecodeGeom {
  int m_rows;         // rows of x in this pass (<= 16)
  int splits;         // S: warps sharing one column block (split-K)
  int units_per_warp; // batches per warp = K / (U * STEP_K * S)
  int total_warps;    // (N / 32) * S
  int bpg;            // batches per scale group (g / (U * STEP_K))
  int pdl;            // launched with programmatic stream serialization
  int x_row_bytes;    // per-CTA smem: bytes per staged x row (slab bytes + 64 B pad against bank conflicts)
  int meta_row_bytes; // per-warp smem: bytes per group row (64 B of scales + zero-points)
  int ngroups;        // scale groups touched by one warp
  int warp_smem;      // per-warp smem bytes (group rows + mbarrier)
  int cta_x_bytes;    // per-CTA smem bytes of the shared x slab (the 4 warps of a CTA work on the same K slab)
  int cb_padded;      // column blocks rounded up to a multiple of 4 (whole CTAs per slab; padding warps only sync)
  int cta_aux_bytes;  // per-CTA smem after the x slab: chunk sums [m_rows][slab_k / 8] + group sums [ngroups][8] (fp32)
  int scalar_push;    // debug (cb200_set_decode_tuning vec_words = 1): element stores to the peers instead of 8-byte vectors
#ifdef CB200_TC_TRACE
  long long* trace;   // [CTAs][8] globaltimer stamps (debug builds only)
#endif
};

// A: activation type, BITS: 4|8, MT: 8-row m tiles (1: M<=8, 2: M<=16), U: steps per batch (group boundaries fall on
// batch boundaries), ZK: cb200_zp_kind (compile time), FUSED: feed magic-number operands to the MMA (see above).
template <class A, int BITS, int MT, int U, int ZK, bool FUSED, bool XSUM = false>
__global__ void __launch_bounds__(128, MT == 2 ? 3 : (FUSED ? 4 : 5))
    decode_kernel(const Problem p, const DecodeGeom geo, float* __restrict__ ws_partial, unsigned int* __restrict__ ws_counter) {
  constexpr int PF = 32 / BITS;     // k per packed word
  constexpr int VEC = 4;            // packed words per lane per load (128-bit)
  constexpr int NT = VEC / 2;       // 16-column MMA tiles per warp
  constexpr int CW = 8 * VEC;       // columns per warp
  constexpr int PAIRS = 16 / BITS;  // pair registers per word
  constexpr int WIN = 2 * U;        // weight loads in flight per lane (steps)
  constexpr Order ORD = Order::Lop;
  constexpr bool FZP = ZK == CB200_ZP_ACT;

  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);  // (the grid is whole CTAs: total_warps is a multiple of 4)
  DEC_STAMP(0);
  const int g8 = lane >> 2, tig = lane & 3;
  const int M = geo.m_rows;
  const int N = p.N;
  const int S = geo.splits;
  // neighbouring warps take neighbouring column blocks of the SAME K slab: the four warps of a CTA read 512 contiguous
  // bytes of every packed row between them and share ONE staged copy of the slab's activations
  const int C = N / CW;               // column blocks
  const int si = w / geo.cb_padded;   // slab index inside the column block (the same for the whole CTA)
  const int cb = w - si * geo.cb_padded;  // column block
  if (cb >= C) {                      // padding warp (N / 32 not a multiple of 4): only takes part in the CTA barrier
    __syncthreads();
    return;
  }
  const int nunits = geo.units_per_warp;
  const int nsteps = nunits * U;
  const int b0 = si * nunits;         // first batch of this warp
  const int bpg = geo.bpg;
  const int col = cb * CW + VEC * g8; // first of this lane's four weight columns
  const bool zp_per_group = ZK != CB200_ZP_NONE && !p.zp_scalar;

  // ---- weight window: WIN loads in flight before anything else (weights are static: legal before griddepcontrol.wait) ----
  const char* wp = reinterpret_cast<const char*>(p.wq + (size_t(b0) * (U * 4) + tig) * N + col);
  const size_t step_bytes = size_t(16) * N;  // 4 packed rows
  uint4 win[WIN];
#pragma unroll
  for (int j = 0; j < WIN; ++j) {
    win[j] = make_uint4(0u, 0u, 0u, 0u);
    if (j < nsteps) win[j] = ldg_stream_v4(wp + j * step_bytes);
  }
  wp += WIN * step_bytes;  // next load of slot j: wp + j * step_bytes
  // Programmatic dependent launch: release the NEXT kernel now.  Its CTAs become resident as ours retire, issue their
  // own weight window and metadata copies, and then block in griddepcontrol.wait until this grid has completed -- so the
  // next kernel's head (launch, prologue, first DRAM round trip) hides behind this kernel's tail.
  if (geo.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  DEC_STAMP(2);

  // ---- shared memory: ONE x slab [m_rows][x_row_bytes] per CTA, then per warp: group rows [ngroups][meta_row_bytes] + mbarrier,
  //      then the CTA's x barrier -----------------------------------------------------------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t xs_base = smem0;
  const uint32_t aux_base = smem0 + geo.cta_x_bytes;                       // chunk sums, then group sums (XS path)
  const uint32_t meta_base = aux_base + geo.cta_aux_bytes + (threadIdx.x >> 5) * geo.warp_smem;
  const uint32_t bar = meta_base + geo.ngroups * geo.meta_row_bytes;
  const uint32_t xbar = aux_base + geo.cta_aux_bytes + 4 * geo.warp_smem;
  const int g0 = b0 / bpg;  // a slab either covers whole groups or lies inside one group
  const int slab_k = nsteps * 4 * PF;
  constexpr uint32_t ZROW = ZK == CB200_ZP_I32 ? 128u : ZK == CB200_ZP_U8 ? 32u : ZK == CB200_ZP_ACT ? 64u : 0u;  // 32 columns
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t total = uint32_t(geo.ngroups) * (64u + (zp_per_group ? ZROW : 0u));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xbar), "r"(uint32_t(M) * uint32_t(slab_k * 2))
                   : "memory");
    }
  }
  __syncthreads();  // the x barrier is initialised for everybody (the only CTA-wide synchronisation of the kernel)
  DEC_STAMP(1);

  auto bulk = [&](uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
  };
  // scales / zero-points of this warp's 32 columns, one row per group (static data: legal before griddepcontrol.wait)
  for (int g = lane; g < geo.ngroups; g += 32) {
    const size_t idx = size_t(g0 + g) * N + cb * CW;
    bulk(meta_base + g * geo.meta_row_bytes, reinterpret_cast<const A*>(p.s) + idx, 64u, bar);
    if (zp_per_group) {
      const size_t zes = ZK == CB200_ZP_I32 ? 4 : ZK == CB200_ZP_U8 ? 1 : 2;
      bulk(meta_base + g * geo.meta_row_bytes + 64, reinterpret_cast<const char*>(p.zp) + idx * zes, ZROW, bar);
    }
  }

  DEC_STAMP(3);
  if (geo.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");  // x / out / workspace belong to the stream order

  // activations: row m of x, columns [b0 * U * STEP_K, + slab_k), one bulk copy per row, ONCE per CTA (warp 0 always exists)
  if (threadIdx.x < M) {
    bulk(xs_base + lane * geo.x_row_bytes, reinterpret_cast<const A*>(p.x) + size_t(lane) * p.K + size_t(b0) * (U * 4 * PF),
         uint32_t(slab_k * 2), xbar);
  }
  bool m_ok[MT];
  uint32_t lane_x[MT];  // + step * (4 * PF * 2)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int m = mt * 8 + g8;
    m_ok[mt] = m < M;
    lane_x[mt] = xs_base + (m_ok[mt] ? m : 0) * geo.x_row_bytes + tig * (PF * 2);
  }
  uint32_t meta_cur = meta_base;  // row of the current group
  // wait for the staged operands (they landed long before the first weight load returns)
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const uint32_t mbar = which ? xbar : bar;
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(mbar)
                   : "memory");
    }
  }

  DEC_STAMP(6);
  constexpr bool HALF = std::is_same<A, __half>::value;
  // fp16 x int4: nibbles at bits 4-7 arrive as 16 q.  They accumulate in their OWN accumulator (accq), which is scaled by 1/16
  // in fp32 when the group closes (round 1 pre-multiplied their activations by 1/16 in fp16 instead, which lost bits for
  // |x| < 2^-10 -- ADVICE r1)
  constexpr bool SPLITQ = FUSED && HALF && BITS == 4;
  // XS: the group sums X[g][m] = sum_k x[m,k] are computed ONCE per CTA from the staged slab (deterministic two-pass
  // reduction below) instead of by an all-ones MMA per operand pair in every warp, and (int4) the slab is rewritten in the
  // operand order the main loop wants -- pairs (k0,k4),(k1,k5),(k2,k6),(k3,k7), the x16 partners pre-scaled by 1/16 -- so a
  // step reads its activations with one ld.shared.v4 and nothing else.  4 instead of 6 MMAs per step.
  // (XSUM is chosen by the host for M <= 2 and long slabs: the CTA-wide prologue costs two barriers and a pass over x,
  //  measured +0.5 us on the 2-batch slabs of the 4096 x 4096 layer and slower for M = 8, -1.2 us on the 58.7 MB layer)
  constexpr bool XS = FUSED && MT == 1 && XSUM;
  const uint32_t xsum_base = aux_base + uint32_t(M) * uint32_t(slab_k / 8) * 4u;  // [ngroups][8] fp32 after the chunk sums
  if constexpr (XS) {
    const int nchunks = slab_k / 8;  // 16-byte chunks (8 activations) per row
    // threads taking part: the CTA's padding warps (column block >= C, always the last ones) have already left
    const int nthr = 32 * min(4, C - (cb - int(threadIdx.x >> 5)));
    auto sum2 = [](uint32_t v) -> float {
      if constexpr (HALF) {
        const float2 f = __half22float2(bits_h2(v));
        return f.x + f.y;
      } else {
        const float2 f = __bfloat1622float2(bits_b2(v));
        return f.x + f.y;
      }
    };
    for (int i = threadIdx.x; i < geo.ngroups * 8; i += nthr) {  // rows >= M stay zero (their outputs are never stored)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum_base + 4u * i), "f"(0.f) : "memory");
    }
    for (int m = 0; m < M; ++m) {
      for (int cc = threadIdx.x; cc < nchunks; cc += nthr) {
        const uint32_t addr = xs_base + m * geo.x_row_bytes + cc * 16;
        const uint4 v = lds128(addr);
        const float cs = (sum2(v.x) + sum2(v.y)) + (sum2(v.z) + sum2(v.w));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(aux_base + 4u * (m * nchunks + cc)), "f"(cs) : "memory");
        if constexpr (BITS == 4) {
          uint32_t o0 = __byte_perm(v.x, v.z, 0x5410);  // (k0,k4)
          uint32_t o1 = __byte_perm(v.x, v.z, 0x7632);  // (k1,k5)
          uint32_t o2 = __byte_perm(v.y, v.w, 0x5410);  // (k2,k6)
          uint32_t o3 = __byte_perm(v.y, v.w, 0x7632);  // (k3,k7)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
      }
    }
    __syncthreads();
    // second pass: one (row, group) per thread, chunk sums added in order
    const int cpg = geo.ngroups > 1 ? (bpg * U * 4 * PF) / 8 : nchunks;  // chunks per group inside this slab
    for (int t = threadIdx.x; t < M * geo.ngroups; t += nthr) {
      const int m = t / geo.ngroups, gi = t - m * geo.ngroups;
      float acc_x = 0.f;
      const uint32_t src = aux_base + 4u * (m * nchunks + gi * cpg);
      for (int c = 0; c < cpg; ++c) {
        float f;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(src + 4u * c));
        acc_x += f;
      }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum_base + 4u * (gi * 8 + m)), "f"(acc_x) : "memory");
    }
    __syncthreads();
  }
  uint32_t xsum_cur = xsum_base;  // XS: group sums of the current group
  // ---- consume-side state ---------------------------------------------------------------------------------------------------
  float zscalar = 0.f;
  A zfscalar = A(0.f);
  if (ZK != CB200_ZP_NONE && p.zp_scalar) {  // one zero-point for the whole matrix: fetch it once
    if constexpr (ZK == CB200_ZP_I32) zscalar = float(__ldg(reinterpret_cast<const int32_t*>(p.zp)));
    if constexpr (ZK == CB200_ZP_U8) zscalar = float(__ldg(reinterpret_cast<const uint8_t*>(p.zp)));
    if constexpr (ZK == CB200_ZP_ACT) zfscalar = *reinterpret_cast<const A*>(p.zp);
  }
  int batches_left = bpg - (b0 % bpg);  // batches until the current group ends
  bool refresh = true;
  const float fbias = float(p.bias);
  constexpr float MAGIC = HALF ? 0.f : 128.f;           // bf16 operands are 128 + q; fp16 operands are q * 2^-24
  constexpr float OPSCALE = HALF ? 16777216.f : 1.f;    // 2^24 undoes the subnormal scaling
  constexpr uint32_t ONE2 = HALF ? 0x3C003C00u : 0x3F803F80u;      // (1.0, 1.0)

  ColQ cq[VEC];          // !FUSED: per-column dequant constants
  float sa[VEC];         // FUSED: s[n] * 2^24 (fp16) | s[n] (bf16): multiplies G = sum_k x * operand
  float sb[VEC];         // FUSED: -s[n] * zb (fp16) | -s[n] * (128 + zb) (bf16): multiplies X = sum_k x

  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  float acc[NT][MT][4];   // running (scaled) sum of this warp's slab
  float accg[NT][MT][4];  // FUSED: G of the current group (exact products, fp32 sum)
  float accq[SPLITQ ? NT : 1][MT][4];  // SPLITQ: the x16 operands' share of G
  float accx[MT][4];      // FUSED: X[m] = sum_k x[m,k] of the current group, in the accumulator layout
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[t][mt][i] = accg[t][mt][i] = accx[mt][i] = 0.f;
        if constexpr (SPLITQ) accq[t][mt][i] = 0.f;
      }
  bool group_open = false;  // FUSED: accg holds a partial group

  auto close_group = [&]() {  // FUSED: acc += sa[n] * G[m,n] + sb[n] * X[m]
    // C fragment: c0,c1 <- weight column 2t (m = 2tig, 2tig+1); c2,c3 <- weight column 2t+1 of this lane;
    // the ones-MMA leaves X[2tig], X[2tig+1] in c0,c1 (and again in c2,c3); the XS path reads them from shared memory
    if constexpr (XS) {
      float x0, x1;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x0), "=f"(x1) : "r"(xsum_cur + 8u * tig));
      accx[0][0] = accx[0][2] = x0;
      accx[0][1] = accx[0][3] = x1;
      xsum_cur += 32u;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float off = fmaf(sb[2 * t + (i >> 1)], accx[mt][i], acc[t][mt][i]);
          float G = accg[t][mt][i];
          if constexpr (SPLITQ) G = fmaf(accq[t][mt][i], 0.0625f, G);
          acc[t][mt][i] = fmaf(sa[2 * t + (i >> 1)], G, off);
        }
  };

  // ---- one batch: U steps out of window slots [W0, W0 + U); `step` is the index of its first step ------------------------------
  auto process = [&](auto w0_tag, int step) {
    constexpr int W0 = decltype(w0_tag)::value;
    if (refresh) {  // group boundary (warp-uniform): this group's scales / zero-points from the staged rows
      A s4[VEC];
      *reinterpret_cast<uint2*>(s4) = lds64(meta_cur + g8 * 8);
      float z[VEC];
      A zf[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        z[j] = zscalar;
        zf[j] = zfscalar;
      }
      if (zp_per_group) {
        if constexpr (ZK == CB200_ZP_I32) {
          const uint4 zr = lds128(meta_cur + 64 + g8 * 16);
          z[0] = float(int(zr.x)); z[1] = float(int(zr.y)); z[2] = float(int(zr.z)); z[3] = float(int(zr.w));
        } else if constexpr (ZK == CB200_ZP_U8) {
          const uint32_t zr = lds32(meta_cur + 64 + g8 * 4);
#pragma unroll
          for (int j = 0; j < VEC; ++j) z[j] = float((zr >> (8 * j)) & 0xffu);
        } else if constexpr (ZK == CB200_ZP_ACT) {
          *reinterpret_cast<uint2*>(zf) = lds64(meta_cur + 64 + g8 * 8);
        }
      }
      meta_cur += geo.meta_row_bytes;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if constexpr (FUSED) {
          const float sj = ActOps<A>::to_float(s4[j]);
          const float zbj = (FZP ? ActOps<A>::to_float(zf[j]) : z[j]) + fbias;
          sa[j] = sj * OPSCALE;
          sb[j] = -sj * (MAGIC + zbj);
        } else {
          cq[j] = make_colq<A, BITS, ORD, FZP>(s4[j], z[j] + fbias, zf[j]);
        }
      }
      refresh = false;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // this step's weights leave the window; its slot is re-loaded with the step WIN ahead right away
      const uint4 wv = win[W0 + u];
      if (step + u + WIN < nsteps) win[W0 + u] = ldg_stream_v4(wp + (W0 + u) * step_bytes);
      // activations of this step from the staged slab, permuted to the lop3 nibble order
      uint32_t xb[MT][PAIRS];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if constexpr (BITS == 4) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (m_ok[mt]) v = lds128(lane_x[mt] + (step + u) * (4 * PF * 2));
          if constexpr (XS) {  // the slab was rewritten in operand order (and pre-scaled) by the CTA prologue
            xb[mt][0] = v.x; xb[mt][1] = v.y; xb[mt][2] = v.z; xb[mt][3] = v.w;
          } else {
            xb[mt][0] = __byte_perm(v.x, v.z, 0x5410);  // (k0,k4)
            xb[mt][1] = __byte_perm(v.x, v.z, 0x7632);  // (k1,k5)
            xb[mt][2] = __byte_perm(v.y, v.w, 0x5410);  // (k2,k6)
            xb[mt][3] = __byte_perm(v.y, v.w, 0x7632);  // (k3,k7)
          }
        } else {
          uint2 v = make_uint2(0u, 0u);
          if (m_ok[mt]) v = lds64(lane_x[mt] + (step + u) * (4 * PF * 2));
          xb[mt][0] = v.x;
          xb[mt][1] = v.y;
        }
      }
      const uint32_t words[VEC] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t d[VEC][PAIRS];
      const bool fresh = !group_open && u == 0;  // first MMA of a group starts from zero instead of clearing registers
      if constexpr (FUSED) {
        // operands without arithmetic: fp16 0x0000|v = v * 2^-24 (v = q, or 16 q for nibbles at bits 4-7), bf16 0x4300|q
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if constexpr (BITS == 4 && HALF) {
            const uint32_t t8 = words[j] >> 8;
            d[j][0] = words[j] & 0x000f000fu;  // (k0   , k4)
            d[j][1] = words[j] & 0x00f000f0u;  // (16 k1, 16 k5)
            d[j][2] = t8 & 0x000f000fu;        // (k2, k6)
            d[j][3] = t8 & 0x00f000f0u;        // (16 k3, 16 k7)
          } else if constexpr (BITS == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) d[j][i] = lop3_and_or(words[j] >> (4 * i), 0x000f000fu, 0x43004300u);  // (128+k_i, 128+k_i+4)
          } else {
            d[j][0] = prmt(words[j], 0u, 0x4140u);  // (k0, k1) * 2^-24
            d[j][1] = prmt(words[j], 0u, 0x4342u);  // (k2, k3) * 2^-24
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if constexpr (SPLITQ) {
            // plain operands (k0,k4),(k2,k6) -> accg; x16 operands (k1,k5),(k3,k7) -> accq; X from both activation pairs
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              if (fresh) {
                mma16816<A>(accg[t][mt], zero4, d[2 * t][0], d[2 * t + 1][0], d[2 * t][2], d[2 * t + 1][2], xb[mt][0], xb[mt][2]);
                mma16816<A>(accq[t][mt], zero4, d[2 * t][1], d[2 * t + 1][1], d[2 * t][3], d[2 * t + 1][3], xb[mt][1], xb[mt][3]);
              } else {
                mma16816<A>(accg[t][mt], accg[t][mt], d[2 * t][0], d[2 * t + 1][0], d[2 * t][2], d[2 * t + 1][2], xb[mt][0], xb[mt][2]);
                mma16816<A>(accq[t][mt], accq[t][mt], d[2 * t][1], d[2 * t + 1][1], d[2 * t][3], d[2 * t + 1][3], xb[mt][1], xb[mt][3]);
              }
            }
            if constexpr (!XS) {
              if (fresh) mma16816<A>(accx[mt], zero4, ONE2, ONE2, ONE2, ONE2, xb[mt][0], xb[mt][2]);
              else mma16816<A>(accx[mt], accx[mt], ONE2, ONE2, ONE2, ONE2, xb[mt][0], xb[mt][2]);
              mma16816<A>(accx[mt], accx[mt], ONE2, ONE2, ONE2, ONE2, xb[mt][1], xb[mt][3]);
            }
          } else {
#pragma unroll
            for (int h = 0; h < PAIRS / 2; ++h) {
#pragma unroll
              for (int t = 0; t < NT; ++t) {
                // data: rows g8 <- column 2t, rows g8+8 <- column 2t+1
                if (fresh && h == 0) {
                  mma16816<A>(accg[t][mt], zero4, d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1], d[2 * t + 1][2 * h + 1],
                              xb[mt][2 * h], xb[mt][2 * h + 1]);
                } else {
                  mma16816<A>(accg[t][mt], accg[t][mt], d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1],
                              d[2 * t + 1][2 * h + 1], xb[mt][2 * h], xb[mt][2 * h + 1]);
                }
              }
              // X[m] += sum_k x[m,k] over this operand pair's 16 k (all 16 rows of the result are equal)
              if constexpr (!XS) {
                if (fresh && h == 0) {
                  mma16816<A>(accx[mt], zero4, ONE2, ONE2, ONE2, ONE2, xb[mt][2 * h], xb[mt][2 * h + 1]);
                } else {
                  mma16816<A>(accx[mt], accx[mt], ONE2, ONE2, ONE2, ONE2, xb[mt][2 * h], xb[mt][2 * h + 1]);
                }
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) dequant_word<A, BITS, ORD, FZP>(words[j], cq[j], d[j]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            // rows g8 of the MMA <- column 2t of this lane, rows g8+8 <- column 2t+1
#pragma unroll
            for (int h = 0; h < PAIRS / 2; ++h) {
              mma16816<A>(acc[t][mt], acc[t][mt], d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1],
                          d[2 * t + 1][2 * h + 1], xb[mt][2 * h], xb[mt][2 * h + 1]);
            }
          }
        }
      }
    }
    group_open = true;
    if (--batches_left == 0) {  // group finished
      if constexpr (FUSED) close_group();
      group_open = false;
      batches_left = bpg;
      refresh = true;
    }
  };

  // ---- main loop: two batches (the whole window) per iteration ------------------------------------------------------------------
  for (int i = 0; i < nunits; i += 2) {
    process(std::integral_constant<int, 0>{}, i * U);
    if (i == 0) DEC_STAMP(4);
    if (i + 1 < nunits) process(std::integral_constant<int, U>{}, (i + 1) * U);
    wp += WIN * step_bytes;
  }
  if constexpr (FUSED) {
    if (group_open) close_group();  // slab ended inside a group (per-channel scales split across warps)
  }
  DEC_STAMP(5);

  // ---- publish -----------------------------------------------------------------------------------------------------------------
  if (S == 1) {  // this warp saw the whole K range: store directly
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int m = mt * 8 + 2 * tig;
        const int n = col + 2 * t;
        const float b0 = bias_at(p, n), b1 = bias_at(p, n + 1);
        if (m < M) {
          store_out_all<A>(p, size_t(m) * N + n, acc[t][mt][0] + b0);
          store_out_all<A>(p, size_t(m) * N + n + 1, acc[t][mt][2] + b1);
        }
        if (m + 1 < M) {
          store_out_all<A>(p, size_t(m + 1) * N + n, acc[t][mt][1] + b0);
          store_out_all<A>(p, size_t(m + 1) * N + n + 1, acc[t][mt][3] + b1);
        }
      }
    push_signal_block_done(p, ws_counter + (CB200_WORKSPACE_COUNTER_BYTES / 4 - 1), unsigned(C), lane);
    return;
  }
  float* mine = ws_partial + (size_t(cb) * S + si) * (M * CW);
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int m = mt * 8 + 2 * tig;
      const int lc = VEC * g8 + 2 * t;
      if (m < M) {
        __stcg(mine + m * CW + lc, acc[t][mt][0]);
        __stcg(mine + m * CW + lc + 1, acc[t][mt][2]);
      }
      if (m + 1 < M) {
        __stcg(mine + (m + 1) * CW + lc, acc[t][mt][1]);
        __stcg(mine + (m + 1) * CW + lc + 1, acc[t][mt][3]);
      }
    }
  __threadfence();
  __syncwarp();
  unsigned prev = 0;
  if (lane == 0) prev = atomicAdd(ws_counter + cb, 1u);
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev != unsigned(S - 1)) return;
  __threadfence();  // last arriver: every other slab's partial is visible now
  const float* base = ws_partial + size_t(cb) * S * (M * CW);
  if (p.n_peers > 0 && !geo.scalar_push) {
    // multi-GPU push: each lane owns 8 consecutive columns of one row (4 lanes per row) and the result leaves as two 8-byte
    // vector stores per destination (fp16 / bf16) -- 4x fewer NVLink writes per peer than element stores
    for (int e = lane; e < M * (CW / 8); e += 32) {
      const int m = e / (CW / 8), lc = (e - m * (CW / 8)) * 8;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      for (int s = 0; s < S; ++s) {
        const float4* src = reinterpret_cast<const float4*>(base + size_t(s) * (M * CW) + m * CW + lc);
        const float4 a = __ldcg(src), b = __ldcg(src + 1);
        v0.x += a.x; v0.y += a.y; v0.z += a.z; v0.w += a.w;
        v1.x += b.x; v1.y += b.y; v1.z += b.z; v1.w += b.w;
      }
      if (p.out_bias != nullptr) {
        const int n = cb * CW + lc;
        v0.x += bias_at(p, n); v0.y += bias_at(p, n + 1); v0.z += bias_at(p, n + 2); v0.w += bias_at(p, n + 3);
        v1.x += bias_at(p, n + 4); v1.y += bias_at(p, n + 5); v1.z += bias_at(p, n + 6); v1.w += bias_at(p, n + 7);
      }
      store_out4_all(p, size_t(m) * N + cb * CW + lc, v0);
      store_out4_all(p, size_t(m) * N + cb * CW + lc + 4, v1);
    }
  } else {
    for (int e = lane; e < M * CW; e += 32) {
      float v = 0.f;
      for (int s = 0; s < S; ++s) v += __ldcg(base + size_t(s) * (M * CW) + e);
      const int m = e / CW, lc = e - m * CW;
      store_out_all<A>(p, size_t(m) * N + cb * CW + lc, v + bias_at(p, cb * CW + lc));
    }
  }
  if (lane == 0) ws_counter[cb] = 0u;  // leave the counter clean for the next launch
  push_signal_block_done(p, ws_counter + (CB200_WORKSPACE_COUNTER_BYTES / 4 - 1), unsigned(C), lane);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <class A, int BITS, int MT, int U, int ZK>
static int launch_fused(bool fused, const Problem& p, const DecodeGeom& geo, int blocks, size_t smem, float* wsp,
                        unsigned* wsc, cudaStream_t stream) {
  void (*kern)(const Problem, const DecodeGeom, float*, unsigned int*);
  // bf16 x int8: 128 + q does not fit a bf16 significand, so THIS kernel has no magic-number operand -> per-weight dequant
  // (the streaming kernel, AUTO's choice, splits the byte into two nibble planes instead)
  if (fused && !(std::is_same<A, __nv_bfloat16>::value && BITS == 8)) {
    kern = decode_kernel<A, BITS, MT, U, ZK, true>;
    if constexpr (MT == 1) {
      if (geo.cta_aux_bytes > 0) kern = decode_kernel<A, BITS, MT, U, ZK, true, true>;  // group sums of x once per CTA
    }
  } else {
    kern = decode_kernel<A, BITS, MT, U, ZK, false>;
  }
  if (smem > 48 * 1024) CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = geo.pdl ? 1 : 0;
  CB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p, geo, wsp, wsc));
  return 0;
}

template <class A, int BITS, int MT, int U>
static int launch_zk(bool fused, const Problem& p, const DecodeGeom& geo, int blocks, size_t smem, float* wsp, unsigned* wsc,
                     cudaStream_t stream) {
  switch (p.zp_kind) {
    case CB200_ZP_NONE: return launch_fused<A, BITS, MT, U, CB200_ZP_NONE>(fused, p, geo, blocks, smem, wsp, wsc, stream);
    case CB200_ZP_I32: return launch_fused<A, BITS, MT, U, CB200_ZP_I32>(fused, p, geo, blocks, smem, wsp, wsc, stream);
    case CB200_ZP_U8: return launch_fused<A, BITS, MT, U, CB200_ZP_U8>(fused, p, geo, blocks, smem, wsp, wsc, stream);
    default: return launch_fused<A, BITS, MT, U, CB200_ZP_ACT>(fused, p, geo, blocks, smem, wsp, wsc, stream);
  }
}

template <class A, int BITS>
static int launch_typed(int u, bool fused, const Problem& p, const DecodeGeom& geo, int blocks, size_t smem, float* wsp,
                        unsigned* wsc, cudaStream_t stream) {
  const bool mt2 = geo.m_rows > 8;
  if (u == 4) {
    return mt2 ? launch_zk<A, BITS, 2, 4>(fused, p, geo, blocks, smem, wsp, wsc, stream)
               : launch_zk<A, BITS, 1, 4>(fused, p, geo, blocks, smem, wsp, wsc, stream);
  }
  return mt2 ? launch_zk<A, BITS, 2, 2>(fused, p, geo, blocks, smem, wsp, wsc, stream)
             : launch_zk<A, BITS, 1, 2>(fused, p, geo, blocks, smem, wsp, wsc, stream);
}

// one entry per translation unit (defined by CB200_DEFINE_DECODE_ENTRY in decode_*.cu)
#define CB200_DECODE_ENTRY(name) \
  int name(int u, bool fused, const Problem& p, const DecodeGeom& geo, int blocks, size_t smem, float* wsp, unsigned* wsc, \
           cudaStream_t stream)
CB200_DECODE_ENTRY(launch_decode_f16_w4);
CB200_DECODE_ENTRY(launch_decode_f16_w8);
CB200_DECODE_ENTRY(launch_decode_bf16_w4);
CB200_DECODE_ENTRY(launch_decode_bf16_w8);

}  // namespace cb200
