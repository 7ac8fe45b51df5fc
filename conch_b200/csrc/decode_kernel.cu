// Decode-regime kernel (M <= 16 per pass): split-K skinny GEMM, HBM-bound.
//
// Replaces the reference's single 128x64x64 Triton tile (conch/kernels/quantization/gemm.py:136-145, 491-494),
// which at M=1 launches N/64 CTAs with no split-K.  Design (B200):
//   * one WARP owns one (32-column block) x (1/S of K) slab; S is a divisor of the batches per column chosen so that the
//     whole grid is one co-resident wave.  Neighbouring warps take neighbouring column blocks of the same K slab, so a
//     CTA's four warps share one staged copy of the activations (one __syncthreads, for that barrier's init); apart
//     from that the warps are independent;
//   * the packed weights are the only large stream: every lane keeps a rotating window of 2*U (= 8) 128-bit read-once
//     loads (ld.global.nc, L1 no-allocate) in flight -- the register that held step j is re-loaded with step j + 2U as
//     soon as it has been read -- so with ~24 resident warps per SM about 100 KB per SM is outstanding, which is what it
//     takes to approach the HBM peak (measured: benchmarks/micro/stream_pattern.cu).  cp.async was measured and
//     rejected: with 4 row segments per instruction LDGSTS serialises into 32 shared-memory wavefronts
//     (benchmarks/micro/decode_steps.cu).  NOTHING else may go through the load/store queue: loads return in issue
//     order, so an L1-hit activation load queued behind DRAM-missing weight loads stalls the step for a full DRAM
//     latency (measured: +8..16 us on a 58.7 MB layer, profiles/r01_micro_decode_features.log).  The warp's slice of
//     x and its scale / zero-point rows are therefore fetched ONCE, at kernel start, by TMA bulk copies
//     (cp.async.bulk -> per-warp shared memory, mbarrier completion) and read back with ld.shared.
//     All cursors are linear pointer increments;
//   * FUSED=false: nibbles are unpacked and W = (q - bias - zp) * s is formed per weight with dequant_word() -- the
//     same bits as the dequant kernel and the reference's w_ref -- before the MMA;
//   * FUSED=true: NO floating-point op per weight and no offset MMA.  fp16: the masked nibble / byte itself is the MMA
//     operand -- the bit pattern 0x0000|v (v < 2048) read as fp16 is the subnormal (or first-binade) number v * 2^-24,
//     exactly, and the tensor core multiplies subnormals exactly -- so a word costs 1 shift + 4 AND.  bf16 int4: the
//     magic-number operand 0x4300|q = 128 + q (bf16 subnormals would push the fp32 products into the fp32 subnormal
//     range).  The zero-point (and the 128) is removed per GROUP instead of per weight: one extra MMA per operand pair
//     with an all-ones 16-row operand yields X[m] = sum_k x[m,k] of the group in the accumulator layout, and
//         acc[m,n] += (s[n] * 2^24) * G[m,n] - (s[n] * zb[n]) * X[m]          (G = sum_k x q 2^-24, fp32, exact products)
//     fp16 int4's nibbles at bits 4-7 arrive as 16 q; their activations are pre-multiplied by 1/16 (and the ones operand
//     that meets them is 16.0).  The result differs from the reference only by not rounding each W to fp16/bf16 -- far
//     inside the 1e-2 tolerance; one-hot activations still reproduce the reference's W bit for bit;
//   * the dot product runs on the tensor pipe as mma.sync.m16n8k16 with the WEIGHTS as the 16-row operand (swap-AB) and
//     x^T as the 8-column operand; the four lanes of a quad hold four different packed rows, so the K-reduction
//     across lanes happens inside the MMA, fp32 accumulate; x is permuted in registers to the nibble order the cheap
//     `lop3` extraction produces ((k0,k4),(k1,k5),(k2,k6),(k3,k7));
//   * S == 1: the warp stores its result directly.  S > 1: it writes an fp32 partial, bumps the column block's arrival
//     counter, and the last arriver sums the S partials in slab order (deterministic, self-resetting counters: no
//     memset, no second launch);
//   * programmatic dependent launch (opt-in, cb200_set_pdl): a warp calls griddepcontrol.launch_dependents as soon as its
//     own weight window is in flight, so the next kernel's CTAs become resident while this kernel's stragglers finish,
//     issue THEIR weight / metadata prefetch (static data: legal before griddepcontrol.wait) and then block in
//     griddepcontrol.wait; x, the output and the workspace are only touched after the wait.
//   Measured anatomy of one launch (benchmarks/decode_trace.py, 58.7 MB layer): head 3.4 us (launch skew, 9.7 MB of
//   first-window requests ramping the HBM up), main loop 12.0 us (5 TB/s), straggler tail + split-K arrival 3-4 us.  A
//   Plain 16-byte loads for the scale / zero-point rows were measured too: behind the weight window they block at issue
//   until the window drains (+2.9 us), in front of it they push the window out of the load queue (~100 512-byte requests
//   per SM is all the LSU tracks) -- hence the TMA bulk copies, which bypass that queue, at ~40 ns of issue time each.
//   `prefetch.global.L2` running 1..16 iterations ahead of the window (to get more bytes in flight than that queue
//   allows) made every layer 5-15% slower, and so did a one-shot prefetch of the next 1-4 iterations issued before
//   griddepcontrol.wait (meant to use the previous kernel's idle tail).  A stream-K variant (equal ranges of the linearised (column block, batch) space, every SM equally loaded) was built and
//   measured: -8% on that layer, +8% on the smaller ones (second publish per straddling warp, longer prologue), no net
//   gain on the bench -- not kept.
#include "common.cuh"
#include "dequant.cuh"

namespace cb200 {

DecodeTuning g_decode_tuning = {0, 0, 0, 0};

#ifdef CB200_TC_TRACE  // debug builds only (benchmarks/decode_trace.py): globaltimer stamps of warp 0 of every CTA
__device__ long long* g_dec_trace = nullptr;
extern "C" __attribute__((visibility("default"))) void cb200_debug_set_decode_trace(void* p) {
  cudaMemcpyToSymbol(g_dec_trace, &p, sizeof(p));
}
__device__ __forceinline__ void dec_stamp(int slot) {
  if (g_dec_trace && (threadIdx.x & 127) == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_dec_trace[blockIdx.x * 8 + slot] = t;
    if (slot == 0) {
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      g_dec_trace[blockIdx.x * 8 + 7] = sm;
    }
  }
}
#define DEC_STAMP(slot) dec_stamp(slot)
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
};

template <class A>
__device__ __forceinline__ void mma16816(float (&d)[4], const float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<A, __half>::value) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
  }
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
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

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
  constexpr bool PRESCALE = FUSED && HALF && BITS == 4;  // nibbles at bits 4-7 arrive as 16 q: their activations get x 1/16
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
          if constexpr (PRESCALE) {
            o1 = ActOps<A>::mul(o1, 0x2C002C00u);
            o3 = ActOps<A>::mul(o3, 0x2C002C00u);
          }
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
  constexpr uint32_t ONE2_ODD = PRESCALE ? 0x4C004C00u : ONE2;     // (16.0, 16.0) against activations scaled by 1/16

  ColQ cq[VEC];          // !FUSED: per-column dequant constants
  float sa[VEC];         // FUSED: s[n] * 2^24 (fp16) | s[n] (bf16): multiplies G = sum_k x * operand
  float sb[VEC];         // FUSED: -s[n] * zb (fp16) | -s[n] * (128 + zb) (bf16): multiplies X = sum_k x

  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  float acc[NT][MT][4];   // running (scaled) sum of this warp's slab
  float accg[NT][MT][4];  // FUSED: G of the current group (exact products, fp32 sum)
  float accx[MT][4];      // FUSED: X[m] = sum_k x[m,k] of the current group, in the accumulator layout
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][mt][i] = accg[t][mt][i] = accx[mt][i] = 0.f;
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
          acc[t][mt][i] = fmaf(sa[2 * t + (i >> 1)], accg[t][mt][i], off);
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
          if constexpr (PRESCALE && !XS) {  // activations that meet the x16 operands: x * 1/16
            xb[mt][1] = ActOps<A>::mul(xb[mt][1], 0x2C002C00u);
            xb[mt][3] = ActOps<A>::mul(xb[mt][3], 0x2C002C00u);
          }
#pragma unroll
          for (int h = 0; h < PAIRS / 2; ++h) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              // data: rows g8 <- column 2t, rows g8+8 <- column 2t+1; operand pair 2h is plain, 2h+1 is the x16 kind
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
                mma16816<A>(accx[mt], zero4, ONE2, ONE2, ONE2_ODD, ONE2_ODD, xb[mt][2 * h], xb[mt][2 * h + 1]);
              } else {
                mma16816<A>(accx[mt], accx[mt], ONE2, ONE2, ONE2_ODD, ONE2_ODD, xb[mt][2 * h], xb[mt][2 * h + 1]);
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
        if (m < M) {
          store_out_all<A>(p, size_t(m) * N + n, acc[t][mt][0]);
          store_out_all<A>(p, size_t(m) * N + n + 1, acc[t][mt][2]);
        }
        if (m + 1 < M) {
          store_out_all<A>(p, size_t(m + 1) * N + n, acc[t][mt][1]);
          store_out_all<A>(p, size_t(m + 1) * N + n + 1, acc[t][mt][3]);
        }
      }
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
  for (int e = lane; e < M * CW; e += 32) {
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += __ldcg(base + size_t(s) * (M * CW) + e);
    const int m = e / CW, lc = e - m * CW;
    store_out_all<A>(p, size_t(m) * N + cb * CW + lc, v);
  }
  if (lane == 0) ws_counter[cb] = 0u;  // leave the counter clean for the next launch
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int pick(int forced, int heuristic) { return forced > 0 ? forced : heuristic; }

int g_decode_pdl = 0;

struct DecodePlan {
  int u;               // steps per batch (4, or 2 for int4 with g = 64)
  int bpg;             // batches per scale group
  int splits, units_per_warp, col_blocks, total_warps;
  int ctas_per_sm;
  int x_row_bytes, meta_row_bytes, ngroups, warp_smem, max_units, cta_x_bytes, cta_aux_bytes, cta_smem, cb_padded;
  bool fused;
};

// Every knob is overridable through cb200_set_decode_tuning(warps_n -> 1: per-weight dequant before the MMA (W bits
// identical to the reference's w_ref), 2 (default, measured 5-10% faster): fused magic-number operands; warps_k -> CTAs
// per SM assumed resident;
// split_k -> forced split count S (rounded down to a legal divisor); vec_words unused).
static DecodePlan decode_plan(int mrows, int K, int N, int bits, int g) {
  DecodePlan pl;
  const int step_k = 4 * (32 / bits);
  pl.u = (g % (4 * step_k) == 0) ? 4 : 2;
  const int B = K / (pl.u * step_k);  // batches per column block
  pl.bpg = g / (pl.u * step_k);
  pl.col_blocks = N / 32;
  pl.fused = g_decode_tuning.warps_n != 1;
  // __launch_bounds__(128, MT == 2 ? 3 : (FUSED ? 4 : 5))
  pl.ctas_per_sm = pick(g_decode_tuning.warps_k, mrows > 8 ? 3 : (pl.fused ? 4 : 5));
  const int by_regs = pl.ctas_per_sm;
  const int zrow = 128;  // worst case (int32 zero-points) keeps the layout independent of the zero-point kind
  // shared memory of one CTA: the x slab (shared by its 4 warps) + per warp the group rows and a barrier + the x barrier
  auto warp_smem_for = [&](int per) {
    const int ng = (per + pl.bpg - 1) / pl.bpg;
    return ((ng * (64 + zrow) + 8 + 127) / 128) * 128;
  };
  auto cta_x_for = [&](int per) { return ((mrows * (per * pl.u * step_k * 2 + 64) + 127) / 128) * 128; };
  // chunk sums [mrows][slab_k / 8] + group sums [ngroups][8], fp32 (used by the fused path for M <= 8)
  auto cta_aux_for = [&](int per) {
    const int ng = (per + pl.bpg - 1) / pl.bpg;
    return ((mrows * (per * pl.u * step_k / 8) * 4 + ng * 32 + 127) / 128) * 128;
  };
  auto cta_smem_for = [&](int per) { return cta_x_for(per) + cta_aux_for(per) + 4 * warp_smem_for(per) + 16; };
  // the CTA's staging budget bounds the slab: x rows (2 B per k per row) within ~24 KB (48 KB for the 16-row variant)
  {
    const int budget = mrows > 8 ? 48 * 1024 : 24 * 1024;
    int mu = budget / (mrows * 2 * pl.u * step_k);
    if (mu < 1) mu = 1;
    pl.max_units = mu;
  }
  // S must divide B, and a slab must not cut a group unless the group spans whole slabs (per-channel / huge groups):
  // legal S: (B / S) % bpg == 0  or  bpg % (B / S) == 0.  Pick the largest legal S that keeps one co-resident wave.
  int best = 1;
  bool best_ok = false;
  for (int sdiv = 1; sdiv <= B; ++sdiv) {
    if (B % sdiv) continue;
    const int per = B / sdiv;
    if (!(per % pl.bpg == 0 || pl.bpg % per == 0)) continue;
    if (g_decode_tuning.split_k > 0) {
      if (sdiv <= g_decode_tuning.split_k || per > pl.max_units) best = sdiv;
      continue;
    }
    if (per > pl.max_units) {  // slab too long for the staging buffer: must split further
      best = sdiv;
      continue;
    }
    int by_smem = (227 * 1024) / (cta_smem_for(per) + 1024);
    if (by_smem > by_regs) by_smem = by_regs;
    const int64_t capacity = int64_t(148) * by_smem * 4;  // co-resident warps at this split
    const int64_t warps = int64_t((pl.col_blocks + 3) / 4 * 4) * sdiv;
    if (best_ok && warps > capacity && sdiv > 1) break;  // one co-resident wave (measured: more waves do not help)
    if (best_ok && per < 2 && sdiv > 1) break;
    best = sdiv;
    best_ok = true;
  }
  pl.splits = best;
  pl.units_per_warp = B / best;
  pl.cb_padded = (pl.col_blocks + 3) / 4 * 4;  // whole CTAs per K slab
  pl.total_warps = pl.cb_padded * best;
  pl.x_row_bytes = pl.units_per_warp * pl.u * step_k * 2 + 64;
  pl.meta_row_bytes = 64 + zrow;
  pl.ngroups = (pl.units_per_warp + pl.bpg - 1) / pl.bpg;
  pl.warp_smem = warp_smem_for(pl.units_per_warp);
  pl.cta_x_bytes = cta_x_for(pl.units_per_warp);
  // group sums of x once per CTA (4 instead of 6 MMAs per step): pays for M <= 2 on long slabs only (measured)
  const bool xsum = pl.fused && mrows <= 2 && pl.units_per_warp >= 8 ;
  pl.cta_aux_bytes = xsum ? cta_aux_for(pl.units_per_warp) : 0;
  pl.cta_smem = cta_smem_for(pl.units_per_warp) - (xsum ? 0 : cta_aux_for(pl.units_per_warp));
  return pl;
}

size_t decode_workspace_bytes(int64_t M, int64_t K, int64_t N, int bits) {
  // the group size changes the plan (U, legal splits): take the largest footprint over the group sizes the op accepts
  size_t need = 0;
  for (int rows : {int(M < 16 ? M : 16), int(M > 16 ? ((M - 1) % 16) + 1 : (M < 16 ? M : 16))}) {
    for (int64_t g = 64; g <= K; g *= 2) {
      if (K % g) continue;
      const DecodePlan pl = decode_plan(rows, int(K), int(N), bits, int(g));
      const size_t b = pl.splits > 1 ? size_t(pl.total_warps) * rows * 32 * sizeof(float) : 0;
      need = b > need ? b : need;
    }
    const DecodePlan pk = decode_plan(rows, int(K), int(N), bits, int(K));
    const size_t b = pk.splits > 1 ? size_t(pk.total_warps) * rows * 32 * sizeof(float) : 0;
    need = b > need ? b : need;
  }
  return size_t(CB200_WORKSPACE_COUNTER_BYTES) + need;
}

template <class A, int BITS, int MT, int U, int ZK>
static int launch_fused(bool fused, const Problem& p, const DecodeGeom& geo, int blocks, size_t smem, float* wsp,
                        unsigned* wsc, cudaStream_t stream) {
  void (*kern)(const Problem, const DecodeGeom, float*, unsigned int*);
  // bf16 x int8: 128 + q does not fit a bf16 significand, so there is no magic-number operand -> per-weight dequant
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

int launch_decode(const Problem& p_in, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream,
                  int* launches) {
  for (int m0 = 0; m0 < p_in.M; m0 += 16) {
    Problem p = p_in;
    const int mrows = (p_in.M - m0) < 16 ? (p_in.M - m0) : 16;
    p.x = reinterpret_cast<const char*>(p_in.x) + size_t(m0) * p_in.K * 2;
    const size_t out_es = p_in.out_dtype == CB200_F32 ? 4 : 2;
    p.out = reinterpret_cast<char*>(p_in.out) + size_t(m0) * p_in.N * out_es;
    for (int i = 0; i < p_in.n_peers; ++i) {
      p.peer_out[i] = reinterpret_cast<char*>(p_in.peer_out[i]) + size_t(m0) * p_in.N * out_es;
    }
    p.M = mrows;

    const DecodePlan pl = decode_plan(mrows, p.K, p.N, bits, p.g);
    const size_t need =
        size_t(CB200_WORKSPACE_COUNTER_BYTES) + (pl.splits > 1 ? size_t(pl.total_warps) * mrows * 32 * sizeof(float) : 0);
    if (pl.splits > 1 && (ws == nullptr || ws_bytes < need)) {
      set_error("decode kernel needs %zu workspace bytes, got %zu", need, ws_bytes);
      return CB200_ERR_WORKSPACE;
    }
    if (size_t(pl.col_blocks) * sizeof(unsigned) > CB200_WORKSPACE_COUNTER_BYTES) {
      set_error("N=%d too large for the split-K counter area", p.N);
      return CB200_ERR_SHAPE;
    }
    if (size_t(p.K / p.g) * p.N >= (size_t(1) << 31)) {
      set_error("scale matrix too large");
      return CB200_ERR_SHAPE;
    }
    DecodeGeom geo;
    geo.m_rows = mrows;
    geo.splits = pl.splits;
    geo.units_per_warp = pl.units_per_warp;
    geo.total_warps = pl.total_warps;
    geo.bpg = pl.bpg;
    geo.pdl = g_decode_pdl;
    geo.x_row_bytes = pl.x_row_bytes;
    geo.meta_row_bytes = pl.meta_row_bytes;
    geo.ngroups = pl.ngroups;
    geo.warp_smem = pl.warp_smem;
    geo.cta_x_bytes = pl.cta_x_bytes;
    geo.cb_padded = pl.cb_padded;
    geo.cta_aux_bytes = pl.cta_aux_bytes;
    const int blocks = (pl.total_warps + 3) / 4;
    float* wsp = ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES) : nullptr;
    unsigned* wsc = reinterpret_cast<unsigned*>(ws);
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_typed<__half, 4>(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream)
                     : launch_typed<__half, 8>(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream);
    } else {
      rc = bits == 4 ? launch_typed<__nv_bfloat16, 4>(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream)
                     : launch_typed<__nv_bfloat16, 8>(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
