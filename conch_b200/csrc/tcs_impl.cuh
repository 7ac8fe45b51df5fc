// tcgen05 streaming decode kernel (CB200_DECODE_TCS): the HBM-bound regime (M <= 32) on the 5th-generation tensor cores.
// Design notes are at the top of tcs_kernel.cu (host plan + dispatch); this header holds the kernel template and the launch
// ladder shared by tcs_f16_w4.cu / tcs_f16_w8.cu / tcs_bf16_w4.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dequant.cuh"
#include "tc_ptx.cuh"

namespace cb200 {

constexpr int kTcsSets = 2;                          // dequant / fold warp sets (4 warps = 128 TMEM lanes each)
constexpr int kTcsThreads = 32 + 64 * kTcsSets + 128 * kTcsSets;  // warp 0 producer; per set: MMA issuer (1, 2), x operand warp (3, 4), 4 dequant warps (5..12)
constexpr int kTcsStageW = 8192;                     // packed weight bytes per stage: 16 packed rows x 128 columns
constexpr int kTcsTmemCols = 256;                    // per CTA: two CTAs (this kernel and its successor) share an SM's 512

struct TcsGeom {
  int m_rows;          // rows of x in this pass (<= MT)
  int spc;             // stages per 128-column block = K / KS (KS = 16 packed rows = 128 k for int4, 64 k for int8)
  int ncb;             // 128-column blocks = ceil(N / 128)
  int q, r;            // run partition: CTA c owns stages [c*q + min(c,r), (c+1)*q + min(c+1,r)) of the k-fastest stage list
  int ring;            // shared-memory ring depth in stages (even)
  int stage_bytes;     // multiple of 1024
  int x_off;           // byte offset of the raw x tile inside a stage
  int x_bytes;         // bytes one x box delivers (m_rows * KS * 2)
  int meta_off;        // byte offset of the scale / zero-point row inside a stage
  int spg;             // stages per scale group = g / KS (>= 1)
  int maxslots;        // split-K partial slots per 128-column block
  int pdl;             // launched with programmatic stream serialization
  long long* trace;    // [CTAs][16] globaltimer stamps when non-null
};

__device__ __forceinline__ int tcs_cta_of(int s, int q, int r) {
  const int big = r * (q + 1);
  return s < big ? s / (q + 1) : r + (s - big) / q;
}
__device__ __forceinline__ void tcs_stamp(long long* trace, int slot) {
  if (trace) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 32 + slot] = t;
  }
}
// phase accounting (trace slots 10..31): accumulated SM cycles a role spent in each of its phases, written once at the end
struct TcsPhase {
  long long t, acc[6];
  bool on;
  __device__ __forceinline__ void begin(bool enable) {
    on = enable;
    for (int i = 0; i < 6; ++i) acc[i] = 0;
    if (on) t = clock64();
  }
  __device__ __forceinline__ void mark(int phase) {
    if (on) {
      const long long n = clock64();
      acc[phase] += n - t;
      t = n;
    }
  }
  __device__ __forceinline__ void dump(long long* trace, int slot0) {
    if (on) {
      for (int i = 0; i < 6; ++i) trace[blockIdx.x * 32 + slot0 + i] = acc[i];
    }
  }
};
__device__ __forceinline__ void tcs_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void tcs_wait(uint32_t bar, uint32_t parity) {  // one probe first, the bounded spin only if it fails
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
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

// Walks the stages a SET owns (every second stage of the CTA's run: set s takes s, s + 2, ...) and knows where the set's share
// of a scale group begins and ends.  A scale group is `spg` consecutive k-slabs of one 128-column block; the two sets are
// independent pipelines (own accumulator, own X sums), each folds ITS stages of a group with the group's scale / zero-point
// -- the fold is linear, so the two shares add up to the group.  Every role of a set steps its own copy of the cursor.
struct TcsCursor {
  int i, kt, cb, gpos, gidx;  // stage of the run, its k slab and column block, position inside / index of its scale group
  int spc, spg, n_run;
  __device__ __forceinline__ void init(int run_a, int n_run_, int spc_, int spg_, int set) {
    spc = spc_;
    spg = spg_;
    n_run = n_run_;
    i = 0;
    cb = run_a / spc;
    kt = run_a - cb * spc;
    gidx = kt / spg;
    gpos = kt - gidx * spg;
    if (set == 1) step1();
  }
  __device__ __forceinline__ void step1() {
    ++i;
    if (++kt == spc) {
      kt = 0;
      ++cb;
      gpos = 0;
      gidx = 0;
    } else if (++gpos == spg) {
      gpos = 0;
      ++gidx;
    }
  }
  __device__ __forceinline__ void next() {  // this set's next stage
    step1();
    step1();
  }
  __device__ __forceinline__ bool valid() const { return i < n_run; }
  __device__ __forceinline__ bool same_group(const TcsCursor& o) const { return cb == o.cb && gidx == o.gidx; }
  __device__ __forceinline__ bool last() const {  // no further stage of this set (two k slabs on) in the same scale group
    return gpos + 2 >= spg || kt + 2 >= spc || i + 2 >= n_run;
  }
};

// A: activation type, BITS: 4|8, MT: tensor-core N (16 or 32), MR: upper bound of the rows of x in the pass (4, 16 or 32: sizes
// the per-thread accumulators; a 1-row GEMV must not carry 16 of them), ZK: cb200_zp_kind.
template <class A, int BITS, int MT, int MR, int ZK>
__global__ void __launch_bounds__(kTcsThreads, MT == 16 ? 2 : 1)
    tcs_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Problem p, const TcsGeom geo,
               float* __restrict__ ws_partial, unsigned int* __restrict__ ws_counter) {
  constexpr int PF = 32 / BITS;   // k per packed word
  constexpr int KS = 16 * PF;     // k per stage: 128 | 64
  constexpr int NH = KS / 64;     // 64-k halves per stage (one 128-byte operand row each): 2 | 1
  constexpr int WPH = 64 / PF;    // packed words per half and column: 8 | 16
  constexpr bool HALF = std::is_same<A, __half>::value;
  constexpr bool EO = HALF && BITS == 4;  // fp16 int4: nibbles at bits 4-7 arrive as 16 q -> own MMAs into their own accumulator
  constexpr bool FZP = ZK == CB200_ZP_ACT;
  constexpr uint32_t ZES = ZK == CB200_ZP_I32 ? 4u : ZK == CB200_ZP_U8 ? 1u : ZK == CB200_ZP_ACT ? 2u : 0u;
  constexpr int DW = (EO ? 2 : 1) * MT;             // TMEM columns of one set's group accumulator(s)
  constexpr int A_COL0 = kTcsSets * DW;             // A operand: one 64-column slot per set (32 per half)
  constexpr int XB_HALF = MT * 128;                 // one operand half: MT rows x 128 B, 128-byte swizzle
  constexpr int XB_BYTES = NH * XB_HALF;            // one x operand slot; two per set
  static_assert(HALF || BITS == 4, "bf16 x int8 is not a tcs_kernel configuration");
  static_assert(A_COL0 + kTcsSets * 64 <= kTcsTmemCols, "TMEM over-subscribed");

  // the next kernel in the stream may start now: it prefetches its own (static) weights while this one runs
  if (geo.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) tcs_stamp(geo.trace, 0);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  const uint32_t smem0 = smem_u32(smem_raw) + smem_pad;
  unsigned char* const smem_gen = smem_raw + smem_pad;
  // control page (2 KB): barriers @0 (8 B each), TMEM base @1016, X sums [2 sets][MT] fp32 @1024
  const int ring = geo.ring;
  auto full = [&](int s) { return smem0 + 8u * uint32_t(s); };                          // TMA landed (tx)
  auto empty = [&](int s) { return smem0 + 8u * uint32_t(32 + s); };                    // 4 dequant warps + the x warp of the slot's set left it
  auto a_ready = [&](int set) { return smem0 + 8u * uint32_t(64 + set); };              // the set stored its A operand (4 warps)
  auto xb_full = [&](int set, int b) { return smem0 + 8u * uint32_t(66 + 2 * set + b); };   // x operand slot b of the set written
  auto mma_done = [&](int set, int b) { return smem0 + 8u * uint32_t(70 + 2 * set + b); };  // MMAs of the set's stage j (b = j & 1) retired
  auto d_full = [&](int set) { return smem0 + 8u * uint32_t(74 + set); };               // the set's group accumulator is final
  auto d_empty = [&](int set) { return smem0 + 8u * uint32_t(76 + set); };              // ... and has been folded (4 warps)
  auto xs_ready = [&](int set) { return smem0 + 8u * uint32_t(78 + set); };             // the set's X sums of the group are written
  const uint32_t tmem_slot = smem0 + 1016u;
  float* const xsum = reinterpret_cast<float*>(smem_gen + 1024);                        // [2 sets][MT]
  const uint32_t xb0 = smem0 + 2048u;                                                   // x operand slots [set][2], 1024-aligned (2 KB control page)
  const uint32_t ring0 = xb0 + 4u * XB_BYTES;
  const unsigned char* const ring_ptr = smem_gen + 2048 + 4 * XB_BYTES;
  unsigned char* const xb_ptr = smem_gen + 2048;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = geo.m_rows, N = p.N;
  const int spc = geo.spc;
  const int cta = blockIdx.x;
  const int run_a = cta * geo.q + min(cta, geo.r);
  const int n_run = geo.q + (cta < geo.r ? 1 : 0);
  const bool zp_per_group = ZK != CB200_ZP_NONE && !p.zp_scalar;
  const int stage_bytes = geo.stage_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ring; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 5);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_ready(s), 4);
      mbar_init(d_full(s), 1);
      mbar_init(d_empty(s), 4);
      mbar_init(xs_ready(s), 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(xb_full(s, b), 1);
        mbar_init(mma_done(s, b), 1);
      }
    }
    fence_barrier_init();
    prefetch_tensormap(&map_w);
    prefetch_tensormap(&map_x);
  }
  // rows >= M of the x operand slots are never written by the x warps: zero them once (their MMA columns are never read back,
  // but a NaN pattern in uninitialised shared memory must not meet a real weight either way)
  for (int ofs = threadIdx.x * 16; ofs < 4 * XB_BYTES; ofs += kTcsThreads * 16) {
    *reinterpret_cast<uint4*>(xb_ptr + ofs) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 1) tmem_alloc<kTcsTmemCols>(tmem_slot);
  fence_proxy_async_smem();  // the zero fill above is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + 1016);
  if (threadIdx.x == 0) tcs_stamp(geo.trace, 1);

  if (warp == 0) {
    // =========================================== TMA producer ===========================================================
    // Weights, scales and zero-points are static: the first `ring` stages are requested BEFORE griddepcontrol.wait (under
    // programmatic dependent launch: while the previous kernel in the stream is still running).  x belongs to the stream order.
    if (elect_one()) {
      TcsCursor c;
      c.init(run_a, n_run, spc, geo.spg, 0);
      auto issue_static = [&](int slot) {
        const uint32_t sbase = ring0 + uint32_t(slot) * uint32_t(stage_bytes);
        const int cols = min(128, N - c.cb * 128);
        const uint32_t meta_bytes = uint32_t(cols) * (2u + (zp_per_group ? ZES : 0u));
        mbar_expect_tx(full(slot), uint32_t(kTcsStageW) + uint32_t(geo.x_bytes) + meta_bytes);
        tma_load_2d(sbase, &map_w, c.cb * 128, c.kt * 16, full(slot));  // columns beyond N are zero-filled
        const size_t idx = size_t(c.gidx) * N + size_t(c.cb) * 128;
        const uint32_t dst = sbase + uint32_t(geo.meta_off);
        tcs_bulk_g2s(dst, reinterpret_cast<const A*>(p.s) + idx, uint32_t(cols) * 2u, full(slot));
        if (zp_per_group) tcs_bulk_g2s(dst + 256u, reinterpret_cast<const char*>(p.zp) + idx * ZES, uint32_t(cols) * ZES, full(slot));
      };
      const int pre = min(ring, n_run);
      for (int i = 0; i < pre; ++i) {
        issue_static(i);
        c.step1();
      }
      tcs_stamp(geo.trace, 2);
      if (geo.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      tcs_stamp(geo.trace, 3);
      {
        TcsCursor cx;
        cx.init(run_a, n_run, spc, geo.spg, 0);
        for (int i = 0; i < pre; ++i) {
          tma_load_2d(ring0 + uint32_t(i) * uint32_t(stage_bytes) + uint32_t(geo.x_off), &map_x, cx.kt * KS, 0, full(i));
          cx.step1();
        }
      }
      int slot = 0;
      uint32_t par = 0;
      for (int i = pre; i < n_run; ++i) {
        tcs_wait(empty(slot), par);
        issue_static(slot);
        tma_load_2d(ring0 + uint32_t(slot) * uint32_t(stage_bytes) + uint32_t(geo.x_off), &map_x, c.kt * KS, 0, full(slot));
        c.step1();
        if (++slot == ring) {
          slot = 0;
          par ^= 1u;
        }
      }
      tcs_stamp(geo.trace, 4);
    }
  } else if (warp <= 2) {
    // =========================================== MMA issuer of set (warp - 1) ===========================================
    // D^T[n, m] (+)= W^T[n, k] x[m, k]: the 128 TMEM lanes are the stage's 128 weight columns, MMA-N = MT rows of x.
    // A comes from tensor memory (written by the set's dequant warps), B from the set's x operand slots in shared memory.
    const int set = warp - 1;
    constexpr uint32_t fmt = HALF ? 0u : 1u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(MT >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t d_e = tmem_base + uint32_t(set * DW), d_o = d_e + MT;
    const uint32_t a_col = tmem_base + uint32_t(A_COL0 + set * 64);
    TcsCursor c;
    c.init(run_a, n_run, spc, geo.spg, set);
    uint32_t groups = 0;          // groups started on this set's accumulator
    bool fresh = true;            // the next stage opens a group
    // operand descriptors of the set's two x operand slots (x NH halves): constants of the whole run
    uint64_t bdesc[2][NH];
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int h = 0; h < NH; ++h) bdesc[b][h] = make_sw128_desc(xb0 + uint32_t(2 * set + b) * XB_BYTES + uint32_t(h) * XB_HALF);
    // one stage = NH x 4 K=16 MMAs, straight-line, accumulate flags known at compile time (FR: the stage opens a group)
    auto issue_stage = [&](auto fr_tag, int b) {
      constexpr bool FR = decltype(fr_tag)::value;
#pragma unroll
      for (int h = 0; h < NH; ++h) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {  // four K=16 MMAs per 64-k half; EO: 0,1 plain nibbles, 2,3 the x16 kind
          const uint32_t a_addr = a_col + uint32_t(h * 32 + k4 * 8);
          const uint64_t bd = (b ? bdesc[1][h] : bdesc[0][h]) + uint64_t(k4 * 2);
          if (EO && k4 >= 2) {
            umma_ts(d_o, a_addr, bd, idesc, (FR && h == 0 && k4 == 2) ? 0u : 1u);
          } else {
            umma_ts(d_e, a_addr, bd, idesc, (FR && h == 0 && k4 == 0) ? 0u : 1u);
          }
        }
      }
    };
    TcsPhase ph;
    ph.begin(geo.trace != nullptr && set == 0 && lane == 0);
    for (int j = 0; c.valid(); ++j, c.next()) {
      const int b = j & 1;
      {
        // lanes 0..2 probe the three barriers at once (a probe costs ~100 cycles even when the barrier has completed)
        const bool need_d = fresh && groups > 0;
        if (lane == 0) tcs_wait(a_ready(set), uint32_t(j) & 1u);
        if (lane == 1) tcs_wait(xb_full(set, b), uint32_t(j >> 1) & 1u);
        if (lane == 2 && need_d) tcs_wait(d_empty(set), (groups - 1u) & 1u);
        __syncwarp();
      }
      ph.mark(0);
      const bool last = c.last();
      tc_fence_after();
      if (elect_one()) {
        if (fresh) {
          issue_stage(std::true_type{}, b);
        } else {
          issue_stage(std::false_type{}, b);
        }
        if (last) umma_commit(d_full(set));
        umma_commit(mma_done(set, b));
      }
      __syncwarp();
      ph.mark(1);
      if (fresh) ++groups;
      fresh = last;
    }
    ph.dump(geo.trace, 10);
  } else if (warp <= 4) {
    // =========================================== x operand warp of set (warp - 3) =======================================
    // Rewrites the raw [M x KS] slice of x the TMA delivered into the K-major, 128-byte-swizzled B operand -- in the k order
    // the mask-only nibble extraction produces -- and accumulates X[m] = sum_k x[m, k] over the set's stages of the open
    // scale group in fp32.
    const int set = warp - 3;
    constexpr int CPR = KS / 8;        // 16-byte chunks per row
    constexpr int RPI = 32 / CPR;      // rows per iteration: 2 | 4
    constexpr int NIT = MT / RPI;
    const int c = lane % CPR, rsub = lane / CPR;
    const int nit = min(NIT, (M + RPI - 1) / RPI);  // iterations that touch an existing row
    float xacc[NIT];  // per lane: sum over this lane's chunk position, all stages of the open group (reduced across lanes at its end)
#pragma unroll
    for (int it = 0; it < NIT; ++it) xacc[it] = 0.f;
    TcsCursor cur;
    cur.init(run_a, n_run, spc, geo.spg, set);
    uint32_t groups = 0;
    int slot = set;     // ring slot of this set's current stage (the ring depth is even: a slot always belongs to the same set)
    uint32_t fpar = 0;  // phase parity of its `full` barrier
    TcsPhase ph;
    ph.begin(geo.trace != nullptr && set == 0 && lane == 0);
    for (int j = 0; cur.valid(); ++j, cur.next()) {
      const int b = j & 1;
      if (lane == 0) tcs_wait(full(slot), fpar);
      if (lane == 1 && j >= 2) tcs_wait(mma_done(set, b), uint32_t((j >> 1) - 1) & 1u);  // the MMAs that read this operand slot retired
      __syncwarp();
      ph.mark(0);
      const unsigned char* xr = ring_ptr + slot * stage_bytes + geo.x_off;
      unsigned char* xb = xb_ptr + (2 * set + b) * XB_BYTES;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        if (it >= nit) break;
        const int m = it * RPI + rsub;
        if (m < M) {
          const uint4 v = *reinterpret_cast<const uint4*>(xr + m * (KS * 2) + c * 16);
          {
            float2 f0, f1, f2, f3;
            if constexpr (HALF) {
              f0 = __half22float2(bits_h2(v.x)); f1 = __half22float2(bits_h2(v.y));
              f2 = __half22float2(bits_h2(v.z)); f3 = __half22float2(bits_h2(v.w));
            } else {
              f0 = __bfloat1622float2(bits_b2(v.x)); f1 = __bfloat1622float2(bits_b2(v.y));
              f2 = __bfloat1622float2(bits_b2(v.z)); f3 = __bfloat1622float2(bits_b2(v.w));
            }
            xacc[it] += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
          }
          const int h = c >> 3, c8 = c & 7;
          unsigned char* row = xb + h * XB_HALF + m * 128;
          const uint32_t sw = uint32_t(m & 7);
          if constexpr (BITS == 4) {
            const uint32_t k04 = __byte_perm(v.x, v.z, 0x5410), k15 = __byte_perm(v.x, v.z, 0x7632);
            const uint32_t k26 = __byte_perm(v.y, v.w, 0x5410), k37 = __byte_perm(v.y, v.w, 0x7632);
            if constexpr (EO) {
              // row = [plain of words 0-3 | plain of words 4-7 | x16 of words 0-3 | x16 of words 4-7], 8 bytes per word and kind
              const uint32_t o1 = uint32_t((c8 < 4 ? 0 : 32) + (c8 & 3) * 8), o2 = o1 + 64u;
              *reinterpret_cast<uint2*>(row + ((((o1 >> 4) ^ sw) << 4) | (o1 & 15u))) = make_uint2(k04, k26);
              *reinterpret_cast<uint2*>(row + ((((o2 >> 4) ^ sw) << 4) | (o2 & 15u))) = make_uint2(k15, k37);
            } else {
              *reinterpret_cast<uint4*>(row + ((uint32_t(c8) ^ sw) << 4)) = make_uint4(k04, k15, k26, k37);
            }
          } else {
            *reinterpret_cast<uint4*>(row + ((uint32_t(c8) ^ sw) << 4)) = v;  // int8 words unpack in natural k order
          }
        }
      }
      ph.mark(1);
      fence_proxy_async_smem();  // generic-proxy writes of the operand -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(xb_full(set, b));
        mbar_arrive(empty(slot));
      }
      if (cur.last()) {
        // the operand is on its way; now the group's X sums.  The buffer is single: the fold of the set's previous group must
        // have read it (one phase behind at most: a parity wait is only meaningful then).
        if (groups > 0) tcs_wait(d_empty(set), (groups - 1u) & 1u);
        float* xs = xsum + set * MT;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          if (it >= nit) break;
          // the row's CPR per-position sums, always in the same tree order whatever M is (batch-invariant)
          float v = xacc[it];
#pragma unroll
          for (int d = CPR / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
          if (c == 0) xs[it * RPI + rsub] = v;
          xacc[it] = 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(xs_ready(set));
        ++groups;
      }
      ph.mark(2);
      slot += 2;
      if (slot >= ring) {
        slot -= ring;
        fpar ^= 1u;
      }
    }
    ph.dump(geo.trace, 16);
  } else {
    // =========================================== dequant + fold warps ===================================================
    const int set = (warp - 5) >> 2;   // takes stages set, set + 2, ... of the run
    const int q = warp & 3;            // TMEM lane quarter this warp may touch (hardware: warp id % 4)
    const int nl = q * 32 + lane;      // weight column inside the stage == TMEM lane
    const bool tracer = geo.trace != nullptr && warp == 5 && lane == 0;
    if (geo.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");  // out / workspace belong to the stream order
    if (tracer) tcs_stamp(geo.trace, 5);

    float zscalar = 0.f;
    if (ZK != CB200_ZP_NONE && p.zp_scalar) {
      if constexpr (ZK == CB200_ZP_I32) zscalar = float(__ldg(reinterpret_cast<const int32_t*>(p.zp)));
      if constexpr (ZK == CB200_ZP_U8) zscalar = float(__ldg(reinterpret_cast<const uint8_t*>(p.zp)));
      if constexpr (ZK == CB200_ZP_ACT) zscalar = ActOps<A>::to_float(*reinterpret_cast<const A*>(p.zp));
    }
    const float fbias = float(p.bias);
    constexpr float MAGIC = HALF ? 0.f : 128.f;         // bf16 operands are 128 + q; fp16 operands are q * 2^-24 (subnormals)
    constexpr float OPSCALE = HALF ? 16777216.f : 1.f;  // 2^24 undoes the subnormal scaling

    float acc[MR];  // this column's running sums for the current column block (rows >= M stay zero)
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m] = 0.f;
    int cur_cb = run_a / spc;
    const int last_cb = (run_a + n_run - 1) / spc;
    // the group share this set still has to fold: column block + scale / zero-point terms of this thread's column
    bool pending = false;
    int pend_cb = 0;
    float pend_sa = 0.f, pend_sb = 0.f;
    uint32_t folds = 0;  // groups folded by this set (phase of its d_full barrier)

    auto publish = [&](int cb) {  // this set's partial of column block cb -> workspace; the last arriver reduces and stores
      const int col0 = cb * 128 + 32 * q;
      if (col0 < N) {
        const int first_cta = tcs_cta_of(cb * spc, geo.q, geo.r);
        const int last_cta = tcs_cta_of(cb * spc + spc - 1, geo.q, geo.r);
        const int nslots = (last_cta - first_cta + 1) * kTcsSets;
        const int myslot = (cta - first_cta) * kTcsSets + set;
        float* base = ws_partial + size_t(cb) * size_t(geo.maxslots) * (M * 128);
        float* mine = base + size_t(myslot) * (M * 128);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
          if (m < M) __stcg(mine + m * 128 + nl, acc[m]);
        }
        __threadfence();
        __syncwarp();
        unsigned prev = 0;
        if (lane == 0) prev = atomicAdd(ws_counter + cb * 4 + q, 1u);
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == unsigned(nslots - 1)) {
          __threadfence();  // last arriver: every other partial of this quarter is visible now
          const bool n_ok = col0 + lane < N;
          for (int m = 0; m < M; ++m) {
            float v = 0.f;
            for (int s = 0; s < nslots; ++s) v += __ldcg(base + size_t(s) * (M * 128) + m * 128 + nl);
            if (n_ok) store_out_all<A>(p, size_t(m) * N + col0 + lane, v);
          }
          if (lane == 0) ws_counter[cb * 4 + q] = 0u;  // leave the counter clean for the next launch
        }
      }
#pragma unroll
      for (int m = 0; m < MR; ++m) acc[m] = 0.f;
    };
    auto fold = [&]() {  // acc[m] += sa * G[n, m] + sb * X[m] for the pending group share, from this set's TMEM accumulator
      while (cur_cb < pend_cb) publish(cur_cb++);
      tcs_wait(d_full(set), folds & 1u);
      tcs_wait(xs_ready(set), folds & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(set * DW);
      const float* xs = xsum + set * MT;
      // four rows of x at a time (keeps the live registers of a 16- or 32-row fold small)
#pragma unroll
      for (int m0 = 0; m0 < MR; m0 += 4) {
        if (m0 < M) {
          uint32_t ge[32], go[32];
          tmem_ld4(taddr + uint32_t(m0), ge);
          if constexpr (EO) tmem_ld4(taddr + uint32_t(MT + m0), go);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (m0 + i < M) {
              float g = __uint_as_float(ge[i]);
              if constexpr (EO) g = fmaf(__uint_as_float(go[i]), 0.0625f, g);  // the odd nibbles were 16 q: exact power-of-two scaling
              acc[m0 + i] = fmaf(pend_sa, g, fmaf(pend_sb, xs[m0 + i], acc[m0 + i]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty(set));
      ++folds;
      pending = false;
    };

    TcsCursor c;
    c.init(run_a, n_run, spc, geo.spg, set);
    int slot = set;     // ring slot of this set's current stage (the ring depth is even: a slot always belongs to the same set)
    uint32_t fpar = 0;  // phase parity of its `full` barrier
    TcsPhase ph;
    ph.begin(tracer);
    for (int j = 0; c.valid(); ++j, c.next()) {
      tcs_wait(full(slot), fpar);
      ph.mark(1);
      if (tracer && j == 0) tcs_stamp(geo.trace, 6);
      const unsigned char* sp = ring_ptr + slot * stage_bytes;
      // this column's 16 packed words (one per packed row) and -- if the set's share of the group ends here -- its scale / zero-point
      uint32_t w[2][8];
      uint32_t (&w2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&w[0][0]);  // (int8: one half of 16 words)
#pragma unroll
      for (int r = 0; r < 16; ++r) w2[r] = *reinterpret_cast<const uint32_t*>(sp + r * 512 + nl * 4);
      const bool ends = c.last();
      float new_sa = 0.f, new_sb = 0.f;
      if (ends) {
        const unsigned char* row = sp + geo.meta_off;
        const float sj = ActOps<A>::to_float(*reinterpret_cast<const A*>(row + nl * 2));
        float z = zscalar;
        if (zp_per_group) {
          if constexpr (ZK == CB200_ZP_I32) z = float(*reinterpret_cast<const int32_t*>(row + 256 + nl * 4));
          if constexpr (ZK == CB200_ZP_U8) z = float(*reinterpret_cast<const uint8_t*>(row + 256 + nl));
          if constexpr (FZP) z = ActOps<A>::to_float(*reinterpret_cast<const A*>(row + 256 + nl * 2));
        }
        new_sa = sj * OPSCALE;
        new_sb = -sj * (MAGIC + z + fbias);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(slot));  // the packed tile is in registers: the producer may refill the slot
      ph.mark(2);
      // the set's single A operand slot: the MMAs of its previous stage must have retired (they were kicked off a whole
      // stage load ago: this wait is normally free)
      if (j >= 1) tcs_wait(mma_done(set, (j - 1) & 1), uint32_t((j - 1) >> 1) & 1u);
      tc_fence_after();
      ph.mark(3);
      constexpr int WR = 16 / NH;  // packed rows per half: 8 | 16  (int8: the single half spans all 16 rows)
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        uint32_t cells[32];  // 64 k of this column, two per 32-bit TMEM cell, in the order the MMAs consume them
#pragma unroll
        for (int jj = 0; jj < WPH; ++jj) {
          const uint32_t word = WR == 8 ? w[h][jj & 7] : w2[jj];
          if constexpr (EO) {  // fp16: 0x0000 | v read as fp16 is the subnormal v * 2^-24, exactly
            const uint32_t t8 = word >> 8;
            cells[2 * jj + 0] = word & 0x000f000fu;       // (k0, k4)      plain: MMAs 0, 1
            cells[2 * jj + 1] = t8 & 0x000f000fu;         // (k2, k6)
            cells[16 + 2 * jj + 0] = word & 0x00f000f0u;  // (16 k1, 16 k5) x16: MMAs 2, 3
            cells[16 + 2 * jj + 1] = t8 & 0x00f000f0u;    // (16 k3, 16 k7)
          } else if constexpr (BITS == 4) {  // bf16: 0x4300 | q = 128 + q, cells (k0,k4) (k1,k5) (k2,k6) (k3,k7)
#pragma unroll
            for (int i = 0; i < 4; ++i) cells[4 * jj + i] = lop3_and_or(word >> (4 * i), 0x000f000fu, 0x43004300u);
          } else {  // fp16 int8: bytes as subnormals, natural order
            cells[2 * jj + 0] = prmt(word, 0u, 0x4140u);
            cells[2 * jj + 1] = prmt(word, 0u, 0x4342u);
          }
        }
        tmem_st32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(A_COL0 + set * 64 + h * 32), cells);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(set));
      ph.mark(4);
      // Fold the group share that ended at an EARLIER stage of this set -- only now, after this stage's operand is on its way
      // to the tensor core: its accumulator is final (d_full was committed before the mma_done waited for above), so the fold
      // never waits, and the issuer (which needs d_empty before it may open the next group) gets it right behind a_ready.
      if (pending) fold();
      if (ends) {
        pend_sa = new_sa;
        pend_sb = new_sb;
        pend_cb = c.cb;
        pending = true;
      }
      ph.mark(0);
      slot += 2;
      if (slot >= ring) {
        slot -= ring;
        fpar ^= 1u;
      }
    }
    if (pending) fold();
    ph.mark(0);
    ph.dump(geo.trace, 22);
    if (tracer) tcs_stamp(geo.trace, 7);
    while (cur_cb <= last_cb) publish(cur_cb++);
    if (tracer) {
      tcs_stamp(geo.trace, 8);
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      geo.trace[blockIdx.x * 32 + 9] = sm;
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTcsTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <class A, int BITS, int MT, int MR, int ZK>
static int launch_tcs_cfg(const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const TcsGeom& geo, int ctas, size_t smem,
                          float* wsp, unsigned* wsc, cudaStream_t stream) {
  auto kern = tcs_kernel<A, BITS, MT, MR, ZK>;
  CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kTcsThreads);
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

template <class A, int BITS, int MT, int MR>
static int launch_tcs_zk(const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const TcsGeom& geo, int ctas, size_t smem,
                         float* wsp, unsigned* wsc, cudaStream_t stream) {
  switch (p.zp_kind) {
    case CB200_ZP_NONE: return launch_tcs_cfg<A, BITS, MT, MR, CB200_ZP_NONE>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
    case CB200_ZP_I32: return launch_tcs_cfg<A, BITS, MT, MR, CB200_ZP_I32>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
    case CB200_ZP_U8: return launch_tcs_cfg<A, BITS, MT, MR, CB200_ZP_U8>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
    default: return launch_tcs_cfg<A, BITS, MT, MR, CB200_ZP_ACT>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
  }
}

template <class A, int BITS>
static int launch_tcs_typed(int mt, const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const TcsGeom& geo, int ctas,
                            size_t smem, float* wsp, unsigned* wsc, cudaStream_t stream) {
  if (mt == 32) return launch_tcs_zk<A, BITS, 32, 32>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
  return geo.m_rows <= 4 ? launch_tcs_zk<A, BITS, 16, 4>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream)
                         : launch_tcs_zk<A, BITS, 16, 16>(mw, mx, p, geo, ctas, smem, wsp, wsc, stream);
}

#define CB200_TCS_ENTRY(name)                                                                                                     \
  int name(int mt, const CUtensorMap& mw, const CUtensorMap& mx, const Problem& p, const TcsGeom& geo, int ctas, size_t smem, \
           float* wsp, unsigned* wsc, cudaStream_t stream)
CB200_TCS_ENTRY(launch_tcs_f16_w4);
CB200_TCS_ENTRY(launch_tcs_f16_w8);
CB200_TCS_ENTRY(launch_tcs_bf16_w4);

}  // namespace cb200
