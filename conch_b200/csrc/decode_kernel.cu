// Decode-regime kernel (M <= 16 per pass): stream-K skinny GEMM, HBM-bound.
//
// Replaces the reference's single 128x64x64 Triton tile (conch/kernels/quantization/gemm.py:136-145, 491-494),
// which at M=1 launches N/64 CTAs with no split-K.  Design (B200):
//   * work = (32-column block) x (batch of U steps, one step = 4 packed rows) units, flattened column-major and cut
//     into W equal contiguous ranges, one per WARP (stream-K at warp granularity).  W is chosen so that the whole grid
//     is one co-resident wave; warps are completely independent (no shared memory, no __syncthreads);
//   * the packed weights are the only large stream: every lane issues 128-bit coalesced, read-once cp.async.cg copies
//     (L2 -> shared, bypassing L1 and the register file) into a private per-warp ring LOOK batches ahead of the one
//     being consumed, so ~100 KB per SM is in flight regardless of register pressure; a lane only ever reads back
//     its own 16 bytes, so the ring needs no barrier, just cp.async.wait_group;
//   * nibbles are unpacked and the exact integer (q - bias - zp) is formed in registers with the magic-number tricks
//     of dequant.cuh; with SA=false the scale multiply is done per weight (dequant_word(): bit-identical W to
//     the dequant kernel and the reference's w_ref), with SA=true the per-group scale is applied once to the
//     fp32 group sum instead (sum_k x_k (q_k - zp) is exact in the tensor core, so this differs from the reference
//     only by NOT rounding each W to fp16/bf16 -- well inside the 1e-2 tolerance, and ~25% fewer instructions on a
//     kernel that is issue-bound right at the HBM roofline);
//   * the dot product runs on the tensor pipe as mma.sync.m16n8k16 with the WEIGHTS as the 16-row operand (swap-AB) and
//     x^T as the 8-column operand; the four lanes of a quad hold four different packed rows, so the K-reduction
//     across lanes happens inside the MMA, fp32 accumulate;
//   * x is tiny and shared by every warp of the SM: read through L1 (16 B per lane per step) and permuted in registers
//     to the nibble order the cheap `lop3` extraction produces ((k0,k4),(k1,k5),(k2,k6),(k3,k7));
//   * a warp whose range covers a whole column block stores the result directly; otherwise it writes an fp32 partial,
//     bumps the block's arrival counter, and the last arriver sums the partials in slot order (deterministic,
//     self-resetting counters: no memset, no second launch).
#include "common.cuh"
#include "dequant.cuh"

namespace cb200 {

DecodeTuning g_decode_tuning = {0, 0, 0, 0};

struct DecodeGeom {
  int m_rows;         // rows of x in this pass (<= 16)
  int units_per_col;  // B: batches per column block = K / (U * STEP_K)
  int q, r;           // warp w owns q + (w < r) consecutive units starting at w*q + min(w, r)
  int total_warps;    // W
  int col_blocks;     // N / 32
  int slots_max;      // partial slots reserved per column block
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

constexpr int kDecodeRingSteps = 12;  // weight ring depth per warp, in steps (512 B each): 6 KB per warp

// unit index -> owning warp (inverse of the range function above)
__device__ __forceinline__ int unit_owner(int u, int q, int r) {
  const int head = r * (q + 1);
  return u < head ? u / (q + 1) : r + (u - head) / q;
}

// A: activation type, BITS: 4|8, MT: 8-row m tiles (1: M<=8, 2: M<=16), U: steps per batch (group boundaries fall on
// batch boundaries), ZK: cb200_zp_kind (compile time), SA: apply the group scale after the group's dot product.
template <class A, int BITS, int MT, int U, int ZK, bool SA>
__global__ void __launch_bounds__(128, MT == 1 ? 6 : 4)
    decode_kernel(Problem p, DecodeGeom geo, float* __restrict__ ws_partial, unsigned int* __restrict__ ws_counter) {
  constexpr int PF = 32 / BITS;     // k per packed word
  constexpr int STEP_K = 4 * PF;    // k per step (4 packed rows, one per lane of a quad)
  constexpr int VEC = 4;            // packed words per lane per load (128-bit)
  constexpr int NT = VEC / 2;       // 16-column MMA tiles per warp
  constexpr int CW = 8 * VEC;       // columns per warp
  constexpr int PAIRS = 16 / BITS;  // pair registers per word
  constexpr Order ORD = Order::Lop;
  constexpr bool FZP = ZK == CB200_ZP_ACT;
  constexpr int LOOK = kDecodeRingSteps / U - 1;  // batches in flight ahead of the one being consumed
  constexpr int SLOTS = LOOK + 1;

  extern __shared__ __align__(16) unsigned char ring_raw[];  // [warps][SLOTS][U][32 lanes][16 B]

  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= geo.total_warps) return;
  const int g8 = lane >> 2, tig = lane & 3;
  const int M = geo.m_rows;
  const int N = p.N;
  const int B = geo.units_per_col;

  const int n_units = geo.q + (w < geo.r ? 1 : 0);
  const int u0 = w * geo.q + min(w, geo.r);
  int cb = u0 / B;       // column block
  int b = u0 - cb * B;   // batch inside the column block
  const int lane_col = VEC * g8;

  // ---- weight ring: get LOOK batches in flight before anything else ------------------------------------------------
  const size_t step_stride = size_t(4) * N;      // words between consecutive steps of one lane
  const size_t unit_stride = step_stride * U;    // words between consecutive batches
  const uint32_t ring = static_cast<uint32_t>(__cvta_generic_to_shared(ring_raw)) +
                        (threadIdx.x >> 5) * (SLOTS * U * 512) + lane * 16;
  // issue cursor: runs LOOK units ahead of the consume cursor (cb, b)
  int icb = cb, ib = b;
  const uint32_t* iwp = p.wq + (size_t(b) * (U * 4) + tig) * N + cb * CW + lane_col;
  auto issue_unit = [&](int slot) {
    const uint32_t dst = ring + slot * (U * 512);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * 512), "l"(iwp + u * step_stride) : "memory");
    }
    ++ib;
    iwp += unit_stride;
    if (ib == B) {
      ib = 0;
      ++icb;
      iwp = p.wq + size_t(tig) * N + icb * CW + lane_col;
    }
  };
#pragma unroll
  for (int j = 0; j < LOOK; ++j) {
    if (j < n_units) issue_unit(j);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- per-group constants ---------------------------------------------------------------------------------------
  const int bpg = p.g / (U * STEP_K);  // batches per group (host guarantees divisibility)
  const int ngrp = p.K / p.g;
  A raw_s[VEC];
  float raw_z[VEC];
  A raw_zf[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    raw_z[j] = 0.f;
    raw_zf[j] = A(0.f);
  }
  if (ZK != CB200_ZP_NONE && p.zp_scalar) {  // one zero-point for the whole matrix: fetch it once
    float z = 0.f;
    A zf = A(0.f);
    if constexpr (ZK == CB200_ZP_I32) z = float(__ldg(reinterpret_cast<const int32_t*>(p.zp)));
    if constexpr (ZK == CB200_ZP_U8) z = float(__ldg(reinterpret_cast<const uint8_t*>(p.zp)));
    if constexpr (ZK == CB200_ZP_ACT) zf = *reinterpret_cast<const A*>(p.zp);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      raw_z[j] = z;
      raw_zf[j] = zf;
    }
  }
  const bool zp_per_group = ZK != CB200_ZP_NONE && !p.zp_scalar;
  // meta = element index of (group, first column of this lane) in the [K/g, N] scale / zero-point matrices
  auto fetch_meta = [&](uint32_t meta) {
    *reinterpret_cast<uint2*>(raw_s) = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const A*>(p.s) + meta));
    if (zp_per_group) {
      if constexpr (ZK == CB200_ZP_I32) {
        const int4 z = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const int32_t*>(p.zp) + meta));
        raw_z[0] = float(z.x); raw_z[1] = float(z.y); raw_z[2] = float(z.z); raw_z[3] = float(z.w);
      } else if constexpr (ZK == CB200_ZP_U8) {
        const uint32_t z = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(p.zp) + meta));
#pragma unroll
        for (int j = 0; j < VEC; ++j) raw_z[j] = float((z >> (8 * j)) & 0xffu);
      } else if constexpr (ZK == CB200_ZP_ACT) {
        *reinterpret_cast<uint2*>(raw_zf) = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const A*>(p.zp) + meta));
      }
    }
  };
  int grp = b / bpg;
  int batches_left = bpg - (b - grp * bpg);  // batches until the current group ends
  bool refresh = true;
  uint32_t meta = uint32_t(grp) * N + cb * CW + lane_col;
  fetch_meta(meta);
  ColQ cq[VEC];
  float sc[VEC];  // SA: group scales as fp32
  const float fbias = float(p.bias);

  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  float acc[NT][MT][4];   // running sum of this column segment
  float accg[NT][MT][4];  // SA: sum of the current group, still unscaled
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][mt][i] = accg[t][mt][i] = 0.f;
  bool group_open = false;  // SA: accg holds a partial group

  // x rows this lane feeds to the MMA B operand (column g8 of the 8-wide m tile); xp follows the current batch
  const A* xp[MT];
  bool m_ok[MT];
  uint32_t xb[MT][PAIRS];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int m = mt * 8 + g8;
    m_ok[mt] = m < M;
    xp[mt] = reinterpret_cast<const A*>(p.x) + size_t(m_ok[mt] ? m : 0) * p.K + tig * PF + size_t(b) * (U * 4 * PF);
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) xb[mt][i] = 0u;
  }

  auto close_group = [&]() {  // SA: acc += scale (per weight column) * group sum
    // rows of the C fragment are weight columns: c0,c1 <- column 2t, c2,c3 <- column 2t+1 of this lane
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        acc[t][mt][0] = fmaf(accg[t][mt][0], sc[2 * t], acc[t][mt][0]);
        acc[t][mt][1] = fmaf(accg[t][mt][1], sc[2 * t], acc[t][mt][1]);
        acc[t][mt][2] = fmaf(accg[t][mt][2], sc[2 * t + 1], acc[t][mt][2]);
        acc[t][mt][3] = fmaf(accg[t][mt][3], sc[2 * t + 1], acc[t][mt][3]);
      }
  };

  // ---- flush one finished (or cut) column segment ------------------------------------------------------------------
  auto flush = [&](int cbi) {
    const int first_w = unit_owner(cbi * B, geo.q, geo.r);
    const int last_w = unit_owner(cbi * B + B - 1, geo.q, geo.r);
    const int ncol0 = cbi * CW + lane_col;
    if (first_w == last_w) {  // this warp saw the whole K range: store directly
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m = mt * 8 + 2 * tig;
          const int n = ncol0 + 2 * t;
          if (m < M) {
            store_out<A>(p.out, p.out_dtype, size_t(m) * N + n, acc[t][mt][0]);
            store_out<A>(p.out, p.out_dtype, size_t(m) * N + n + 1, acc[t][mt][2]);
          }
          if (m + 1 < M) {
            store_out<A>(p.out, p.out_dtype, size_t(m + 1) * N + n, acc[t][mt][1]);
            store_out<A>(p.out, p.out_dtype, size_t(m + 1) * N + n + 1, acc[t][mt][3]);
          }
        }
      return;
    }
    const int contributors = last_w - first_w + 1;
    float* mine = ws_partial + (size_t(cbi) * geo.slots_max + (w - first_w)) * (M * CW);
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int m = mt * 8 + 2 * tig;
        const int lc = lane_col + 2 * t;
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
    if (lane == 0) prev = atomicAdd(ws_counter + cbi, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != unsigned(contributors - 1)) return;
    __threadfence();  // last arriver: every other contributor's partial is visible now
    const float* base = ws_partial + size_t(cbi) * geo.slots_max * (M * CW);
    for (int e = lane; e < M * CW; e += 32) {
      float v = 0.f;
      for (int s = 0; s < contributors; ++s) v += __ldcg(base + size_t(s) * (M * CW) + e);
      const int m = e / CW, lc = e - m * CW;
      store_out<A>(p.out, p.out_dtype, size_t(m) * N + cbi * CW + lc, v);
    }
    if (lane == 0) ws_counter[cbi] = 0u;  // leave the counter clean for the next launch
  };

  // ---- one unit: U steps of dequant + MMA on the weights already in registers ------------------------------------------
  auto process = [&](const uint4 (&wreg)[U]) {
    if (refresh) {  // group boundary (warp-uniform): rebuild the per-column constants, prefetch the next group's
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        cq[j] = make_colq<A, BITS, ORD, FZP>(raw_s[j], raw_z[j] + fbias, raw_zf[j]);
        sc[j] = ActOps<A>::to_float(raw_s[j]);
      }
      refresh = false;
      // next group: next row of the metadata, or row 0 of the next column block
      uint32_t nmeta = meta + N;
      if (grp + 1 == ngrp) nmeta = uint32_t(cb + 1) * CW + lane_col;
      if (grp + 1 < ngrp || cb + 1 < geo.col_blocks) fetch_meta(nmeta);
      meta = nmeta;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // activations of this step: 16 B (8 B for int8) per lane through L1, permuted to the lop3 nibble order
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (m_ok[mt]) {
          if constexpr (BITS == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(xp[mt] + u * (4 * PF)));
            xb[mt][0] = __byte_perm(v.x, v.z, 0x5410);  // (k0,k4)
            xb[mt][1] = __byte_perm(v.x, v.z, 0x7632);  // (k1,k5)
            xb[mt][2] = __byte_perm(v.y, v.w, 0x5410);  // (k2,k6)
            xb[mt][3] = __byte_perm(v.y, v.w, 0x7632);  // (k3,k7)
          } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(xp[mt] + u * (4 * PF)));
            xb[mt][0] = v.x;
            xb[mt][1] = v.y;
          }
        }
      }
      const uint32_t words[VEC] = {wreg[u].x, wreg[u].y, wreg[u].z, wreg[u].w};
      uint32_t d[VEC][PAIRS];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if constexpr (SA) {
          dequant_word_int<A, BITS, ORD, FZP>(words[j], cq[j], d[j]);
        } else {
          dequant_word<A, BITS, ORD, FZP>(words[j], cq[j], d[j]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          // rows g8 of the MMA <- column 2t of this lane, rows g8+8 <- column 2t+1
#pragma unroll
          for (int h = 0; h < PAIRS / 2; ++h) {
            if constexpr (SA) {
              if (!group_open && u == 0 && h == 0) {
                mma16816<A>(accg[t][mt], zero4, d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1],
                            d[2 * t + 1][2 * h + 1], xb[mt][2 * h], xb[mt][2 * h + 1]);
              } else {
                mma16816<A>(accg[t][mt], accg[t][mt], d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1],
                            d[2 * t + 1][2 * h + 1], xb[mt][2 * h], xb[mt][2 * h + 1]);
              }
            } else {
              mma16816<A>(acc[t][mt], acc[t][mt], d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1],
                          d[2 * t + 1][2 * h + 1], xb[mt][2 * h], xb[mt][2 * h + 1]);
            }
          }
        }
      }
    }
    group_open = true;
    if (--batches_left == 0) {  // group finished
      if constexpr (SA) close_group();
      group_open = false;
      batches_left = bpg;
      refresh = true;
      ++grp;
    }
  };

  // advance (cb, b, xp) to the next unit; returns true when that unit starts a new column block
  auto advance = [&]() {
    ++b;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) xp[mt] += U * 4 * PF;
    if (b != B) return false;
    b = 0;
    ++cb;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) xp[mt] -= p.K;
    return true;
  };
  auto end_column = [&](int cbi) {  // always on a group boundary
    flush(cbi);
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[t][mt][k] = 0.f;
    grp = 0;
  };

  // ---- main loop: unit i is consumed from ring slot i % SLOTS while unit i + LOOK is being fetched ---------------------
  int slot = 0, islot = LOOK % SLOTS;
  for (int i = 0; i < n_units; ++i) {
    if (i + LOOK < n_units) issue_unit(islot);
    asm volatile("cp.async.commit_group;" ::: "memory");  // always commit (possibly empty): keeps the group count static
    islot = islot + 1 == SLOTS ? 0 : islot + 1;
    asm volatile("cp.async.wait_group %0;" ::"n"(LOOK) : "memory");  // everything but the newest LOOK groups has landed
    uint4 wreg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(wreg[u].x), "=r"(wreg[u].y), "=r"(wreg[u].z), "=r"(wreg[u].w)
                   : "r"(ring + (slot * U + u) * 512));
    }
    slot = slot + 1 == SLOTS ? 0 : slot + 1;
    process(wreg);
    const int done_cb = cb;
    if (advance()) end_column(done_cb);
  }
  if (b != 0) {  // range ended inside a column block: close the open group (SA) and publish the partial
    if constexpr (SA) {
      if (group_open) close_group();
    }
    flush(cb);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int pick(int forced, int heuristic) { return forced > 0 ? forced : heuristic; }

struct DecodePlan {
  int u;            // steps per batch (4, or 2 for int4 with g = 64)
  int units_per_col, col_blocks, total_units;
  int total_warps, q, r, slots_max;
  bool scale_after;
};

// Every knob is overridable through cb200_set_decode_tuning(warps_n -> 1: force in-loop scaling (bit-exact W),
// 2: force scale-after; warps_k -> CTAs per SM assumed resident; split_k -> units per warp; vec_words unused).
static DecodePlan decode_plan(int mrows, int K, int N, int bits, int g) {
  DecodePlan pl;
  const int step_k = 4 * (32 / bits);
  pl.u = (g % (4 * step_k) == 0) ? 4 : 2;
  pl.units_per_col = K / (pl.u * step_k);
  pl.col_blocks = N / 32;
  pl.total_units = pl.units_per_col * pl.col_blocks;
  const int ctas_per_sm = pick(g_decode_tuning.warps_k, mrows > 8 ? 4 : 6);
  const int capacity = 148 * ctas_per_sm * 4;  // co-resident warps
  int per_warp = (pl.total_units + capacity - 1) / capacity;
  if (per_warp < 2 && pl.total_units >= 2 * 148 * 4) per_warp = 2;  // tiny problems: still two batches per warp
  per_warp = pick(g_decode_tuning.split_k, per_warp);
  if (per_warp > pl.total_units) per_warp = pl.total_units;
  pl.total_warps = (pl.total_units + per_warp - 1) / per_warp;
  pl.q = pl.total_units / pl.total_warps;
  pl.r = pl.total_units % pl.total_warps;
  // a column block of B units is covered by at most ceil(B / q) + 1 warps
  pl.slots_max = (pl.units_per_col + pl.q - 1) / pl.q + 1;
  pl.scale_after = g_decode_tuning.warps_n == 1 ? false : true;
  return pl;
}

size_t decode_workspace_bytes(int64_t M, int64_t K, int64_t N, int bits) {
  const int mrows = int(M < 16 ? M : 16);
  // the group size changes the plan only through U; take the larger footprint of the two
  size_t need = 0;
  for (int g : {64, 128}) {
    if (K % g) continue;
    const DecodePlan pl = decode_plan(mrows, int(K), int(N), bits, g);
    const size_t b = size_t(pl.col_blocks) * pl.slots_max * mrows * 32 * sizeof(float);
    need = b > need ? b : need;
  }
  return size_t(CB200_WORKSPACE_COUNTER_BYTES) + need;
}

template <class A, int BITS, int MT, int U, int ZK>
static int launch_sa(bool sa, const Problem& p, const DecodeGeom& geo, int blocks, float* wsp, unsigned* wsc,
                     cudaStream_t stream) {
  constexpr size_t smem = size_t(4) * kDecodeRingSteps * 512;  // 4 warps x ring
  if (sa) {
    decode_kernel<A, BITS, MT, U, ZK, true><<<blocks, 128, smem, stream>>>(p, geo, wsp, wsc);
  } else {
    decode_kernel<A, BITS, MT, U, ZK, false><<<blocks, 128, smem, stream>>>(p, geo, wsp, wsc);
  }
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <class A, int BITS, int MT, int U>
static int launch_zk(bool sa, const Problem& p, const DecodeGeom& geo, int blocks, float* wsp, unsigned* wsc,
                     cudaStream_t stream) {
  switch (p.zp_kind) {
    case CB200_ZP_NONE: return launch_sa<A, BITS, MT, U, CB200_ZP_NONE>(sa, p, geo, blocks, wsp, wsc, stream);
    case CB200_ZP_I32: return launch_sa<A, BITS, MT, U, CB200_ZP_I32>(sa, p, geo, blocks, wsp, wsc, stream);
    case CB200_ZP_U8: return launch_sa<A, BITS, MT, U, CB200_ZP_U8>(sa, p, geo, blocks, wsp, wsc, stream);
    default: return launch_sa<A, BITS, MT, U, CB200_ZP_ACT>(false, p, geo, blocks, wsp, wsc, stream);
  }
}

template <class A, int BITS>
static int launch_typed(int u, bool sa, const Problem& p, const DecodeGeom& geo, int blocks, float* wsp, unsigned* wsc,
                        cudaStream_t stream) {
  const bool mt2 = geo.m_rows > 8;
  if (u == 4) {
    return mt2 ? launch_zk<A, BITS, 2, 4>(sa, p, geo, blocks, wsp, wsc, stream)
               : launch_zk<A, BITS, 1, 4>(sa, p, geo, blocks, wsp, wsc, stream);
  }
  return mt2 ? launch_zk<A, BITS, 2, 2>(sa, p, geo, blocks, wsp, wsc, stream)
             : launch_zk<A, BITS, 1, 2>(sa, p, geo, blocks, wsp, wsc, stream);
}

int launch_decode(const Problem& p_in, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream,
                  int* launches) {
  for (int m0 = 0; m0 < p_in.M; m0 += 16) {
    Problem p = p_in;
    const int mrows = (p_in.M - m0) < 16 ? (p_in.M - m0) : 16;
    p.x = reinterpret_cast<const char*>(p_in.x) + size_t(m0) * p_in.K * 2;
    const size_t out_es = p_in.out_dtype == CB200_F32 ? 4 : 2;
    p.out = reinterpret_cast<char*>(p_in.out) + size_t(m0) * p_in.N * out_es;
    p.M = mrows;

    const DecodePlan pl = decode_plan(mrows, p.K, p.N, bits, p.g);
    const size_t need =
        size_t(CB200_WORKSPACE_COUNTER_BYTES) + size_t(pl.col_blocks) * pl.slots_max * mrows * 32 * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
      set_error("decode kernel needs %zu workspace bytes, got %zu", need, ws_bytes);
      return CB200_ERR_WORKSPACE;
    }
    if (size_t(pl.col_blocks) * sizeof(unsigned) > CB200_WORKSPACE_COUNTER_BYTES) {
      set_error("N=%d too large for the split-K counter area", p.N);
      return CB200_ERR_SHAPE;
    }
    DecodeGeom geo;
    geo.m_rows = mrows;
    geo.units_per_col = pl.units_per_col;
    geo.q = pl.q;
    geo.r = pl.r;
    geo.total_warps = pl.total_warps;
    geo.col_blocks = pl.col_blocks;
    geo.slots_max = pl.slots_max;
    const int blocks = (pl.total_warps + 3) / 4;
    float* wsp = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES);
    unsigned* wsc = reinterpret_cast<unsigned*>(ws);
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_typed<__half, 4>(pl.u, pl.scale_after, p, geo, blocks, wsp, wsc, stream)
                     : launch_typed<__half, 8>(pl.u, pl.scale_after, p, geo, blocks, wsp, wsc, stream);
    } else {
      rc = bits == 4 ? launch_typed<__nv_bfloat16, 4>(pl.u, pl.scale_after, p, geo, blocks, wsp, wsc, stream)
                     : launch_typed<__nv_bfloat16, 8>(pl.u, pl.scale_after, p, geo, blocks, wsp, wsc, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
