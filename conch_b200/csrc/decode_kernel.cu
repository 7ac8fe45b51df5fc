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
#include "decode_impl.cuh"

namespace cb200 {

DecodeTuning g_decode_tuning = {0, 0, 0, 0};

#ifdef CB200_TC_TRACE
static long long* g_dec_trace = nullptr;  // host copy of the trace buffer pointer, passed to the kernels in DecodeGeom
extern "C" __attribute__((visibility("default"))) void cb200_debug_set_decode_trace(void* p) {
  g_dec_trace = reinterpret_cast<long long*>(p);
}
#endif

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
// identical to the reference's w_ref), 0 / 2 (default, measured 5-10% faster): fused magic-number operands -- for fp16 x int4
// the x16 nibbles accumulate separately and that accumulator is scaled by 1/16 in fp32 (round 1 scaled their activations in
// fp16 and lost bits for |x| < 2^-10);
// warps_k -> CTAs per SM assumed resident; split_k -> forced split count S (rounded down to a legal divisor); vec_words unused).
static DecodePlan decode_plan(int mrows, int K, int N, int bits, int g, bool exact_default = false) {
  DecodePlan pl;
  const int step_k = 4 * (32 / bits);
  pl.u = (g % (4 * step_k) == 0) ? 4 : 2;
  const int B = K / (pl.u * step_k);  // batches per column block
  pl.bpg = g / (pl.u * step_k);
  pl.col_blocks = N / 32;
  (void)exact_default;
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
      for (bool exact : {false, true}) {
        const DecodePlan pl = decode_plan(rows, int(K), int(N), bits, int(g), exact);
        const size_t b = pl.splits > 1 ? size_t(pl.total_warps) * rows * 32 * sizeof(float) : 0;
        need = b > need ? b : need;
      }
    }
    for (bool exact : {false, true}) {
      const DecodePlan pk = decode_plan(rows, int(K), int(N), bits, int(K), exact);
      const size_t b = pk.splits > 1 ? size_t(pk.total_warps) * rows * 32 * sizeof(float) : 0;
      need = b > need ? b : need;
    }
  }
  return size_t(CB200_WORKSPACE_COUNTER_BYTES) + need;
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

    const DecodePlan pl = decode_plan(mrows, p.K, p.N, bits, p.g, act_dtype == CB200_F16 && bits == 4);
    const size_t need =
        size_t(CB200_WORKSPACE_COUNTER_BYTES) + (pl.splits > 1 ? size_t(pl.total_warps) * mrows * 32 * sizeof(float) : 0);
    if (pl.splits > 1 && (ws == nullptr || ws_bytes < need)) {
      set_error("decode kernel needs %zu workspace bytes, got %zu", need, ws_bytes);
      return CB200_ERR_WORKSPACE;
    }
    if (size_t(pl.col_blocks + 1) * sizeof(unsigned) > CB200_WORKSPACE_COUNTER_BYTES) {  // (+1: the push-signal counter in the last slot)
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
    geo.scalar_push = g_decode_tuning.vec_words == 1;
#ifdef CB200_TC_TRACE
    geo.trace = g_dec_trace;
#endif
    const int blocks = (pl.total_warps + 3) / 4;
    float* wsp = ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES) : nullptr;
    unsigned* wsc = reinterpret_cast<unsigned*>(ws);
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_decode_f16_w4(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream)
                     : launch_decode_f16_w8(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream);
    } else {
      rc = bits == 4 ? launch_decode_bf16_w4(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream)
                     : launch_decode_bf16_w8(pl.u, pl.fused, p, geo, blocks, size_t(pl.cta_smem), wsp, wsc, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
