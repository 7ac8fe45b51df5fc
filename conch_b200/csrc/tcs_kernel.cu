// tcgen05 streaming decode kernel (CB200_DECODE_TCS, AUTO's choice for M <= 32): host-side plan and dispatch.
//
// Replaces the reference's single 128x64x64 Triton tile (conch/kernels/quantization/gemm.py:136-145, 491-494) for the
// HBM-bound regime.  Why a THIRD decode kernel: profiling the two mma.sync formulations (decode_kernel.cu, stream_kernel.cu)
// showed that a B200 SM is ISSUE-bound, not HBM-bound, when every warp has to unpack nibbles, permute activations, feed
// mma.sync and fold scales itself: ~340 instructions per (32 columns x 128 k) against a budget of ~350 issue slots per
// scheduler at full HBM rate (profiles/r02_stream_*).  Here the only per-weight work left on the CUDA cores is the
// unpack -- 1 shift + 4 AND per packed word -- and everything else moves to fixed-function hardware:
//
//   * TMA ring (as in stream_kernel.cu): a producer warp streams [128 columns x 16 packed rows] weight tiles, the matching
//     [M x 128 k] slice of x and the group's scale / zero-point row into shared-memory stages; static data is requested
//     before griddepcontrol.wait, so under programmatic dependent launch the next GEMV's first ~100 KB per SM arrive while
//     this one is still running (a CTA is half an SM: 352 threads, <= 80 registers, ~110 KB, 256 of the 512 TMEM columns).
//   * dequant warps (2 sets x 4 warps; a thread owns ONE weight column = one TMEM lane): packed words -> registers -> mask
//     nibbles into fp16 SUBNORMAL operands (0x000q = q * 2^-24; bf16: 0x4300 | q = 128 + q) -> tcgen05.st straight into
//     tensor memory as the MMA's A operand.  No floating-point instruction, no shared-memory round trip.
//   * one thread issues tcgen05.mma.kind::f16 (M = 128 weight columns, N = 16 | 32 rows of x, K = 16) with A from TMEM and B
//     = x from shared memory; the accumulator G[n, m] = sum_k x[m,k] * operand[k,n] of the open SCALE GROUP lives in TMEM.
//     fp16 int4's odd nibbles arrive as 16 q: they get their own MMAs into their own accumulator, scaled by 1/16 in fp32 at
//     the fold (no pre-scaling of x: small-magnitude activations keep their bits).
//   * an "x operand" warp rewrites each raw x slice into the K-major, 128-byte-swizzled B operand in the k order the
//     mask-only extraction produces (so the weights need no permutes), and accumulates X[m] = sum_k x[m,k] per group.
//   * fold: when a scale group is complete the set that dequantised its last stage reads G from TMEM (tcgen05.ld) and does
//     acc[n, m] += (s * 2^24) * G - (s * (zp + bias)) * X[m]   in fp32 -- the only place scales and zero-points appear.
//     It is deferred by one stage, so the set never waits for the tensor pipe.
//   * stream-K over #SMs (or 2 x #SMs) equal runs of the k-fastest stage list, deterministic last-arriver split-K reduction
//     with self-resetting counters -- as in stream_kernel.cu.
// The same kernel serves M = 1 .. 32 with the same per-weight cost: it closes the mid-M hole below the prefill kernel.
#include <atomic>

#include "host_util.cuh"
#include "tcs_impl.cuh"

namespace cb200 {

struct TcsPlan {
  int mt, ctas, q, r, spc, ncb, total;
  int ring, stage_bytes, x_off, x_bytes, meta_off, spg, maxslots;
  size_t smem;
};

static int tcs_sms() {
  const DeviceInfo* di = current_device_info();
  return di && di->sms > 0 ? di->sms : 148;
}

static TcsPlan tcs_plan(int mrows, int K, int N, int bits, int g, int zp_kind, bool zp_scalar) {
  TcsPlan pl{};
  const int pf = 32 / bits;
  const int ks = 16 * pf;
  pl.mt = mrows <= 16 ? 16 : 32;
  pl.spc = K / ks;
  pl.ncb = (N + 127) / 128;
  pl.total = pl.spc * pl.ncb;
  pl.spg = g / ks;
  const int sms = tcs_sms();
  // a CTA is half an SM.  Small layers: one CTA per SM, the NEXT kernel's CTAs are co-resident from the start and prefetch
  // under it.  Long runs (>= 16 stages = 128 KB per SM): two CTAs per SM, all warps on this layer's streaming.  The 32-row
  // variant needs more registers: one CTA per SM with the whole shared memory as its ring.
  int ctas = g_stream_tuning.ctas > 0 ? g_stream_tuning.ctas : (pl.mt == 16 && pl.total >= 16 * sms ? 2 * sms : sms);
  if (ctas > pl.total) ctas = pl.total;
  pl.ctas = ctas;
  pl.q = pl.total / ctas;
  pl.r = pl.total % ctas;
  const int zrow = zp_scalar ? 0 : zp_kind == CB200_ZP_I32 ? 512 : zp_kind == CB200_ZP_U8 ? 128 : zp_kind == CB200_ZP_ACT ? 256 : 0;
  pl.x_bytes = mrows * ks * 2;
  pl.x_off = kTcsStageW;
  pl.meta_off = pl.x_off + ((pl.x_bytes + 127) / 128) * 128;
  pl.stage_bytes = ((pl.meta_off + 256 + zrow + 1023) / 1024) * 1024;
  const int xb_bytes = (ks / 64) * pl.mt * 128;
  const int budget = (g_stream_tuning.ring_kb > 0 ? g_stream_tuning.ring_kb : (pl.mt == 16 ? 110 : 200)) * 1024;
  int ring = (budget - 3072 - 4 * xb_bytes) / pl.stage_bytes;
  if (ring > 32) ring = 32;
  if (ring > pl.q + 1) ring = pl.q + 1;
  ring -= ring % kTcsSets;  // a slot is always consumed by the same set (mbarrier phase order)
  if (ring < kTcsSets) ring = kTcsSets;
  pl.ring = ring;
  pl.smem = size_t(3072) + size_t(4 * xb_bytes) + size_t(ring) * pl.stage_bytes;  // 1 KB alignment slack + 2 KB control page
  int maxc = (pl.spc + pl.q - 1) / pl.q + 1;
  if (maxc > ctas) maxc = ctas;
  pl.maxslots = maxc * kTcsSets;
  return pl;
}

bool tcs_supported(const Problem& p, int bits, int act_dtype) {
  if (encode_fn() == nullptr) return false;
  if (act_dtype == CB200_BF16 && bits == 8) return false;  // no magic-number operand: decode_kernel's per-weight dequant
  const int ks = 16 * (32 / bits);
  return p.K % ks == 0 && p.N % 64 == 0 && p.g % ks == 0;  // scale groups are whole stages (int4: g % 128 == 0)
}

static size_t tcs_partial_bytes(const TcsPlan& pl, int mrows) { return size_t(pl.ncb) * pl.maxslots * mrows * 128 * sizeof(float); }

size_t tcs_workspace_bytes(int64_t M, int64_t K, int64_t N, int bits) {
  const int ks = 16 * (32 / bits);
  if (K % ks != 0) return 0;
  const int rows = M < 32 ? int(M) : 32;
  const TcsPlan pl = tcs_plan(rows, int(K), int(N), bits, int(K), CB200_ZP_I32, false);
  return size_t(CB200_WORKSPACE_COUNTER_BYTES) + tcs_partial_bytes(pl, rows);
}

static int make_tcs_tmaps(const Problem& p, int bits, int act_dtype, int mrows, CUtensorMap& mw, CUtensorMap& mx) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return CB200_ERR_DEVICE;
  }
  const int pf = 32 / bits;
  {  // packed W [K/pf, N] int32: box {128 columns, 16 packed rows}, no swizzle (a thread reads one column: conflict free)
    const cuuint64_t dims[2] = {cuuint64_t(p.N), cuuint64_t(p.K / pf)};
    const cuuint64_t strides[1] = {cuuint64_t(p.N) * 4};
    const cuuint32_t box[2] = {128, 16};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<uint32_t*>(p.wq), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(w_q_packed, tcs) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  {  // x [M, K]: box {KS k, M rows}, no swizzle (the x operand warp re-lays it out)
    const cuuint64_t dims[2] = {cuuint64_t(p.K), cuuint64_t(mrows)};
    const cuuint64_t strides[1] = {cuuint64_t(p.K) * 2};
    const cuuint32_t box[2] = {cuuint32_t(16 * pf), cuuint32_t(mrows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&mx, act_dtype == CB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(x, tcs) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  return 0;
}

int launch_tcs(const Problem& p_in, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches) {
  for (int m0 = 0; m0 < p_in.M; m0 += 32) {
    Problem p = p_in;
    const int mrows = (p_in.M - m0) < 32 ? (p_in.M - m0) : 32;
    const size_t out_es = p_in.out_dtype == CB200_F32 ? 4 : 2;
    p.x = reinterpret_cast<const char*>(p_in.x) + size_t(m0) * p_in.K * 2;
    p.out = reinterpret_cast<char*>(p_in.out) + size_t(m0) * p_in.N * out_es;
    for (int i = 0; i < p_in.n_peers; ++i) p.peer_out[i] = reinterpret_cast<char*>(p_in.peer_out[i]) + size_t(m0) * p_in.N * out_es;
    p.M = mrows;

    const TcsPlan pl = tcs_plan(mrows, p.K, p.N, bits, p.g, p.zp_kind, p.zp_scalar != 0);
    const size_t need = size_t(CB200_WORKSPACE_COUNTER_BYTES) + tcs_partial_bytes(pl, mrows);
    if (ws == nullptr || ws_bytes < need) {
      set_error("tcgen05 streaming decode kernel needs %zu workspace bytes, got %zu", need, ws_bytes);
      return CB200_ERR_WORKSPACE;
    }
    if (size_t(pl.ncb) * 4 * sizeof(unsigned) > CB200_WORKSPACE_COUNTER_BYTES) {
      set_error("N=%d too large for the split-K counter area", p.N);
      return CB200_ERR_SHAPE;
    }
    if (size_t(p.K / p.g) * p.N >= (size_t(1) << 31)) {
      set_error("scale matrix too large");
      return CB200_ERR_SHAPE;
    }
    CUtensorMap mw, mx;
    if (int rc = make_tcs_tmaps(p, bits, act_dtype, mrows, mw, mx)) return rc;
    TcsGeom geo;
    geo.m_rows = mrows;
    geo.spc = pl.spc;
    geo.ncb = pl.ncb;
    geo.q = pl.q;
    geo.r = pl.r;
    geo.ring = pl.ring;
    geo.stage_bytes = pl.stage_bytes;
    geo.x_off = pl.x_off;
    geo.x_bytes = pl.x_bytes;
    geo.meta_off = pl.meta_off;
    geo.spg = pl.spg;
    geo.maxslots = pl.maxslots;
    geo.pdl = g_decode_pdl;
    geo.trace = g_stream_trace;
    float* wsp = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES);
    unsigned* wsc = reinterpret_cast<unsigned*>(ws);
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_tcs_f16_w4(pl.mt, mw, mx, p, geo, pl.ctas, pl.smem, wsp, wsc, stream)
                     : launch_tcs_f16_w8(pl.mt, mw, mx, p, geo, pl.ctas, pl.smem, wsp, wsc, stream);
    } else {
      rc = launch_tcs_bf16_w4(pl.mt, mw, mx, p, geo, pl.ctas, pl.smem, wsp, wsc, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
