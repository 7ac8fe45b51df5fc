// Streaming decode kernel (CB200_DECODE_STREAM, AUTO's choice for M <= 16): host-side plan and dispatch.
//
// Replaces the reference's single 128x64x64 Triton tile (conch/kernels/quantization/gemm.py:136-145, 491-494), which at M=1
// launches N/64 CTAs with no split-K, and supersedes round 1's register-window kernel (decode_kernel.cu, kept for bf16 x int8
// and as the per-weight-dequant variant) as the default decode path.
//
// What bounds a decode GEMV on a B200 is not the streaming loop but what happens AROUND it: round 1 measured a 12 us loop
// inside a 19.6 us launch for the 58.7 MB layer (head: launch skew + HBM ramp-up; tail: stragglers + split-K arrival), and
// with four 128-register CTAs per SM the successor kernel cannot become resident before a predecessor CTA retires, so
// programmatic dependent launch had nothing to overlap.  This kernel is shaped so that TWO consecutive GEMVs share every SM:
//
//   * persistent: one CTA per SM (grid = #SMs; two for long layers), 256 threads, 128 registers, ~110 KB of shared memory -- half an SM.
//     The successor's CTAs become co-resident as soon as every CTA of this kernel has executed griddepcontrol.launch_dependents
//     (its first instruction), request THEIR first ring of weight tiles (static data: legal before griddepcontrol.wait), and only
//     then block on the stream order.  The HBM therefore keeps streaming across the kernel boundary: the next layer's first
//     ~100 KB per SM (the whole 8.4 MB o-proj layer is 57 KB per SM) arrive while this layer's tail, split-K reduction and the
//     launch latency of the next grid play out.
//   * the packed weights never touch the register file on their way in: TMA tile loads (cp.async.bulk.tensor, SWIZZLE_128B,
//     [4 x 32 columns x 16 packed rows] = 8 KB per stage) land in a ring of shared-memory stages guarded by `full` mbarriers.
//     There is no producer warp (a 9th warp would cap the kernel at 96 registers): warp 0 requests the first ring, after that
//     the LAST of a stage's four consumer warps to leave it re-arms the slot and requests the stage `ring` ahead.  The stage also carries what the consumers need for it: the
//     matching [M x 128 k] slice of x (one TMA box) and the scale / zero-point row(s) of its group (bulk copies), so the
//     consumers issue nothing but ld.shared.  Ring depth is whatever fits the ~105 KB budget (8-11 stages).
//   * work is the k-fastest list of (128-column block, 16-packed-row slab) stages, cut into #SMs contiguous runs of equal
//     length (+-1): every SM streams the same number of bytes whatever the shape (a stream-K split; no wave quantisation:
//     4096 x 4096 is 1024 stages = 6.9 per SM).
//   * two consumer groups of four warps take alternate stages of the run; inside a group warp q owns columns [32q, 32q+32).
//     One step = one ld.shared.v4 of packed words (bank-conflict free thanks to the hardware swizzle plus a lane -> column
//     permutation), 1 shift + 4 AND per word to turn nibbles into fp16 SUBNORMAL operands (0x000q = q * 2^-24, exact on the
//     tensor pipe), mma.sync.m16n8k16 with the weights as the 16-row operand and x^T as the 8-column operand.  No
//     floating-point instruction per weight: scale and zero-point are applied once per scale group in fp32,
//     acc += (s * 2^24) * G - (s * zb) * X  with  G = sum_k x q 2^-24  and  X = sum_k x  (an all-ones MMA).
//     fp16 int4's odd nibbles arrive as 16 q: they accumulate into their OWN accumulator which is scaled by 1/16 in fp32 at
//     group close (round 1 pre-multiplied x by 1/16 in fp16, which loses bits for |x| < 2^-10 -- ADVICE r1).
//   * split-K: a column block's stages are spread over 1..6 CTAs x 2 groups.  Every contributor writes an fp32 partial slot,
//     bumps the block's arrival counter, and the last arriver sums the slots in slot order (deterministic, counters reset
//     themselves) and stores the result with 8-byte vector stores -- to `out` and, for cb200_mpgemm_push, to every peer.
#include <atomic>

#include "host_util.cuh"
#include "stream_impl.cuh"

namespace cb200 {

StreamTuning g_stream_tuning = {0, 0, 0};

struct StreamPlan {
  int ctas, q, r, spc, ncb, total;
  int ring, stage_bytes, x_off, x_bytes, meta_off, meta_row_bytes, gps, hpg, hpg_shift, maxslots;
  size_t smem;
};

static int stream_sms() {
  const DeviceInfo* di = current_device_info();
  return di && di->sms > 0 ? di->sms : 148;
}

// (zp_kind < 0: worst case over the zero-point kinds, for workspace sizing, which does not depend on it anyway)
static StreamPlan stream_plan(int mrows, int K, int N, int bits, int g, int zp_kind, bool zp_scalar) {
  StreamPlan pl{};
  const int pf = 32 / bits;
  const int ks = 16 * pf;
  pl.spc = K / ks;
  pl.ncb = (N + 127) / 128;
  pl.total = pl.spc * pl.ncb;
  // CTAs: a CTA is half an SM (8 warps, 128 registers, ~110 KB).  Small layers run one CTA per SM so that the NEXT kernel's CTAs
  // are co-resident from the start and prefetch under it; layers with a long run per SM (>= 16 stages = 128 KB) run two per SM
  // -- all 16 warps on their own streaming, the successor moves in as the first CTAs retire.  The 16-row variant needs 184
  // registers: one CTA per SM, with the whole shared memory as its ring.
  const int sms = stream_sms();
  int ctas = g_stream_tuning.ctas > 0 ? g_stream_tuning.ctas : (mrows <= 8 && pl.total >= 16 * sms ? 2 * sms : sms);
  if (ctas > pl.total) ctas = pl.total;
  pl.ctas = ctas;
  pl.q = pl.total / ctas;
  pl.r = pl.total % ctas;
  pl.hpg = g / (8 * pf);  // half-stages (8 packed rows) per scale group
  pl.hpg_shift = -1;
  for (int sh = 0; sh < 30; ++sh) {
    if ((1 << sh) == pl.hpg) pl.hpg_shift = sh;
  }
  pl.gps = (bits == 4 && g % 128 != 0) ? 2 : 1;
  const int zrow = zp_scalar ? 0 : zp_kind == CB200_ZP_I32 ? 512 : zp_kind == CB200_ZP_U8 ? 128 : zp_kind == CB200_ZP_ACT ? 256 : 0;
  pl.meta_row_bytes = 256 + zrow;
  pl.x_bytes = mrows * ks * 2;
  pl.x_off = kStreamStageW;
  pl.meta_off = pl.x_off + ((pl.x_bytes + 127) / 128) * 128;
  pl.stage_bytes = ((pl.meta_off + pl.gps * pl.meta_row_bytes + 1023) / 1024) * 1024;
  // two co-resident CTAs per SM: (227 KB - 2 x 1 KB reserved) / 2, minus the alignment slack and the barrier page
  const int budget = (g_stream_tuning.ring_kb > 0 ? g_stream_tuning.ring_kb : (mrows <= 8 ? 110 : 200)) * 1024;
  int ring = (budget - 2048) / pl.stage_bytes;
  if (ring > 32) ring = 32;
  if (ring > pl.q + 1) ring = pl.q + 1;  // never deeper than the longest run
  // a multiple of the number of consumer groups: then a slot is always consumed by the SAME group, which therefore observes
  // every phase of the slot's `full` barrier in order (a group that skipped a phase could pass a parity wait one phase early)
  ring -= ring % kStreamGroups;
  if (ring < kStreamGroups) ring = kStreamGroups;
  pl.ring = ring;
  pl.smem = size_t(2048) + size_t(ring) * pl.stage_bytes;
  // a column block's spc stages intersect at most ceil(spc / q) + 1 runs
  int maxc = (pl.spc + pl.q - 1) / pl.q + 1;
  if (maxc > ctas) maxc = ctas;
  pl.maxslots = maxc * kStreamGroups;
  return pl;
}

bool stream_supported(const Problem& p, int bits, int act_dtype) {
  if (encode_fn() == nullptr) return false;
  if (act_dtype == CB200_BF16 && bits == 8) return false;  // no magic-number operand: decode_kernel's per-weight dequant
  const int ks = 16 * (32 / bits);                         // a stage is 16 packed rows: K must be a whole number of stages
  return p.K % ks == 0 && p.N % 64 == 0 && p.g % 64 == 0;
}

static size_t stream_partial_bytes(const StreamPlan& pl, int mrows) {
  return size_t(pl.ncb) * 4 * pl.maxslots * mrows * 32 * sizeof(float);
}

size_t stream_workspace_bytes(int64_t M, int64_t K, int64_t N, int bits) {
  if (K % (16 * (32 / bits)) != 0) return 0;  // not a streaming-kernel shape (stream_supported)
  const int rows = M < 16 ? int(M) : 16;
  const StreamPlan pl = stream_plan(rows, int(K), int(N), bits, int(K), CB200_ZP_I32, false);
  return size_t(CB200_WORKSPACE_COUNTER_BYTES) + stream_partial_bytes(pl, rows);
}

long long* g_stream_trace = nullptr;

// packed W [K/pf, N] int32.  Preferred: ONE 3-D map {32 columns, K/pf rows, N/32 column blocks} with strides {N*4, 128} and box
// {32, 16, 4} -- the smem image is four [16 rows x 128 B] blocks, each 128-byte swizzled on its own row index (what makes the
// kernel's ld.shared.v4 conflict free), fetched by a single TMA instruction per stage.  The dimension order is not stride-
// monotonic; if the driver's encoder refuses it, fall back to a 2-D map {N, K/pf} with box {32, 16} and four TMAs per stage.
static int make_stream_tmaps(const Problem& p, int bits, int act_dtype, int mrows, CUtensorMap& mw, CUtensorMap& mx, int& w3d) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return CB200_ERR_DEVICE;
  }
  const int pf = 32 / bits;
  w3d = 0;
  static std::atomic<int> w3d_refused{0};
  if (!(g_stream_tuning.dbg & 16) && !w3d_refused.load(std::memory_order_relaxed) && p.N > 32) {
    const cuuint64_t dims[3] = {32, cuuint64_t(p.K / pf), cuuint64_t(p.N / 32)};
    const cuuint64_t strides[2] = {cuuint64_t(p.N) * 4, 128};
    const cuuint32_t box[3] = {32, 16, 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<uint32_t*>(p.wq), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      w3d = 1;
    } else {
      w3d_refused.store(1, std::memory_order_relaxed);
    }
  }
  if (!w3d) {
    const cuuint64_t dims[2] = {cuuint64_t(p.N), cuuint64_t(p.K / pf)};
    const cuuint64_t strides[1] = {cuuint64_t(p.N) * 4};
    const cuuint32_t box[2] = {32, 16};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<uint32_t*>(p.wq), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(w_q_packed, stream) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  {  // x [M, K]: box {KS k, M rows}, no swizzle
    const cuuint64_t dims[2] = {cuuint64_t(p.K), cuuint64_t(mrows)};
    const cuuint64_t strides[1] = {cuuint64_t(p.K) * 2};
    const cuuint32_t box[2] = {cuuint32_t(16 * pf), cuuint32_t(mrows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&mx, act_dtype == CB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(x, stream) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  return 0;
}

int launch_stream(const Problem& p_in, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches) {
  for (int m0 = 0; m0 < p_in.M; m0 += 16) {
    Problem p = p_in;
    const int mrows = (p_in.M - m0) < 16 ? (p_in.M - m0) : 16;
    const size_t out_es = p_in.out_dtype == CB200_F32 ? 4 : 2;
    p.x = reinterpret_cast<const char*>(p_in.x) + size_t(m0) * p_in.K * 2;
    p.out = reinterpret_cast<char*>(p_in.out) + size_t(m0) * p_in.N * out_es;
    for (int i = 0; i < p_in.n_peers; ++i) p.peer_out[i] = reinterpret_cast<char*>(p_in.peer_out[i]) + size_t(m0) * p_in.N * out_es;
    p.M = mrows;

    const StreamPlan pl = stream_plan(mrows, p.K, p.N, bits, p.g, p.zp_kind, p.zp_scalar != 0);
    const size_t need = size_t(CB200_WORKSPACE_COUNTER_BYTES) + stream_partial_bytes(pl, mrows);
    if (ws == nullptr || ws_bytes < need) {
      set_error("streaming decode kernel needs %zu workspace bytes, got %zu", need, ws_bytes);
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
    int w3d = 0;
    if (int rc = make_stream_tmaps(p, bits, act_dtype, mrows, mw, mx, w3d)) return rc;
    StreamGeom geo;
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
    geo.meta_row_bytes = pl.meta_row_bytes;
    geo.gps = pl.gps;
    geo.hpg = pl.hpg;
    geo.hpg_shift = pl.hpg_shift;
    geo.w3d = w3d;
    geo.trace = g_stream_trace;
    geo.maxslots = pl.maxslots;
    geo.pdl = g_decode_pdl;
    geo.dbg = g_stream_tuning.dbg;
    float* wsp = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES);
    unsigned* wsc = reinterpret_cast<unsigned*>(ws);
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_stream_f16_w4(mw, mx, p, geo, pl.ctas, pl.smem, wsp, wsc, stream)
                     : launch_stream_f16_w8(mw, mx, p, geo, pl.ctas, pl.smem, wsp, wsc, stream);
    } else {
      rc = launch_stream_bf16_w4(mw, mx, p, geo, pl.ctas, pl.smem, wsp, wsc, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
