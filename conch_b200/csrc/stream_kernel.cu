// Streaming ("ring") decode kernel (CB200_DECODE_STREAM, AUTO's choice for M <= 8): host-side plan and dispatch.
//
// Replaces the reference's single 128x64x64 Triton tile (conch/kernels/quantization/gemm.py:136-145, 491-494), which at M=1
// launches N/64 CTAs with no split-K, and supersedes round 1's register-window kernel (decode_kernel.cu, kept for odd group sizes
// and the per-weight-dequant variant) as the default decode path.
//
// Shape of the kernel (stream_impl.cuh), each item from a measurement (profiles/r02_ring_*.log):
//   * ONE CTA per SM, warp specialised: a producer warp + 2..4 groups of consumer warps (12-16 consumer warps in all).  The
//     producer never waits for the stream order: launched with programmatic dependent launch, a CTA starts as soon as an SM is
//     free and has its ring full of WEIGHTS (static data) by the time the previous kernel completes.
//   * the packed weights never pass through registers on their way in: one TMA tile load per stage -- a 3-D box
//     {32 columns, 16 packed rows, nw column slices}, 128-byte swizzled, so that every consumer warp reads its [16 x 128 B] slice
//     with conflict-free ld.shared.v4 -- plus two bulk copies for the stage's scale / zero-point rows.  The producer's loop is
//     warp-uniform (operands in uniform registers, one elected lane issues): as a single diverged thread it was the bottleneck
//     (1373 -> 1187 us for the 128-GEMV decode step).
//   * a column block is 32 * nw columns wide (nw = 4..8 consumer warps per group, one 32-column slice each); its K range is cut
//     into S slabs (S = 1, 2, 4) taken by the S CTAs of a thread-block CLUSTER.  No workspace, no atomics, no last-arriver: every
//     (CTA, consumer group) sends its fp32 partial to the CTA that owns the column slice with st.async into distributed shared
//     memory (an mbarrier in the owner counts the bytes); the owner adds the partials in a fixed order and stores.  0.3 us instead
//     of the 1-2 us fence + atomic + re-read of the global-memory scheme, and deterministic.
//   * consumer group h takes the stages h, h + G, ... of the CTA's run: twice to four times the warps in flight for the same ring.
//     One step of a warp = one ld.shared.v4 of packed words + one of activations, 1 shift + 4 AND per word to turn nibbles into fp16
//     SUBNORMAL operands (0x000q = q * 2^-24, exact on the tensor pipe), mma.sync.m16n8k16 with the weights as the 16-row operand
//     and x^T as the 8-column operand.  Scale and zero-point are applied once per scale group in fp32:
//         acc += (s * 2^24) * G - (s * zb) * X,   G = sum_k x q 2^-24,   X = sum_k x.
//     X comes from a table the CTA fills while it stages its slab of x (one pass, shuffle tree, no extra MMAs).  Nibbles at bits
//     4-7 arrive as 16 q and accumulate in their OWN accumulator, scaled by 1/16 in fp32 at group close (round 1 pre-multiplied
//     x by 1/16 in fp16, which loses bits for |x| < 2^-10 -- ADVICE r1).
#include <atomic>

#include "host_util.cuh"
#include "stream_impl.cuh"

namespace cb200 {

StreamTuning g_stream_tuning = {0, 0, 0};
long long* g_stream_trace = nullptr;

struct StreamPlan {
  int nw, S, ngrp, ring;
  int pf, QPB, qpg, one_group, gps, zes;
  int stage_bytes, x_pitch, x_bytes, red_bytes, xsum_bytes, win_stages;
  size_t smem;
  dim3 grid;
};

static int stream_sms() {
  const DeviceInfo* di = current_device_info();
  return di && di->sms > 0 ? di->sms : 148;
}

bool stream_supported(const Problem& p, int bits, int act_dtype) {
  if (encode_fn() == nullptr) return false;
  (void)act_dtype;  // (bf16 x int8 runs as two nibble planes since round 2)
  const int quad_k = 4 * (32 / bits);                      // k per quad (4 packed rows)
  if (p.K % quad_k != 0 || p.N % 32 != 0 || p.g % quad_k != 0) return false;
  const int qpg = p.g / quad_k;
  // a stage is an aligned block of 4 quads: scale groups either tile it (1 or 2 quads) or are whole stages (or the whole of K)
  return p.g == p.K || qpg == 1 || qpg == 2 || qpg % 4 == 0;
}

// Column-block width (nw consumer warps per group) and K slabs (S, the cluster size): as many evenly loaded CTAs as fit ONE wave
// of clusters.  The plan does not depend on M (rows of a batch see the arithmetic of a single-row call).
static StreamPlan stream_plan(int mrows, int K, int N, int bits, int g, int zp_kind, bool zp_scalar) {
  StreamPlan pl{};
  pl.pf = 32 / bits;
  const int quad_k = 4 * pl.pf;
  pl.QPB = K / quad_k;
  pl.one_group = g >= K;
  pl.qpg = pl.one_group ? pl.QPB : g / quad_k;
  pl.gps = !pl.one_group && pl.qpg < 4 ? 4 / pl.qpg : 1;
  pl.zes = (zp_kind == CB200_ZP_NONE || zp_scalar) ? 0 : zp_kind == CB200_ZP_I32 ? 4 : zp_kind == CB200_ZP_U8 ? 1 : 2;
  const int sms = stream_sms();
  const int limit = g_stream_tuning.ctas > 0 ? g_stream_tuning.ctas : sms;
  // co-resident clusters of one-CTA-per-SM kernels: 2-clusters pack perfectly, 4-clusters strand a few SMs per GPC
  const int cap[5] = {0, limit, limit & ~1, 0, (limit >= 148 ? 128 : limit) & ~3};
  double best = -1.0;
  pl.nw = 8;
  pl.S = 1;
  for (int nw = 8; nw >= 4; --nw) {
    for (int S = 1; S <= 4; S *= 2) {
      const int ncb = (N + 32 * nw - 1) / (32 * nw);
      const int ctas = ncb * S;
      if (ctas > cap[S] || pl.QPB / S < 4) continue;
      const double eff = double(N) * S / (32 * nw);  // SMs' worth of evenly loaded CTAs
      if (eff > best + 1e-9) {
        best = eff;
        pl.nw = nw;
        pl.S = S;
      }
    }
  }
  const int max_warps = mrows > 8 ? 14 : kStreamMaxWarps;
  pl.ngrp = max_warps / pl.nw;
  if (pl.ngrp < 1) pl.ngrp = 1;
  const int BN = 32 * pl.nw;
  pl.stage_bytes = ((2048 * pl.nw + pl.gps * BN * (2 + pl.zes) + 1023) / 1024) * 1024;
  const int nxq = (pl.QPB + pl.S - 1) / pl.S + 1;  // quads of the longest slab
  const int xq = 4 * pl.pf * 2;                    // bytes of one x row per quad
  const int nstages = (nxq + 3) / 4 + 1;
  // 16-row variant: x is staged in windows of win_stages stages (<= 64 KB), so that the ring keeps its depth
  pl.win_stages = nstages;
  if (mrows > 8 && mrows * nxq * xq > 64 * 1024) {
    pl.win_stages = (64 * 1024) / (mrows * 4 * xq);
    if (pl.win_stages < 2) pl.win_stages = 2;
    const int nwin = (nstages + pl.win_stages - 1) / pl.win_stages;
    pl.win_stages = (nstages + nwin - 1) / nwin;  // even windows
  }
  const int wxq = pl.win_stages < nstages ? 4 * pl.win_stages : nxq;  // quads of x staged at a time
  pl.x_pitch = wxq * xq + ((wxq * xq) % 128 == 0 ? 64 : 0);
  pl.x_bytes = mrows * pl.x_pitch;
  pl.xsum_bytes = nstages * pl.gps * (mrows > 8 ? 64 : 32);
  pl.red_bytes = ((pl.nw + pl.S - 1) / pl.S) * pl.S * pl.ngrp * mrows * 128 + pl.xsum_bytes;
  const int budget = (g_stream_tuning.ring_kb > 0 ? g_stream_tuning.ring_kb : 224) * 1024;
  int ring = (budget - kStreamHdr - pl.x_bytes - pl.red_bytes - 1024) / pl.stage_bytes;
  if (ring > kStreamMaxRing) ring = kStreamMaxRing;
  if (ring > nstages) ring = nstages;
  if (ring < 1) ring = 1;
  // a multiple of the number of consumer groups: then a slot is always consumed by the SAME group, which therefore observes
  // every phase of the slot's `full` barrier in order (a group that skipped a phase could pass a parity wait one phase early)
  if (pl.ngrp > ring) pl.ngrp = ring;
  ring -= ring % pl.ngrp;
  pl.ring = ring;
  pl.smem = size_t(kStreamHdr) + pl.x_bytes + pl.red_bytes + 1024 + size_t(ring) * pl.stage_bytes;
  pl.grid = dim3(pl.S, (N + BN - 1) / BN);
  return pl;
}

// partial sums travel through distributed shared memory; only the completion-flag counter of cb200_mpgemm_push_signal lives there
size_t stream_workspace_bytes(int64_t, int64_t, int64_t, int) { return CB200_WORKSPACE_COUNTER_BYTES; }

// packed W [K/pf, N] int32 as a 3-D tensor {32 columns, K/pf rows, N/32 column slices}, strides {N*4, 128}, box {32, 16, nw}: the
// smem image is nw [16 rows x 128 B] blocks, each 128-byte swizzled on its own row index.  Encoding costs a few microseconds of
// host time, so the maps are cached per thread, keyed by (pointer, shape, box).
static int stream_tmap(const Problem& p, int pf, int nw, const CUtensorMap** out) {
  struct Entry {
    const void* ptr;
    int K, N, nw, pf;
    CUtensorMap map;
  };
  constexpr int kEntries = 256;
  thread_local Entry cache[kEntries] = {};
  const size_t h = (reinterpret_cast<uintptr_t>(p.wq) >> 8) * 0x9E3779B97F4A7C15ull;
  Entry& e = cache[(h >> 40) % kEntries];
  if (e.ptr != p.wq || e.K != p.K || e.N != p.N || e.nw != nw || e.pf != pf) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
      set_error("cuTensorMapEncodeTiled driver entry point not available");
      return CB200_ERR_DEVICE;
    }
    const cuuint64_t dims[3] = {32, cuuint64_t(p.K / pf), cuuint64_t(p.N / 32)};
    const cuuint64_t strides[2] = {cuuint64_t(p.N) * 4, 128};
    const cuuint32_t box[3] = {32, 16, cuuint32_t(nw)};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<uint32_t*>(p.wq), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      e.ptr = nullptr;
      set_error("cuTensorMapEncodeTiled(w_q_packed) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
    e.ptr = p.wq;
    e.K = p.K;
    e.N = p.N;
    e.nw = nw;
    e.pf = pf;
  }
  *out = &e.map;
  return 0;
}

int launch_stream(const Problem& p_in, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches) {
  if ((p_in.local_flag != nullptr || p_in.wait_flags != nullptr) && (ws == nullptr || ws_bytes < size_t(CB200_WORKSPACE_COUNTER_BYTES))) {
    set_error("cb200_mpgemm_push_signal / cb200_mpgemm_chain need the %d-byte counter area of the workspace", CB200_WORKSPACE_COUNTER_BYTES);
    return CB200_ERR_WORKSPACE;
  }
  if (p_in.wait_flags != nullptr && p_in.M > 8) {
    set_error("cb200_mpgemm_chain: one 8-row pass of the streaming decode kernel (M <= 8)");
    return CB200_ERR_UNSUPPORTED;
  }
  if (size_t(p_in.K / p_in.g) * p_in.N >= (size_t(1) << 31)) {
    set_error("scale matrix too large");
    return CB200_ERR_SHAPE;
  }
  for (int m0 = 0; m0 < p_in.M; m0 += 16) {
    Problem p = p_in;
    const int mrows = (p_in.M - m0) < 16 ? (p_in.M - m0) : 16;
    const size_t out_es = p_in.out_dtype == CB200_F32 ? 4 : 2;
    p.x = reinterpret_cast<const char*>(p_in.x) + size_t(m0) * p_in.K * 2;
    p.out = reinterpret_cast<char*>(p_in.out) + size_t(m0) * p_in.N * out_es;
    for (int i = 0; i < p_in.n_peers; ++i) p.peer_out[i] = reinterpret_cast<char*>(p_in.peer_out[i]) + size_t(m0) * p_in.N * out_es;
    p.M = mrows;

    const StreamPlan pl = stream_plan(mrows, p.K, p.N, bits, p.g, p.zp_kind, p.zp_scalar != 0);
    if (pl.smem > size_t(227) * 1024) {
      set_error("streaming decode kernel: %d rows of K=%d do not fit shared memory next to the weight ring", mrows, p.K);
      return CB200_ERR_SHAPE;
    }
    const CUtensorMap* mw = nullptr;
    if (int rc = stream_tmap(p, pl.pf, pl.nw, &mw)) return rc;
    StreamGeom geo{};
    geo.m_rows = mrows;
    geo.nw = pl.nw;
    geo.ngrp = pl.ngrp;
    geo.S = pl.S;
    geo.QPB = pl.QPB;
    geo.qpg = pl.qpg;
    geo.one_group = pl.one_group;
    geo.gps = pl.gps;
    geo.ngroups = p.K / p.g;
    geo.zes = pl.zes;
    geo.w_bytes = 2048 * pl.nw;
    geo.stage_bytes = pl.stage_bytes;
    geo.ring = pl.ring;
    geo.x_pitch = pl.x_pitch;
    geo.x_bytes = pl.x_bytes;
    geo.red_bytes = pl.red_bytes;
    geo.xsum_bytes = pl.xsum_bytes;
    geo.win_stages = pl.win_stages;
    geo.done = ws ? reinterpret_cast<unsigned int*>(ws) + (CB200_WORKSPACE_COUNTER_BYTES / 4 - 1) : nullptr;
    geo.no_hint = g_stream_tuning.dbg & 1;
    geo.trace = g_stream_trace;
    const bool pdl = g_decode_pdl != 0;
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_stream_f16_w4(*mw, p, geo, pl.grid, pl.smem, pdl, stream)
                     : launch_stream_f16_w8(*mw, p, geo, pl.grid, pl.smem, pdl, stream);
    } else {
      rc = bits == 4 ? launch_stream_bf16_w4(*mw, p, geo, pl.grid, pl.smem, pdl, stream)
                     : launch_stream_bf16_w8(*mw, p, geo, pl.grid, pl.smem, pdl, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
