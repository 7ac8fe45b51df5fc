// Decode-regime kernel (M <= 16 per pass): split-K skinny GEMM, HBM-bound.
//
// Replaces the reference's single 128x64x64 Triton tile (conch/kernels/quantization/gemm.py:136-145, 491-494),
// which at M=1 launches N/64 CTAs with no split-K.  Design (B200):
//   * the packed weights are the only large stream: every lane issues 64- or 128-bit coalesced, read-once
//     loads (ld.global.nc, L1 no-allocate, L2 evict-first) straight into registers, U steps deep, so
//     >= 32 KB per SM is in flight;
//   * nibbles are unpacked and (q - bias - zp) * s is formed in registers with dequant_word() -- the same
//     routine, hence the same bits, as the standalone dequant kernel;
//   * the dot product runs on the tensor pipe as mma.sync.m16n8k16 with the WEIGHTS as the 16-row operand
//     (swap-AB) and x^T as the 8-column operand: on B200 an FFMA dot product would be issue-bound
//     (~4 instr/weight vs a budget of ~2.7 at the HBM roofline), HMMA costs 1 instr / 256 weights;
//     the four lanes of a quad hold four different packed rows, so the K-reduction across lanes happens
//     inside the MMA, fp32 accumulate;
//   * K is split across the warps of a CTA (and optionally across CTAs); warps reduce through shared
//     memory, CTAs through an fp32 workspace + arrival counter (last CTA sums, no memset, no second launch);
//   * x is staged once per CTA in shared memory, pre-permuted to the nibble order the cheap `lop3`
//     extraction produces ((k0,k4),(k1,k5),(k2,k6),(k3,k7)), so no byte permutes are spent on weights.
#include "common.cuh"
#include "dequant.cuh"

namespace cb200 {

DecodeTuning g_decode_tuning = {0, 0, 0, 0};

struct DecodeGeom {
  int warps_n;      // warps along N per CTA
  int warps_k;      // warps along K per CTA
  int split_k;      // CTAs along K (gridDim.y)
  int steps_cta;    // steps (4 packed rows each) per CTA
  int xs_stride;    // bytes between consecutive m rows of the staged x
  int m_rows;       // rows of x in this pass (<= 16)
};

template <class A>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<A, __half>::value) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

template <int VEC>
struct WVec;
template <>
struct WVec<2> {
  uint32_t w[2];
  __device__ __forceinline__ void load(const uint32_t* p) {
    const uint2 v = ldg_stream_v2(p);
    w[0] = v.x; w[1] = v.y;
  }
};
template <>
struct WVec<4> {
  uint32_t w[4];
  __device__ __forceinline__ void load(const uint32_t* p) {
    const uint4 v = ldg_stream_v4(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
};

// A: activation type, BITS: 4|8, MT: number of 8-row m tiles (1: M<=8, 2: M<=16), VEC: packed words per lane
// per load (2 -> 16 columns per warp, 4 -> 32), FZP: zero-points are in the activation dtype.
template <class A, int BITS, int MT, int VEC, bool FZP>
__global__ void __launch_bounds__(512) decode_kernel(Problem p, DecodeGeom geo, float* __restrict__ ws_partial,
                                                     unsigned int* __restrict__ ws_counter) {
  constexpr int PF = 32 / BITS;        // k per packed word
  constexpr int STEP_K = 4 * PF;       // k per step (4 packed rows, one per lane of a quad)
  constexpr int CH = PF * 2;           // bytes of x per packed row
  constexpr int NT = VEC / 2;          // 16-column MMA tiles per warp
  constexpr int CW = 8 * VEC;          // columns per warp
  constexpr int U = (VEC == 2) ? 8 : 4;  // steps per load batch
  constexpr int PAIRS = 16 / BITS;     // pair registers per word
  constexpr Order ORD = Order::Lop;

  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* xs = smem;                                               // [m_rows][xs_stride]
  float* red = reinterpret_cast<float*>(smem + size_t(geo.m_rows) * geo.xs_stride);  // [warps][MT*8][CW]
  __shared__ int s_is_last;

  const int tid = threadIdx.x;
  const int wid = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, tig = lane & 3;
  const int wn = wid % geo.warps_n, wk = wid / geo.warps_n;
  const int M = geo.m_rows;
  const int N = p.N;
  const int nb = (blockIdx.x * geo.warps_n + wn) * CW;  // warp's first column
  const int c0 = nb + VEC * g8;                          // lane's first column
  const bool col_ok = c0 < N;                            // N % CW != 0 tail (whole-lane granularity)

  const int step_cta0 = blockIdx.y * geo.steps_cta;
  const int s0 = step_cta0 + int((int64_t(geo.steps_cta) * wk) / geo.warps_k);
  const int s1 = step_cta0 + int((int64_t(geo.steps_cta) * (wk + 1)) / geo.warps_k);

  // ---- 1. get the first batch of weight loads in flight before anything else --------------------------
  WVec<VEC> buf[U];
  const uint32_t* wbase = p.wq + size_t(tig) * N + (col_ok ? c0 : 0);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (s0 + u < s1) buf[u].load(wbase + size_t(4) * (s0 + u) * N);
  }

  // ---- 2. stage x[:, K-range of this CTA] into shared memory, permuted to the lop3 nibble order --------
  {
    const int chunks = geo.steps_cta * 4;  // packed rows in this CTA's K range
    const A* xg = reinterpret_cast<const A*>(p.x);
    const size_t k0 = size_t(step_cta0) * STEP_K;
    for (int idx = tid; idx < M * chunks; idx += blockDim.x) {
      const int m = idx / chunks, r = idx - m * chunks;
      const A* src = xg + size_t(m) * p.K + k0 + size_t(r) * PF;
      unsigned char* dst = xs + size_t(m) * geo.xs_stride + size_t(r) * CH;
      if constexpr (BITS == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(src);
        uint4 o;
        o.x = __byte_perm(v.x, v.z, 0x5410);  // (k0,k4)
        o.y = __byte_perm(v.x, v.z, 0x7632);  // (k1,k5)
        o.z = __byte_perm(v.y, v.w, 0x5410);  // (k2,k6)
        o.w = __byte_perm(v.y, v.w, 0x7632);  // (k3,k7)
        *reinterpret_cast<uint4*>(dst) = o;
      } else {
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
      }
    }
  }

  // ---- 3. per-group constants (scale, bias + zero-point), one group prefetched ahead -------------------
  const int spg = p.g / STEP_K;  // steps per group (g % 64 == 0 guaranteed by the host)
  const A* sp = reinterpret_cast<const A*>(p.s);
  A raw_s[VEC];
  float raw_z[VEC];
  A raw_zf[VEC];
  auto fetch_group = [&](int grp) {
    const int ngrp = p.K / p.g;
    if (grp >= ngrp) grp = ngrp - 1;
    const size_t base = size_t(grp) * N + (col_ok ? c0 : 0);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      raw_s[j] = sp[base + j];
      const size_t zi = p.zp_scalar ? 0 : base + j;
      if constexpr (FZP) {
        raw_zf[j] = reinterpret_cast<const A*>(p.zp)[zi];
        raw_z[j] = 0.f;
      } else {
        raw_zf[j] = A(0.f);
        raw_z[j] = load_zp_int<A>(p.zp, p.zp_kind, zi);
      }
    }
  };
  ColQ cq[VEC];
  int cur_grp = (s0 * STEP_K) / p.g;
  fetch_group(cur_grp);

  float acc[NT][MT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][mt][i] = 0.f;

  __syncthreads();  // x staged

  int grp_end_step = s0;  // forces a constant refresh on the first step
  const unsigned char* xrow[MT];
  bool m_ok[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int m = mt * 8 + g8;
    m_ok[mt] = m < M;
    xrow[mt] = xs + size_t(m_ok[mt] ? m : 0) * geo.xs_stride + size_t(tig) * CH - size_t(step_cta0) * 4 * CH;
  }

  // ---- 4. main loop: U steps per batch, next batch's loads issued before this batch is consumed --------
  for (int sb = s0; sb < s1; sb += U) {
    WVec<VEC> cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = buf[u];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (sb + U + u < s1) buf[u].load(wbase + size_t(4) * (sb + U + u) * N);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int st = sb + u;
      if (st >= s1) break;  // warp-uniform
      if (st >= grp_end_step) {
        cur_grp = (st * STEP_K) / p.g;
        grp_end_step = (cur_grp + 1) * spg;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          cq[j] = make_colq<A, BITS, ORD, FZP>(raw_s[j], raw_z[j] + float(p.bias), raw_zf[j]);
        }
        fetch_group(cur_grp + 1);
      }
      uint32_t d[VEC][PAIRS];
#pragma unroll
      for (int j = 0; j < VEC; ++j) dequant_word<A, BITS, ORD, FZP>(cur[u].w[j], cq[j], d[j]);

#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t xb[PAIRS];
        if constexpr (BITS == 4) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (m_ok[mt]) v = *reinterpret_cast<const uint4*>(xrow[mt] + size_t(st) * 4 * CH);
          xb[0] = v.x; xb[1] = v.y; xb[2] = v.z; xb[3] = v.w;
        } else {
          uint2 v = make_uint2(0u, 0u);
          if (m_ok[mt]) v = *reinterpret_cast<const uint2*>(xrow[mt] + size_t(st) * 4 * CH);
          xb[0] = v.x; xb[1] = v.y;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          // rows g8 of the MMA <- column 2t of this lane, rows g8+8 <- column 2t+1
#pragma unroll
          for (int h = 0; h < PAIRS / 2; ++h) {
            mma16816<A>(acc[t][mt], d[2 * t][2 * h], d[2 * t + 1][2 * h], d[2 * t][2 * h + 1], d[2 * t + 1][2 * h + 1],
                        xb[2 * h], xb[2 * h + 1]);
          }
        }
      }
    }
  }

  // ---- 5. reduce the K-split warps through shared memory -------------------------------------------------
  {
    float* mine = red + size_t(wid) * (MT * 8) * CW;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int m = mt * 8 + 2 * tig;
        const int lc = VEC * g8 + 2 * t;
        mine[(m)*CW + lc] = acc[t][mt][0];
        mine[(m + 1) * CW + lc] = acc[t][mt][1];
        mine[(m)*CW + lc + 1] = acc[t][mt][2];
        mine[(m + 1) * CW + lc + 1] = acc[t][mt][3];
      }
  }
  __syncthreads();

  const int cols_cta = geo.warps_n * CW;
  const int n_cta0 = blockIdx.x * cols_cta;
  const int outputs = M * cols_cta;
  const bool split = geo.split_k > 1;
  for (int e = tid; e < outputs; e += blockDim.x) {
    const int m = e / cols_cta, c = e - m * cols_cta;
    const int w_n = c / CW, lc = c - w_n * CW;
    float v = 0.f;
    for (int k = 0; k < geo.warps_k; ++k) v += red[(size_t(k * geo.warps_n + w_n) * (MT * 8) + m) * CW + lc];
    const int n = n_cta0 + c;
    if (n < N) {
      if (!split) {
        store_out<A>(p.out, p.out_dtype, size_t(m) * N + n, v);
      } else {
        __stcg(ws_partial + (size_t(blockIdx.y) * M + m) * N + n, v);
      }
    }
  }
  if (!split) return;

  // ---- 6. cross-CTA split-K: the last CTA to arrive for this column block sums the partials ---------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(ws_counter + blockIdx.x, 1u);
    s_is_last = (prev == unsigned(geo.split_k - 1));
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  for (int e = tid; e < outputs; e += blockDim.x) {
    const int m = e / cols_cta, c = e - m * cols_cta;
    const int n = n_cta0 + c;
    if (n >= N) continue;
    float v = 0.f;
    for (int k = 0; k < geo.split_k; ++k) v += __ldcg(ws_partial + (size_t(k) * M + m) * N + n);
    store_out<A>(p.out, p.out_dtype, size_t(m) * N + n, v);
  }
  if (tid == 0) ws_counter[blockIdx.x] = 0u;  // leave the counter clean for the next launch
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
size_t decode_workspace_bytes(int64_t M, int64_t K, int64_t N) {
  (void)K;
  const int64_t mrows = M < 16 ? M : 16;
  // up to 32 K-splits of fp32 partials for one pass of <= 16 rows
  return size_t(CB200_WORKSPACE_COUNTER_BYTES) + size_t(32) * mrows * N * sizeof(float);
}

static int pick(int forced, int heuristic) { return forced > 0 ? forced : heuristic; }

template <class A, int BITS, int MT, int VEC, bool FZP>
static int launch_one(const Problem& p, const DecodeGeom& geo, size_t smem, dim3 grid, int threads, float* wsp,
                      unsigned* wsc, cudaStream_t stream) {
  auto kern = decode_kernel<A, BITS, MT, VEC, FZP>;
  if (smem > 48 * 1024) {
    CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  }
  kern<<<grid, threads, smem, stream>>>(p, geo, wsp, wsc);
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <class A, int BITS>
static int launch_typed(const Problem& p, const DecodeGeom& geo, int vec, size_t smem, dim3 grid, int threads,
                        float* wsp, unsigned* wsc, cudaStream_t stream) {
  const bool fzp = p.zp_kind == CB200_ZP_ACT;
  const int mt = geo.m_rows > 8 ? 2 : 1;
  if (fzp) {
    return mt == 1 ? launch_one<A, BITS, 1, 2, true>(p, geo, smem, grid, threads, wsp, wsc, stream)
                   : launch_one<A, BITS, 2, 2, true>(p, geo, smem, grid, threads, wsp, wsc, stream);
  }
  if (vec == 4) {
    return mt == 1 ? launch_one<A, BITS, 1, 4, false>(p, geo, smem, grid, threads, wsp, wsc, stream)
                   : launch_one<A, BITS, 2, 4, false>(p, geo, smem, grid, threads, wsp, wsc, stream);
  }
  return mt == 1 ? launch_one<A, BITS, 1, 2, false>(p, geo, smem, grid, threads, wsp, wsc, stream)
                 : launch_one<A, BITS, 2, 2, false>(p, geo, smem, grid, threads, wsp, wsc, stream);
}

int launch_decode(const Problem& p_in, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream,
                  int* launches) {
  const int pf = 32 / bits;
  const int step_k = 4 * pf;
  const int steps_total = p_in.K / step_k;
  const int ch = pf * 2;
  const bool fzp = p_in.zp_kind == CB200_ZP_ACT;
  int vec = fzp ? 2 : pick(g_decode_tuning.vec_words, 2);
  if (vec != 2 && vec != 4) vec = 2;
  const int cw = 8 * vec;

  for (int m0 = 0; m0 < p_in.M; m0 += 16) {
    Problem p = p_in;
    const int mrows = (p_in.M - m0) < 16 ? (p_in.M - m0) : 16;
    p.x = reinterpret_cast<const char*>(p_in.x) + size_t(m0) * p_in.K * 2;
    const size_t out_es = p_in.out_dtype == CB200_F32 ? 4 : 2;
    p.out = reinterpret_cast<char*>(p_in.out) + size_t(m0) * p_in.N * out_es;
    p.M = mrows;

    // ---- tiling heuristic (every knob overridable through cb200_set_decode_tuning) ----------------------
    const int col_blocks_w1 = (p.N + cw - 1) / cw;  // CTAs if one warp-column per CTA
    int warps_n = pick(g_decode_tuning.warps_n, col_blocks_w1 > 148 * 6 ? 2 : 1);
    int warps_k = pick(g_decode_tuning.warps_k, 16 / warps_n);
    if (warps_n * warps_k > 16) warps_k = 16 / warps_n;
    // split K across CTAs only when x for the whole K does not fit the smem budget or the grid is too small
    int split_k = 1;
    const size_t x_budget = 96 * 1024;
    while (size_t(mrows) * (size_t(p.K) / split_k) * 2 > x_budget && split_k < 32) split_k *= 2;
    const int col_blocks = (p.N + cw * warps_n - 1) / (cw * warps_n);
    while (col_blocks * split_k < 148 && split_k < 32 && steps_total / (split_k * 2 * warps_k) >= 2) split_k *= 2;
    split_k = pick(g_decode_tuning.split_k, split_k);
    while (split_k > 1 && steps_total % split_k != 0) --split_k;
    if (split_k > 32) split_k = 32;
    if (split_k > 1) {
      const size_t need = size_t(CB200_WORKSPACE_COUNTER_BYTES) + size_t(split_k) * mrows * p.N * sizeof(float);
      if (ws == nullptr || ws_bytes < need) {
        set_error("decode split-K=%d needs %zu workspace bytes, got %zu", split_k, need, ws_bytes);
        return CB200_ERR_WORKSPACE;
      }
      if (col_blocks * sizeof(unsigned) > CB200_WORKSPACE_COUNTER_BYTES) {
        set_error("N=%d too large for the split-K counter area", p.N);
        return CB200_ERR_SHAPE;
      }
    }

    DecodeGeom geo;
    geo.warps_n = warps_n;
    geo.warps_k = warps_k;
    geo.split_k = split_k;
    geo.steps_cta = steps_total / split_k;
    const int chunks = geo.steps_cta * 4;
    // pad the m-row stride so the 8 (or 4) different m rows a load phase touches fall in distinct banks
    int stride = chunks * ch;
    stride += (bits == 4) ? ((stride % 128 == 64) ? 0 : (64 + 128 - stride % 128) % 128)
                          : ((stride % 128 == 32) ? 0 : (32 + 128 - stride % 128) % 128);
    geo.xs_stride = stride;
    geo.m_rows = mrows;
    const int mt = mrows > 8 ? 2 : 1;
    const size_t smem = size_t(mrows) * stride + size_t(warps_n * warps_k) * (mt * 8) * cw * sizeof(float);
    if (smem > 200 * 1024) {
      set_error("decode kernel shared memory %zu exceeds budget (M=%d K=%d)", smem, mrows, p.K);
      return CB200_ERR_SHAPE;
    }
    dim3 grid(col_blocks, split_k);
    const int threads = 32 * warps_n * warps_k;
    float* wsp = ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + CB200_WORKSPACE_COUNTER_BYTES) : nullptr;
    unsigned* wsc = reinterpret_cast<unsigned*>(ws);
    int rc;
    if (act_dtype == CB200_F16) {
      rc = bits == 4 ? launch_typed<__half, 4>(p, geo, vec, smem, grid, threads, wsp, wsc, stream)
                     : launch_typed<__half, 8>(p, geo, vec, smem, grid, threads, wsp, wsc, stream);
    } else {
      rc = bits == 4 ? launch_typed<__nv_bfloat16, 4>(p, geo, vec, smem, grid, threads, wsp, wsc, stream)
                     : launch_typed<__nv_bfloat16, 8>(p, geo, vec, smem, grid, threads, wsp, wsc, stream);
    }
    if (rc != 0) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace cb200
