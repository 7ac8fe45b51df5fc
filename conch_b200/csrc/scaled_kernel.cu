// scaled_gemm: out[M,N] = cast( (a[M,K] @ b[K,N]) * (scale_a[m] * scale_b[n]) ) (+ bias[n]) for 8-bit operands -- int8 x int8
// with int32 accumulation, fp8 e4m3 x e4m3 with fp32 accumulation.  SURVEY.md section 8(f) row 3.
//
// Replaces: scaled_gemm_launcher -> the Triton `_gemm_kernel` with WEIGHT_AND_ACTIVATION channel scales
//           (reference conch/kernels/quantization/gemm.py:564-627, epilogue :408-440) and the `output.add_(bias)` kernel that
//           follows it (conch/ops/quantization/gemm.py:249-250).
//
// Design: a plain warp-specialised tcgen05 GEMM -- there is nothing to dequantise, both operands go from HBM to the tensor
// core untouched.
//   * 128 x BN output tile per CTA (BN = 256, or 128 for small problems), K in 128-byte blocks; warp 0 = TMA producer (a tile
//     [128 m x 128 B] and b^T tile [BN n x 128 B], both K-major, 128-byte swizzled, a 3/4-stage ring), warp 1 = MMA issuer
//     (tcgen05.mma kind::i8 / kind::f8f6f4, M = 128, N = BN, K = 32 per instruction, accumulator in TMEM), warps 2-5 = epilogue
//     (tcgen05.ld, scales, rounding, 16-byte stores).
//   * BN = 128: two CTAs per SM (96 KB of ring + 128 TMEM columns each), one CTA's epilogue overlaps the other's main loop;
//     BN = 256: one CTA per SM with a 192 KB ring -- 175 instead of 128 flop per operand byte, which an 8-bit tile needs.
//   * arithmetic of the epilogue is the reference kernel's, operation for operation: acc -> fp32, times (scale_a * scale_b)
//     (the scale PRODUCT first, :431-440), one rounding to the output dtype; a bias is then added to the ROUNDED value and
//     rounded again, exactly what `output.add_(bias)` computes -- so int8 results are bit-identical to the reference op's.
#include "host_util.cuh"
#include "tc_ptx.cuh"

namespace cb200 {

struct ScaledParams {
  const float* sa;   // [M] or [1]
  const float* sb;   // [N] or [1]
  const void* bias;  // [N] output dtype, or null
  void* out;         // [M, N]
  int M, N, K;
  int sa_scalar, sb_scalar;
  int out_dtype;
  int n_tiles, kblocks;
};

namespace scaled {
constexpr int BM = 128, BK = 128;
template <int BN, int STAGES>
struct Tile {
  static constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
};

template <int KIND>
__device__ __forceinline__ void umma_ss8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// RN to the output dtype and back: the value the reference op holds after its single cast
__device__ __forceinline__ float round_out(float v, int out_dtype) {
  if (out_dtype == CB200_F16) return __half2float(__float2half_rn(v));
  if (out_dtype == CB200_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}
}  // namespace scaled

// KIND 0: int8 (s8 x s8 -> s32), 1: fp8 e4m3 (-> f32); BN: tile width (UMMA N), STAGES: ring depth
template <int KIND, int BN, int STAGES>
__global__ void __launch_bounds__(192, BN == 128 ? 2 : 1)
    scaled_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ScaledParams prm) {
  using namespace scaled;
  constexpr int A_BYTES = Tile<BN, STAGES>::A_BYTES, STAGE_BYTES = Tile<BN, STAGES>::STAGE_BYTES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bars + 8 * s; };
  auto empty = [&](int s) { return bars + 8 * (STAGES + s); };
  const uint32_t d_full = bars + 8 * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x % prm.n_tiles) * BN, m0 = (blockIdx.x / prm.n_tiles) * BM;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);  // BN fp32 / int32 columns x 128 lanes: the accumulator
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer ----------------------------------------
    if (elect_one()) {
      for (int kb = 0; kb < prm.kblocks; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(empty(s), ((kb / STAGES) - 1) & 1);
        mbar_expect_tx(full(s), STAGE_BYTES);
        tma_load_2d(base + s * STAGE_BYTES, &tmap_a, kb * BK, m0, full(s));
        tma_load_2d(base + s * STAGE_BYTES + A_BYTES, &tmap_b, kb * BK, n0, full(s));
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer ------------------------------------------
    // c_format: S32 (2) / F32 (1); a / b format: signed 8 bit (1) / E4M3 (0); both K-major; N >> 3, M >> 4
    constexpr uint32_t idesc = ((KIND == 0 ? 2u : 1u) << 4) | ((KIND == 0 ? 1u : 0u) << 7) | ((KIND == 0 ? 1u : 0u) << 10) |
                               (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    if (elect_one()) {
      for (int kb = 0; kb < prm.kblocks; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(full(s), (kb / STAGES) & 1);
        tc_fence_after();
        const uint64_t adesc = make_sw128_desc(base + s * STAGE_BYTES), bdesc = make_sw128_desc(base + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 32; ++k)
          umma_ss8<KIND>(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty(s));  // the stage may be refilled once these MMAs have read it
      }
      umma_commit(d_full);
    }
  } else {
    // ---------------------------------------------------------------- epilogue --------------------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = m0 + q * 32 + lane;
    mbar_wait(d_full, 0);
    tc_fence_after();
    const float sa = m < prm.M ? __ldg(prm.sa + (prm.sa_scalar ? 0 : m)) : 0.f;
    const size_t es = prm.out_dtype == CB200_F32 ? 4 : 2;
    char* orow = reinterpret_cast<char*>(prm.out) + (size_t(m) * prm.N + n0) * es;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c), r);
      tmem_ld_wait();
      if (m >= prm.M) continue;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        const int n = n0 + c + j0;
        if (n >= prm.N) break;  // N % 8 == 0: a group of eight columns is inside or outside as a whole
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float acc = KIND == 0 ? __int2float_rn(int(r[j0 + j])) : __uint_as_float(r[j0 + j]);
          const float sb = __ldg(prm.sb + (prm.sb_scalar ? 0 : n + j));
          v[j] = acc * (sa * sb);
          if (prm.bias != nullptr) {
            float b;
            if (prm.out_dtype == CB200_F16) b = __half2float(reinterpret_cast<const __half*>(prm.bias)[n + j]);
            else if (prm.out_dtype == CB200_BF16) b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(prm.bias)[n + j]);
            else b = reinterpret_cast<const float*>(prm.bias)[n + j];
            v[j] = round_out(v[j], prm.out_dtype) + b;
          }
        }
        if (prm.out_dtype == CB200_F32) {
          float4* dst = reinterpret_cast<float4*>(orow + size_t(c + j0) * 4);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          uint4 pk;
          uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (prm.out_dtype == CB200_F16) {
              const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              w[j] = *reinterpret_cast<const uint32_t*>(&h);
            } else {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
          }
          *reinterpret_cast<uint4*>(orow + size_t(c + j0) * 2) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

template <int KIND, int BN, int STAGES>
static int launch_scaled_kind(const CUtensorMap& ma, const CUtensorMap& mb, const ScaledParams& prm, int tiles, cudaStream_t stream) {
  static unsigned long long attr_mask = 0;
  constexpr int SMEM = scaled::Tile<BN, STAGES>::SMEM_BYTES;
  if (int rc = set_max_smem_once(scaled_kernel<KIND, BN, STAGES>, SMEM, attr_mask)) return rc;
  scaled_kernel<KIND, BN, STAGES><<<tiles, 192, SMEM, stream>>>(ma, mb, prm);
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int g_scaled_tile = 0;  // debug: 128 / 256 forces the tile width (0 = heuristic)

int launch_scaled(const void* a, const void* b_t, const float* scale_a, const float* scale_b, const void* bias, void* out, int M,
                  int K, int N, int in_kind, int out_dtype, int sa_scalar, int sb_scalar, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return CB200_ERR_DEVICE;
  }
  // tile width: 256 columns halve the re-reads of `a` (175 instead of 128 flop per byte of operand traffic -- at 8 bits a
  // 128 x 128 tile wants more bytes per clock than an SM can pull from L2)
  const int sms = current_device_info() ? current_device_info()->sms : 148;
  const int m_tiles = (M + scaled::BM - 1) / scaled::BM;
  // (measured, profiles/r02_scaled_gemm_bench.log: with one CTA per SM the 256-wide tile's un-overlapped epilogue costs more
  //  than its operand traffic saves -- 1097 vs 1449 TFLOP/s at 4096 x 4096 x 14336 -- so 128 stays the default until the
  //  kernel is persistent with a second TMEM accumulator)
  (void)sms;
  int bn = 128;
  if (g_scaled_tile == 128 || g_scaled_tile == 256) bn = g_scaled_tile;
  CUtensorMap ma, mb;
  const cuuint32_t box[2] = {scaled::BK, scaled::BM};
  const cuuint32_t box_b[2] = {scaled::BK, cuuint32_t(bn)};
  const cuuint32_t estr[2] = {1, 1};
  const cuuint64_t strides[1] = {cuuint64_t(K)};
  {
    const cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(M)};
    const CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(a), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(a) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  {
    const cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(N)};
    const CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(b_t), dims, strides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(b) failed: %d", int(r));
      return CB200_ERR_ARG;
    }
  }
  ScaledParams prm{};
  prm.sa = scale_a;
  prm.sb = scale_b;
  prm.bias = bias;
  prm.out = out;
  prm.M = M;
  prm.N = N;
  prm.K = K;
  prm.sa_scalar = sa_scalar;
  prm.sb_scalar = sb_scalar;
  prm.out_dtype = out_dtype;
  prm.n_tiles = (N + bn - 1) / bn;
  prm.kblocks = (K + scaled::BK - 1) / scaled::BK;
  const int tiles = prm.n_tiles * m_tiles;
  if (bn == 256) {
    return in_kind == CB200_IN_INT8 ? launch_scaled_kind<0, 256, 4>(ma, mb, prm, tiles, stream)
                                    : launch_scaled_kind<1, 256, 4>(ma, mb, prm, tiles, stream);
  }
  return in_kind == CB200_IN_INT8 ? launch_scaled_kind<0, 128, 3>(ma, mb, prm, tiles, stream)
                                  : launch_scaled_kind<1, 128, 3>(ma, mb, prm, tiles, stream);
}

}  // namespace cb200
