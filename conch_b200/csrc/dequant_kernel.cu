// Standalone dequantise kernel: w_q_packed[K/pf, N] (+ scales, zero-points) -> W[K, N] in the activation dtype.
// This is the `_dequantize` stage of the reference (conch/kernels/quantization/gemm.py:176-216) on its own;
// it exists so the bit-exactness of the shared dequant_word() routine can be tested against the reference's
// `w_ref` (conch/third_party/vllm/quant_utils.py:74) independently of any GEMM.
//
// HBM-bound: reads bits/8 B and writes 2 B per weight.  One thread owns 4 adjacent columns of one packed row
// (one 128-bit load), produces pf rows x 4 columns and stores 8-byte row segments; a warp therefore writes
// 256 contiguous bytes per k row.
#include "common.cuh"
#include "dequant.cuh"

namespace cb200 {

template <class A, int BITS, bool FZP>
__global__ void __launch_bounds__(256) dequant_kernel(Problem p) {
  constexpr int PF = 32 / BITS;
  const int n4 = p.N >> 2;                       // 4-column groups per row
  const int64_t total = int64_t(p.K / PF) * n4;  // one item = (packed row, 4 columns)
  const A* sp = reinterpret_cast<const A*>(p.s);
  A* outp = reinterpret_cast<A*>(p.out);
  for (int64_t item = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; item < total;
       item += int64_t(gridDim.x) * blockDim.x) {
    const int kp = int(item / n4);
    const int n0 = int(item % n4) * 4;
    const int grp = (kp * PF) / p.g;
    const uint4 w4 = *reinterpret_cast<const uint4*>(p.wq + size_t(kp) * p.N + n0);
    const uint32_t words[4] = {w4.x, w4.y, w4.z, w4.w};
    uint32_t regs[4][16 / BITS];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t midx = p.zp_scalar ? 0 : size_t(grp) * p.N + n0 + j;
      float zb = float(p.bias);
      A zpf = A(0.f);
      if constexpr (FZP) {
        zpf = reinterpret_cast<const A*>(p.zp)[midx];
      } else {
        zb += load_zp_int<A>(p.zp, p.zp_kind, midx);
      }
      const ColQ c = make_colq<A, BITS, Order::Natural, FZP>(sp[size_t(grp) * p.N + n0 + j], zb, zpf);
      dequant_word<A, BITS, Order::Natural, FZP>(words[j], c, regs[j]);
    }
    // regs[j][i] = (k = 2i, 2i+1) of column n0+j  ->  row k gets 4 consecutive columns (8 bytes)
#pragma unroll
    for (int i = 0; i < PF / 2; ++i) {
      uint2 lo, hi;
      lo.x = __byte_perm(regs[0][i], regs[1][i], 0x5410);  // (col0.k_even, col1.k_even)
      lo.y = __byte_perm(regs[2][i], regs[3][i], 0x5410);
      hi.x = __byte_perm(regs[0][i], regs[1][i], 0x7632);  // odd k
      hi.y = __byte_perm(regs[2][i], regs[3][i], 0x7632);
      const size_t row = size_t(kp) * PF + 2 * i;
      *reinterpret_cast<uint2*>(outp + row * p.N + n0) = lo;
      *reinterpret_cast<uint2*>(outp + (row + 1) * p.N + n0) = hi;
    }
  }
}

template <class A, int BITS>
static int launch_dq(const Problem& p, cudaStream_t stream) {
  const int64_t items = int64_t(p.K / (32 / BITS)) * (p.N / 4);
  int blocks = int((items + 255) / 256);
  const int cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (p.zp_kind == CB200_ZP_ACT) {
    dequant_kernel<A, BITS, true><<<blocks, 256, 0, stream>>>(p);
  } else {
    dequant_kernel<A, BITS, false><<<blocks, 256, 0, stream>>>(p);
  }
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_dequant(const Problem& p, int bits, int act_dtype, cudaStream_t stream) {
  if (act_dtype == CB200_F16) {
    return bits == 4 ? launch_dq<__half, 4>(p, stream) : launch_dq<__half, 8>(p, stream);
  }
  return bits == 4 ? launch_dq<__nv_bfloat16, 4>(p, stream) : launch_dq<__nv_bfloat16, 8>(p, stream);
}

// ---- cb200_wait_flags: the consumer side of cb200_mpgemm_push_signal ------------------------------------------------------------
// One warp.  `*expected` counts the waits done so far on this flag array (device memory, advanced by the kernel itself, so a
// replayed CUDA graph keeps counting); lane i spins (system-scope acquire loads) until counter flags[i] has reached it.
// Bounded: a lost signal becomes a trap (launch error), not a hung GPU.
__global__ void wait_flags_kernel(const uint32_t* flags, int count, uint32_t* expected) {
  const int lane = threadIdx.x;
  const uint32_t want = *expected + 1u;
  if (lane < count) {
    const uint32_t* f = flags + lane;
    for (unsigned spin = 0;; ++spin) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (int32_t(v - want) >= 0) break;
      if (spin > (1u << 28)) __trap();
    }
  }
  __syncwarp();
  if (lane == 0) *expected = want;
}

int launch_wait_flags(const uint32_t* flags, int count, uint32_t* expected, cudaStream_t stream) {
  wait_flags_kernel<<<1, 32, 0, stream>>>(flags, count, expected);
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cb200
