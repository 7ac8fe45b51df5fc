// Shared host/device plumbing for libconch_b200: problem descriptor, error reporting, small PTX wrappers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/conch_b200.h"

namespace cb200 {

// One mixed-precision GEMM problem, as the kernels see it (all device pointers).
constexpr int CB200_MAX_PEERS = 7;  // 8 GPUs per node

struct Problem {
  const void* x;       // [M, K] act
  const uint32_t* wq;  // [K/pf, N] packed
  const void* s;       // [K/g, N] act
  const void* zp;      // [K/g, N] / [1] / null
  void* out;           // [M, N]
  int M, K, N;
  int bias;         // weight_bias
  int g;            // group size in k (already resolved: -1 -> K)
  int zp_kind;      // cb200_zp_kind
  int zp_scalar;    // w_zp.numel() == 1
  int out_dtype;    // cb200_dtype
  int n_peers;      // cb200_mpgemm_push: the result is also stored into peer_out[0 .. n_peers) (same [M, N] layout)
  void* peer_out[CB200_MAX_PEERS];
  // cb200_mpgemm_push_signal: once the WHOLE result has been pushed, every peer_flag (a 32-bit counter in the peer's memory) and
  // local_flag are incremented by one: per-GEMM visibility without a cross-GPU barrier, safe under CUDA-graph replay
  uint32_t* peer_flag[CB200_MAX_PEERS];
  uint32_t* local_flag;
  // cb200_mpgemm_chain: x is itself a gathered buffer other GPUs push into -- it is complete once wait_flags[0 .. wait_count) have
  // all reached *wait_expected + 1; the kernel's consumers spin on them (after griddepcontrol.wait, while the producer warp keeps
  // filling the weight ring) and the grid's last block advances *wait_expected.  null: x belongs to the stream order.
  const uint32_t* wait_flags;
  uint32_t* wait_expected;
  int wait_count;
  const void* out_bias;  // cb200_mpgemm_bias: [N] in the OUTPUT dtype, added to the fp32 sum before the one rounding (or null)
};

// bias[n] as fp32 (0 when there is none): every kernel adds it where it forms the final fp32 value of out[m, n]
__device__ __forceinline__ float bias_at(const Problem& p, int n) {
  if (p.out_bias == nullptr) return 0.f;
  if (p.out_dtype == CB200_F16) return __half2float(reinterpret_cast<const __half*>(p.out_bias)[n]);
  if (p.out_dtype == CB200_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.out_bias)[n]);
  return reinterpret_cast<const float*>(p.out_bias)[n];
}

// Called by a warp that has finished (stored and pushed) one 32-column block of a cb200_mpgemm_push_signal GEMV: system-scope
// fence, count the block, and -- as the last block of the grid -- raise the flags.  `done` is a self-resetting counter.
__device__ __forceinline__ void push_signal_block_done(const Problem& p, unsigned int* done, unsigned int total_blocks, int lane) {
  if (p.local_flag == nullptr && p.wait_expected == nullptr) return;
  if (p.local_flag != nullptr) __threadfence_system();  // this warp's stores to the peers are performed system-wide before the count
  __syncwarp();
  if (lane == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == total_blocks - 1u) {
      *done = 0u;
      // cb200_mpgemm_chain: every block of the grid has finished, hence read *wait_expected: count this wait
      if (p.wait_expected != nullptr) *p.wait_expected = *p.wait_expected + 1u;
      if (p.local_flag != nullptr) {
        __threadfence_system();
        for (int i = 0; i < p.n_peers; ++i) atomicAdd_system(p.peer_flag[i], 1u);
        atomicAdd_system(p.local_flag, 1u);
      }
    }
  }
}

void set_error(const char* fmt, ...);

#define CB200_CUDA_CHECK(expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::cb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return static_cast<int>(_e);                                                    \
    }                                                                                 \
  } while (0)

// ---- launchers implemented in the per-kernel translation units ----------------------------------------
int launch_dequant(const Problem& p, int bits, int act_dtype, cudaStream_t stream);
int launch_decode(const Problem& p, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream,
                  int* launches);
size_t decode_workspace_bytes(int64_t M, int64_t K, int64_t N, int bits);
int launch_tc(const Problem& p, int bits, int act_dtype, int variant, void* ws, size_t ws_bytes,
              cudaStream_t stream, int* launches);
int launch_tc2(const Problem& p, int bits, int act_dtype, cudaStream_t stream, int* launches);
size_t tc_workspace_bytes(int64_t M, int64_t K, int64_t N);
bool tc_supported(const Problem& p, int bits);

int launch_stream(const Problem& p, int bits, int act_dtype, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches);
size_t stream_workspace_bytes(int64_t M, int64_t K, int64_t N, int bits);
bool stream_supported(const Problem& p, int bits, int act_dtype);
int launch_scaled(const void* a, const void* b_t, const float* scale_a, const float* scale_b, const void* bias, void* out, int M,
                  int K, int N, int in_kind, int out_dtype, int sa_scalar, int sb_scalar, cudaStream_t stream);
int launch_static_quant(const void* x, const float* scale, void* out, int64_t n, int in_dtype, int out_kind, cudaStream_t stream);
int launch_wait_flags(const uint32_t* flags, int count, uint32_t* expected, cudaStream_t stream);
int launch_push_bulk(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* multicast_dst, int max_ctas,
                     cudaStream_t stream);
int launch_push(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* const* peer_flags, void* local_flag,
                cudaStream_t stream);
struct StreamTuning {
  int ring_kb;   // shared memory the kernel may use per CTA, in KiB (0 = all of it)
  int ctas;      // upper bound on the CTAs launched (0 = one per SM)
  int dbg;       // bit 0: no L2 cache policy on the weight tiles (A/B timing)
};
extern StreamTuning g_stream_tuning;
extern long long* g_stream_trace;  // device buffer the next stream_kernel launches stamp (nullptr = off)

struct DecodeTuning {
  int warps_n, warps_k, split_k, vec_words;
};
extern DecodeTuning g_decode_tuning;
extern int g_decode_pdl;
struct TcTuning {
  int m_tile, split_k;
};
extern TcTuning g_tc_tuning;
extern int g_tc_cluster;

// ---- device helpers -----------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void store_out(void* out, int out_dtype, size_t idx, float v) {
  if (out_dtype == CB200_F16) {
    reinterpret_cast<__half*>(out)[idx] = __float2half_rn(v);
  } else if (out_dtype == CB200_BF16) {
    reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(v);
  } else {
    reinterpret_cast<float*>(out)[idx] = v;
  }
}

// out[idx] = v, and the same element of every peer buffer (multi-GPU epilogue: P2P stores over NVLink)
template <class T>
__device__ __forceinline__ void store_out_all(const Problem& p, size_t idx, float v) {
  store_out<T>(p.out, p.out_dtype, idx, v);
  for (int i = 0; i < p.n_peers; ++i) store_out<T>(p.peer_out[i], p.out_dtype, idx, v);
}

// four consecutive outputs (idx % 4 == 0) as ONE vector store per destination: 8 bytes for fp16 / bf16, 16 for fp32 --
// the shape the multi-GPU push wants (one NVLink write per peer instead of four 2-byte ones)
__device__ __forceinline__ void store_out4(void* out, int out_dtype, size_t idx, float4 v) {
  if (out_dtype == CB200_F16) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + idx) = pk;
  } else if (out_dtype == CB200_BF16) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + idx) = pk;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + idx) = v;
  }
}
__device__ __forceinline__ void store_out4_all(const Problem& p, size_t idx, float4 v) {
  store_out4(p.out, p.out_dtype, idx, v);
  for (int i = 0; i < p.n_peers; ++i) store_out4(p.peer_out[i], p.out_dtype, idx, v);
}

// streaming (read-once) global loads: bypass L1 allocation
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
               : "=r"(r.x), "=r"(r.y)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace cb200
