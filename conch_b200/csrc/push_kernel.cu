// cb200_push: a finished [M, N/G] slice goes to every peer's gather buffer from a SIDE stream -- the decoupled form of the fused
// gather.  The GEMV chain of a token then never contains a kernel with outstanding NVLink stores (a grid that wrote to peers is
// not complete, and its programmatic dependents not released, until those stores are acknowledged), while per-GEMM visibility is
// kept through the same completion counters cb200_mpgemm_push_signal raises.
//
// No shared memory, 32 registers: the CTAs of this kernel become resident NEXT TO the one-CTA-per-SM decode kernel of the main
// stream.  HBM-free on the sending side (the slice was just written: L2 hits); bound by the NVLink round trip.
#include "common.cuh"
#include "host_util.cuh"

namespace cb200 {

struct PushArgs {
  void* dst[CB200_MAX_PEERS];
  uint32_t* flag[CB200_MAX_PEERS];
  uint32_t* local_flag;
  int n_peers;
};

__global__ void __launch_bounds__(512) push_kernel(const uint4* __restrict__ src, size_t n16, const PushArgs a) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  constexpr int U = 4;  // chunks in flight per thread
  for (size_t base = size_t(blockIdx.x) * blockDim.x + threadIdx.x; base < n16; base += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + u * stride;
      if (i < n16) v[u] = __ldcg(src + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + u * stride;
      if (i < n16) {
        for (int p = 0; p < a.n_peers; ++p) reinterpret_cast<uint4*>(a.dst[p])[i] = v[u];
      }
    }
  }
  if (a.local_flag != nullptr) {  // (single CTA: see launch_push)
    __threadfence_system();       // this thread's stores are performed at the peers before the block counts itself done
    __syncthreads();
    if (int(threadIdx.x) < a.n_peers) atomicAdd_system(a.flag[threadIdx.x], 1u);
    if (int(threadIdx.x) == a.n_peers) atomicAdd_system(a.local_flag, 1u);
  }
}

// The large-M leg (cb200_push_bulk): a few CTAs on a HIGH-PRIORITY side stream move a whole [M, N/G] slice while the next GEMM owns
// the rest of the SMs.  With a multicast address (NVLS: one store, the switch replicates it into every GPU's buffer) the sender's
// egress is the slice once instead of once per peer.
__global__ void __launch_bounds__(512) push_bulk_kernel(const uint4* __restrict__ src, size_t n16, const PushArgs a, uint4* mc) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  constexpr int U = 8;  // 128 bytes in flight per thread
  for (size_t base = size_t(blockIdx.x) * blockDim.x + threadIdx.x; base < n16; base += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + u * stride;
      if (i < n16) v[u] = __ldcs(src + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + u * stride;
      if (i < n16) {
        if (mc != nullptr) {
          asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i), "f"(__uint_as_float(v[u].x)),
                       "f"(__uint_as_float(v[u].y)), "f"(__uint_as_float(v[u].z)), "f"(__uint_as_float(v[u].w))
                       : "memory");
        } else {
          for (int p = 0; p < a.n_peers; ++p) reinterpret_cast<uint4*>(a.dst[p])[i] = v[u];
        }
      }
    }
  }
}

int launch_push_bulk(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* multicast_dst, int max_ctas,
                     cudaStream_t stream) {
  PushArgs a{};
  a.n_peers = multicast_dst ? 0 : num_peers;
  for (int i = 0; i < a.n_peers; ++i) a.dst[i] = peer_dst[i];
  const size_t n16 = bytes / 16;
  const DeviceInfo* di = current_device_info();
  size_t cap = di ? size_t(di->sms) : 148;
  if (max_ctas > 0 && size_t(max_ctas) < cap) cap = size_t(max_ctas);
  const size_t want = (n16 + 8 * 512 - 1) / (8 * 512);
  const int ctas = int(want < 1 ? 1 : want > cap ? cap : want);
  push_bulk_kernel<<<ctas, 512, 0, stream>>>(reinterpret_cast<const uint4*>(src), n16, a, reinterpret_cast<uint4*>(multicast_dst));
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_push(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* const* peer_flags, void* local_flag,
                cudaStream_t stream) {
  PushArgs a{};
  a.n_peers = num_peers;
  for (int i = 0; i < num_peers; ++i) {
    a.dst[i] = peer_dst[i];
    a.flag[i] = local_flag ? reinterpret_cast<uint32_t*>(peer_flags[i]) : nullptr;
  }
  a.local_flag = reinterpret_cast<uint32_t*>(local_flag);
  const size_t n16 = bytes / 16;
  int ctas = 1;
  if (local_flag == nullptr) {  // no completion counters: as many CTAs as the bytes are worth, one per SM at most
    const DeviceInfo* di = current_device_info();
    const size_t want = (n16 + 4 * 512 - 1) / (4 * 512);
    const size_t cap = di ? size_t(di->sms) : 148;
    ctas = int(want < 1 ? 1 : want > cap ? cap : want);
  }
  push_kernel<<<ctas, 512, 0, stream>>>(reinterpret_cast<const uint4*>(src), n16, a);
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cb200
