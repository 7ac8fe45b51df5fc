// Micro-benchmark: does a 256-bit load per lane get more bytes through the load/store unit than a 128-bit one?
// Decode-kernel access shape: a warp owns a column block (lane = (row-in-step tig, column chunk g8)), walks K in steps of
// 4 packed rows with a rotating window of WIN loads in flight.  [R, C] int32 = the packed gate_up weights by default.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_width load_width.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int W>
struct Vec { unsigned v[W]; };

template <int W>
__device__ __forceinline__ Vec<W> ldg(const void* p);
template <>
__device__ __forceinline__ Vec<4> ldg<4>(const void* p) {
  Vec<4> r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p));
  return r;
}
template <>
__device__ __forceinline__ Vec<8> ldg<8>(const void* p) {
  Vec<8> r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

// W words per lane per load; warp covers 8 * W columns x 4 rows per step; WIN steps in flight; S slabs along K
template <int W, int WIN>
__global__ void __launch_bounds__(128) stream(const unsigned* __restrict__ w, int R, int C, int S, int total_warps, unsigned* out) {
  const int lane = threadIdx.x & 31, g8 = lane >> 2, tig = lane & 3;
  const int wi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (wi >= total_warps) return;
  const int CB = C / (8 * W);
  const int si = wi / CB, cb = wi - si * CB;
  const int steps = R / 4 / S;
  const char* p = reinterpret_cast<const char*>(w + (size_t(si) * steps * 4 + tig) * C + cb * 8 * W + W * g8);
  const size_t step_bytes = size_t(16) * C;
  Vec<W> win[WIN];
#pragma unroll
  for (int j = 0; j < WIN; ++j) if (j < steps) win[j] = ldg<W>(p + j * step_bytes);
  unsigned acc = 0;
  for (int s0 = 0; s0 < steps; s0 += WIN) {
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
      if (s0 + j < steps) {
        Vec<W> v = win[j];
        if (s0 + j + WIN < steps) win[j] = ldg<W>(p + size_t(s0 + j + WIN) * step_bytes);
#pragma unroll
        for (int i = 0; i < W; ++i) acc += v.v[i] * 0x9E3779B1u ^ (acc >> 7);
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

__global__ void fill_random(unsigned* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned v = unsigned(i) * 2654435761u + seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; v *= 3266489917u; v ^= v >> 16;
    p[i] = v;
  }
}

int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 512, C = argc > 2 ? atoi(argv[2]) : 28672;
  const size_t bytes = size_t(R) * C * 4;
  const int copies = int((400ull << 20) / bytes) + 2;
  std::vector<unsigned*> bufs(copies);
  { unsigned sd = 1; for (auto& b : bufs) { cudaMalloc(&b, bytes); fill_random<<<1184, 256>>>(b, bytes / 4, sd++); } cudaDeviceSynchronize(); }
  unsigned* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](auto launch, const char* name) {
    for (int i = 0; i < copies; ++i) launch(bufs[i]);
    cudaDeviceSynchronize();
    const int reps = 5;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) for (int i = 0; i < copies; ++i) launch(bufs[i]);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / (reps * copies);
    printf("%-64s %8.2f us  %8.1f GB/s  (%s)\n", name, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  printf("matrix [%d x %d] int32 = %.1f MB, %d copies\n", R, C, bytes / 1e6, copies);
  for (int S : {1, 2, 4, 8}) {
    char name[160];
    { const int warps = C / 32 * S; snprintf(name, sizeof name, "128-bit loads, window 8, S=%d: %5d warps (%.1f per SM)", S, warps, warps / 148.0);
      timeit([&](unsigned* w) { stream<4, 8><<<(warps + 3) / 4, 128>>>(w, R, C, S, warps, out); }, name); }
    { const int warps = C / 64 * S; snprintf(name, sizeof name, "256-bit loads, window 8, S=%d: %5d warps (%.1f per SM)", S, warps, warps / 148.0);
      timeit([&](unsigned* w) { stream<8, 8><<<(warps + 3) / 4, 128>>>(w, R, C, S, warps, out); }, name); }
    { const int warps = C / 64 * S; snprintf(name, sizeof name, "256-bit loads, window 4, S=%d: %5d warps (%.1f per SM)", S, warps, warps / 148.0);
      timeit([&](unsigned* w) { stream<8, 4><<<(warps + 3) / 4, 128>>>(w, R, C, S, warps, out); }, name); }
  }
  return 0;
}
