// Micro-benchmark: achievable HBM read bandwidth for the access patterns a W4 GEMV can use on a [R, C] int32 matrix.
//   pattern 0: flat streaming (each thread 16 B, consecutive, grid-stride)                  -- upper bound
//   pattern 1: warp owns `seg` bytes of a row-segment x `rows` consecutive rows (column-block major, like decode_kernel)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint4 ldg_na(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__global__ void flat(const uint4* __restrict__ w, size_t n16, unsigned* out) {
  unsigned acc = 0;
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 a = ldg_na(w + i), b = ldg_na(w + i + stride), c = ldg_na(w + i + 2 * stride), d = ldg_na(w + i + 3 * stride);
    acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < n16; i += stride) { uint4 a = ldg_na(w + i); acc += a.x ^ a.y ^ a.z ^ a.w; }
  if (acc == 0x12345678u) out[0] = acc;
}

// Each warp: column segment of seg16 x 16 bytes (seg16 lanes-worth; if seg16 < 32 the warp covers 32/seg16 rows per load),
// walks `rows_per_warp` rows.  Work units (colseg, rowchunk) flattened column-major over warps.
__global__ void tiled(const uint4* __restrict__ w, int R, int C16, int seg16, int rows_per_unit, int units_per_col, int total_units,
                      int units_per_warp, unsigned* out) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int rows_per_load = 32 / seg16;        // rows covered by one warp-wide load
  const int lr = lane / seg16, lc = lane % seg16;
  unsigned acc = 0;
  for (int uu = 0; uu < units_per_warp; ++uu) {
    const int unit = warp * units_per_warp + uu;
    if (unit >= total_units) break;
    const int cs = unit / units_per_col, rc = unit % units_per_col;
    const uint4* base = w + size_t(rc) * rows_per_unit * C16 + size_t(cs) * seg16 + lc;
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int row = j * rows_per_load + lr;
      if (row < rows_per_unit) v[j] = ldg_na(base + size_t(row) * C16); else v[j] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
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
  const int R = argc > 1 ? atoi(argv[1]) : 512, C = argc > 2 ? atoi(argv[2]) : 28672;  // gate_up packed: [K/8, N]
  const size_t bytes = size_t(R) * C * 4;
  const int copies = int((400ull << 20) / bytes) + 2;
  std::vector<uint4*> bufs(copies);
  { unsigned sd = 1; for (auto& b : bufs) { cudaMalloc(&b, bytes); fill_random<<<1184, 256>>>(reinterpret_cast<unsigned*>(b), bytes / 4, sd++); } cudaDeviceSynchronize(); }
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
    printf("%-44s %8.2f us  %8.1f GB/s  (%s)\n", name, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  printf("matrix [%d x %d] int32 = %.1f MB, %d copies\n", R, C, bytes / 1e6, copies);
  for (int blocks : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
    char name[128]; snprintf(name, sizeof name, "flat streaming, %d CTAs x 256", blocks);
    timeit([&](uint4* w) { flat<<<blocks, 256>>>(w, bytes / 16, out); }, name);
  }
  const int C16 = C / 4;
  for (int seg16 : {8, 16, 32}) {           // 128 B, 256 B, 512 B per row per warp
    for (int rows_per_unit : {16, 32, 64}) {
      const int rows_per_load = 32 / seg16;
      if (rows_per_unit > 8 * rows_per_load) continue;
      const int units_per_col = R / rows_per_unit;
      const int total_units = units_per_col * (C16 / seg16);
      for (int warps_target : {148 * 20, 148 * 32, 148 * 48}) {
        int upw = (total_units + warps_target - 1) / warps_target;
        const int warps = (total_units + upw - 1) / upw;
        char name[160];
        snprintf(name, sizeof name, "tiled seg=%4dB rows/unit=%2d units/warp=%2d warps=%5d", seg16 * 16, rows_per_unit, upw, warps);
        timeit([&](uint4* w) { tiled<<<(warps + 3) / 4, 128>>>(w, R, C16, seg16, rows_per_unit, units_per_col, total_units, upw, out); }, name);
      }
    }
  }
  return 0;
}
