// mma.sync.m16n8k16 (f16 x f16 -> f32) on sm_100a: latency of a dependent accumulator chain and issue rate of independent
// chains, with normal and with SUBNORMAL fp16 operands (the streaming decode kernel feeds nibbles as subnormals).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/micro/hmma_latency benchmarks/micro/hmma_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void kern(uint32_t aval, uint32_t bval, int iters, long long* out, float* sink) {
  float acc[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c)
    for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  uint32_t a = aval + (threadIdx.x & 1), b = bval;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(acc[c], a, a, a, a, b, b);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c)
    for (int i = 0; i < 4; ++i) s += acc[c][i];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void run(const char* what, uint32_t aval, uint32_t bval, int warps) {
  long long* d;
  float* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2048;
  kern<CHAINS><<<148, 32 * warps>>>(aval, bval, iters, d, sink);
  kern<CHAINS><<<148, 32 * warps>>>(aval, bval, iters, d, sink);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = double(h[0]) / iters;
  printf("%-28s chains=%d warps/SM=%2d: %7.1f cycles per loop, %6.2f cycles per HMMA per warp, %6.2f cycles per HMMA per SM scheduler\n", what,
         CHAINS, warps, cyc, cyc / CHAINS, cyc / CHAINS / ((warps + 3) / 4));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  const uint32_t normal = 0x3C003C00u, sub = 0x000f000fu, sub16 = 0x00f000f0u, x = 0x3555B8CDu;
  run<1>("normal x normal", normal, normal, 1);
  run<1>("subnormal A x normal B", sub, x, 1);
  run<1>("subnormal(16q) A x normal B", sub16, x, 1);
  run<2>("subnormal A", sub, x, 1);
  run<4>("subnormal A", sub, x, 1);
  run<8>("subnormal A", sub, x, 1);
  run<8>("normal A", normal, x, 1);
  run<4>("subnormal A", sub, x, 4);
  run<4>("subnormal A", sub, x, 8);
  run<4>("subnormal A", sub, x, 16);
  run<4>("normal A", normal, x, 16);
  run<1>("subnormal A", sub, x, 8);
  run<1>("subnormal A", sub, x, 16);
  return 0;
}
