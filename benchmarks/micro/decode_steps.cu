// Incremental micro-benchmark: start from a pure streaming read with the decode kernel's exact access shape
// (warp = 32 columns x 16 packed rows per batch, 128-bit loads, 4 rows x 128 B per instruction) and add the decode
// kernel's ingredients one at a time, to find which one costs the time.
//   LOADER 0: registers, double-buffered (next batch loaded while this one is consumed)
//   LOADER 1: cp.async.cg ring, LOOK batches ahead, read back with ld.shared.v4
//   COMPUTE 0: xor-reduce   1: lop3 unpack (fp16 magic)   2: lop3 + 6 HMMA per step (fused decode inner loop)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <vector>

#include "../../conch_b200/csrc/dequant.cuh"

__device__ __forceinline__ uint4 ldg_na(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t m, uint32_t o) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(m), "r"(o));
  return r;
}
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int COMPUTE>
__device__ __forceinline__ void consume(const uint4 (&w)[4], float (&acc)[6][4], unsigned& x) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t words[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
    if constexpr (COMPUTE == 0) {
      x ^= words[0] ^ words[1] ^ words[2] ^ words[3];
    } else {
      uint32_t d[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t t8 = words[j] >> 8;
        d[j][0] = lop(words[j], 0x000f000fu, 0x64006400u);
        d[j][1] = lop(words[j], 0x00f000f0u, 0x64006400u);
        d[j][2] = lop(t8, 0x000f000fu, 0x64006400u);
        d[j][3] = lop(t8, 0x00f000f0u, 0x64006400u);
      }
      if constexpr (COMPUTE == 3) {
        cb200::ColQ cq;
        cq.s = 0x1C001C00u; cq.c0 = 0x64086408u; cq.c1 = 0xD480D480u; cq.zf = 0;
        uint32_t e[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cb200::dequant_word<__half, 4, cb200::Order::Lop, false>(words[j], cq, e[j]);
        const uint32_t xb0 = 0x3C003C00u + u, xb1 = 0x3C003C00u, xb2 = 0x3C003C01u, xb3 = 0x3C003C02u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mma(acc[t], e[2 * t][0], e[2 * t + 1][0], e[2 * t][1], e[2 * t + 1][1], xb0, xb1);
          mma(acc[t], e[2 * t][2], e[2 * t + 1][2], e[2 * t][3], e[2 * t + 1][3], xb2, xb3);
        }
      } else if constexpr (COMPUTE == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x ^= d[j][0] ^ d[j][1] ^ d[j][2] ^ d[j][3];
      } else {
        const uint32_t xb0 = 0x3C003C00u + u, xb1 = 0x3C003C00u, xb2 = 0x3C003C01u, xb3 = 0x3C003C02u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mma(acc[2 * t], d[2 * t][0], d[2 * t + 1][0], d[2 * t][2], d[2 * t + 1][2], xb0, xb2);
          mma(acc[2 * t + 1], d[2 * t][1], d[2 * t + 1][1], d[2 * t][3], d[2 * t + 1][3], xb1, xb3);
        }
        mma(acc[4], 0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u, xb0, xb2);
        mma(acc[5], 0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u, xb1, xb3);
      }
    }
  }
}

// one warp = column block cb (32 int32 columns), slab si of S: units [si*n, (si+1)*n), unit = 16 packed rows
template <int LOADER, int COMPUTE, int LOOK>
__global__ void __launch_bounds__(128) k(const uint32_t* __restrict__ wq, int N, int S, int nunits, int total_warps, float* out) {
  extern __shared__ __align__(16) unsigned char ring_raw[];
  const int lane = threadIdx.x & 31, g8 = lane >> 2, tig = lane & 3;
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= total_warps) return;
  const int cb = w / S, si = w - cb * S;
  const char* p = reinterpret_cast<const char*>(wq + (size_t(si) * nunits * 16 + tig) * N + cb * 32 + 4 * g8);
  const size_t step_bytes = size_t(16) * N, unit_bytes = 4 * step_bytes;
  float acc[6][4] = {};
  unsigned x = 0;
  if constexpr (LOADER == 0) {
    uint4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = ldg_na(p + u * step_bytes);
    p += unit_bytes;
    for (int i = 0; i < nunits; i += 2) {
      if (i + 1 < nunits) {
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = ldg_na(p + u * step_bytes);
        p += unit_bytes;
      }
      consume<COMPUTE>(a, acc, x);
      if (i + 1 >= nunits) break;
      if (i + 2 < nunits) {
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = ldg_na(p + u * step_bytes);
        p += unit_bytes;
      }
      consume<COMPUTE>(b, acc, x);
    }
  } else if constexpr (LOADER == 2) {
    const int nsteps = nunits * 4;
    uint4 win[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { win[j] = make_uint4(0, 0, 0, 0); if (j < nsteps) win[j] = ldg_na(p + j * step_bytes); }
    p += 8 * step_bytes;
    for (int i = 0; i < nunits; i += 2) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (i + h < nunits) {
          uint4 a[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            a[u] = win[4 * h + u];
            if ((i + h) * 4 + u + 8 < nsteps) win[4 * h + u] = ldg_na(p + (4 * h + u) * step_bytes);
          }
          consume<COMPUTE>(a, acc, x);
        }
      }
      p += 8 * step_bytes;
    }
  } else {
    constexpr int SLOTS = LOOK + 1;
    const uint32_t ring0 = uint32_t(__cvta_generic_to_shared(ring_raw)) + (threadIdx.x >> 5) * (SLOTS * 2048) + lane * 16;
    const uint32_t ring_end = ring0 + SLOTS * 2048;
    auto issue = [&](uint32_t dst) {
#pragma unroll
      for (int u = 0; u < 4; ++u) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * 512), "l"(p + u * step_bytes) : "memory");
      p += unit_bytes;
    };
#pragma unroll
    for (int j = 0; j < LOOK; ++j) {
      if (j < nunits) issue(ring0 + j * 2048);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t islot = ring0 + (LOOK % SLOTS) * 2048, cslot = ring0;
    for (int i = 0; i < nunits; ++i) {
      if (i + LOOK < nunits) issue(islot);
      asm volatile("cp.async.commit_group;" ::: "memory");
      islot += 2048; if (islot == ring_end) islot = ring0;
      asm volatile("cp.async.wait_group %0;" ::"n"(LOOK) : "memory");
      uint4 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[u].x), "=r"(a[u].y), "=r"(a[u].z), "=r"(a[u].w) : "r"(cslot + u * 512));
      cslot += 2048; if (cslot == ring_end) cslot = ring0;
      consume<COMPUTE>(a, acc, x);
    }
  }
  float s = float(x);
#pragma unroll
  for (int i = 0; i < 6; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 12345.678f || x == 0x9e3779b9u) out[w] = s;
}

__global__ void fill_random(unsigned* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned v = unsigned(i) * 2654435761u + seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; v *= 3266489917u; v ^= v >> 16;
    p[i] = v;
  }
}

template <bool FX, bool FM, bool FP>
__global__ void __launch_bounds__(128, 5) kreal(const uint32_t* __restrict__ wq, const __half* __restrict__ xg, const __half* __restrict__ sg,
                                             const int* __restrict__ zg, int N, int S, int nunits, int total_warps, float* part,
                                             unsigned* cnt, __half* out) {
  const int lane = threadIdx.x & 31, g8 = lane >> 2, tig = lane & 3;
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= total_warps) return;
  const int cb = w / S, si = w - cb * S;
  const int col = cb * 32 + 4 * g8;
  const int b0 = si * nunits;
  const char* p = reinterpret_cast<const char*>(wq + (size_t(b0) * 16 + tig) * N + col);
  const size_t step_bytes = size_t(16) * N;
  const int nsteps = nunits * 4;
  uint4 win[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { win[j] = make_uint4(0, 0, 0, 0); if (j < nsteps) win[j] = ldg_na(p + j * step_bytes); }
  p += 8 * step_bytes;
  const __half* sp = sg + size_t(b0) * N + col;
  const int* zp = zg + size_t(b0) * N + col;
  uint2 raw_s = make_uint2(0x1C001C00u, 0x1C001C00u);
  uint4 raw_z = make_uint4(8, 8, 8, 8);
  if (FM) { raw_s = __ldg(reinterpret_cast<const uint2*>(sp)); raw_z = __ldg(reinterpret_cast<const uint4*>(zp)); sp += N; zp += N; }
  const __half* xp = xg + tig * 8 + size_t(b0) * 128;
  uint4 xnext = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
  const bool m_ok = g8 < 1;
  if (FX && m_ok) xnext = __ldg(reinterpret_cast<const uint4*>(xp));
  xp += 32;
  float acc[2][4] = {};
  cb200::ColQ cq[4];
  for (int i = 0; i < nunits; ++i) {
    {
      __half s4[4]; *reinterpret_cast<uint2*>(s4) = raw_s;
      const int z[4] = {int(raw_z.x), int(raw_z.y), int(raw_z.z), int(raw_z.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) cq[j] = cb200::make_colq<__half, 4, cb200::Order::Lop, false>(s4[j], float(z[j]), __half(0.f));
      if (FM && i + 1 < nunits) { raw_s = __ldg(reinterpret_cast<const uint2*>(sp)); raw_z = __ldg(reinterpret_cast<const uint4*>(zp)); sp += N; zp += N; }
    }
    const int h = i & 1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint4 wv;
      if (h == 0) { wv = win[u]; if (i * 4 + u + 8 < nsteps) win[u] = ldg_na(p + u * step_bytes); }
      else        { wv = win[4 + u]; if (i * 4 + u + 8 < nsteps) win[4 + u] = ldg_na(p + (4 + u) * step_bytes); }
      uint32_t xb[4];
      xb[0] = __byte_perm(xnext.x, xnext.z, 0x5410); xb[1] = __byte_perm(xnext.x, xnext.z, 0x7632);
      xb[2] = __byte_perm(xnext.y, xnext.w, 0x5410); xb[3] = __byte_perm(xnext.y, xnext.w, 0x7632);
      if (FX && m_ok && i * 4 + u + 1 < nsteps) xnext = __ldg(reinterpret_cast<const uint4*>(xp));
      xp += 32;
      const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t e[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cb200::dequant_word<__half, 4, cb200::Order::Lop, false>(words[j], cq[j], e[j]);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mma(acc[t], e[2 * t][0], e[2 * t + 1][0], e[2 * t][1], e[2 * t + 1][1], xb[0], xb[1]);
        mma(acc[t], e[2 * t][2], e[2 * t + 1][2], e[2 * t][3], e[2 * t + 1][3], xb[2], xb[3]);
      }
    }
    if (h == 1) p += 8 * step_bytes;
  }
  if (!FP) {
    if (tig == 0) { out[col] = __float2half(acc[0][0]); out[col + 1] = __float2half(acc[0][2]); out[col + 2] = __float2half(acc[1][0]); out[col + 3] = __float2half(acc[1][2]); }
    return;
  }
  float* mine = part + (size_t(cb) * S + si) * 32;
  if (tig == 0) { __stcg(mine + 4 * g8, acc[0][0]); __stcg(mine + 4 * g8 + 1, acc[0][2]); __stcg(mine + 4 * g8 + 2, acc[1][0]); __stcg(mine + 4 * g8 + 3, acc[1][2]); }
  __threadfence();
  __syncwarp();
  unsigned prev = 0;
  if (lane == 0) prev = atomicAdd(cnt + cb, 1u);
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev != unsigned(S - 1)) return;
  __threadfence();
  float v = 0.f;
  for (int s2 = 0; s2 < S; ++s2) v += __ldcg(part + (size_t(cb) * S + s2) * 32 + lane);
  out[cb * 32 + lane] = __float2half(v);
  if (lane == 0) cnt[cb] = 0u;
}

int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 512, N = argc > 2 ? atoi(argv[2]) : 28672;
  const size_t bytes = size_t(R) * N * 4;
  const int copies = int((400ull << 20) / bytes) + 2;
  std::vector<uint32_t*> bufs(copies);
  { unsigned sd = 1; for (auto& b : bufs) { cudaMalloc(&b, bytes); fill_random<<<1184, 256>>>(reinterpret_cast<unsigned*>(b), bytes / 4, sd++); } cudaDeviceSynchronize(); }
  float* out; cudaMalloc(&out, 1 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int B = R / 16, col_blocks = N / 32;
  printf("packed matrix [%d x %d] int32 = %.1f MB, %d copies, %d column blocks x %d batches\n", R, N, bytes / 1e6, copies, col_blocks, B);
  auto run = [&](auto kern, const char* name, int S, int look, size_t smem_per_warp) {
    const int nunits = B / S, warps = col_blocks * S, blocks = (warps + 3) / 4;
    const size_t smem = 4 * smem_per_warp;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    for (int i = 0; i < copies; ++i) kern<<<blocks, 128, smem>>>(bufs[i], N, S, nunits, warps, out);
    cudaDeviceSynchronize();
    const int reps = 5;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) for (int i = 0; i < copies; ++i) kern<<<blocks, 128, smem>>>(bufs[i], N, S, nunits, warps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / (reps * copies);
    printf("%-34s S=%2d warps=%5d units/warp=%3d: %7.2f us %7.1f GB/s (%s)\n", name, S, warps, nunits, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  {
    __half *xg, *sg, *og; int* zg; float* part; unsigned* cnt;
    cudaMalloc(&xg, 65536); cudaMemset(xg, 0x3c, 65536);
    cudaMalloc(&sg, size_t(B) * N * 2); cudaMemset(sg, 0x1c, size_t(B) * N * 2);
    cudaMalloc(&zg, size_t(B) * N * 4); cudaMemset(zg, 0, size_t(B) * N * 4);
    cudaMalloc(&og, size_t(N) * 2); cudaMalloc(&part, size_t(N) * 64 * 4); cudaMalloc(&cnt, 65536); cudaMemset(cnt, 0, 65536);
    auto runreal = [&](auto kern, const char* name, int S) {
      const int nunits = B / S, warps = col_blocks * S, blocks = (warps + 3) / 4;
      for (int i = 0; i < copies; ++i) kern<<<blocks, 128>>>(bufs[i], xg, sg, zg, N, S, nunits, warps, part, cnt, og);
      cudaDeviceSynchronize();
      const int reps = 5;
      cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) for (int i = 0; i < copies; ++i) kern<<<blocks, 128>>>(bufs[i], xg, sg, zg, N, S, nunits, warps, part, cnt, og);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double us = ms * 1e3 / (reps * copies);
      printf("%-34s S=%2d warps=%5d units/warp=%3d: %7.2f us %7.1f GB/s (%s)\n", name, S, warps, nunits, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    for (int S : {2, 4, 8, 16}) {
      if (B % S) continue;
      runreal(kreal<false, false, false>, "real: window+dequant+4HMMA", S);
      runreal(kreal<true, false, false>, "real: + x loads", S);
      runreal(kreal<false, true, false>, "real: + per-group meta", S);
      runreal(kreal<false, false, true>, "real: + split-K publish", S);
      runreal(kreal<true, true, true>, "real: all", S);
    }
  }
  for (int S : {4, 8}) {
    if (B % S) continue;
    run(k<0, 0, 1>, "regs dbuf, xor", S, 1, 0);
    run(k<0, 1, 1>, "regs dbuf, lop3", S, 1, 0);
    run(k<0, 2, 1>, "regs dbuf, lop3+6 HMMA", S, 1, 0);
    run(k<0, 3, 1>, "regs dbuf, dequant_word+4 HMMA", S, 1, 0);
    run(k<2, 0, 1>, "regs window8, xor", S, 1, 0);
    run(k<2, 2, 1>, "regs window8, lop3+6 HMMA", S, 1, 0);
    run(k<2, 3, 1>, "regs window8, dequant_word+4 HMMA", S, 1, 0);
    run(k<1, 0, 1>, "cp.async LOOK=1, xor", S, 1, 2 * 2048);
    run(k<1, 0, 2>, "cp.async LOOK=2, xor", S, 2, 3 * 2048);
    run(k<1, 0, 4>, "cp.async LOOK=4, xor", S, 4, 5 * 2048);
    run(k<1, 2, 2>, "cp.async LOOK=2, lop3+6 HMMA", S, 2, 3 * 2048);
    run(k<1, 2, 4>, "cp.async LOOK=4, lop3+6 HMMA", S, 4, 5 * 2048);
  }
  return 0;
}
