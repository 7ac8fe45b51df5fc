// Static per-tensor scaled quantisation of activations to int8 / fp8 e4m3: the producers of scaled_gemm's `a` operand.
// SURVEY.md section 8(f) row 4.
//
// Replaces: _static_scaled_int8_quant_kernel / _static_scaled_fp8_quant_kernel
//           (reference conch/kernels/quantization/int8.py:12-58, fp8.py:14-64; oracles conch/reference/quantization/int8.py:12-18,
//           fp8.py:12-18).
//
// Arithmetic, operation for operation: q = x * (1 / scale) in fp32 (the reciprocal is formed once, then a multiply -- not a
// division per element), clamp to the type's finite range, then
//   int8: conversion with truncation toward zero (what torch's and Triton's float -> int casts do),
//   fp8 : round-to-nearest-even to e4m3fn.
// HBM-bound elementwise work: the scale is per TENSOR, so the [tokens, hidden] shape does not matter and the kernel walks the
// flat array with 16-byte loads (8 halfs / 4 floats per thread and step), a grid of 8 CTAs per SM, scalar tail.
#include <cuda_fp8.h>

#include "common.cuh"
#include "host_util.cuh"

namespace cb200 {

template <int OUT_KIND>
__device__ __forceinline__ uint8_t quant_one(float x, float inv) {
  const float v = x * inv;
  if constexpr (OUT_KIND == CB200_IN_INT8) {
    const float c = fminf(fmaxf(v, -128.f), 127.f);
    return uint8_t(int8_t(__float2int_rz(c)));  // (NaN -> 0, like the reference's cast on CUDA)
  } else {
    const float c = fminf(fmaxf(v, -448.f), 448.f);
    return uint8_t(__nv_cvt_float_to_fp8(c, __NV_SATFINITE, __NV_E4M3));
  }
}

template <class T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }

template <class T, int OUT_KIND>
__global__ void __launch_bounds__(256) static_quant_kernel(const T* __restrict__ x, const float* __restrict__ scale, uint8_t* __restrict__ out,
                                                           int64_t n) {
  constexpr int VEC = 16 / int(sizeof(T));  // elements per 16-byte load
  const float inv = 1.0f / __ldg(scale);
  const int64_t nvec = n / VEC;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(x) + i);  // read once: streaming load
    const T* e = reinterpret_cast<const T*>(&raw);
    uint8_t q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) q[j] = quant_one<OUT_KIND>(to_f32<T>(e[j]), inv);
    if constexpr (VEC == 8) {
      *reinterpret_cast<uint2*>(out + i * VEC) = *reinterpret_cast<const uint2*>(q);
    } else {
      *reinterpret_cast<uint32_t*>(out + i * VEC) = *reinterpret_cast<const uint32_t*>(q);
    }
  }
  for (int64_t i = nvec * VEC + blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride)
    out[i] = quant_one<OUT_KIND>(to_f32<T>(x[i]), inv);
}

int launch_static_quant(const void* x, const float* scale, void* out, int64_t n, int in_dtype, int out_kind, cudaStream_t stream) {
  const DeviceInfo* di = current_device_info();
  const int sms = di && di->sms > 0 ? di->sms : 148;
  const int64_t per_block = 256 * 8;
  int64_t blocks = (n + per_block - 1) / per_block;
  if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
  if (blocks < 1) blocks = 1;
  uint8_t* o = reinterpret_cast<uint8_t*>(out);
#define CB200_QUANT_CASE(T)                                                                                                       \
  if (out_kind == CB200_IN_INT8) static_quant_kernel<T, CB200_IN_INT8><<<int(blocks), 256, 0, stream>>>(reinterpret_cast<const T*>(x), scale, o, n); \
  else static_quant_kernel<T, CB200_IN_FP8_E4M3><<<int(blocks), 256, 0, stream>>>(reinterpret_cast<const T*>(x), scale, o, n)
  if (in_dtype == CB200_F16) {
    CB200_QUANT_CASE(__half);
  } else if (in_dtype == CB200_BF16) {
    CB200_QUANT_CASE(__nv_bfloat16);
  } else {
    CB200_QUANT_CASE(float);
  }
#undef CB200_QUANT_CASE
  CB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cb200
