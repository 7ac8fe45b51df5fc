// In-register unpack + dequantise of GemLite-style packed int4/int8 words.  Shared by the standalone
// dequant kernel, the decode (HMMA) kernel and the tcgen05 kernels, so all three produce the SAME bits.
//
// Semantics restated from the reference (conch/kernels/quantization/gemm.py:192-216):
//     v = float_act(q - weight_bias)            exact small integer
//     W = RN_act( (v - zp) * s )                one rounding (the multiply) when zp is an integer
// A fused fma(q, s, -zp*s) would round differently, so the integer difference is formed exactly first
// (magic-number tricks below), then ONE multiply in the activation dtype.
//
// A packed word holds pf = 32/bits consecutive k of ONE column (conch/third_party/vllm/quant_utils.py:119-124).
// dequant_word() turns one word into pf/2 registers, each a pair of activation-dtype values:
//     Order::Natural : reg i = (k = 2i, 2i+1)                      -- what an MMA K-major operand wants
//     Order::Lop     : int4: reg i = (k = i, i+4), i = 0..3        -- cheapest extraction (no byte permutes);
//                      usable wherever the OTHER operand can be permuted the same way (decode path)
//                      int8: identical to Natural
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <type_traits>

namespace cb200 {

enum class Order { Natural = 0, Lop = 1 };

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & mask) | orv
  return r;
}
// (the intrinsic, not inline asm with an "r" selector: with a compile-time selector ptxas then encodes it as an immediate
//  instead of moving it into a register before every PRMT -- 32 wasted instructions per 64-deep K block and thread)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// Magic-number bytes for PRMT, deliberately NOT compile-time constants: SASS PRMT takes either the selector or the second
// source as an immediate, not both.  With a literal second source ptxas keeps the (compile-time) selectors in registers and
// re-materialises one before EVERY PRMT -- 32 wasted moves per 64-deep K block and thread in the tcgen05 dequant loop.
// Read from the constant bank, the magic is a plain operand and the selector becomes the immediate.
static __constant__ uint32_t kMagic64 = 0x64646464u;  // fp16: 0x64 next to a payload byte = 1024 + payload
static __constant__ uint32_t kMagic4B = 0x4B000000u;  // fp32: 2^23 + payload

__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ uint32_t b2_bits(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __nv_bfloat162 bits_b2(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }

// ---- per-(column, group) constants -----------------------------------------------------------------
// s   : scale, broadcast to both halves (fp32 bits for the bf16/int8 route)
// c0/c1: pre-combined magic + bias + zero-point terms, meaning depends on <A,BITS,ORDER> (see dequant_word)
// zf  : zero-point pair in the activation dtype, only for the float-zero-point route (FZP)
struct ColQ {
  uint32_t s, c0, c1, zf;
};

template <class A>
struct ActOps;

template <>
struct ActOps<__half> {
  static __device__ __forceinline__ uint32_t pair(float lo, float hi) {
    return h2_bits(__floats2half2_rn(lo, hi));
  }
  static __device__ __forceinline__ uint32_t bcast(__half v) { return h2_bits(__half2half2(v)); }
  static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return h2_bits(__hsub2(bits_h2(a), bits_h2(b))); }
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) { return h2_bits(__hmul2(bits_h2(a), bits_h2(b))); }
  static __device__ __forceinline__ uint32_t fma(uint32_t a, uint32_t b, uint32_t c) {
    return h2_bits(__hfma2(bits_h2(a), bits_h2(b), bits_h2(c)));
  }
  static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
};

template <>
struct ActOps<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t pair(float lo, float hi) {
    return b2_bits(__floats2bfloat162_rn(lo, hi));
  }
  static __device__ __forceinline__ uint32_t bcast(__nv_bfloat16 v) { return b2_bits(__bfloat162bfloat162(v)); }
  static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return b2_bits(__hsub2(bits_b2(a), bits_b2(b))); }
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) { return b2_bits(__hmul2(bits_b2(a), bits_b2(b))); }
  static __device__ __forceinline__ uint32_t fma(uint32_t a, uint32_t b, uint32_t c) {
    return b2_bits(__hfma2(bits_b2(a), bits_b2(b), bits_b2(c)));
  }
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
};

// zb = weight_bias + integer zero-point (0 when there is none); zpf = float zero-point (FZP route only).
template <class A, int BITS, Order ORD, bool FZP>
__device__ __forceinline__ ColQ make_colq(A scale, float zb, A zpf) {
  ColQ c;
  c.zf = FZP ? ActOps<A>::bcast(zpf) : 0u;
  if constexpr (std::is_same<A, __half>::value) {
    c.s = ActOps<A>::bcast(scale);
    if constexpr (BITS == 4 && ORD == Order::Lop) {
      c.c0 = ActOps<A>::pair(1024.f + zb, 1024.f + zb);  // subtracted from 1024+q
      c.c1 = ActOps<A>::pair(-(64.f + zb), -(64.f + zb));  // added to (1024+16q)/16
    } else if constexpr (BITS == 4) {
      c.c0 = ActOps<A>::pair(-(1024.f + zb), -(64.f + zb));  // fma addend for (1024+q_even, 1024+16 q_odd)*(1,1/16)
      c.c1 = 0u;
    } else {
      c.c0 = ActOps<A>::pair(1024.f + zb, 1024.f + zb);
      c.c1 = 0u;
    }
  } else {
    if constexpr (BITS == 4) {
      c.s = ActOps<A>::bcast(scale);
      c.c0 = ActOps<A>::pair(128.f + zb, 128.f + zb);  // subtracted from 128+q   (bf16: 8-bit significand)
      c.c1 = 0u;
    } else {
      // int8 x bf16: 128+q does not fit a bf16 significand -> go through fp32 (exact), round once at the end.
      c.s = FZP ? ActOps<A>::bcast(scale) : __float_as_uint(ActOps<A>::to_float(scale));
      c.c0 = __float_as_uint(8388608.f + zb);
      c.c1 = 0u;
    }
  }
  return c;
}

// One packed word -> pf/2 pair registers.  SCALE=true: dequantised weights RN((q - zb) * s) (the reference's W);
// SCALE=false: the exact integers (q - zb) in the activation dtype (the caller applies the scale elsewhere).
template <class A, int BITS, Order ORD, bool FZP, bool SCALE>
__device__ __forceinline__ void dequant_word_core(uint32_t w, const ColQ& c, uint32_t (&out)[16 / BITS]) {
  struct Ops : ActOps<A> {
    static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) { return SCALE ? ActOps<A>::mul(a, b) : a; }
  };
  if constexpr (std::is_same<A, __half>::value && BITS == 4 && ORD == Order::Lop) {
    // fp16 magic: 0x6400 | q = 1024 + q ; 0x6400 | (q << 4) = 1024 + 16 q   (both exact in fp16)
    const uint32_t t = w >> 8;
    const uint32_t p0 = lop3_and_or(w, 0x000f000fu, 0x64006400u);  // (1024+k0   , 1024+k4)
    const uint32_t p1 = lop3_and_or(w, 0x00f000f0u, 0x64006400u);  // (1024+16k1 , 1024+16k5)
    const uint32_t p2 = lop3_and_or(t, 0x000f000fu, 0x64006400u);  // (k2, k6)
    const uint32_t p3 = lop3_and_or(t, 0x00f000f0u, 0x64006400u);  // (16k3, 16k7)
    const uint32_t sixteenth = 0x2C002C00u;                        // (1/16, 1/16)
    uint32_t v0 = Ops::sub(p0, c.c0);
    uint32_t v1 = Ops::fma(p1, sixteenth, c.c1);
    uint32_t v2 = Ops::sub(p2, c.c0);
    uint32_t v3 = Ops::fma(p3, sixteenth, c.c1);
    if constexpr (FZP) {
      v0 = Ops::sub(v0, c.zf); v1 = Ops::sub(v1, c.zf); v2 = Ops::sub(v2, c.zf); v3 = Ops::sub(v3, c.zf);
    }
    out[0] = Ops::mul(v0, c.s); out[1] = Ops::mul(v1, c.s); out[2] = Ops::mul(v2, c.s); out[3] = Ops::mul(v3, c.s);
  } else if constexpr (std::is_same<A, __half>::value && BITS == 4) {
    // byte i of the word = (k=2i low nibble, k=2i+1 high nibble): replicate it into both halves next to 0x64,
    // keep the low nibble in the low half and the high nibble in the high half: (1024+q_2i, 1024+16 q_2i+1).
    const uint32_t one_sixteenth = 0x2C003C00u;  // (1.0, 1/16)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t b = prmt(w, kMagic64, 0x4040u + 0x0101u * i);
      const uint32_t p = b & 0x64f0640fu;
      uint32_t v = Ops::fma(p, one_sixteenth, c.c0);
      if constexpr (FZP) v = Ops::sub(v, c.zf);
      out[i] = Ops::mul(v, c.s);
    }
  } else if constexpr (std::is_same<A, __half>::value && BITS == 8) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t p = prmt(w, kMagic64, i == 0 ? 0x4140u : 0x4342u);  // (1024+k_2i, 1024+k_2i+1)
      uint32_t v = Ops::sub(p, c.c0);
      if constexpr (FZP) v = Ops::sub(v, c.zf);
      out[i] = Ops::mul(v, c.s);
    }
  } else if constexpr (BITS == 4) {
    // bf16 magic: 0x4300 | q = 128 + q (q < 128).  Only bits 0..6 of each half are significand, so every
    // nibble has to be shifted down to bit 0 first.
    if constexpr (ORD == Order::Lop) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t p = lop3_and_or(w >> (4 * i), 0x000f000fu, 0x43004300u);  // (128+k_i, 128+k_i+4)
        uint32_t v = Ops::sub(p, c.c0);
        if constexpr (FZP) v = Ops::sub(v, c.zf);
        out[i] = Ops::mul(v, c.s);
      }
    } else {
      const uint32_t t = w >> 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t b = prmt(w, t, 0x0400u + 0x0101u * i);            // byte0 = w.byte_i, byte2 = (w>>4).byte_i
        const uint32_t p = lop3_and_or(b, 0x000f000fu, 0x43004300u);     // (128+k_2i, 128+k_2i+1)
        uint32_t v = Ops::sub(p, c.c0);
        if constexpr (FZP) v = Ops::sub(v, c.zf);
        out[i] = Ops::mul(v, c.s);
      }
    }
  } else {
    // int8 x bf16: 0x4B000000 | q = 2^23 + q in fp32; (q - zb) and (q - zb)*s are exact in fp32
    // (9-bit integer x 8-bit significand), cvt.rn packs with the single rounding the reference performs.
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t b = prmt(w, kMagic4B, 0x7540u + i);
      f[i] = __uint_as_float(b) - __uint_as_float(c.c0);
    }
    if constexpr (FZP) {
      // float zero-point: the reference subtracts in bf16 (one extra rounding) before the multiply
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint32_t v = Ops::pair(f[2 * i], f[2 * i + 1]);
        v = Ops::sub(v, c.zf);
        out[i] = Ops::mul(v, c.s);
      }
    } else {
      const float s = SCALE ? __uint_as_float(c.s) : 1.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) out[i] = Ops::pair(f[2 * i] * s, f[2 * i + 1] * s);
    }
  }
}

template <class A, int BITS, Order ORD, bool FZP>
__device__ __forceinline__ void dequant_word(uint32_t w, const ColQ& c, uint32_t (&out)[16 / BITS]) {
  dequant_word_core<A, BITS, ORD, FZP, true>(w, c, out);
}
template <class A, int BITS, Order ORD, bool FZP>
__device__ __forceinline__ void dequant_word_int(uint32_t w, const ColQ& c, uint32_t (&out)[16 / BITS]) {
  dequant_word_core<A, BITS, ORD, FZP, false>(w, c, out);
}

// ---- zero-point / scale fetch helpers -----------------------------------------------------------------
// zp_kind values follow cb200_zp_kind in include/conch_b200.h
template <class A>
__device__ __forceinline__ float load_zp_int(const void* zp, int zp_kind, size_t idx) {
  if (zp_kind == 1) return static_cast<float>(__ldg(reinterpret_cast<const int32_t*>(zp) + idx));
  if (zp_kind == 2) return static_cast<float>(__ldg(reinterpret_cast<const uint8_t*>(zp) + idx));
  return 0.f;
}

}  // namespace cb200
