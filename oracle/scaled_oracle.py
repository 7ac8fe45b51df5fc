"""CPU oracle of the scaled-GEMM sibling path and its activation quantisers.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module; the product
path (conch_b200/) never does.  Every function restates a reference function and cites it; tests/test_oracle_scaled.py pins
them bit for bit to fixtures produced by the real reference (oracle/make_golden_scaled.py, run in the build container).

  scaled_gemm_reference  conch/reference/quantization/scaled_gemm.py:12-27   (fp32 matmul, scale_a, scale_b, cast, + bias)
  scaled_gemm_op         conch/kernels/quantization/gemm.py:431-440 + conch/ops/quantization/gemm.py:249-250
                         (the Triton kernel's epilogue order: acc * (scale_a * scale_b), cast, then output.add_(bias))
  scaled_int8_quant      conch/reference/quantization/int8.py:12-18  ==  conch/kernels/quantization/int8.py:46-58
  scaled_fp8_quant       conch/reference/quantization/fp8.py:12-18
"""

from __future__ import annotations

import torch

FP8 = torch.float8_e4m3fn


def scaled_gemm_reference(a, b, scale_a, scale_b, out_dtype, bias=None):
    """The reference test's oracle (tests/scaled_gemm_test.py:75): fp32 matmul of the 8-bit operands, times scale_a, times
    scale_b^T, one cast, bias added in the output dtype."""
    out = torch.matmul(a.to(torch.float32), b.to(torch.float32))
    out = scale_a * out
    out = scale_b.T * out
    out = out.to(out_dtype)
    if bias is not None:
        out = out + bias
    return out


def scaled_gemm_op(a, b, scale_a, scale_b, out_dtype, bias=None):
    """What the reference OP computes (the kernel under test there): exact integer / fp32 accumulation, accumulator to fp32,
    multiplied by the PRODUCT scale_a[m] * scale_b[n], one rounding to the output dtype; bias added by a second kernel, i.e.
    to the rounded value, and rounded again.  For int8 the accumulation is exact, so this is what the kernels must reproduce
    bit for bit; for fp8 the tensor core's fp32 summation order is its own and only the tolerance applies."""
    if a.dtype == torch.int8:
        acc = torch.matmul(a.to(torch.float64), b.to(torch.float64)).to(torch.int64).to(torch.float32)  # exact, then RN to fp32
    else:
        acc = torch.matmul(a.to(torch.float32), b.to(torch.float32))
    sa = scale_a.to(torch.float32).reshape(-1, 1)
    sb = scale_b.to(torch.float32).reshape(1, -1)
    out = (acc * (sa * sb)).to(out_dtype)
    if bias is not None:
        out = out + bias.to(out_dtype)
    return out


def scaled_int8_quant(x, scale):
    """(x * 1/scale) in fp32, clamped to [-128, 127], converted with truncation toward zero (torch's and Triton's float -> int
    cast): conch/reference/quantization/int8.py:12-18."""
    inv = scale.to(torch.float32).reciprocal()
    return (x.to(torch.float32) * inv).clamp(min=-128, max=127).to(torch.int8)


def scaled_fp8_quant(x, scale):
    """(x * 1/scale) in fp32, clamped to +-448, rounded to nearest-even e4m3: conch/reference/quantization/fp8.py:12-18."""
    finfo = torch.finfo(FP8)
    inv = scale.to(torch.float32).reciprocal()
    return (x.to(torch.float32) * inv).clamp(min=finfo.min, max=finfo.max).to(FP8)


def make_scaled_problem(m, k, n, kind, out_dtype, scalar_a, scalar_b, use_bias, seed=0):
    """The reference test's recipe (tests/scaled_gemm_test.py:55-73), seeded on the CPU."""
    g = torch.Generator().manual_seed(seed)
    scale_a = torch.rand((1, 1), generator=g, dtype=torch.float32) if scalar_a else 0.25 * torch.rand((m, 1), generator=g)
    scale_b = torch.rand((1, 1), generator=g, dtype=torch.float32) if scalar_b else 0.25 * torch.rand((n, 1), generator=g)
    if kind == "fp8":
        a = (0.25 * torch.rand((m, k), generator=g, dtype=torch.float32)).to(FP8)
        b = (0.25 * torch.rand((n, k), generator=g, dtype=torch.float32)).to(FP8).T
    else:
        a = torch.randint(-32, 32, (m, k), generator=g, dtype=torch.int8)
        b = torch.randint(-32, 32, (n, k), generator=g, dtype=torch.int8).T
    bias = torch.rand((n,), generator=g).to(out_dtype) if use_bias else None
    return a, b, scale_a, scale_b, bias
