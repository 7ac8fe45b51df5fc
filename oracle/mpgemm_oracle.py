"""CPU oracle for Conch's mixed-precision GEMM -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module restates, on the CPU, the algorithm of the reference hot path
(`conch.ops.quantization.gemm.mixed_precision_gemm`) so that the CUDA kernels can be
checked against it.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The shipped package
(`conch_b200/`) never does: there is no CPU fallback in the product.

Pinned against the reference itself (not "parity unpinned"): `oracle/make_golden.py`
imports the real reference from /root/reference in the build container, runs its
quantiser / packer / `w_ref` on seeded CPU inputs and commits the resulting tensors under
`tests/golden/`; `tests/test_oracle_golden.py` replays them through this file and demands
bit-equality (the reference ships no golden vectors of its own).

Reference lines restated (paths relative to /root/reference):
  * weight types ............ conch/third_party/vllm/scalar_type.py:153-169, 298-318
  * group quantiser ......... conch/third_party/vllm/quant_utils.py:14-101
  * K-direction packing ..... conch/third_party/vllm/quant_utils.py:104-125
  * unpack word/shift ....... conch/kernels/quantization/gemm.py:319-326
  * dequant op order ........ conch/kernels/quantization/gemm.py:176-216
  * group index ............. conch/kernels/quantization/gemm.py:329,354
  * mode / dtype defaults ... conch/ops/quantization/gemm.py:34-38, 113-130
  * GEMM (fp32 accumulate) .. conch/kernels/quantization/gemm.py:335,401,443-457
"""

from __future__ import annotations

import dataclasses

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# Weight types  (scalar_type.py:298-318: uint4, uint8, uint4b8, uint8b128)
# --------------------------------------------------------------------------------------


@dataclasses.dataclass(frozen=True)
class WeightType:
    """A stored-unsigned integer weight type: `bits` wide, logical value = stored - bias."""

    name: str
    bits: int
    bias: int

    @property
    def pack_factor(self) -> int:
        # quant_utils.py:9-11 -- always packed into 32-bit words
        return 32 // self.bits

    @property
    def qmin(self) -> int:
        # scalar_type.py:157-162 (min of the *logical* type: raw min 0, minus the bias)
        return -self.bias

    @property
    def qmax(self) -> int:
        # scalar_type.py:164-169
        return (1 << self.bits) - 1 - self.bias


WEIGHT_TYPES = {
    "uint4": WeightType("uint4", 4, 0),
    "uint8": WeightType("uint8", 8, 0),
    "uint4b8": WeightType("uint4b8", 4, 8),
    "uint8b128": WeightType("uint8b128", 8, 128),
}


# --------------------------------------------------------------------------------------
# Packing layout  (quant_utils.py:104-125 ; inverse = kernels/quantization/gemm.py:319-326)
# --------------------------------------------------------------------------------------


def pack_rows(w_q: torch.Tensor, bits: int) -> torch.Tensor:
    """[K, N] stored values -> [K/pf, N] int32, nibble/byte i of a word = row pf*kp + i (LSB first)."""
    pf = 32 // bits
    k, n = w_q.shape
    if k % pf:
        raise ValueError(f"K={k} is not a multiple of the pack factor {pf}")
    q = w_q.detach().cpu().numpy().astype(np.uint32).reshape(k // pf, pf, n)
    shifts = (np.arange(pf, dtype=np.uint32) * np.uint32(bits)).reshape(1, pf, 1)
    words = np.bitwise_or.reduce(q << shifts, axis=1)
    return torch.from_numpy(words.view(np.int32).copy())


def unpack_rows(w_q_packed: torch.Tensor, bits: int) -> torch.Tensor:
    """[K/pf, N] int32 -> [K, N] int32 stored values: (word >> bits*(k % pf)) & mask."""
    pf = 32 // bits
    words = w_q_packed.detach().cpu().contiguous().numpy().view(np.uint32)
    kp, n = words.shape
    shifts = (np.arange(pf, dtype=np.uint32) * np.uint32(bits)).reshape(1, pf, 1)
    q = (words[:, None, :] >> shifts) & np.uint32((1 << bits) - 1)
    return torch.from_numpy(q.reshape(kp * pf, n).astype(np.int32))


# --------------------------------------------------------------------------------------
# Quantiser  (quant_utils.py:14-101) -- used only to manufacture test inputs
# --------------------------------------------------------------------------------------


def quantize_weights(
    w: torch.Tensor, wtype: WeightType, group_size: int, zero_points: bool
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor | None]:
    """Group-wise min/max quantiser.  Returns (w_ref[K,N] dt, w_q[K,N] int32 stored, w_s[K/g,N] dt, w_zp[K/g,N] int32|None).

    All arithmetic happens in `w.dtype`, exactly as the reference does it, so scales and
    zero-points come out bit-identical to the reference's on the same CPU input.
    """
    k, n = w.shape
    dt = w.dtype
    g = k if group_size == -1 else group_size  # quant_utils.py:32-33
    if k % g:
        raise ValueError(f"K={k} not divisible by group_size={g}")
    wg = w.reshape(k // g, g, n)
    hi = wg.amax(dim=1, keepdim=True)  # :42
    lo = wg.amin(dim=1, keepdim=True)  # :43
    qmin, qmax = wtype.qmin, wtype.qmax
    if zero_points:
        # :51-54
        s = (hi - lo).clamp(min=1e-5) / qmax
        zp = torch.round(torch.abs(lo / s)).clamp(qmin, qmax).to(torch.int32)
    else:
        # :55-61 -- a zero bound is replaced by infinity so the quotient is 0
        d_hi = float(qmax) if qmax != 0 else float("inf")
        d_lo = float(qmin) if qmin != 0 else float("inf")
        s = torch.maximum((hi / d_hi).abs(), (lo / d_lo).abs())
        zp = None
    q = torch.round(wg / s).to(torch.int32)  # :64
    if zp is not None:
        q = q + zp
    q = q.clamp(qmin, qmax)  # :65
    centred = q - zp if zp is not None else q
    w_ref = centred.to(dt) * s  # :74
    q = q + wtype.bias  # :76-77
    return (
        w_ref.reshape(k, n).contiguous(),
        q.reshape(k, n).contiguous(),
        s.reshape(k // g, n).contiguous(),
        None if zp is None else zp.reshape(k // g, n).contiguous(),
    )


# --------------------------------------------------------------------------------------
# Dequantisation  (kernels/quantization/gemm.py:176-216)
# --------------------------------------------------------------------------------------


def dequantize(
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    bits: int,
    bias: int,
    group_size: int,
    dtype: torch.dtype,
) -> torch.Tensor:
    """W[k,n] = RN_dtype( (q - bias - zp[k//g,n]) * s[k//g,n] ), one rounding (the multiply)."""
    q = unpack_rows(w_q_packed, bits)  # :192
    k, n = q.shape
    g = k if group_size == -1 else group_size
    v = (q - bias).to(dtype)  # :192-194, exact small integer
    s = w_s.detach().cpu().to(dtype).reshape(-1, n)
    s = s.repeat_interleave(g, dim=0)[:k]  # group index k//g, :329,354
    if w_zp is None:
        return v * s  # SYMMETRIC_NO_SHIFT :201-202
    zp = w_zp.detach().cpu()
    if zp.numel() == 1:
        z = zp.reshape(()).to(dtype)  # scalar zero-point :207-208
        return (v - z).to(dtype) * s
    z = zp.reshape(-1, n).to(dtype).repeat_interleave(g, dim=0)[:k]  # :371 cast to meta dtype
    return (v - z) * s  # SYMMETRIC_WITH_SHIFT :209-210


def mixed_precision_gemm_ref(
    x: torch.Tensor,
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    weight_size_bits: int,
    weight_bias: int,
    group_size: int,
    output_dtype: torch.dtype | None = None,
) -> torch.Tensor:
    """out = cast( fp32-accumulated x @ dequant(W) ); same argument order as the reference op."""
    w = dequantize(w_q_packed, w_s, w_zp, weight_size_bits, weight_bias, group_size, x.dtype)
    out = x.detach().cpu().float() @ w.float()  # acc_dtype fp32, gemm.py:335,401
    return out.to(output_dtype or x.dtype)  # :443-457


def reference_test_expression(x: torch.Tensor, w_ref: torch.Tensor) -> torch.Tensor:
    """The reference test's own oracle, tests/mixed_precision_gemm_test.py:70 -- matmul in x.dtype."""
    return torch.matmul(x, w_ref)


# --------------------------------------------------------------------------------------
# Seeded input recipes
# --------------------------------------------------------------------------------------


@dataclasses.dataclass
class Problem:
    x: torch.Tensor
    w_q_packed: torch.Tensor
    w_s: torch.Tensor
    w_zp: torch.Tensor | None
    w_ref: torch.Tensor
    wtype: WeightType
    group_size: int

    @property
    def op_args(self):
        return (self.x, self.w_q_packed, self.w_s, self.w_zp, self.wtype.bits, self.wtype.bias, self.group_size)


def make_problem(
    m: int,
    k: int,
    n: int,
    wtype_name: str = "uint4",
    group_size: int = 128,
    dtype: torch.dtype = torch.float16,
    zero_points: bool = True,
    seed: int = 0,
    recipe: str = "llama",
) -> Problem:
    """Build a seeded problem on the CPU.

    recipe="reference": x, w = 10*(U[0,1) - 0.3)   (tests/mixed_precision_gemm_test.py:60-61)
    recipe="llama":     x = N(0,1), w = 0.02*N(0,1) (zero mean: no fp16 overflow at K=14336)
    """
    gen = torch.Generator().manual_seed(seed)
    if recipe == "reference":
        x = (10 * (torch.rand((m, k), generator=gen, dtype=torch.float32) - 0.3)).to(dtype)
        w = (10 * (torch.rand((k, n), generator=gen, dtype=torch.float32) - 0.3)).to(dtype)
    elif recipe == "llama":
        x = torch.randn((m, k), generator=gen, dtype=torch.float32).to(dtype)
        w = (0.02 * torch.randn((k, n), generator=gen, dtype=torch.float32)).to(dtype)
    else:
        raise ValueError(recipe)
    wtype = WEIGHT_TYPES[wtype_name]
    g = k if group_size == -1 else group_size
    w_ref, w_q, w_s, w_zp = quantize_weights(w, wtype, g, zero_points)
    return Problem(x, pack_rows(w_q, wtype.bits), w_s, w_zp, w_ref, wtype, group_size)
