"""Weight types, group quantiser and K-direction packer -- the producers of `mixed_precision_gemm`'s inputs.

The reference vendors these from vLLM as test helpers (conch/third_party/vllm/scalar_type.py:298-318,
conch/third_party/vllm/quant_utils.py:14-125); a user switching to conch_b200 needs the same producers, so they
are part of the package.  Pure torch, device-agnostic (runs on the GPU for Llama-sized weights), same
arithmetic in `w.dtype` as the reference so the tensors are interchangeable with the reference's.
"""

from __future__ import annotations

import dataclasses

import torch


@dataclasses.dataclass(frozen=True)
class ScalarType:
    """Unsigned stored integer of `size_bits` with an implicit `bias` (logical value = stored - bias)."""

    size_bits: int
    bias: int = 0

    def min(self) -> int:
        return -self.bias

    def max(self) -> int:
        return (1 << self.size_bits) - 1 - self.bias

    def has_bias(self) -> bool:
        return self.bias != 0

    def __str__(self) -> str:
        return f"uint{self.size_bits}" + (f"b{self.bias}" if self.bias else "")


class scalar_types:  # noqa: N801  (name kept from the reference for drop-in imports)
    uint4 = ScalarType(4, 0)
    uint8 = ScalarType(8, 0)
    uint4b8 = ScalarType(4, 8)
    uint8b128 = ScalarType(8, 128)


def get_pack_factor(num_bits: int) -> int:
    if 32 % num_bits:
        raise ValueError(f"Unsupported num_bits = {num_bits}")
    return 32 // num_bits


def quantize_weights(
    w: torch.Tensor,
    quant_type: ScalarType,
    group_size: int | None,
    zero_points: bool = False,
    ref_zero_points_after_scales: bool = False,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor | None]:
    """Min/max group quantiser.  Returns (w_ref, w_q, w_s, w_zp) like the reference helper
    (conch/third_party/vllm/quant_utils.py:14-101): w_ref [K,N] w.dtype, w_q [K,N] int32 stored values
    (bias added), w_s [K/g,N] w.dtype, w_zp [K/g,N] int32 or None.  ref_zero_points_after_scales computes w_ref as
    q * s - zp * s (two roundings, reference :71-74) instead of (q - zp) * s."""
    if not w.is_floating_point() or w.dim() != 2:
        raise ValueError("w must be a 2-D floating point tensor")
    size_k, size_n = w.shape
    g = size_k if group_size in (None, -1) else group_size
    if size_k % g:
        raise ValueError(f"group_size {g} must divide K={size_k}")
    groups = w.view(size_k // g, g, size_n) if w.is_contiguous() else w.reshape(size_k // g, g, size_n)
    w_max = groups.amax(dim=1, keepdim=True)
    w_min = groups.amin(dim=1, keepdim=True)
    q_lo, q_hi = quant_type.min(), quant_type.max()
    if zero_points:
        w_s = (w_max - w_min).clamp(min=1e-5) / q_hi
        w_zp = torch.round(torch.abs(w_min / w_s)).clamp(q_lo, q_hi).to(torch.int32)
        w_q = (torch.round(groups / w_s).to(torch.int32) + w_zp).clamp(q_lo, q_hi)
        if ref_zero_points_after_scales:
            w_ref = w_q.to(w.dtype) * w_s - w_zp.to(w.dtype) * w_s
        else:
            w_ref = (w_q - w_zp).to(w.dtype) * w_s
    else:
        hi_div = float(q_hi) if q_hi != 0 else float("inf")
        lo_div = float(q_lo) if q_lo != 0 else float("inf")
        w_s = torch.maximum((w_max / hi_div).abs(), (w_min / lo_div).abs())
        w_zp = None
        w_q = torch.round(groups / w_s).to(torch.int32).clamp(q_lo, q_hi)
        w_ref = w_q.to(w.dtype) * w_s
    w_q = w_q + quant_type.bias
    return (
        w_ref.reshape(size_k, size_n),
        w_q.reshape(size_k, size_n),
        w_s.reshape(size_k // g, size_n),
        None if w_zp is None else w_zp.reshape(size_k // g, size_n),
    )


def pack_rows(q_w: torch.Tensor, num_bits: int, size_k: int | None = None, size_n: int | None = None) -> torch.Tensor:
    """[K,N] stored values -> [K/pf,N] int32; value k = pf*row + i occupies bits [num_bits*i, num_bits*(i+1)).
    Same layout as the reference packer (conch/third_party/vllm/quant_utils.py:104-125), computed on q_w's device."""
    k, n = q_w.shape
    if (size_k is not None and size_k != k) or (size_n is not None and size_n != n):
        raise ValueError("q_w shape does not match size_k/size_n")
    pf = get_pack_factor(num_bits)
    if k % pf:
        raise ValueError(f"K={k} not divisible by pack factor {pf}")
    q = q_w.to(torch.int64).view(k // pf, pf, n) & ((1 << num_bits) - 1)
    shifts = (torch.arange(pf, device=q_w.device, dtype=torch.int64) * num_bits).view(1, pf, 1)
    words = (q << shifts).sum(dim=1)  # disjoint bit fields: sum == or
    words = torch.where(words >= (1 << 31), words - (1 << 32), words)  # reinterpret as signed int32
    return words.to(torch.int32)


def random_packed_weights(
    k: int, n: int, quant_type: ScalarType, group_size: int, dtype: torch.dtype, device, zero_points: bool = True,
    seed: int = 0, zp_dtype: torch.dtype = torch.int32,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor | None]:  # fmt: skip
    """Synthetic (w_q_packed, w_s, w_zp) drawn directly in the packed domain -- for timing runs where the
    values do not matter but the layout, dtypes and magnitudes (|W| ~ 0.02) do."""
    g = k if group_size == -1 else group_size
    gen = torch.Generator(device=device).manual_seed(seed)
    pf = get_pack_factor(quant_type.size_bits)
    packed = torch.randint(-(1 << 31), (1 << 31) - 1, (k // pf, n), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    w_s = (0.004 + 0.004 * torch.rand((k // g, n), generator=gen, device=device, dtype=torch.float32)).to(dtype)
    w_zp = None
    if zero_points:
        mid = (quant_type.max() + quant_type.min()) // 2
        w_zp = (mid + torch.randint(-2, 3, (k // g, n), generator=gen, device=device)).clamp(quant_type.min(), quant_type.max()).to(zp_dtype)
    return packed, w_s, w_zp
