"""FP8 (e4m3fn) quantisation of activations (the producer of scaled_gemm's `a`): the reference's conch/ops/quantization/fp8.py."""

from __future__ import annotations

import torch

from conch_b200.kernels.quantization.gemm import static_scaled_quant_launcher


def static_scaled_fp8_quant(output_tensor: torch.Tensor, input_tensor: torch.Tensor, scale: torch.Tensor) -> None:
    """output = e4m3fn(clamp(input * (1 / scale), -448, 448)), round to nearest even; one scale for the whole tensor
    (reference conch/ops/quantization/fp8.py:12-30)."""
    if output_tensor.shape != input_tensor.shape:
        raise ValueError(f"output shape {tuple(output_tensor.shape)} != input shape {tuple(input_tensor.shape)}")
    if scale.numel() != 1:
        raise ValueError("static quantisation takes ONE scale")
    if output_tensor.dtype != torch.float8_e4m3fn:
        raise ValueError(f"output_tensor must be float8_e4m3fn, got {output_tensor.dtype}")
    static_scaled_quant_launcher(output_tensor, input_tensor, scale)


def scaled_fp8_quant(input_tensor: torch.Tensor, scale: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Scaled fp8 quantisation; only the static form exists, as in the reference (conch/ops/quantization/fp8.py:33-58)."""
    if scale is None:
        raise NotImplementedError("Dynamic quantization not implemented yet")
    output_tensor = torch.empty_like(input_tensor, dtype=torch.float8_e4m3fn, memory_format=torch.contiguous_format)
    static_scaled_fp8_quant(output_tensor, input_tensor, scale)
    return output_tensor, scale
