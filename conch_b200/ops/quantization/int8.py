"""Int8 quantisation of activations (the producer of scaled_gemm's `a`): the reference's conch/ops/quantization/int8.py."""

from __future__ import annotations

import torch

from conch_b200.kernels.quantization.gemm import static_scaled_quant_launcher


def static_scaled_int8_quant(output_tensor: torch.Tensor, input_tensor: torch.Tensor, scale: torch.Tensor) -> None:
    """output = int8(clamp(input * (1 / scale), -128, 127)), truncating toward zero; one scale for the whole tensor
    (reference conch/ops/quantization/int8.py:11-26)."""
    if output_tensor.shape != input_tensor.shape:
        raise ValueError(f"output shape {tuple(output_tensor.shape)} != input shape {tuple(input_tensor.shape)}")
    if scale.numel() != 1:
        raise ValueError("static quantisation takes ONE scale")
    if output_tensor.dtype != torch.int8:
        raise ValueError(f"output_tensor must be int8, got {output_tensor.dtype}")
    static_scaled_quant_launcher(output_tensor, input_tensor, scale)


def scaled_int8_quant(input_tensor: torch.Tensor, scale: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Scaled int8 quantisation; only the static form exists, as in the reference (conch/ops/quantization/int8.py:29-48)."""
    if scale is None:
        raise NotImplementedError("Dynamic int8 quantization not yet implemented")
    output_tensor = torch.empty_like(input_tensor, dtype=torch.int8, memory_format=torch.contiguous_format)
    static_scaled_int8_quant(output_tensor, input_tensor, scale)
    return output_tensor, scale
