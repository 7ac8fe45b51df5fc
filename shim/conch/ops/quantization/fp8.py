"""conch.ops.quantization.fp8 -> conch_b200 (reference conch/ops/quantization/fp8.py:12-58)."""
from conch_b200.ops.quantization.fp8 import scaled_fp8_quant, static_scaled_fp8_quant  # noqa: F401
