"""conch.ops.quantization.int8 -> conch_b200 (reference conch/ops/quantization/int8.py:11-48)."""
from conch_b200.ops.quantization.int8 import scaled_int8_quant, static_scaled_int8_quant  # noqa: F401
