"""conch.ops.quantization.gemm -> conch_b200 (reference conch/ops/quantization/gemm.py:20-168)."""
from conch_b200.ops.quantization.gemm import (  # noqa: F401
    create_mixed_precision_metadata,
    create_scaled_metadata,
    dequantize_weights,
    mixed_precision_gemm,
    scaled_gemm,
)
