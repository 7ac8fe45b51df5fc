"""conch_b200: B200-native (sm_100a) drop-in for Conch's `mixed_precision_gemm`.

One operator, one backend: hand-written CUDA behind a C ABI (`libconch_b200.so`, see include/conch_b200.h).
There is no Triton path, no multi-backend dispatch and no CPU fallback: calling the op without the built
library or without a B200 raises.
"""

__version__ = "1.0.0"

from conch_b200.ops.quantization.gemm import (  # noqa: E402,F401
    create_mixed_precision_metadata,
    dequantize_weights,
    mixed_precision_gemm,
)
