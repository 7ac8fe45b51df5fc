"""conch.kernels.quantization.gemm -> conch_b200 (reference conch/kernels/quantization/gemm.py:20-133, 482-545)."""
from conch_b200.kernels.quantization.gemm import *  # noqa: F401,F403
